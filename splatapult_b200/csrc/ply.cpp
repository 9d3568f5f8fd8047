// ply.cpp -- binary_little_endian PLY, vertex element only.
// Accept/reject behaviour follows the reference reader (src/ply.cpp:140-254): the first three non-comment
// lines must be exactly "ply", "format binary_little_endian 1.0", "element vertex <N>"; then only
// "property <type> <name>" lines until "end_header"; LF line ends (a stray '\r' makes a keyword line mismatch);
// no list properties; property offsets accumulate in declaration order; the body is N*vertexSize bytes.
// Held to the reference's own ply.cpp by a differential fuzz (tests/test_ref_pin.py), accidents included: see ParseHeader.
#include "splatapult_b200/ply.h"

#include <algorithm>
#include <exception>
#include <istream>
#include <ostream>
#include <sstream>

namespace
{
using Type = BinaryAttribute::Type;

struct TypeName { const char* longName; const char* alias; Type type; };
const TypeName kTypeNames[] = {
    {"char", "int8", Type::Char},     {"uchar", "uint8", Type::UChar}, {"short", "int16", Type::Short},
    {"ushort", "uint16", Type::UShort}, {"int", "int32", Type::Int},   {"uint", "uint32", Type::UInt},
    {"float", "float32", Type::Float}, {"double", "float64", Type::Double},
};

Type TypeFromToken(const std::string& tok)
{
    for (const TypeName& t : kTypeNames)
        if (tok == t.longName || tok == t.alias) return t.type;
    return Type::Unknown;
}

const char* TokenFromType(Type type)
{
    for (const TypeName& t : kTypeNames)
        if (t.type == type) return t.longName;
    return "unknown";
}

// next line that does not start with "comment"; false at EOF
bool NextLine(std::istream& in, std::string& line)
{
    while (std::getline(in, line))
        if (line.compare(0, 7, "comment") != 0) return true;
    return false;
}
}  // namespace

bool Ply::ParseHeader(std::istream& in)
{
    std::string line;
    if (!NextLine(in, line) || line != "ply") { error = "not a ply file"; return false; }
    if (!NextLine(in, line)) { error = "truncated header"; return false; }
    if (line != "format binary_little_endian 1.0")
    {
        error = (line == "format binary_big_endian 1.0") ? "only binary_little_endian is supported" : "expected a format line";
        return false;
    }
    if (!NextLine(in, line)) { error = "truncated header"; return false; }
    // One stream and three tokens for every header line, as in the reference (ply.cpp:176-209): an extraction that finds
    // nothing leaves its token AS IT WAS.  So a blank (or "\r"-only) line after a property line repeats that property -- its
    // size is added again, the name keeps its first offset -- and "property float" re-uses the previous name.  The reference's
    // own reader accepts such files and lays the vertex out that way (tests/test_ref_pin.py, differential fuzz).
    std::istringstream iss(line);
    std::string token1, token2, token3;
    if (!((iss >> token1 >> token2 >> vertexCount) && token1 == "element" && token2 == "vertex"))
    {
        error = "expected \"element vertex <count>\"";
        return false;
    }
    for (;;)
    {
        if (!NextLine(in, line)) { error = "missing end_header"; return false; }
        if (line == "end_header") return true;
        iss.str(line);
        iss.clear();
        iss >> token1 >> token2 >> token3;
        if (token1 != "property") { error = "expected a property line, got \"" + line + "\""; return false; }
        const Type type = TypeFromToken(token2);
        if (type == Type::Unknown) { error = "unsupported property type \"" + token2 + "\""; return false; }
        AddProperty(token3, type);
    }
}

bool Ply::Parse(std::istream& in)
{
    properties.clear();
    nameIndex.clear();
    vertexCount = vertexSize = 0;
    error.clear();
    if (!ParseHeader(in)) return false;
    // a header may claim any vertex count: the reference's `new uint8_t[size * count]` (ply.cpp:83) would throw out
    // of the application; a library reports it
    // (also with vertexSize == 0, i.e. no property lines: GaussianCloud::Allocate multiplies the count by ITS stride)
    constexpr size_t kMaxVertices = (size_t(1) << 31) - 1;  // the renderer refuses more than 2^31 - 1 gaussians anyway
    if (vertexCount > kMaxVertices || (vertexSize && vertexCount > static_cast<size_t>(-1) / vertexSize))
    {
        error = "vertex count " + std::to_string(vertexCount) + " is out of range (at most 2^31 - 1 vertices)";
        vertexCount = 0;
        return false;
    }
    try
    {
        AllocData(vertexCount);
    }
    catch (const std::exception&)
    {
        error = "cannot allocate " + std::to_string(vertexCount) + " vertices of " + std::to_string(vertexSize) + " bytes";
        vertexCount = 0;
        body.clear();
        return false;
    }
    in.read(reinterpret_cast<char*>(body.data()), static_cast<std::streamsize>(body.size()));
    return true;  // like the reference, a short body is not an error: missing bytes stay zero
}

void Ply::Dump(std::ostream& out) const
{
    out << "ply\nformat binary_little_endian 1.0\nelement vertex " << vertexCount << "\n";
    for (const Property& p : properties) out << "property " << TokenFromType(p.attrib.type) << " " << p.name << "\n";
    out << "end_header\n";
    out.write(reinterpret_cast<const char*>(body.data()), static_cast<std::streamsize>(body.size()));
}

bool Ply::GetProperty(const std::string& key, BinaryAttribute& attributeOut) const
{
    auto it = nameIndex.find(key);
    if (it == nameIndex.end()) return false;
    attributeOut = properties[it->second].attrib;
    return true;
}

void Ply::AddProperty(const std::string& key, BinaryAttribute::Type type)
{
    BinaryAttribute attrib(type, vertexSize);
    nameIndex.emplace(key, properties.size());  // emplace: a duplicate name keeps its first declaration
    properties.push_back({key, attrib});
    vertexSize += attrib.size;
}

void Ply::AllocData(size_t numVertices)
{
    vertexCount = numVertices;
    body.assign(vertexSize * numVertices, 0);
}

void Ply::ForEachVertex(const VertexCallback& cb) const
{
    const uint8_t* p = body.data();
    for (size_t i = 0; i < vertexCount; i++, p += vertexSize) cb(p, vertexSize);
}

void Ply::ForEachVertexMut(const VertexCallbackMut& cb)
{
    uint8_t* p = body.data();
    for (size_t i = 0; i < vertexCount; i++, p += vertexSize) cb(p, vertexSize);
}
