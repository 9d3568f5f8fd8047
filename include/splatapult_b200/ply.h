// ply.h -- binary_little_endian PLY reader/writer, vertex element only.
// Same surface and accept/reject behaviour as the reference's Ply (src/ply.h:19-46, src/ply.cpp:72-281).
#pragma once
#include <cstdint>
#include <functional>
#include <iosfwd>
#include <string>
#include <unordered_map>
#include <vector>

#include "binaryattribute.h"
#include "splatmath.h"  // SPLAT_CPP_API

class SPLAT_CPP_API Ply
{
public:
    Ply() = default;

    bool Parse(std::istream& plyFile);       // header + one bulk read of the body
    void Dump(std::ostream& plyFile) const;  // header (properties in offset order) + body

    bool GetProperty(const std::string& key, BinaryAttribute& attributeOut) const;
    void AddProperty(const std::string& key, BinaryAttribute::Type type);
    void AllocData(size_t numVertices);

    using VertexCallback = std::function<void(const void*, size_t)>;
    void ForEachVertex(const VertexCallback& cb) const;
    using VertexCallbackMut = std::function<void(void*, size_t)>;
    void ForEachVertexMut(const VertexCallbackMut& cb);

    size_t GetVertexCount() const { return vertexCount; }
    size_t GetVertexSize() const { return vertexSize; }
    const uint8_t* GetData() const { return body.data(); }
    uint8_t* GetData() { return body.data(); }
    const std::string& GetError() const { return error; }

protected:
    bool ParseHeader(std::istream& plyFile);

    struct Property { std::string name; BinaryAttribute attrib; };
    std::vector<Property> properties;                  // declaration order == offset order
    std::unordered_map<std::string, size_t> nameIndex;  // first declaration of a name wins
    std::vector<uint8_t> body;
    size_t vertexCount = 0;
    size_t vertexSize = 0;
    std::string error;
};
