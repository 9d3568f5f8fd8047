// Minimal stand-in for the parts of nlohmann/json 3.11 that the reference's src/camerasconfig.cpp uses:
// json::parse(std::istream&), range-for over an array, operator[] with a key or an index on a NON-const value (a missing
// key / index yields null, as in nlohmann), get<int>() / get<float>() (a number converts with static_cast; anything else
// throws type_error), json::exception.  TEST INFRASTRUCTURE for oracle/_ref only.
#pragma once
#include <cstdint>
#include <cerrno>
#include <cstdlib>
#include <istream>
#include <iterator>
#include <map>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace nlohmann
{
class json
{
public:
    class exception : public std::exception
    {
    public:
        explicit exception(const std::string& m) : msg(m) {}
        const char* what() const noexcept override { return msg.c_str(); }

    private:
        std::string msg;
    };
    class parse_error : public exception { public: using exception::exception; };
    class type_error : public exception { public: using exception::exception; };

    enum class kind { null, boolean, number_integer, number_unsigned, number_float, string, array, object };

    json() = default;

    static json parse(std::istream& in)
    {
        std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        size_t p = 0;
        json v = parse_value(text, p, 0);
        skip_ws(text, p);
        if (p != text.size()) throw parse_error("[json.exception.parse_error.101] unexpected trailing characters");
        return v;
    }

    // non-const access, nlohmann semantics: null becomes an object / array, a missing member is created as null
    json& operator[](const char* key) { return (*this)[std::string(key)]; }
    json& operator[](const std::string& key)
    {
        if (k == kind::null) k = kind::object;
        if (k != kind::object) throw type_error("[json.exception.type_error.305] cannot use operator[] with a string argument");
        return obj[key];
    }
    json& operator[](size_t idx)
    {
        if (k == kind::null) k = kind::array;
        if (k != kind::array) throw type_error("[json.exception.type_error.305] cannot use operator[] with a numeric argument");
        if (idx >= arr.size()) arr.resize(idx + 1);
        return arr[idx];
    }
    json& operator[](int idx) { return (*this)[(size_t)idx]; }

    template <typename T> T get() const
    {
        static_assert(std::is_arithmetic<T>::value, "stand-in: arithmetic get<T>() only");
        switch (k)
        {
        case kind::number_integer: return static_cast<T>(i);
        case kind::number_unsigned: return static_cast<T>(u);
        case kind::number_float: return static_cast<T>(d);
        case kind::boolean: return static_cast<T>(b);
        default: throw type_error("[json.exception.type_error.302] type must be number");
        }
    }

    // range-for: array elements (an object iterates its values, a primitive itself once, null nothing -- as nlohmann does)
    std::vector<json>::iterator begin() { flatten(); return view->begin(); }
    std::vector<json>::iterator end() { flatten(); return view->end(); }

private:
    kind k = kind::null;
    bool b = false;
    int64_t i = 0;
    uint64_t u = 0;
    double d = 0.0;
    std::string s;
    std::vector<json> arr;
    std::map<std::string, json> obj;
    std::vector<json> tmp;
    std::vector<json>* view = nullptr;

    void flatten()
    {
        if (k == kind::array) { view = &arr; return; }
        tmp.clear();
        if (k == kind::object) for (auto& kv : obj) tmp.push_back(kv.second);
        else if (k != kind::null) { json c; c.k = k; c.b = b; c.i = i; c.u = u; c.d = d; c.s = s; tmp.push_back(c); }
        view = &tmp;
    }
    static void skip_ws(const std::string& t, size_t& p) { while (p < t.size() && (t[p] == ' ' || t[p] == '\t' || t[p] == '\n' || t[p] == '\r')) p++; }
    [[noreturn]] static void bad(const char* what) { throw parse_error(std::string("[json.exception.parse_error.101] ") + what); }

    static std::string parse_string(const std::string& t, size_t& p)
    {
        std::string out;
        p++;  // opening quote
        for (;;)
        {
            if (p >= t.size()) bad("unterminated string");
            const unsigned char c = (unsigned char)t[p++];
            if (c == '"') return out;
            if (c < 0x20) bad("control character in string");
            if (c != '\\') { out.push_back((char)c); continue; }
            if (p >= t.size()) bad("unterminated escape");
            const char e = t[p++];
            switch (e)
            {
            case '"': out.push_back('"'); break;
            case '\\': out.push_back('\\'); break;
            case '/': out.push_back('/'); break;
            case 'b': out.push_back('\b'); break;
            case 'f': out.push_back('\f'); break;
            case 'n': out.push_back('\n'); break;
            case 'r': out.push_back('\r'); break;
            case 't': out.push_back('\t'); break;
            case 'u':
            {
                if (p + 4 > t.size()) bad("short \\u escape");
                unsigned cp = 0;
                for (int j = 0; j < 4; j++)
                {
                    const char h = t[p++];
                    cp <<= 4;
                    if (h >= '0' && h <= '9') cp |= (unsigned)(h - '0');
                    else if (h >= 'a' && h <= 'f') cp |= (unsigned)(h - 'a' + 10);
                    else if (h >= 'A' && h <= 'F') cp |= (unsigned)(h - 'A' + 10);
                    else bad("bad \\u escape");
                }
                if (cp < 0x80) out.push_back((char)cp);
                else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
                else { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
                break;
            }
            default: bad("bad escape");
            }
        }
    }

    static json parse_number(const std::string& t, size_t& p)
    {
        const size_t start = p;
        bool isFloat = false;
        if (p < t.size() && t[p] == '-') p++;
        if (p >= t.size() || t[p] < '0' || t[p] > '9') bad("bad number");
        if (t[p] == '0') p++;
        else while (p < t.size() && t[p] >= '0' && t[p] <= '9') p++;
        if (p < t.size() && t[p] == '.')
        {
            isFloat = true;
            p++;
            if (p >= t.size() || t[p] < '0' || t[p] > '9') bad("bad fraction");
            while (p < t.size() && t[p] >= '0' && t[p] <= '9') p++;
        }
        if (p < t.size() && (t[p] == 'e' || t[p] == 'E'))
        {
            isFloat = true;
            p++;
            if (p < t.size() && (t[p] == '+' || t[p] == '-')) p++;
            if (p >= t.size() || t[p] < '0' || t[p] > '9') bad("bad exponent");
            while (p < t.size() && t[p] >= '0' && t[p] <= '9') p++;
        }
        const std::string tok = t.substr(start, p - start);
        json v;
        if (!isFloat)
        {
            errno = 0;
            if (tok[0] == '-')
            {
                const long long x = std::strtoll(tok.c_str(), nullptr, 10);
                if (errno == 0) { v.k = kind::number_integer; v.i = x; return v; }
            }
            else
            {
                const unsigned long long x = std::strtoull(tok.c_str(), nullptr, 10);
                if (errno == 0) { v.k = kind::number_unsigned; v.u = x; return v; }
            }
        }
        v.k = kind::number_float;  // nlohmann: strtod of the token (also integers that overflow 64 bits)
        v.d = std::strtod(tok.c_str(), nullptr);
        return v;
    }

    static json parse_value(const std::string& t, size_t& p, int depth)
    {
        if (depth > 512) bad("nesting too deep");
        skip_ws(t, p);
        if (p >= t.size()) bad("unexpected end of input");
        json v;
        const char c = t[p];
        if (c == '{')
        {
            v.k = kind::object;
            p++;
            skip_ws(t, p);
            if (p < t.size() && t[p] == '}') { p++; return v; }
            for (;;)
            {
                skip_ws(t, p);
                if (p >= t.size() || t[p] != '"') bad("object key must be a string");
                const std::string key = parse_string(t, p);
                skip_ws(t, p);
                if (p >= t.size() || t[p] != ':') bad("expected ':'");
                p++;
                v.obj[key] = parse_value(t, p, depth + 1);  // a repeated key keeps the LAST value, as nlohmann does
                skip_ws(t, p);
                if (p < t.size() && t[p] == ',') { p++; continue; }
                if (p < t.size() && t[p] == '}') { p++; return v; }
                bad("expected ',' or '}'");
            }
        }
        if (c == '[')
        {
            v.k = kind::array;
            p++;
            skip_ws(t, p);
            if (p < t.size() && t[p] == ']') { p++; return v; }
            for (;;)
            {
                v.arr.push_back(parse_value(t, p, depth + 1));
                skip_ws(t, p);
                if (p < t.size() && t[p] == ',') { p++; continue; }
                if (p < t.size() && t[p] == ']') { p++; return v; }
                bad("expected ',' or ']'");
            }
        }
        if (c == '"') { v.k = kind::string; v.s = parse_string(t, p); return v; }
        if (t.compare(p, 4, "true") == 0) { p += 4; v.k = kind::boolean; v.b = true; return v; }
        if (t.compare(p, 5, "false") == 0) { p += 5; v.k = kind::boolean; v.b = false; return v; }
        if (t.compare(p, 4, "null") == 0) { p += 4; return v; }
        if (c == '-' || (c >= '0' && c <= '9')) return parse_number(t, p);
        bad("unexpected character");
    }
};
}  // namespace nlohmann
