// binaryattribute.h -- typed (type, size, offset) view of one field inside an interleaved record.
// Same public surface as the reference's BinaryAttribute (src/core/binaryattribute.h:12-111): Type enum,
// Get / Read / Write / ForEach / ForEachMut and the three public fields.
#pragma once
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <functional>

class BinaryAttribute
{
public:
    enum class Type { Unknown, Char, UChar, Short, UShort, Int, UInt, Float, Double, NumTypes };

    static size_t SizeOf(Type t)
    {
        switch (t)
        {
        case Type::Char: case Type::UChar: return 1;
        case Type::Short: case Type::UShort: return 2;
        case Type::Int: case Type::UInt: case Type::Float: return 4;
        case Type::Double: return 8;
        default: return 0;
        }
    }

    BinaryAttribute() = default;
    BinaryAttribute(Type typeIn, size_t offsetIn) : type(typeIn), size(SizeOf(typeIn)), offset(offsetIn) {}

    bool Valid() const { return type != Type::Unknown; }

    // pointer to this field inside the record that starts at `record`; nullptr for an Unknown attribute
    template <typename T> const T* Get(const void* record) const
    {
        if (!Valid()) return nullptr;
        assert(size == sizeof(T));
        return reinterpret_cast<const T*>(static_cast<const uint8_t*>(record) + offset);
    }
    template <typename T> T* Get(void* record) const
    {
        return const_cast<T*>(Get<T>(static_cast<const void*>(record)));
    }

    // value of the field, or T(0) when the attribute is Unknown (a missing PLY property reads as zero)
    template <typename T> T Read(const void* record) const
    {
        const T* p = Get<T>(record);
        return p ? *p : T(0);
    }
    template <typename T> bool Write(void* record, const T& value) const
    {
        T* p = Get<T>(record);
        if (p) *p = value;
        return p != nullptr;
    }

    template <typename T> void ForEach(const void* base, size_t stride, size_t count, const std::function<void(const T*)>& cb) const
    {
        assert(Valid() && base);
        const uint8_t* rec = static_cast<const uint8_t*>(base);
        for (size_t i = 0; i < count; i++, rec += stride) cb(reinterpret_cast<const T*>(rec + offset));
    }
    template <typename T> void ForEachMut(void* base, size_t stride, size_t count, const std::function<void(T*)>& cb) const
    {
        assert(Valid() && base);
        uint8_t* rec = static_cast<uint8_t*>(base);
        for (size_t i = 0; i < count; i++, rec += stride) cb(reinterpret_cast<T*>(rec + offset));
    }

    Type type = Type::Unknown;
    size_t size = 0;
    size_t offset = 0;
};
