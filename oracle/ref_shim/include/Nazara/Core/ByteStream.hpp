// oracle/ref_shim — stand-in for <Nazara/Core/ByteStream.hpp>: a growable byte buffer with the stream operators
// Chunk::Serialize / Deserialize use (src/CommonLib/Chunk.cpp:252-346).
// THIRD-PARTY BEHAVIOUR RESTATED HERE (unpinned): Nazara's SerializationContext defaults to BIG-endian encoding of arithmetic
// types (bytes swapped on little-endian hosts); a std::string is a UInt32 length followed by its bytes; a Vector3 is its three
// components.  Define NZ_SHIM_LITTLE_ENDIAN to flip the byte order in this one place.
#pragma once
#include <Nazara/Math/Vector3.hpp>
#include <cstring>
#include <utility>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace Nz
{
	class ByteStream
	{
		public:
			std::vector<UInt8> data;
			std::size_t cursor = 0;

			void Write(const void* p, std::size_t n)
			{
				const UInt8* b = static_cast<const UInt8*>(p);
				data.insert(data.end(), b, b + n);
			}
			void Read(void* p, std::size_t n)
			{
				if (cursor + n > data.size())
					throw std::runtime_error("ByteStream: read past end");
				std::memcpy(p, data.data() + cursor, n);
				cursor += n;
			}
	};

	template<typename T> inline void ShimToStreamOrder(T& v)
	{
#ifndef NZ_SHIM_LITTLE_ENDIAN
		UInt8* b = reinterpret_cast<UInt8*>(&v);
		for (std::size_t i = 0; i < sizeof(T) / 2; ++i)
			std::swap(b[i], b[sizeof(T) - 1 - i]);
#endif
	}
	template<typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>> ByteStream& operator<<(ByteStream& s, T v) { ShimToStreamOrder(v); s.Write(&v, sizeof(T)); return s; }
	template<typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>> ByteStream& operator>>(ByteStream& s, T& v) { s.Read(&v, sizeof(T)); ShimToStreamOrder(v); return s; }
	inline ByteStream& operator<<(ByteStream& s, const std::string& str) { s << UInt32(str.size()); s.Write(str.data(), str.size()); return s; }
	inline ByteStream& operator>>(ByteStream& s, std::string& str) { UInt32 n; s >> n; str.resize(n); s.Read(str.data(), n); return s; }
	template<typename T> ByteStream& operator<<(ByteStream& s, const Vector3<T>& v) { return s << v.x << v.y << v.z; }
	template<typename T> ByteStream& operator>>(ByteStream& s, Vector3<T>& v) { return s >> v.x >> v.y >> v.z; }
}
