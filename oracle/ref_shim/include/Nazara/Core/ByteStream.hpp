// oracle/ref_shim — stand-in for <Nazara/Core/ByteStream.hpp>: a growable little-endian byte buffer with the stream operators
// Chunk::Serialize / Deserialize use (src/CommonLib/Chunk.cpp:252-346).  Integers are written little-endian (Nazara's default
// ByteStream endianness on x86); strings as UInt32 length + bytes.
#pragma once
#include <Nazara/Math/Vector3.hpp>
#include <cstring>
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

	template<typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>> ByteStream& operator<<(ByteStream& s, T v) { s.Write(&v, sizeof(T)); return s; }
	template<typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>> ByteStream& operator>>(ByteStream& s, T& v) { s.Read(&v, sizeof(T)); return s; }
	inline ByteStream& operator<<(ByteStream& s, const std::string& str) { s << UInt32(str.size()); s.Write(str.data(), str.size()); return s; }
	inline ByteStream& operator>>(ByteStream& s, std::string& str) { UInt32 n; s >> n; str.resize(n); s.Read(str.data(), n); return s; }
	template<typename T> ByteStream& operator<<(ByteStream& s, const Vector3<T>& v) { return s << v.x << v.y << v.z; }
	template<typename T> ByteStream& operator>>(ByteStream& s, Vector3<T>& v) { return s >> v.x >> v.y >> v.z; }
}
