// oracle/ref_shim — TEST INFRASTRUCTURE.  Stand-in for <NazaraUtils/Prerequisites.hpp> (NazaraEngine is not on this box).
// Only what the reference's chunk pipeline sources use.  Written from the public Nazara API, not copied from it.
#pragma once
#include <cassert>
#include <cstddef>
#include <cstdint>

#define NAZARA_EXPORT
#define NAZARA_IMPORT
#define NazaraAssert(cond, msg) assert((cond) && msg)
#define NazaraUnused(x) (void)(x)

namespace Nz
{
	using Int8 = std::int8_t;
	using UInt8 = std::uint8_t;
	using Int16 = std::int16_t;
	using UInt16 = std::uint16_t;
	using Int32 = std::int32_t;
	using UInt32 = std::uint32_t;
	using Int64 = std::int64_t;
	using UInt64 = std::uint64_t;
}
