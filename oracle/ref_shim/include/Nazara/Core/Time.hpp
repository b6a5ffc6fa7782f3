// oracle/ref_shim — stand-in for <Nazara/Core/Time.hpp>
#pragma once
#include <NazaraUtils/Prerequisites.hpp>
namespace Nz
{
	struct Time
	{
		Int64 nanoseconds;
		static constexpr Time TickDuration(Int64 tickRate) { return Time{ 1'000'000'000ll / tickRate }; }
	};
}
