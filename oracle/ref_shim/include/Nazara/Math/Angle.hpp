// oracle/ref_shim — stand-in for <Nazara/Math/Angle.hpp>: only what include/CommonLib/GameConstants.hpp needs to compile
// (constexpr angle constants that the chunk pipeline never reads).
#pragma once

namespace Nz
{
	template<typename T> struct DegreeAngle
	{
		T value = T(0);
		constexpr DegreeAngle() = default;
		constexpr DegreeAngle(T v) : value(v) {}
	};
	using DegreeAnglef = DegreeAngle<float>;
}
