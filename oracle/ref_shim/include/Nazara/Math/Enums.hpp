// oracle/ref_shim — stand-in for <Nazara/Math/Enums.hpp>.  Enum order as listed by the reference at
// src/CommonLib/DeformedChunk.cpp:123-130 (an EnumArray initialised in enum order).
#pragma once
namespace Nz
{
	enum class BoxCorner
	{
		LeftBottomFar,
		LeftTopFar,
		RightBottomFar,
		RightTopFar,
		LeftBottomNear,
		LeftTopNear,
		RightBottomNear,
		RightTopNear,

		Max = RightTopNear
	};

}
