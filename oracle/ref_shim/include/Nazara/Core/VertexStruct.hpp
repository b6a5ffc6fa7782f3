// oracle/ref_shim — stand-in for <Nazara/Core/VertexStruct.hpp> (forward-declared only by the chunk pipeline)
#pragma once
#include <Nazara/Core/Color.hpp>
#include <Nazara/Math/Vector3.hpp>
namespace Nz
{
	struct VertexStruct_XYZ_Color_UV
	{
		Vector3f position;
		Color color;
		Vector2f uv;
	};
}
