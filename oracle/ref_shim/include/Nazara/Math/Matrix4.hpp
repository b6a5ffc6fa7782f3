// oracle/ref_shim — stand-in for <Nazara/Math/Matrix4.hpp>.  Only named by DeformedChunk::ComputeCoordinates (block picking,
// not on the gen/mesh path), whose results are unused there; an opaque value type is enough.
#pragma once
#include <Nazara/Math/Vector3.hpp>

namespace Nz
{
	template<typename T> struct Matrix4
	{
		static Matrix4 Transform(const Vector3<T>&, const Quaternion<T>&) { return {}; }
		static Matrix4 TransformInverse(const Vector3<T>&, const Quaternion<T>&) { return {}; }
	};
	using Matrix4f = Matrix4<float>;
}
