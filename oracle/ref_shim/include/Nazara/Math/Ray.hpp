// oracle/ref_shim — stand-in for <Nazara/Math/Ray.hpp>.  Only DeformedChunk::ComputeCoordinates (block picking, NOT on the
// gen/mesh path) uses it; a plain slab test is enough for the file to link.
#pragma once
#include <Nazara/Math/Box.hpp>
#include <algorithm>
#include <limits>

namespace Nz
{
	template<typename T> struct Ray
	{
		Vector3<T> origin, direction;
		Ray(const Vector3<T>& o, const Vector3<T>& d) : origin(o), direction(d) {}
		Vector3<T> GetPoint(T lambda) const { return origin + direction * lambda; }
		bool Intersect(const Box<T>& box, T* closestHit = nullptr, T* furthestHit = nullptr) const
		{
			T tmin = T(0), tmax = std::numeric_limits<T>::infinity();
			Vector3<T> mn = box.GetMinimum(), mx = box.GetMaximum();
			for (std::size_t i = 0; i < 3; ++i)
			{
				if (direction[i] == T(0))
				{
					if (origin[i] < mn[i] || origin[i] > mx[i])
						return false;
					continue;
				}
				T inv = T(1) / direction[i];
				T t1 = (mn[i] - origin[i]) * inv, t2 = (mx[i] - origin[i]) * inv;
				if (t1 > t2)
					std::swap(t1, t2);
				tmin = std::max(tmin, t1);
				tmax = std::min(tmax, t2);
				if (tmin > tmax)
					return false;
			}
			if (closestHit)
				*closestHit = tmin;
			if (furthestHit)
				*furthestHit = tmax;
			return true;
		}
	};
	using Rayf = Ray<float>;
}
