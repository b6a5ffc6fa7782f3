// oracle/ref_shim — stand-in for <Nazara/Math/Box.hpp>.
// THIRD-PARTY SEMANTICS RESTATED HERE (unpinned): the corner naming.  The enum order is the one the reference
// lists at src/CommonLib/DeformedChunk.cpp:123-130; the label -> side mapping is the "effective" one derived in SURVEY.md §8.1
// (Top/Bottom <-> X with Top = X1, Near/Far <-> Y with Near = Y1, Left/Right <-> Z with Left = Z1), the only one under which
// Chunk::BuildMesh's face labels agree with its own visibility tests.
#pragma once
#include <Nazara/Math/Enums.hpp>
#include <Nazara/Math/Vector3.hpp>
#include <NazaraUtils/EnumArray.hpp>

namespace Nz
{
	template<typename T> struct Box
	{
		T x, y, z, width, height, depth;

		constexpr Box() : x(), y(), z(), width(), height(), depth() {}
		constexpr Box(T X, T Y, T Z, T W, T H, T D) : x(X), y(Y), z(Z), width(W), height(H), depth(D) {}
		constexpr Box(const Vector3<T>& pos, const Vector3<T>& lengths) : x(pos.x), y(pos.y), z(pos.z), width(lengths.x), height(lengths.y), depth(lengths.z) {}

		constexpr Vector3<T> GetCenter() const { return Vector3<T>(x + width / T(2), y + height / T(2), z + depth / T(2)); }
		constexpr Vector3<T> GetLengths() const { return Vector3<T>(width, height, depth); }
		constexpr Vector3<T> GetMinimum() const { return Vector3<T>(x, y, z); }
		constexpr Vector3<T> GetMaximum() const { return Vector3<T>(x + width, y + height, z + depth); }

		constexpr Vector3<T> GetCorner(BoxCorner corner) const
		{
			const T x1 = x + width, y1 = y + height, z1 = z + depth;
			switch (corner)
			{
				case BoxCorner::LeftBottomFar:   return Vector3<T>(x, y, z1);
				case BoxCorner::LeftTopFar:      return Vector3<T>(x1, y, z1);
				case BoxCorner::RightBottomFar:  return Vector3<T>(x, y, z);
				case BoxCorner::RightTopFar:     return Vector3<T>(x1, y, z);
				case BoxCorner::LeftBottomNear:  return Vector3<T>(x, y1, z1);
				case BoxCorner::LeftTopNear:     return Vector3<T>(x1, y1, z1);
				case BoxCorner::RightBottomNear: return Vector3<T>(x, y1, z);
				case BoxCorner::RightTopNear:    return Vector3<T>(x1, y1, z);
			}
			return Vector3<T>();
		}

		constexpr EnumArray<BoxCorner, Vector3<T>> GetCorners() const
		{
			return {
				GetCorner(BoxCorner::LeftBottomFar), GetCorner(BoxCorner::LeftTopFar), GetCorner(BoxCorner::RightBottomFar), GetCorner(BoxCorner::RightTopFar),
				GetCorner(BoxCorner::LeftBottomNear), GetCorner(BoxCorner::LeftTopNear), GetCorner(BoxCorner::RightBottomNear), GetCorner(BoxCorner::RightTopNear)
			};
		}
	};
	using Boxf = Box<float>;
}
