// oracle/ref_shim — stand-in for <Nazara/Math/Vector3.hpp> (+ Vector2, Quaternion): only what the chunk pipeline uses.
// THIRD-PARTY ARITHMETIC RESTATED HERE (not pinned by any reference test): Normalize = multiply by
// 1/length when length > 0; Quaternion::RotationBetween and Quaternion * Vector3 as documented in SURVEY.md.
#pragma once
#include <Nazara/Math/Enums.hpp>
#include <NazaraUtils/Algorithm.hpp>
#include <NazaraUtils/Prerequisites.hpp>
#include <cmath>
#include <cstddef>
#include <functional>
#include <ostream>

namespace Nz
{
	template<typename T> struct Vector2
	{
		T x, y;
		constexpr Vector2() : x(), y() {}
		constexpr Vector2(T X, T Y) : x(X), y(Y) {}
		constexpr explicit Vector2(T v) : x(v), y(v) {}
		constexpr Vector2 operator*(T s) const { return Vector2(x * s, y * s); }
		constexpr Vector2 operator+(const Vector2& o) const { return Vector2(x + o.x, y + o.y); }
	};
	using Vector2f = Vector2<float>;

	template<typename T> struct Vector3
	{
		T x, y, z;

		constexpr Vector3() : x(), y(), z() {}
		constexpr Vector3(T X, T Y, T Z) : x(X), y(Y), z(Z) {}
		constexpr explicit Vector3(T v) : x(v), y(v), z(v) {}
		constexpr Vector3(const Vector2<T>& v, T Z) : x(v.x), y(v.y), z(Z) {}
		template<typename U> constexpr explicit Vector3(const Vector3<U>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
		// Nz::Vector3ui pos(indices.x, indices.z, indices.y) with float arguments (FlatChunk.cpp:42)
		template<typename A, typename B, typename C, typename = std::enable_if_t<!std::is_same_v<A, T> || !std::is_same_v<B, T> || !std::is_same_v<C, T>>>
		constexpr Vector3(A X, B Y, C Z) : x(static_cast<T>(X)), y(static_cast<T>(Y)), z(static_cast<T>(Z)) {}

		constexpr T& operator[](std::size_t i) { return i == 0 ? x : (i == 1 ? y : z); }
		constexpr const T& operator[](std::size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }

		constexpr Vector3 operator+(const Vector3& o) const { return Vector3(x + o.x, y + o.y, z + o.z); }
		constexpr Vector3 operator-(const Vector3& o) const { return Vector3(x - o.x, y - o.y, z - o.z); }
		constexpr Vector3 operator-() const { return Vector3(-x, -y, -z); }
		constexpr Vector3 operator*(const Vector3& o) const { return Vector3(x * o.x, y * o.y, z * o.z); }
		constexpr Vector3 operator*(T s) const { return Vector3(x * s, y * s, z * s); }
		constexpr Vector3 operator/(const Vector3& o) const { return Vector3(x / o.x, y / o.y, z / o.z); }
		constexpr Vector3 operator/(T s) const { return Vector3(x / s, y / s, z / s); }
		constexpr Vector3& operator+=(const Vector3& o) { x += o.x; y += o.y; z += o.z; return *this; }
		constexpr Vector3& operator-=(const Vector3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
		constexpr Vector3& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
		constexpr Vector3& operator/=(T s) { x /= s; y /= s; z /= s; return *this; }
		constexpr bool operator==(const Vector3& o) const { return x == o.x && y == o.y && z == o.z; }
		constexpr bool operator!=(const Vector3& o) const { return !(*this == o); }

		constexpr T DotProduct(const Vector3& o) const { return x * o.x + y * o.y + z * o.z; }
		constexpr Vector3 CrossProduct(const Vector3& o) const { return Vector3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
		constexpr Vector3 GetAbs() const { return Vector3(x < 0 ? -x : x, y < 0 ? -y : y, z < 0 ? -z : z); }
		T GetLength() const { return static_cast<T>(std::sqrt(GetSquaredLength())); }
		constexpr T GetSquaredLength() const { return x * x + y * y + z * z; }
		constexpr T SquaredDistance(const Vector3& o) const { return (o - *this).GetSquaredLength(); }
		constexpr Vector3& Maximize(const Vector3& o)
		{
			if (o.x > x) x = o.x;
			if (o.y > y) y = o.y;
			if (o.z > z) z = o.z;
			return *this;
		}
		Vector3& Normalize(T* length = nullptr)
		{
			T norm = GetLength();
			if (norm > T(0))
			{
				T invNorm = T(1) / norm;
				x *= invNorm;
				y *= invNorm;
				z *= invNorm;
			}
			if (length)
				*length = norm;
			return *this;
		}

		template<typename F> static constexpr Vector3 Apply(const Vector3& v, F&& f) { return Vector3(f(v.x), f(v.y), f(v.z)); }
		static constexpr Vector3 Clamp(const Vector3& v, const Vector3& mn, const Vector3& mx)
		{
			auto c = [](T a, T lo, T hi) { return a < lo ? lo : (a > hi ? hi : a); }; // std::clamp
			return Vector3(c(v.x, mn.x, mx.x), c(v.y, mn.y, mx.y), c(v.z, mn.z, mx.z));
		}
		static Vector3 Normalize(const Vector3& v) { Vector3 r = v; r.Normalize(); return r; }

		// Nazara's right-handed frame: Forward = -Z, Up = +Y, Right = +X
		static constexpr Vector3 Backward() { return Vector3(0, 0, 1); }
		static constexpr Vector3 Down() { return Vector3(0, -1, 0); }
		static constexpr Vector3 Forward() { return Vector3(0, 0, -1); }
		static constexpr Vector3 Left() { return Vector3(-1, 0, 0); }
		static constexpr Vector3 Right() { return Vector3(1, 0, 0); }
		static constexpr Vector3 Up() { return Vector3(0, 1, 0); }
		static constexpr Vector3 UnitX() { return Vector3(1, 0, 0); }
		static constexpr Vector3 UnitY() { return Vector3(0, 1, 0); }
		static constexpr Vector3 Zero() { return Vector3(0, 0, 0); }
	};

	template<typename T> constexpr Vector3<T> operator*(T s, const Vector3<T>& v) { return Vector3<T>(s * v.x, s * v.y, s * v.z); }
	template<typename T> std::ostream& operator<<(std::ostream& os, const Vector3<T>& v) { return os << "Vector3(" << v.x << ", " << v.y << ", " << v.z << ")"; }

	using Vector3f = Vector3<float>;
	using Vector3i = Vector3<int>;
	using Vector3ui = Vector3<unsigned int>;
	using Vector3i32 = Vector3<Int32>;

	template<typename T> struct Quaternion
	{
		T w, x, y, z;
		constexpr Quaternion() : w(1), x(0), y(0), z(0) {}
		constexpr Quaternion(T W, T X, T Y, T Z) : w(W), x(X), y(Y), z(Z) {}

		static constexpr Quaternion Identity() { return Quaternion(1, 0, 0, 0); }

		Quaternion& Normalize()
		{
			T norm = std::sqrt(w * w + x * x + y * y + z * z);
			if (norm > T(0))
			{
				T inv = T(1) / norm;
				w *= inv; x *= inv; y *= inv; z *= inv;
			}
			return *this;
		}

		// angle (radians) around a normalised axis
		static Quaternion FromAxisAngle(const Vector3<T>& axis, T angle)
		{
			T half = angle / T(2);
			T s = std::sin(half);
			Vector3<T> a = Vector3<T>::Normalize(axis);
			return Quaternion(std::cos(half), a.x * s, a.y * s, a.z * s);
		}

		static Quaternion RotationBetween(const Vector3<T>& from, const Vector3<T>& to)
		{
			T dot = from.DotProduct(to);
			if (dot < T(-0.999999))
			{
				Vector3<T> crossProduct;
				if (from.DotProduct(Vector3<T>::UnitX()) < T(0.999999))
					crossProduct = Vector3<T>::UnitX().CrossProduct(from);
				else
					crossProduct = Vector3<T>::UnitY().CrossProduct(from);
				return FromAxisAngle(crossProduct, T(3.14159265358979323846));
			}
			else if (dot > T(0.999999))
				return Identity();
			else
			{
				T norm = std::sqrt(from.GetSquaredLength() * to.GetSquaredLength());
				Vector3<T> c = from.CrossProduct(to);
				Quaternion q(norm + dot, c.x, c.y, c.z);
				q.Normalize();
				return q;
			}
		}

		Vector3<T> operator*(const Vector3<T>& vec) const
		{
			Vector3<T> quatVec(x, y, z);
			Vector3<T> uv = quatVec.CrossProduct(vec);
			Vector3<T> uuv = quatVec.CrossProduct(uv);
			uv *= T(2) * w;
			uuv *= T(2);
			return vec + uv + uuv;
		}
	};
	using Quaternionf = Quaternion<float>;
}

namespace std
{
	template<typename T> struct hash<Nz::Vector3<T>>
	{
		size_t operator()(const Nz::Vector3<T>& v) const
		{
			size_t h = std::hash<T>()(v.x);
			h ^= std::hash<T>()(v.y) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
			h ^= std::hash<T>()(v.z) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
			return h;
		}
	};
}
