// oracle/ref_shim — stand-in for <Nazara/Core/Color.hpp> (debug colours of Direction.hpp)
#pragma once
namespace Nz
{
	struct Color
	{
		float r, g, b, a;
		constexpr Color(float R = 1.f, float G = 1.f, float B = 1.f, float A = 1.f) : r(R), g(G), b(B), a(A) {}
		static constexpr Color Blue() { return Color(0.f, 0.f, 1.f); }
		static constexpr Color Cyan() { return Color(0.f, 1.f, 1.f); }
		static constexpr Color Gray() { return Color(0.5f, 0.5f, 0.5f); }
		static constexpr Color Green() { return Color(0.f, 1.f, 0.f); }
		static constexpr Color Magenta() { return Color(1.f, 0.f, 1.f); }
		static constexpr Color Red() { return Color(1.f, 0.f, 0.f); }
		static constexpr Color White() { return Color(1.f, 1.f, 1.f); }
	};
}
