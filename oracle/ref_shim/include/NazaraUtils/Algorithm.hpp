// oracle/ref_shim — stand-in for <NazaraUtils/Algorithm.hpp>
#pragma once
#include <NazaraUtils/Prerequisites.hpp>
#include <limits>
#include <stdexcept>
#include <type_traits>

namespace Nz
{
	// Nz::SafeCast range-checks in debug builds only; out-of-range here means the reference would have asserted.
	// The harness wants to know (the GPU library reports TSOM_ERR_UNSUPPORTED for such planets), so it throws.
	struct SafeCastError : std::runtime_error
	{
		SafeCastError() : std::runtime_error("Nz::SafeCast out of range") {}
	};

	template<typename To, typename From> constexpr To SafeCast(From v)
	{
		if constexpr (std::is_integral_v<To> && std::is_integral_v<From>)
		{
			if constexpr (std::is_signed_v<From> && std::is_unsigned_v<To>)
			{
				if (v < 0)
					throw SafeCastError();
				if (static_cast<std::make_unsigned_t<From>>(v) > std::numeric_limits<To>::max())
					throw SafeCastError();
			}
			else if constexpr (std::is_unsigned_v<From> && std::is_signed_v<To>)
			{
				if (v > static_cast<std::make_unsigned_t<To>>(std::numeric_limits<To>::max()))
					throw SafeCastError();
			}
			else
			{
				if (v < std::numeric_limits<To>::lowest() || v > std::numeric_limits<To>::max())
					throw SafeCastError();
			}
		}
		return static_cast<To>(v);
	}

	template<typename From> struct SafeCaster
	{
		From value;
		constexpr SafeCaster(From v) : value(v) {}
		template<typename To> constexpr operator To() const { return SafeCast<To>(value); }
	};
	template<typename From> SafeCaster(From) -> SafeCaster<From>;

	template<typename T> constexpr T IntegralPow(T base, unsigned int exponent)
	{
		T r = 1;
		for (unsigned int i = 0; i < exponent; ++i)
			r *= base;
		return r;
	}

	template<typename T, typename T2> constexpr T Lerp(const T& from, const T& to, const T2& interpolation)
	{
		return static_cast<T>(from + interpolation * (to - from));
	}
}
