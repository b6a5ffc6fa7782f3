// oracle/ref_shim — stand-in for <PerlinNoise.hpp> (siv::PerlinNoise v3.0.0, xmake package `perlinnoise`, un-vendored).
// THIRD-PARTY ARITHMETIC RESTATED HERE (unpinned): the published algorithm of Reputeless/PerlinNoise v3:
// reseed = iota + the library's own shuffle over std::mt19937, improved-noise 3D with z = 0.34567 for the 2D entry points,
// octaves with persistence 0.5, normalisation by the amplitude sum.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <random>

namespace siv
{
	class PerlinNoise
	{
		public:
			using value_type = double;
			using seed_type = std::mt19937::result_type;

			PerlinNoise() { reseed(std::mt19937::default_seed); }
			explicit PerlinNoise(seed_type seed) { reseed(seed); }

			void reseed(seed_type seed)
			{
				std::iota(m_permutation.begin(), m_permutation.end(), std::uint8_t(0));
				std::mt19937 urbg{ seed };
				// the library's own Fisher-Yates variant (not std::shuffle): element i swaps with urbg() % (i + 1)
				const std::size_t n = m_permutation.size();
				for (std::size_t i = 1; i < n; ++i)
					std::swap(m_permutation[i], m_permutation[static_cast<std::uint64_t>(urbg()) % (i + 1)]);
			}

			value_type noise3D(value_type x, value_type y, value_type z) const
			{
				const value_type _x = std::floor(x), _y = std::floor(y), _z = std::floor(z);
				const std::int32_t ix = static_cast<std::int32_t>(_x) & 255;
				const std::int32_t iy = static_cast<std::int32_t>(_y) & 255;
				const std::int32_t iz = static_cast<std::int32_t>(_z) & 255;
				const value_type fx = x - _x, fy = y - _y, fz = z - _z;
				const value_type u = Fade(fx), v = Fade(fy), w = Fade(fz);

				const std::uint8_t A = (m_permutation[ix & 255] + iy) & 255;
				const std::uint8_t B = (m_permutation[(ix + 1) & 255] + iy) & 255;
				const std::uint8_t AA = (m_permutation[A] + iz) & 255;
				const std::uint8_t AB = (m_permutation[(A + 1) & 255] + iz) & 255;
				const std::uint8_t BA = (m_permutation[B] + iz) & 255;
				const std::uint8_t BB = (m_permutation[(B + 1) & 255] + iz) & 255;

				const value_type p0 = Grad(m_permutation[AA], fx, fy, fz);
				const value_type p1 = Grad(m_permutation[BA], fx - 1, fy, fz);
				const value_type p2 = Grad(m_permutation[AB], fx, fy - 1, fz);
				const value_type p3 = Grad(m_permutation[BB], fx - 1, fy - 1, fz);
				const value_type p4 = Grad(m_permutation[(AA + 1) & 255], fx, fy, fz - 1);
				const value_type p5 = Grad(m_permutation[(BA + 1) & 255], fx - 1, fy, fz - 1);
				const value_type p6 = Grad(m_permutation[(AB + 1) & 255], fx, fy - 1, fz - 1);
				const value_type p7 = Grad(m_permutation[(BB + 1) & 255], fx - 1, fy - 1, fz - 1);

				const value_type q0 = Lerp(p0, p1, u), q1 = Lerp(p2, p3, u), q2 = Lerp(p4, p5, u), q3 = Lerp(p6, p7, u);
				const value_type r0 = Lerp(q0, q1, v), r1 = Lerp(q2, q3, v);
				return Lerp(r0, r1, w);
			}

			value_type noise2D(value_type x, value_type y) const { return noise3D(x, y, static_cast<value_type>(0.34567)); }

			value_type octave2D(value_type x, value_type y, std::int32_t octaves, value_type persistence = value_type(0.5)) const
			{
				value_type result = 0, amplitude = 1;
				for (std::int32_t i = 0; i < octaves; ++i)
				{
					result += noise2D(x, y) * amplitude;
					x *= 2;
					y *= 2;
					amplitude *= persistence;
				}
				return result;
			}

			value_type normalizedOctave2D(value_type x, value_type y, std::int32_t octaves, value_type persistence = value_type(0.5)) const
			{
				return octave2D(x, y, octaves, persistence) / MaxAmplitude(octaves, persistence);
			}

			value_type normalizedOctave2D_01(value_type x, value_type y, std::int32_t octaves, value_type persistence = value_type(0.5)) const
			{
				return normalizedOctave2D(x, y, octaves, persistence) * value_type(0.5) + value_type(0.5);
			}

		private:
			static value_type Fade(value_type t) { return t * t * t * (t * (t * 6 - 15) + 10); }
			static value_type Lerp(value_type a, value_type b, value_type t) { return a + (b - a) * t; }
			static value_type Grad(std::uint8_t hash, value_type x, value_type y, value_type z)
			{
				const std::uint8_t h = hash & 15;
				const value_type u = h < 8 ? x : y;
				const value_type v = h < 4 ? y : ((h == 12 || h == 14) ? x : z);
				return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
			}
			static value_type MaxAmplitude(std::int32_t octaves, value_type persistence)
			{
				value_type result = 0, amplitude = 1;
				for (std::int32_t i = 0; i < octaves; ++i)
				{
					result += amplitude;
					amplitude *= persistence;
				}
				return result;
			}

			std::array<std::uint8_t, 256> m_permutation;
	};
}
