// oracle/tsom_oracle.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of ThisSpaceOfMine's voxel chunk pipeline (generation + face-culled
// meshing + collider triangles + edit/dirty rules).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build, load or call this.
// The shipped path (thisspaceofmine_b200/csrc) never includes or links it.
//
// PARITY STATUS: "parity unpinned" for generation, meshing and collider output — the
// reference holds no golden vectors for them and cannot be compiled here
// (needs xmake + NazaraEngine + siv::PerlinNoise + tsl::hopscotch_map, none on this box).
// The only known-answer vectors the reference ships (tests/UnitTests/Common/ChunkTests.cpp:17-43,
// block-index math) ARE checked against this file (tests/test_oracle_kat.py).
//
// Third-party arithmetic restated here (not present under /root/reference):
//  * siv::PerlinNoise v3.0.0 (xmake package "perlinnoise", unpinned => latest; API used by the
//    reference — reseed(seed), normalizedOctave2D_01(x,y,octaves) — exists only in v3).
//  * Nazara math (nazaraengine >= 2023.08.15): Vector3f, Quaternionf::RotationBetween,
//    Quaternionf * Vector3f, Boxf::GetCorner(s).
//  * libstdc++ <random> is used DIRECTLY (std::minstd_rand, std::bernoulli_distribution,
//    std::mt19937), matching the reference's Linux build.
//
// Build with: g++ -std=c++17 -O2 -ffp-contract=off   (no FMA contraction: discrete decisions
// such as DirectionFromNormal ties and std::round must not depend on fusing).
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <numeric>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace orc
{
	// ---------------------------------------------------------------- BlockIndex.hpp:15-18
	using BlockIndex = std::uint16_t;
	constexpr BlockIndex EmptyBlockIndex = 0;
	constexpr BlockIndex InvalidBlockIndex = 0xFFFF;

	// ChunkContainer.hpp:42 / Planet.hpp
	constexpr unsigned int ChunkSize = 32;
	constexpr unsigned int ChunkBlockCount = ChunkSize * ChunkSize * ChunkSize;

	// ---------------------------------------------------------------- minimal Nazara math stand-in
	struct Vec3i
	{
		std::int32_t x = 0, y = 0, z = 0;
		std::int32_t& operator[](unsigned i) { return i == 0 ? x : (i == 1 ? y : z); }
		std::int32_t operator[](unsigned i) const { return i == 0 ? x : (i == 1 ? y : z); }
		bool operator==(const Vec3i& o) const { return x == o.x && y == o.y && z == o.z; }
		bool operator!=(const Vec3i& o) const { return !(*this == o); }
		bool operator<(const Vec3i& o) const { return std::tie(x, y, z) < std::tie(o.x, o.y, o.z); }
	};

	struct Vec3ui
	{
		unsigned int x = 0, y = 0, z = 0;
		unsigned int& operator[](unsigned i) { return i == 0 ? x : (i == 1 ? y : z); }
		bool operator==(const Vec3ui& o) const { return x == o.x && y == o.y && z == o.z; }
	};

	struct Vec3f
	{
		float x = 0.f, y = 0.f, z = 0.f;

		Vec3f operator+(const Vec3f& o) const { return { x + o.x, y + o.y, z + o.z }; }
		Vec3f operator-(const Vec3f& o) const { return { x - o.x, y - o.y, z - o.z }; }
		Vec3f operator*(float s) const { return { x * s, y * s, z * s }; }
		Vec3f operator/(float s) const { return { x / s, y / s, z / s }; }

		float Dot(const Vec3f& o) const { return x * o.x + y * o.y + z * o.z; }
		Vec3f Cross(const Vec3f& o) const { return { y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x }; }
		float SquaredLength() const { return x * x + y * y + z * z; }

		// Nazara Vector3::Normalize: scale by 1/length when length > 0, else untouched
		static Vec3f Normalize(Vec3f v)
		{
			float norm = std::sqrt(v.SquaredLength());
			if (norm > 0.f)
			{
				float invNorm = 1.f / norm;
				v.x *= invNorm;
				v.y *= invNorm;
				v.z *= invNorm;
			}
			return v;
		}
	};

	struct Quatf
	{
		float w = 1.f, x = 0.f, y = 0.f, z = 0.f;

		Quatf Normalized() const
		{
			Quatf q = *this;
			float norm = std::sqrt(w * w + x * x + y * y + z * z);
			if (norm > 0.f)
			{
				float invNorm = 1.f / norm;
				q.w *= invNorm;
				q.x *= invNorm;
				q.y *= invNorm;
				q.z *= invNorm;
			}
			return q;
		}

		// Nazara Quaternion::RotationBetween (restated; see SURVEY §8c "Third-party arithmetic #3")
		static Quatf RotationBetween(const Vec3f& from, const Vec3f& to)
		{
			float dot = from.Dot(to);
			if (dot < -0.999999f)
			{
				Vec3f axis;
				if (from.Dot(Vec3f{ 1.f, 0.f, 0.f }) < 0.999999f)
					axis = Vec3f{ 1.f, 0.f, 0.f }.Cross(from);
				else
					axis = Vec3f{ 0.f, 1.f, 0.f }.Cross(from);

				axis = Vec3f::Normalize(axis);

				// Quaternion(RadianAngle(pi), axis): half angle, w = cos, xyz = axis * sin
				float half = 3.14159265358979323846f / 2.f;
				float s = std::sin(half);
				float c = std::cos(half);
				return Quatf{ c, axis.x * s, axis.y * s, axis.z * s };
			}
			else if (dot > 0.999999f)
				return Quatf{};
			else
			{
				float norm = std::sqrt(from.SquaredLength() * to.SquaredLength());
				Vec3f c = from.Cross(to);
				return Quatf{ norm + dot, c.x, c.y, c.z }.Normalized();
			}
		}

		// Nazara Quaternion * Vector3
		Vec3f operator*(const Vec3f& v) const
		{
			Vec3f qv{ x, y, z };
			Vec3f uv = qv.Cross(v);
			Vec3f uuv = qv.Cross(uv);
			uv = uv * (2.f * w);
			uuv = uuv * 2.f;
			return v + uv + uuv;
		}
	};

	// ---------------------------------------------------------------- Direction.hpp:18-90
	enum Direction : int { Back = 0, Down = 1, Front = 2, Left = 3, Right = 4, Up = 5 };
	constexpr int DirectionCount = 6;

	// Nazara: Backward=(0,0,1) Down=(0,-1,0) Forward=(0,0,-1) Left=(-1,0,0) Right=(1,0,0) Up=(0,1,0)
	inline const std::array<Vec3f, 6> s_dirNormals = { {
		{ 0.f, 0.f, 1.f }, { 0.f, -1.f, 0.f }, { 0.f, 0.f, -1.f }, { -1.f, 0.f, 0.f }, { 1.f, 0.f, 0.f }, { 0.f, 1.f, 0.f }
	} };

	struct DirectionAxis { unsigned forwardAxis, rightAxis, upAxis; int forwardDir, rightDir, upDir; };
	inline const std::array<DirectionAxis, 6> s_dirAxis = { {
		{ 1, 0, 2,  1,  1,  1 }, // Back
		{ 2, 0, 1,  1,  1, -1 }, // Down
		{ 1, 0, 2, -1,  1, -1 }, // Front
		{ 2, 1, 0, -1,  1, -1 }, // Left
		{ 2, 1, 0, -1, -1,  1 }, // Right
		{ 2, 0, 1, -1,  1,  1 }, // Up
	} };

	// local (x fastest) axes: Up = local z + 1 (Direction.hpp:74-81)
	inline const std::array<Vec3i, 6> s_blockDirOffset = { {
		{ 0, 1, 0 }, { 0, 0, -1 }, { 0, -1, 0 }, { -1, 0, 0 }, { 1, 0, 0 }, { 0, 0, 1 }
	} };

	// world axes (Direction.hpp:83-90)
	inline const std::array<Vec3i, 6> s_chunkDirOffset = { {
		{ 0, 0, 1 }, { 0, -1, 0 }, { 0, 0, -1 }, { -1, 0, 0 }, { 1, 0, 0 }, { 0, 1, 0 }
	} };

	// Direction.inl:7-21 — ">=" keeps the LAST maximal direction
	inline Direction DirectionFromNormal(const Vec3f& outsideNormal)
	{
		Direction closestDir = Back;
		float closestDirDot = -1.f;
		for (int d = 0; d < DirectionCount; ++d)
		{
			if (float dot = outsideNormal.Dot(s_dirNormals[d]); dot >= closestDirDot)
			{
				closestDir = Direction(d);
				closestDirDot = dot;
			}
		}
		return closestDir;
	}

	// ---------------------------------------------------------------- BlockLibrary.cpp:10-143
	struct BlockInfo
	{
		std::string basePath, baseBackPath, baseDownPath, baseFrontPath, baseLeftPath, baseRightPath, baseSidePath, baseUpPath;
		bool hasCollisions = true;
		bool isDoubleSided = false;
		bool isTransparent = false;
		float permeability = 0.f;
	};

	struct BlockData
	{
		std::string name;
		std::array<unsigned int, 6> texIndices{};
		bool hasCollisions = true;
		bool isDoubleSided = false;
		bool isTransparent = false;
		float permeability = 0.f;
	};

	class BlockLibrary
	{
		public:
			BlockLibrary()
			{
				auto info = [] { return BlockInfo{}; };
				{ BlockInfo b = info(); b.hasCollisions = false; b.isTransparent = true; b.permeability = 1.f; RegisterBlock("empty", b); }
				{
					BlockInfo b = info();
					b.baseBackPath = "blocks/debug_back"; b.baseDownPath = "blocks/debug_down"; b.baseFrontPath = "blocks/debug_front";
					b.baseLeftPath = "blocks/debug_left"; b.baseRightPath = "blocks/debug_right"; b.baseUpPath = "blocks/debug_up";
					RegisterBlock("debug", b);
				}
				{ BlockInfo b = info(); b.basePath = "blocks/dirt"; b.permeability = 0.1f; RegisterBlock("dirt", b); }
				{ BlockInfo b = info(); b.basePath = "blocks/grass_top"; b.baseDownPath = "blocks/dirt"; b.baseSidePath = "blocks/grass_side"; b.permeability = 0.1f; RegisterBlock("grass", b); }
				{ BlockInfo b = info(); b.basePath = "blocks/smooth_stone"; RegisterBlock("hull", b); }
				{ BlockInfo b = info(); b.basePath = "blocks/smooth_stone_slab_side"; RegisterBlock("hull2", b); }
				{ BlockInfo b = info(); b.basePath = "blocks/snow"; b.permeability = 0.5f; RegisterBlock("snow", b); }
				{ BlockInfo b = info(); b.basePath = "blocks/cobblestone"; RegisterBlock("stone", b); }
				{ BlockInfo b = info(); b.basePath = "blocks/mossy_cobblestone"; RegisterBlock("stone_mossy", b); }
				{ BlockInfo b = info(); b.basePath = "blocks/forcefield"; b.hasCollisions = false; b.isTransparent = true; RegisterBlock("forcefield", b); }
				{ BlockInfo b = info(); b.basePath = "blocks/planks"; RegisterBlock("planks", b); }
				{ BlockInfo b = info(); b.basePath = "blocks/stone_bricks"; RegisterBlock("stone_bricks", b); }
				{ BlockInfo b = info(); b.basePath = "blocks/copper_block"; RegisterBlock("copper_block", b); }
				{ BlockInfo b = info(); b.basePath = "blocks/glass"; b.isDoubleSided = true; b.isTransparent = true; RegisterBlock("glass", b); }
			}

			const BlockData& GetBlockData(BlockIndex idx) const { return m_blocks[idx]; }

			BlockIndex GetBlockIndex(std::string_view name) const
			{
				auto it = m_blockIndices.find(std::string(name));
				return it == m_blockIndices.end() ? InvalidBlockIndex : it->second;
			}

			std::size_t GetBlockCount() const { return m_blocks.size(); }

			// BlockLibrary.cpp:85-131
			BlockIndex RegisterBlock(std::string name, BlockInfo blockInfo)
			{
				BlockIndex blockIndex = BlockIndex(m_blocks.size());
				BlockData& data = m_blocks.emplace_back();
				data.hasCollisions = blockInfo.hasCollisions;
				data.isDoubleSided = blockInfo.isDoubleSided;
				data.isTransparent = blockInfo.isTransparent;
				data.permeability = blockInfo.permeability;
				data.name = name;

				unsigned int baseTexIndex = blockInfo.basePath.empty() ? 0u : RegisterTexture(blockInfo.basePath);
				unsigned int baseSideTexIndex = blockInfo.baseSidePath.empty() ? baseTexIndex : RegisterTexture(blockInfo.baseSidePath);

				const std::string* dirToStr[6] = { &blockInfo.baseBackPath, &blockInfo.baseDownPath, &blockInfo.baseFrontPath, &blockInfo.baseLeftPath, &blockInfo.baseRightPath, &blockInfo.baseUpPath };
				for (int dir = 0; dir < 6; ++dir)
				{
					if (!dirToStr[dir]->empty())
						data.texIndices[dir] = RegisterTexture(*dirToStr[dir]);
					else if (dir != Up && dir != Down)
						data.texIndices[dir] = baseSideTexIndex;
					else
						data.texIndices[dir] = baseTexIndex;
				}

				m_blockIndices.emplace(std::move(name), blockIndex);
				return blockIndex;
			}

		private:
			// BlockLibrary.cpp:133-143 — slice #0 reserved for empty
			unsigned int RegisterTexture(const std::string& path)
			{
				auto it = m_textureIndices.find(path);
				if (it != m_textureIndices.end())
					return it->second;
				unsigned int texIndex = unsigned(m_textureIndices.size() + 1);
				m_textureIndices.emplace(path, texIndex);
				return texIndex;
			}

			std::unordered_map<std::string, BlockIndex> m_blockIndices;
			std::unordered_map<std::string, unsigned int> m_textureIndices;
			std::vector<BlockData> m_blocks;
	};

	// ---------------------------------------------------------------- siv::PerlinNoise v3.0.0 (restated, double)
	class PerlinNoise
	{
		public:
			// reseed(seed): iota + library-private shuffle driven by std::mt19937{seed}
			void reseed(std::uint32_t seed)
			{
				std::mt19937 urbg{ seed };
				for (int i = 0; i < 256; ++i)
					m_perm[i] = std::uint8_t(i);
				for (std::uint64_t i = 1; i < 256; ++i)
				{
					std::uint64_t j = std::uint64_t(urbg()) % (i + 1);
					std::swap(m_perm[i], m_perm[j]);
				}
			}

			const std::uint8_t* permutation() const { return m_perm.data(); }
			// siv::PerlinNoise::deserialize(state): the permutation table is the whole state
			void deserialize(const std::uint8_t* state) { std::copy(state, state + 256, m_perm.begin()); }

			double noise3D(double x, double y, double z) const
			{
				const double _x = std::floor(x);
				const double _y = std::floor(y);
				const double _z = std::floor(z);

				const std::int32_t ix = std::int32_t(_x) & 255;
				const std::int32_t iy = std::int32_t(_y) & 255;
				const std::int32_t iz = std::int32_t(_z) & 255;

				const double fx = x - _x;
				const double fy = y - _y;
				const double fz = z - _z;

				const double u = Fade(fx);
				const double v = Fade(fy);
				const double w = Fade(fz);

				const std::uint8_t A = std::uint8_t((m_perm[ix & 255] + iy) & 255);
				const std::uint8_t B = std::uint8_t((m_perm[(ix + 1) & 255] + iy) & 255);
				const std::uint8_t AA = std::uint8_t((m_perm[A] + iz) & 255);
				const std::uint8_t AB = std::uint8_t((m_perm[(A + 1) & 255] + iz) & 255);
				const std::uint8_t BA = std::uint8_t((m_perm[B] + iz) & 255);
				const std::uint8_t BB = std::uint8_t((m_perm[(B + 1) & 255] + iz) & 255);

				const double p0 = Grad(m_perm[AA], fx, fy, fz);
				const double p1 = Grad(m_perm[BA], fx - 1, fy, fz);
				const double p2 = Grad(m_perm[AB], fx, fy - 1, fz);
				const double p3 = Grad(m_perm[BB], fx - 1, fy - 1, fz);
				const double p4 = Grad(m_perm[(AA + 1) & 255], fx, fy, fz - 1);
				const double p5 = Grad(m_perm[(BA + 1) & 255], fx - 1, fy, fz - 1);
				const double p6 = Grad(m_perm[(AB + 1) & 255], fx, fy - 1, fz - 1);
				const double p7 = Grad(m_perm[(BB + 1) & 255], fx - 1, fy - 1, fz - 1);

				const double q0 = Lerp(p0, p1, u);
				const double q1 = Lerp(p2, p3, u);
				const double q2 = Lerp(p4, p5, u);
				const double q3 = Lerp(p6, p7, u);

				const double r0 = Lerp(q0, q1, v);
				const double r1 = Lerp(q2, q3, v);

				return Lerp(r0, r1, w);
			}

			double noise2D(double x, double y) const { return noise3D(x, y, 0.34567); }

			double octave2D(double x, double y, std::int32_t octaves, double persistence = 0.5) const
			{
				double result = 0;
				double amplitude = 1;
				for (std::int32_t i = 0; i < octaves; ++i)
				{
					result += noise2D(x, y) * amplitude;
					x *= 2;
					y *= 2;
					amplitude *= persistence;
				}
				return result;
			}

			static double MaxAmplitude(std::int32_t octaves, double persistence)
			{
				double result = 0;
				double amplitude = 1;
				for (std::int32_t i = 0; i < octaves; ++i)
				{
					result += amplitude;
					amplitude *= persistence;
				}
				return result;
			}

			double normalizedOctave2D_01(double x, double y, std::int32_t octaves, double persistence = 0.5) const
			{
				double n = octave2D(x, y, octaves, persistence) / MaxAmplitude(octaves, persistence);
				return n * 0.5 + 0.5;
			}

		private:
			static double Fade(double t) { return t * t * t * (t * (t * 6 - 15) + 10); }
			static double Lerp(double a, double b, double t) { return a + (b - a) * t; }
			static double Grad(std::uint8_t hash, double x, double y, double z)
			{
				const std::uint8_t h = hash & 15;
				const double u = h < 8 ? x : y;
				const double v = h < 4 ? y : (h == 12 || h == 14 ? x : z);
				return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
			}

			std::array<std::uint8_t, 256> m_perm{};
	};

	// ---------------------------------------------------------------- Chunk / ChunkContainer
	class ChunkContainer;

	// Effective Nz::BoxCorner semantics (SURVEY §8.1 — derived, UNPINNED by any reference test):
	// for Boxf(x, y, z, w, h, d):  Top/Bottom <-> X (Top = x+w) · Near/Far <-> Y (Near = y+h) · Left/Right <-> Z (Left = z+d).
	// Enum order follows the explicit list at DeformedChunk.cpp:123-130.
	enum BoxCorner : int { LeftBottomFar, LeftTopFar, RightBottomFar, RightTopFar, LeftBottomNear, LeftTopNear, RightBottomNear, RightTopNear };
	using Corners = std::array<Vec3f, 8>;

	inline Vec3f BoxGetCorner(float x, float y, float z, float w, float h, float d, BoxCorner c)
	{
		bool left = (c == LeftBottomFar || c == LeftTopFar || c == LeftBottomNear || c == LeftTopNear);
		bool top = (c == LeftTopFar || c == RightTopFar || c == LeftTopNear || c == RightTopNear);
		bool near_ = (c == LeftBottomNear || c == LeftTopNear || c == RightBottomNear || c == RightTopNear);
		return Vec3f{ top ? x + w : x, near_ ? y + h : y, left ? z + d : z };
	}

	inline Corners BoxGetCorners(float x, float y, float z, float w, float h, float d)
	{
		Corners c;
		for (int i = 0; i < 8; ++i)
			c[i] = BoxGetCorner(x, y, z, w, h, d, BoxCorner(i));
		return c;
	}

	enum class CornerMode { Flat, Deformed };

	struct MeshOut
	{
		std::vector<std::uint32_t> indices;
		std::vector<Vec3f> position;
		std::vector<Vec3f> normal; // filled iff wantNormal
		std::vector<Vec3f> uvw;    // filled iff wantUv
	};

	class Chunk
	{
		public:
			Chunk(const BlockLibrary& lib, ChunkContainer& owner, Vec3i indices, float blockSize) :
			m_indices(indices), m_blockLibrary(lib), m_owner(owner), m_blockSize(blockSize) {}

			// Chunk.inl:25-32
			static unsigned int GetBlockLocalIndex(const Vec3ui& i) { return ChunkSize * (ChunkSize * i.z + i.y) + i.x; }
			BlockIndex GetBlockContent(const Vec3ui& i) const { return m_blocks[GetBlockLocalIndex(i)]; }
			bool HasContent() const { return !m_blocks.empty(); }
			const Vec3i& GetIndices() const { return m_indices; }
			const std::vector<BlockIndex>& Blocks() const { return m_blocks; }
			const std::vector<std::uint8_t>& CollisionMask() const { return m_collisionCellMask; }
			float GetBlockSize() const { return m_blockSize; }

			// Chunk.inl:109-123 + Chunk.cpp:368-384
			template<typename F> void Reset(F&& f);
			// Chunk.inl:97-107: all Empty, no OnReset signal
			void ResetEmpty()
			{
				m_blocks.assign(ChunkBlockCount, EmptyBlockIndex);
				m_collisionCellMask.assign(ChunkBlockCount, 0);
				m_blockTypeCount.assign(EmptyBlockIndex + 1, 0);
				m_blockTypeCount[EmptyBlockIndex] = ChunkBlockCount;
			}

			// Chunk.cpp:348-366
			void UpdateBlock(const Vec3ui& indices, BlockIndex newBlock);

			// Chunk.cpp:312-346 / 252-310: the save format.  Stream conventions of Nazara's ByteStream as restated in
			// oracle/ref_shim/include/Nazara/Core/ByteStream.hpp (UNPINNED third-party behaviour): arithmetic types big-endian,
			// std::string = UInt32 length + bytes, Vector3ui = three UInt32.
			void Serialize(std::vector<std::uint8_t>& out) const;
			void Deserialize(const std::uint8_t* data, std::size_t size); // throws std::runtime_error like the reference

			// corner providers: FlatChunk.cpp:48-54, DeformedChunk.cpp:115-154
			CornerMode cornerMode = CornerMode::Flat;
			Vec3f deformationCenter{};
			float deformationRadius = 0.f;

			Corners ComputeVoxelCorners(const Vec3ui& indices) const
			{
				if (cornerMode == CornerMode::Flat)
				{
					Vec3f blockPos = (Vec3f{ float(indices.x), float(indices.y), float(indices.z) } - Vec3f{ float(ChunkSize), float(ChunkSize), float(ChunkSize) } * 0.5f) * m_blockSize;
					return BoxGetCorners(blockPos.x, blockPos.z, blockPos.y, m_blockSize, m_blockSize, m_blockSize);
				}
				else
				{
					float fX = indices.x * m_blockSize;
					float fY = indices.y * m_blockSize;
					float fZ = indices.z * m_blockSize;
					Corners c = BoxGetCorners(fX, fZ, fY, m_blockSize, m_blockSize, m_blockSize);
					for (Vec3f& p : c)
						p = DeformPosition(p);
					return c;
				}
			}

			// DeformedChunk.cpp:139-154
			Vec3f DeformPosition(const Vec3f& position) const
			{
				float distToCenter = std::max({
					std::abs(position.x - deformationCenter.x),
					std::abs(position.y - deformationCenter.y),
					std::abs(position.z - deformationCenter.z),
				});

				float innerReductionSize = std::max(distToCenter - deformationRadius, 0.f);
				// Boxf(center - r, Vector3f(2r)): min = center - r, max = min + 2r
				Vec3f boxMin = deformationCenter - Vec3f{ innerReductionSize, innerReductionSize, innerReductionSize };
				float len = innerReductionSize * 2.f;
				Vec3f boxMax{ boxMin.x + len, boxMin.y + len, boxMin.z + len };

				auto clamp = [](float v, float lo, float hi) { return std::max(std::min(v, hi), lo); };
				Vec3f innerPos{ clamp(position.x, boxMin.x, boxMax.x), clamp(position.y, boxMin.y, boxMax.y), clamp(position.z, boxMin.z, boxMax.z) };
				Vec3f normal = Vec3f::Normalize(position - innerPos);

				return innerPos + normal * std::min(deformationRadius, distToCenter);
			}

			// Chunk.cpp:20-250
			void BuildMesh(MeshOut& out, const Vec3f& gravityCenter, bool wantNormal, bool wantUv) const;

			// FlatChunk.cpp:56-153 (greedy box merge) — emits (pos, size) in block units, x,y(up),z world-local order
			struct ColliderBox { Vec3f pos, size; };
			std::vector<ColliderBox> BuildBoxCollider() const;

			std::function<void(Chunk*)> onReset;
			std::function<void(Chunk*, const Vec3ui&, BlockIndex)> onBlockUpdated;

		private:
			void OnChunkReset();

			std::vector<BlockIndex> m_blocks;
			std::vector<std::uint32_t> m_blockTypeCount;
			std::vector<std::uint8_t> m_collisionCellMask; // 1 byte per block here (reference: bitset)
			Vec3i m_indices;
			const BlockLibrary& m_blockLibrary;
			ChunkContainer& m_owner;
			float m_blockSize;
	};

	class ChunkContainer
	{
		public:
			explicit ChunkContainer(float tileSize) : m_tileSize(tileSize) {}
			virtual ~ChunkContainer() = default;

			// ChunkContainer.inl:14-17
			static Vec3i GetBlockIndices(const Vec3i& chunkIndices, const Vec3ui& indices)
			{
				const std::int32_t half = std::int32_t(ChunkSize) / 2;
				return Vec3i{
					chunkIndices.x * std::int32_t(ChunkSize) + std::int32_t(indices.x) - half,
					chunkIndices.y * std::int32_t(ChunkSize) + std::int32_t(indices.z) - half,
					chunkIndices.z * std::int32_t(ChunkSize) + std::int32_t(indices.y) - half
				};
			}

			// ChunkContainer.inl:19-42
			static Vec3i GetChunkIndicesByBlockIndices(const Vec3i& indices, Vec3ui* localIndices = nullptr)
			{
				const std::int32_t cs = std::int32_t(ChunkSize);
				Vec3i adjusted{ indices.x + cs / 2, indices.y + cs / 2, indices.z + cs / 2 };
				auto offs = [cs](int v) { return (v < 0) ? v - cs + 1 : v; };
				Vec3i chunkIndices{ offs(adjusted.x) / cs, offs(adjusted.y) / cs, offs(adjusted.z) / cs };
				if (localIndices)
				{
					auto fix = [cs](int v) { return (v < 0) ? v + cs : v; };
					Vec3i local{ fix(adjusted.x - chunkIndices.x * cs), fix(adjusted.y - chunkIndices.y * cs), fix(adjusted.z - chunkIndices.z * cs) };
					*localIndices = Vec3ui{ unsigned(local.x), unsigned(local.z), unsigned(local.y) };
				}
				return chunkIndices;
			}

			// ChunkContainer.inl:53-56
			Vec3f GetChunkOffset(const Vec3i& indices) const
			{
				return Vec3f{ float(indices.x), float(indices.y), float(indices.z) } * float(ChunkSize) * m_tileSize;
			}

			virtual Vec3f GetCenter() const { return Vec3f{}; } // Planet.inl:7-10
			float GetTileSize() const { return m_tileSize; }

			Chunk* GetChunk(const Vec3i& idx)
			{
				auto it = m_chunks.find(idx);
				return it == m_chunks.end() ? nullptr : it->second.get();
			}
			const Chunk* GetChunk(const Vec3i& idx) const
			{
				auto it = m_chunks.find(idx);
				return it == m_chunks.end() ? nullptr : it->second.get();
			}
			std::size_t GetChunkCount() const { return m_chunks.size(); }

			// ChunkContainer.hpp:44-46 signal + ChunkEntities.cpp:43-48 listener (per-tick OR of masks)
			void OnChunkUpdated(Chunk* chunk, std::uint8_t neighborMask) { invalidated[chunk->GetIndices()] |= neighborMask; }

			// ChunkEntities::Update + ClientChunkEntities::ProcessChunkUpdate (ClientChunkEntities.cpp:243-256):
			// chunks to rebuild this tick = invalidated chunks with content ∪ their masked neighbours that exist and have content.
			std::vector<Vec3i> CollectDirtyChunks(bool clear = true)
			{
				std::map<Vec3i, bool> dirty;
				for (auto&& [idx, mask] : invalidated)
				{
					const Chunk* c = GetChunk(idx);
					if (!c || !c->HasContent())
						continue;
					dirty[idx] = true;
					for (int d = 0; d < 6; ++d)
					{
						if (!(mask & (1u << d)))
							continue;
						Vec3i n{ idx.x + s_chunkDirOffset[d].x, idx.y + s_chunkDirOffset[d].y, idx.z + s_chunkDirOffset[d].z };
						const Chunk* nc = GetChunk(n);
						if (!nc || !nc->HasContent())
							continue;
						dirty[n] = true;
					}
				}
				if (clear)
					invalidated.clear();
				std::vector<Vec3i> res;
				for (auto&& [idx, _] : dirty)
					res.push_back(idx);
				return res;
			}

			std::map<Vec3i, std::uint8_t> invalidated;

		protected:
			std::map<Vec3i, std::unique_ptr<Chunk>> m_chunks;
			float m_tileSize;
	};

	template<typename F> void Chunk::Reset(F&& f)
	{
		if (!HasContent())
		{
			m_blocks.resize(ChunkBlockCount, EmptyBlockIndex);
			m_collisionCellMask.resize(ChunkBlockCount, 0);
			m_blockTypeCount.resize(EmptyBlockIndex + 1);
			m_blockTypeCount[EmptyBlockIndex] = ChunkBlockCount;
		}
		f(m_blocks.data());
		OnChunkReset();
	}

	inline void Chunk::OnChunkReset()
	{
		std::fill(m_blockTypeCount.begin(), m_blockTypeCount.end(), 0);
		for (std::size_t i = 0; i < m_blocks.size(); ++i)
		{
			BlockIndex content = m_blocks[i];
			m_collisionCellMask[i] = m_blockLibrary.GetBlockData(content).hasCollisions;
			if (content >= m_blockTypeCount.size())
				m_blockTypeCount.resize(content + 1);
			m_blockTypeCount[content]++;
		}
		if (onReset)
			onReset(this);
	}

	inline void Chunk::UpdateBlock(const Vec3ui& indices, BlockIndex newBlock)
	{
		const BlockData& data = m_blockLibrary.GetBlockData(newBlock);
		unsigned int bi = GetBlockLocalIndex(indices);
		BlockIndex old = m_blocks[bi];
		m_blocks[bi] = newBlock;
		m_collisionCellMask[bi] = data.hasCollisions;
		m_blockTypeCount[old]--;
		if (newBlock >= m_blockTypeCount.size())
			m_blockTypeCount.resize(newBlock + 1);
		m_blockTypeCount[newBlock]++;
		if (onBlockUpdated)
			onBlockUpdated(this, indices, newBlock);
	}

	namespace stream
	{
		inline void PutU8(std::vector<std::uint8_t>& o, std::uint8_t v) { o.push_back(v); }
		inline void PutU16(std::vector<std::uint8_t>& o, std::uint16_t v) { o.push_back(std::uint8_t(v >> 8)); o.push_back(std::uint8_t(v)); }
		inline void PutU32(std::vector<std::uint8_t>& o, std::uint32_t v) { for (int s = 24; s >= 0; s -= 8) o.push_back(std::uint8_t(v >> s)); }
		struct Reader
		{
			const std::uint8_t* p;
			std::size_t n, cur = 0;
			void Need(std::size_t k) const { if (cur + k > n) throw std::runtime_error("ByteStream: read past end"); }
			std::uint8_t U8() { Need(1); return p[cur++]; }
			std::uint16_t U16() { Need(2); std::uint16_t v = std::uint16_t(p[cur] << 8 | p[cur + 1]); cur += 2; return v; }
			std::uint32_t U32() { Need(4); std::uint32_t v = std::uint32_t(p[cur]) << 24 | std::uint32_t(p[cur + 1]) << 16 | std::uint32_t(p[cur + 2]) << 8 | p[cur + 3]; cur += 4; return v; }
		};
	}

	constexpr std::uint32_t ChunkBinaryVersion = 1; // InternalConstants.hpp:24

	// Chunk.cpp:312-346
	inline void Chunk::Serialize(std::vector<std::uint8_t>& out) const
	{
		using namespace stream;
		PutU32(out, ChunkBinaryVersion);
		PutU32(out, ChunkSize); PutU32(out, ChunkSize); PutU32(out, ChunkSize);

		std::vector<BlockIndex> serializationIndices(m_blockTypeCount.size());
		std::uint16_t nextUniqueIndex = 0;
		for (std::size_t i = 0; i < m_blockTypeCount.size(); ++i)
		{
			if (m_blockTypeCount[i] == 0)
				continue;
			serializationIndices[i] = nextUniqueIndex++;
		}
		PutU16(out, nextUniqueIndex);
		for (std::size_t i = 0; i < m_blockTypeCount.size(); ++i)
		{
			if (m_blockTypeCount[i] == 0)
				continue;
			const std::string& name = m_blockLibrary.GetBlockData(BlockIndex(i)).name;
			PutU32(out, std::uint32_t(name.size()));
			out.insert(out.end(), name.begin(), name.end());
		}
		// "nextUniqueIndex is the number of bits required..." (sic): more than 8 distinct types -> 16-bit indices
		if (nextUniqueIndex > 8)
		{
			for (BlockIndex b : m_blocks)
				PutU16(out, serializationIndices[b]);
		}
		else
		{
			for (BlockIndex b : m_blocks)
				PutU8(out, std::uint8_t(serializationIndices[b]));
		}
	}

	// Chunk.cpp:252-310
	inline void Chunk::Deserialize(const std::uint8_t* data, std::size_t size)
	{
		stream::Reader r{ data, size };
		if (r.U32() != ChunkBinaryVersion)
			throw std::runtime_error("incompatible chunk version");
		std::uint32_t sx = r.U32(), sy = r.U32(), sz = r.U32();
		if (sx != ChunkSize || sy != ChunkSize || sz != ChunkSize)
			throw std::runtime_error("incompatible chunk size");
		std::uint16_t blockTypeCount = r.U16();
		std::vector<BlockIndex> deserializationIndices;
		deserializationIndices.reserve(blockTypeCount);
		for (std::uint16_t i = 0; i < blockTypeCount; ++i)
		{
			std::uint32_t len = r.U32();
			r.Need(len);
			std::string name(reinterpret_cast<const char*>(r.p + r.cur), len);
			r.cur += len;
			BlockIndex idx = m_blockLibrary.GetBlockIndex(name);
			if (idx == InvalidBlockIndex)
				throw std::runtime_error("unknown block " + name);
			deserializationIndices.push_back(idx);
		}
		std::vector<BlockIndex> blocks(ChunkBlockCount);
		for (BlockIndex& b : blocks)
		{
			std::uint16_t v = blockTypeCount > 8 ? r.U16() : r.U8();
			if (v >= deserializationIndices.size())
				throw std::runtime_error("palette index out of range"); // the reference indexes out of bounds here (UB)
			b = deserializationIndices[v];
		}
		Reset([&](BlockIndex* dst) { std::copy(blocks.begin(), blocks.end(), dst); });
	}

	// ---- Chunk::BuildMesh (Chunk.cpp:20-250)
	inline void Chunk::BuildMesh(MeshOut& out, const Vec3f& gravityCenter, bool wantNormal, bool wantUv) const
	{
		auto DrawFace = [&](BlockIndex blockIndex, const Vec3f& blockCenter, const std::array<Vec3f, 4>& pos)
		{
			// addVertices(4): ClientChunkEntities.cpp:146-158 / DeformedChunk.cpp:20-29
			std::uint32_t firstIndex = std::uint32_t(out.position.size());
			for (const Vec3f& p : pos)
				out.position.push_back(p);

			Vec3f faceCenter = std::accumulate(pos.begin(), pos.end(), Vec3f{}) / float(pos.size());
			Vec3f faceDirection = Vec3f::Normalize(faceCenter - blockCenter);

			if (wantNormal)
			{
				for (std::size_t i = 0; i < pos.size(); ++i)
					out.normal.push_back(faceDirection);
			}

			if (wantUv)
			{
				Vec3f faceUp = s_dirNormals[DirectionFromNormal(Vec3f::Normalize(faceCenter - gravityCenter))];
				Quatf upRotation = Quatf::RotationBetween(faceUp, Vec3f{ 0.f, 1.f, 0.f });
				Direction texDirection = DirectionFromNormal(upRotation * faceDirection);

				const BlockData& blockData = m_blockLibrary.GetBlockData(blockIndex);
				std::size_t textureIndex = blockData.texIndices[texDirection];
				float sliceIndex = float(textureIndex);

				for (std::size_t i = 0; i < pos.size(); ++i)
				{
					Vec3f dir = upRotation * (pos[i] - blockCenter);
					Vec3f dirAbs{ std::abs(dir.x), std::abs(dir.y), std::abs(dir.z) };

					float mag = 0.f;
					float u = 0.f, v = 0.f;
					switch (texDirection)
					{
						case Back:
						case Front:
							mag = 0.5f / dirAbs.x;
							u = dir.x < 0.f ? -dir.z : dir.z;
							v = -dir.y;
							break;
						case Down:
						case Up:
							mag = 0.5f / dirAbs.y;
							u = dir.x;
							v = dir.y < 0.f ? -dir.z : dir.z;
							break;
						case Left:
						case Right:
							mag = 0.5f / dirAbs.z;
							u = dir.z < 0.f ? dir.x : -dir.x;
							v = -dir.y;
							break;
					}
					out.uvw.push_back(Vec3f{ u * mag + 0.5f, v * mag + 0.5f, sliceIndex });
				}
			}

			out.indices.push_back(firstIndex);
			out.indices.push_back(firstIndex + 2);
			out.indices.push_back(firstIndex + 1);
			out.indices.push_back(firstIndex + 1);
			out.indices.push_back(firstIndex + 2);
			out.indices.push_back(firstIndex + 3);
		};

		// neighbour chunks, Chunk.cpp:104-113
		std::array<const Chunk*, 6> neighborChunks{};
		for (int d = 0; d < 6; ++d)
		{
			Vec3i n{ m_indices.x + s_chunkDirOffset[d].x, m_indices.y + s_chunkDirOffset[d].y, m_indices.z + s_chunkDirOffset[d].z };
			neighborChunks[d] = static_cast<const ChunkContainer&>(m_owner).GetChunk(n);
		}

		// Chunk.cpp:124-172
		auto GetNeighborBlock = [&](Vec3ui indices, Direction direction) -> std::optional<BlockIndex>
		{
			Vec3i chunkIndices = m_indices;
			std::swap(chunkIndices.y, chunkIndices.z);

			for (unsigned axis = 0; axis < 3; ++axis)
			{
				unsigned int& index = indices[axis];
				int offset = s_blockDirOffset[direction][axis];
				if (offset > 0)
				{
					index += offset;
					if (index >= ChunkSize)
					{
						index -= ChunkSize;
						chunkIndices[axis]++;
					}
				}
				else if (offset < 0)
				{
					unsigned int posOffset = std::abs(offset);
					if (posOffset > index)
					{
						index += ChunkSize;
						chunkIndices[axis]--;
					}
					index -= posOffset;
				}
			}

			std::swap(chunkIndices.y, chunkIndices.z);

			if (chunkIndices != m_indices)
			{
				const Chunk* chunk = neighborChunks[direction];
				if (!chunk)
					return {};
				if (!chunk->HasContent())
					return {};
				return chunk->GetBlockContent(indices);
			}
			else
				return GetBlockContent(indices);
		};

		// face table in the reference's emission order, Chunk.cpp:200-246 (normal quad, double-sided twin)
		struct FaceDef { Direction dir; BoxCorner quad[4]; BoxCorner twin[4]; };
		static const FaceDef faces[6] = {
			{ Up,    { RightTopNear, LeftTopNear, RightBottomNear, LeftBottomNear }, { LeftTopNear, RightTopNear, LeftBottomNear, RightBottomNear } },
			{ Down,  { LeftTopFar, RightTopFar, LeftBottomFar, RightBottomFar },     { RightTopFar, LeftTopFar, RightBottomFar, LeftBottomFar } },
			{ Front, { RightTopFar, RightTopNear, RightBottomFar, RightBottomNear }, { RightTopNear, RightTopFar, RightBottomNear, RightBottomFar } },
			{ Back,  { LeftTopNear, LeftTopFar, LeftBottomNear, LeftBottomFar },     { LeftTopFar, LeftTopNear, LeftBottomFar, LeftBottomNear } },
			{ Left,  { RightBottomNear, LeftBottomNear, RightBottomFar, LeftBottomFar }, { LeftBottomNear, RightBottomNear, LeftBottomFar, RightBottomFar } },
			{ Right, { LeftTopNear, RightTopNear, LeftTopFar, RightTopFar },         { RightTopNear, LeftTopNear, RightTopFar, LeftTopFar } },
		};

		for (unsigned int z = 0; z < ChunkSize; ++z)
		{
			for (unsigned int y = 0; y < ChunkSize; ++y)
			{
				for (unsigned int x = 0; x < ChunkSize; ++x)
				{
					BlockIndex blockIndex = GetBlockContent({ x, y, z });
					if (blockIndex == EmptyBlockIndex)
						continue;

					const BlockData& blockData = m_blockLibrary.GetBlockData(blockIndex);
					Corners corners = ComputeVoxelCorners({ x, y, z });
					Vec3f blockCenter = std::accumulate(corners.begin(), corners.end(), Vec3f{}) / float(corners.size());

					auto IsTransparent = [&](BlockIndex neighborBlockIndex)
					{
						if (blockIndex == neighborBlockIndex)
							return false;
						return m_blockLibrary.GetBlockData(neighborBlockIndex).isTransparent;
					};

					for (const FaceDef& f : faces)
					{
						auto neighborOpt = GetNeighborBlock({ x, y, z }, f.dir);
						if (!neighborOpt || IsTransparent(*neighborOpt))
						{
							DrawFace(blockIndex, blockCenter, { corners[f.quad[0]], corners[f.quad[1]], corners[f.quad[2]], corners[f.quad[3]] });
							if (blockData.isDoubleSided)
								DrawFace(blockIndex, blockCenter, { corners[f.twin[0]], corners[f.twin[1]], corners[f.twin[2]], corners[f.twin[3]] });
						}
					}
				}
			}
		}
	}

	// ---- FlatChunk::BuildCollider greedy merge (FlatChunk.cpp:56-153)
	inline std::vector<Chunk::ColliderBox> Chunk::BuildBoxCollider() const
	{
		std::vector<ColliderBox> boxes;
		std::vector<std::uint8_t> mask = m_collisionCellMask;
		const unsigned dx = ChunkSize, dy = ChunkSize, dz = ChunkSize;
		auto L = [&](unsigned x, unsigned y, unsigned z) { return dx * (dy * z + y) + x; };
		const float half = float(ChunkSize) * 0.5f;

		bool hasStart = false;
		unsigned sx = 0, sy = 0, sz = 0;
		auto Commit = [&](unsigned endX)
		{
			if (!hasStart)
				return;
			unsigned endY = sy, endZ = sz;
			for (endY += 1; endY < dy; ++endY)
			{
				bool ok = true;
				for (unsigned x = sx; x <= endX; ++x)
					if (!mask[L(x, endY, endZ)]) { ok = false; break; }
				if (!ok)
					break;
			}
			endY--;
			for (endZ += 1; endZ < dz; ++endZ)
			{
				bool ok = true;
				for (unsigned y = sy; y <= endY && ok; ++y)
					for (unsigned x = sx; x <= endX; ++x)
						if (!mask[L(x, y, endZ)]) { ok = false; break; }
				if (!ok)
					break;
			}
			endZ--;
			for (unsigned z = sz; z <= endZ; ++z)
				for (unsigned y = sy; y <= endY; ++y)
					for (unsigned x = sx; x <= endX; ++x)
						mask[L(x, y, z)] = 0;

			Vec3f startOffset{ float(sx), float(sz), float(sy) };
			Vec3f endOffset{ float(endX + 1), float(endZ + 1), float(endY + 1) };
			boxes.push_back({ startOffset - Vec3f{ half, half, half }, endOffset - startOffset });
			hasStart = false;
		};

		for (unsigned z = 0; z < dz; ++z)
		{
			for (unsigned y = 0; y < dy; ++y)
			{
				for (unsigned x = 0; x < dx; ++x)
				{
					if (!mask[L(x, y, z)])
					{
						Commit(x - 1);
						continue;
					}
					if (!hasStart)
					{
						hasStart = true;
						sx = x; sy = y; sz = z;
					}
				}
				Commit(dx - 1);
			}
		}
		return boxes;
	}

	// ---------------------------------------------------------------- Planet (Planet.cpp)
	class Planet : public ChunkContainer
	{
		public:
			Planet(float tileSize, float cornerRadius, float gravity) : ChunkContainer(tileSize), m_cornerRadius(cornerRadius), m_gravity(gravity) {}

			// Planet.cpp:28-68 (FlatChunk + dirty-mask rule)
			Chunk& AddChunk(const BlockLibrary& lib, const Vec3i& indices)
			{
				auto chunk = std::make_unique<Chunk>(lib, *this, indices, m_tileSize);
				chunk->onReset = [this](Chunk* c) { OnChunkUpdated(c, 0x3F); };
				chunk->onBlockUpdated = [this](Chunk* c, const Vec3ui& i, BlockIndex)
				{
					std::uint8_t mask = 0;
					if (i.x == 0) mask |= 1u << Left; else if (i.x == ChunkSize - 1) mask |= 1u << Right;
					if (i.y == 0) mask |= 1u << Front; else if (i.y == ChunkSize - 1) mask |= 1u << Back;
					if (i.z == 0) mask |= 1u << Down; else if (i.z == ChunkSize - 1) mask |= 1u << Up;
					OnChunkUpdated(c, mask);
				};
				Chunk* ptr = chunk.get();
				m_chunks[indices] = std::move(chunk);
				return *ptr;
			}

			// Planet.cpp:160-383.  Returns false if the reference would have tripped Nz::SafeCaster (negative depth, debug assert).
			static bool GenerateChunkBlocks(const BlockLibrary& lib, const Vec3i& chunkIndices, std::uint32_t seed, const Vec3ui& chunkCount, BlockIndex* blockIndices)
			{
				constexpr std::size_t freeSpace = 30;
				bool ok = true;

				std::uint32_t chunkSeed = seed + std::uint32_t(chunkIndices.x) + std::uint32_t(chunkIndices.y) + std::uint32_t(chunkIndices.z);
				std::minstd_rand rand(chunkSeed);
				std::bernoulli_distribution dis(0.9);

				BlockIndex dirt = lib.GetBlockIndex("dirt");
				BlockIndex grass = lib.GetBlockIndex("grass");
				BlockIndex stone = lib.GetBlockIndex("stone");
				BlockIndex stoneMossy = lib.GetBlockIndex("stone_mossy");
				BlockIndex snow = lib.GetBlockIndex("snow");

				Vec3i maxHeight{ (std::int32_t(chunkCount.x) + 1) / 2, (std::int32_t(chunkCount.y) + 1) / 2, (std::int32_t(chunkCount.z) + 1) / 2 };
				maxHeight.x *= int(ChunkSize);
				maxHeight.y *= int(ChunkSize);
				maxHeight.z *= int(ChunkSize);

				std::array<PerlinNoise, 6> perlin;
				for (int d = 0; d < 6; ++d)
					perlin[d].reseed(seed + unsigned(d));

				// pass 1, Planet.cpp:191-228
				BlockIndex* ptr = blockIndices;
				for (unsigned z = 0; z < ChunkSize; ++z)
				{
					for (unsigned y = 0; y < ChunkSize; ++y)
					{
						for (unsigned x = 0; x < ChunkSize; ++x)
						{
							Vec3i p = GetBlockIndices(chunkIndices, { x, y, z });
							int sdepth = std::min({ maxHeight.x - std::abs(p.x), maxHeight.y - std::abs(p.z), maxHeight.z - std::abs(p.y) });
							if (sdepth < 0)
								ok = false;
							unsigned int depth = unsigned(sdepth);
							if (depth < freeSpace)
							{
								*ptr++ = EmptyBlockIndex;
								continue;
							}
							depth -= freeSpace;

							BlockIndex b;
							if (depth <= 6)
								b = snow;
							else if (depth <= 18)
								b = dirt;
							else
								b = dis(rand) ? stone : stoneMossy;

							if (std::abs(p.x) <= 2 && std::abs(p.z) <= 2)
								b = EmptyBlockIndex;

							if (b != InvalidBlockIndex)
								*ptr++ = b;
						}
					}
				}

				constexpr double heightScale = 1.5f;
				constexpr double scale = 0.02f;

				// One terrain pass (Planet.cpp:233-381): `top` = local coords of the column's reference cell, `axis` = local axis the
				// column runs along, `sign` = +1 when height grows with the local coordinate.
				auto pass = [&](Direction dir, int maxH, auto&& mapPosOf, auto&& noiseArgs, auto&& blockDepthOf, auto&& cellOf)
				{
					for (unsigned j = 0; j < ChunkSize; ++j)
					{
						for (unsigned i = 0; i < ChunkSize; ++i)
						{
							Vec3i mapPos = mapPosOf(i, j);
							auto [a, b] = noiseArgs(mapPos);
							double height = perlin[dir].normalizedOctave2D_01(a * scale, b * scale, 4) * heightScale;

							int terrainDepth = int(std::round(std::min<double>(height * (maxH / 2 - freeSpace) + freeSpace, maxH / 2)));
							int blockDepth = blockDepthOf(mapPos);
							if (blockDepth < terrainDepth)
								continue;

							unsigned int startHeight = unsigned(blockDepth - terrainDepth);
							if (startHeight >= ChunkSize)
								continue;

							if (BlockIndex& t = blockIndices[Chunk::GetBlockLocalIndex(cellOf(startHeight, i, j))]; t == dirt)
								t = grass;

							for (unsigned int h = startHeight + 1; h < ChunkSize; ++h)
								blockIndices[Chunk::GetBlockLocalIndex(cellOf(h, i, j))] = EmptyBlockIndex;
						}
					}
				};

				const unsigned last = ChunkSize - 1;
				using P = std::pair<int, int>;
				// +X :233-256
				pass(Right, maxHeight.x,
					[&](unsigned x, unsigned y) { return GetBlockIndices(chunkIndices, { 0, x, y }); },
					[](const Vec3i& m) { return P{ m.y, m.z }; },
					[&](const Vec3i& m) { return maxHeight.x - m.x + 1; },
					[](unsigned h, unsigned x, unsigned y) { return Vec3ui{ h, x, y }; });
				// -X :258-281
				pass(Left, maxHeight.x,
					[&](unsigned x, unsigned y) { return GetBlockIndices(chunkIndices, { last, x, y }); },
					[](const Vec3i& m) { return P{ m.y, m.z }; },
					[&](const Vec3i& m) { return maxHeight.x + m.x + 1; },
					[&](unsigned h, unsigned x, unsigned y) { return Vec3ui{ ChunkSize - h - 1, x, y }; });
				// +Y :283-306
				pass(Up, maxHeight.y,
					[&](unsigned x, unsigned z) { return GetBlockIndices(chunkIndices, { x, z, 0 }); },
					[](const Vec3i& m) { return P{ m.x, m.z }; },
					[&](const Vec3i& m) { return maxHeight.y - m.y + 1; },
					[](unsigned h, unsigned x, unsigned z) { return Vec3ui{ x, z, h }; });
				// -Y :308-331
				pass(Down, maxHeight.y,
					[&](unsigned x, unsigned z) { return GetBlockIndices(chunkIndices, { x, z, last }); },
					[](const Vec3i& m) { return P{ m.x, m.z }; },
					[&](const Vec3i& m) { return maxHeight.y + m.y + 1; },
					[&](unsigned h, unsigned x, unsigned z) { return Vec3ui{ x, z, ChunkSize - h - 1 }; });
				// +Z :333-356
				pass(Back, maxHeight.z,
					[&](unsigned x, unsigned y) { return GetBlockIndices(chunkIndices, { x, 0, y }); },
					[](const Vec3i& m) { return P{ m.x, m.y }; },
					[&](const Vec3i& m) { return maxHeight.z - m.z + 1; },
					[](unsigned h, unsigned x, unsigned y) { return Vec3ui{ x, h, y }; });
				// -Z :358-381
				pass(Front, maxHeight.z,
					[&](unsigned x, unsigned y) { return GetBlockIndices(chunkIndices, { x, last, y }); },
					[](const Vec3i& m) { return P{ m.x, m.y }; },
					[&](const Vec3i& m) { return maxHeight.z + m.z + 1; },
					[&](unsigned h, unsigned x, unsigned y) { return Vec3ui{ x, ChunkSize - h - 1, y }; });

				return ok;
			}

			bool GenerateChunk(const BlockLibrary& lib, Chunk& chunk, std::uint32_t seed, const Vec3ui& chunkCount)
			{
				bool ok = true;
				chunk.Reset([&](BlockIndex* blocks) { ok = GenerateChunkBlocks(lib, chunk.GetIndices(), seed, chunkCount, blocks); });
				return ok;
			}

			// Planet.cpp:385-403 (serial here; the threaded variant lives in capi.cpp for the CPU baseline)
			bool GenerateChunks(const BlockLibrary& lib, std::uint32_t seed, const Vec3ui& chunkCount)
			{
				bool ok = true;
				for (int cz = 0; cz < int(chunkCount.z); ++cz)
					for (int cy = 0; cy < int(chunkCount.y); ++cy)
						for (int cx = 0; cx < int(chunkCount.x); ++cx)
						{
							Chunk& chunk = AddChunk(lib, { cx - int(chunkCount.x / 2), cy - int(chunkCount.y / 2), cz - int(chunkCount.z / 2) });
							ok &= GenerateChunk(lib, chunk, seed, chunkCount);
						}
				return ok;
			}

			// Planet.cpp:405-512
			void GeneratePlatform(const BlockLibrary& lib, Direction upDirection, const Vec3i& platformCenter)
			{
				constexpr int platformSize = 15;
				constexpr unsigned int freeHeight = 10;
				const DirectionAxis& dirAxis = s_dirAxis[upDirection];

				Vec3i coordinates = platformCenter;
				int& xPos = coordinates[dirAxis.rightAxis];
				xPos += -dirAxis.rightDir * platformSize / 2;
				int& yPos = coordinates[dirAxis.upAxis];
				int& zPos = coordinates[dirAxis.forwardAxis];
				zPos += -dirAxis.forwardDir * platformSize / 2;

				BlockIndex border = lib.GetBlockIndex("copper_block");
				BlockIndex interior = lib.GetBlockIndex("stone_bricks");

				Vec3i originalCoordinates = coordinates;
				for (unsigned int y = 0; y < freeHeight; ++y)
				{
					unsigned int startingZ = zPos;
					for (unsigned int z = 0; z < platformSize; ++z)
					{
						unsigned int startingX = xPos;
						for (unsigned int x = 0; x < platformSize; ++x)
						{
							BlockIndex b;
							if (y != 0)
								b = EmptyBlockIndex;
							else if (x == 0 || x == platformSize - 1 || z == 0 || z == platformSize - 1)
								b = border;
							else
								b = interior;

							Vec3ui inner;
							Vec3i ci = GetChunkIndicesByBlockIndices(coordinates, &inner);
							if (Chunk* chunk = GetChunk(ci))
								chunk->UpdateBlock(inner, b);

							xPos += dirAxis.rightDir;
						}
						xPos = startingX;
						zPos += dirAxis.forwardDir;
					}

					if (yPos == 0 && dirAxis.upDir < 0)
						break;

					yPos += dirAxis.upDir;
					zPos = startingZ;
				}

				BlockIndex planks = lib.GetBlockIndex("planks");
				coordinates = originalCoordinates;
				for (unsigned int y = 1; /*no cond*/; ++y)
				{
					yPos -= dirAxis.upDir;
					bool hasEmpty = false;

					unsigned int startingZ = zPos;
					for (unsigned int z = 0; z < platformSize; ++z)
					{
						unsigned int startingX = xPos;
						for (unsigned int x = 0; x < platformSize; ++x)
						{
							Vec3ui inner;
							Vec3i ci = GetChunkIndicesByBlockIndices(coordinates, &inner);
							Chunk* chunk = GetChunk(ci);
							if (!chunk)
								continue; // NB: before xPos advances (Planet.cpp:478-482)

							xPos += dirAxis.rightDir;

							if (y % 3 == 0)
							{
								if (x != 0 && x != platformSize - 1 && z != 0 && z != platformSize - 1)
									continue;
							}
							else
							{
								if (chunk->GetBlockContent(inner) != EmptyBlockIndex)
									continue;
								if ((x != 0 && x != platformSize - 1) || (z != 0 && z != platformSize - 1))
									continue;
							}

							hasEmpty = true;
							chunk->UpdateBlock(inner, planks);
						}
						xPos = startingX;
						zPos += dirAxis.forwardDir;
					}

					if (!hasEmpty)
						break;

					zPos = startingZ;
				}
			}

			float GetCornerRadius() const { return m_cornerRadius; }

		private:
			float m_cornerRadius;
			float m_gravity;
	};

	// ---------------------------------------------------------------- Ship (Ship.cpp) — the second ChunkContainer of the reference
	class Ship : public ChunkContainer
	{
		public:
			explicit Ship(float tileSize) : ChunkContainer(tileSize) {}

			// Ship.cpp:21-61: FlatChunk + the dirty-mask rule with Up / Down swapped relative to Planet (z == 0 -> Up, z == 31 -> Down, :48-51)
			Chunk& AddChunk(const BlockLibrary& lib, const Vec3i& indices)
			{
				auto chunk = std::make_unique<Chunk>(lib, *this, indices, m_tileSize);
				chunk->onReset = [this](Chunk* c) { OnChunkUpdated(c, 0x3F); };
				chunk->onBlockUpdated = [this](Chunk* c, const Vec3ui& i, BlockIndex)
				{
					std::uint8_t mask = 0;
					if (i.x == 0) mask |= 1u << Left; else if (i.x == ChunkSize - 1) mask |= 1u << Right;
					if (i.y == 0) mask |= 1u << Front; else if (i.y == ChunkSize - 1) mask |= 1u << Back;
					if (i.z == 0) mask |= 1u << Up; else if (i.z == ChunkSize - 1) mask |= 1u << Down;
					OnChunkUpdated(c, mask);
				};
				Chunk* ptr = chunk.get();
				m_chunks[indices] = std::move(chunk);
				return *ptr;
			}

			// Ship.cpp:115-152
			void Generate(const BlockLibrary& lib, bool small)
			{
				Chunk& chunk = AddChunk(lib, { 0, 0, 0 });
				BlockIndex hull = lib.GetBlockIndex("hull");
				if (hull == InvalidBlockIndex)
					return;
				BlockIndex forcefield = lib.GetBlockIndex("forcefield");
				const unsigned boxSize = small ? 6 : 12, height = small ? 4 : 6;
				const Vec3ui start{ ChunkSize / 2 - boxSize / 2, ChunkSize / 2 - boxSize / 2, ChunkSize / 2 - height / 2 };
				chunk.ResetEmpty(); // Chunk::Reset() without a callback does not fire OnReset (Chunk.inl:97-107)
				for (unsigned z = 0; z < height; ++z)
					for (unsigned y = 0; y < boxSize; ++y)
						for (unsigned x = 0; x < boxSize; ++x)
						{
							if (x != 0 && x != boxSize - 1 && y != 0 && y != boxSize - 1 && z != 0 && z != height - 1)
								continue;
							const Vec3ui pos{ start.x + x, start.y + y, start.z + z };
							if (x == 0 && y == boxSize / 2 && z > 0 && z < height - 1)
								chunk.UpdateBlock(pos, forcefield);
							else
								chunk.UpdateBlock(pos, hull);
						}
			}

			// Ship.cpp:63-93, flattened: (child offset, box centre * blockSize, box lengths * blockSize) per box (FlatChunk.cpp:13-30);
			// one chunk: its own compound (offset zero); several: one child per chunk with a collider, offset = GetChunkOffset
			struct HullBox { Vec3f chunkOffset, center, lengths; };
			std::vector<HullBox> BuildHullCollider() const
			{
				std::vector<HullBox> out;
				const bool compound = m_chunks.size() > 1;
				for (auto&& [idx, chunk] : m_chunks)
				{
					if (!chunk->HasContent())
						continue;
					const Vec3f off = compound ? GetChunkOffset(idx) : Vec3f{};
					for (const Chunk::ColliderBox& b : chunk->BuildBoxCollider())
					{
						const Vec3f c{ b.pos.x + b.size.x / 2.f, b.pos.y + b.size.y / 2.f, b.pos.z + b.size.z / 2.f }; // Nz::Boxf::GetCenter
						out.push_back({ off, c * m_tileSize, b.size * m_tileSize });
					}
				}
				return out;
			}
	};
}
