// gen_kernels.cu — planet chunk generation on sm_100a (replaces Planet::GenerateChunk, reference src/CommonLib/Planet.cpp:160-383)
// plus the small maintenance kernels (x-slab packing, edit scatter, platform script, greedy box collider, P2P halo pull).
//
// B200-first decomposition of GenerateChunk:
//   terrain_kernel  : the Perlin height function depends only on the planet column, not on the chunk.  It is evaluated ONCE per
//                     column and direction (FP64, explicit round-to-nearest intrinsics, no FMA contraction) into int16
//                     `terrainDepth` tiles of 32x32; the reference re-evaluates it in every chunk stacked along the axis.
//   generate_kernel : one CTA per chunk, pure integer work.  Each thread produces 8 consecutive ids and stores them with one
//                     16-byte store (the chunk is written exactly once, 64 KiB, fully coalesced).  The sequential
//                     std::minstd_rand / std::bernoulli_distribution(0.9) stream of Planet.cpp:167-168,219 is reproduced by LCG
//                     skip-ahead: the blocks that consume a draw form an axis-aligned box in the chunk, so the draw index of a
//                     block is closed-form and x_k = 48271^k * x_0 mod (2^31-1) is reached with three table multiplications.
//                     The six terrain passes (Planet.cpp:233-381) commute into "Empty if above any column start, grass if dirt at
//                     a column start", evaluated per block from six per-column start tables in shared memory.
#include "tsom_kernels.h"
#include "tsom_device.cuh"

#include <climits>

namespace tsomk
{
	// ------------------------------------------------------------------------------------------------ siv::PerlinNoise v3.0.0
	__device__ __forceinline__ double pn_fade(double t)
	{
		// t * t * t * (t * (t * 6 - 15) + 10)
		double inner = __dadd_rn(__dmul_rn(t, __dsub_rn(__dmul_rn(t, 6.0), 15.0)), 10.0);
		return __dmul_rn(__dmul_rn(__dmul_rn(t, t), t), inner);
	}

	__device__ __forceinline__ double pn_lerp(double a, double b, double t) { return __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), t)); }

	__device__ __forceinline__ double pn_grad(unsigned hash, double x, double y, double z)
	{
		const unsigned h = hash & 15u;
		const double u = h < 8u ? x : y;
		const double v = h < 4u ? y : ((h == 12u || h == 14u) ? x : z);
		return __dadd_rn((h & 1u) == 0u ? u : -u, (h & 2u) == 0u ? v : -v);
	}

	__device__ double pn_noise3d(const uint8_t* __restrict__ p, double x, double y, double z)
	{
		const double _x = floor(x), _y = floor(y), _z = floor(z);
		const int ix = int(_x) & 255, iy = int(_y) & 255, iz = int(_z) & 255;
		const double fx = __dsub_rn(x, _x), fy = __dsub_rn(y, _y), fz = __dsub_rn(z, _z);
		const double u = pn_fade(fx), v = pn_fade(fy), w = pn_fade(fz);

		const unsigned A = (p[ix & 255] + iy) & 255u;
		const unsigned B = (p[(ix + 1) & 255] + iy) & 255u;
		const unsigned AA = (p[A] + iz) & 255u;
		const unsigned AB = (p[(A + 1) & 255] + iz) & 255u;
		const unsigned BA = (p[B] + iz) & 255u;
		const unsigned BB = (p[(B + 1) & 255] + iz) & 255u;

		const double fx1 = __dsub_rn(fx, 1.0), fy1 = __dsub_rn(fy, 1.0), fz1 = __dsub_rn(fz, 1.0);
		const double p0 = pn_grad(p[AA], fx, fy, fz);
		const double p1 = pn_grad(p[BA], fx1, fy, fz);
		const double p2 = pn_grad(p[AB], fx, fy1, fz);
		const double p3 = pn_grad(p[BB], fx1, fy1, fz);
		const double p4 = pn_grad(p[(AA + 1) & 255], fx, fy, fz1);
		const double p5 = pn_grad(p[(BA + 1) & 255], fx1, fy, fz1);
		const double p6 = pn_grad(p[(AB + 1) & 255], fx, fy1, fz1);
		const double p7 = pn_grad(p[(BB + 1) & 255], fx1, fy1, fz1);

		const double q0 = pn_lerp(p0, p1, u), q1 = pn_lerp(p2, p3, u), q2 = pn_lerp(p4, p5, u), q3 = pn_lerp(p6, p7, u);
		const double r0 = pn_lerp(q0, q1, v), r1 = pn_lerp(q2, q3, v);
		return pn_lerp(r0, r1, w);
	}

	// normalizedOctave2D_01(x, y, 4) with persistence 0.5: sum / 1.875 * 0.5 + 0.5
	__device__ double pn_octave01(const uint8_t* __restrict__ p, double x, double y)
	{
		double result = 0.0, amplitude = 1.0;
#pragma unroll 1
		for (int i = 0; i < 4; ++i)
		{
			result = __dadd_rn(result, __dmul_rn(pn_noise3d(p, x, y, 0.34567), amplitude));
			x = __dmul_rn(x, 2.0);
			y = __dmul_rn(y, 2.0);
			amplitude = __dmul_rn(amplitude, 0.5);
		}
		const double n = __ddiv_rn(result, 1.875);
		return __dadd_rn(__dmul_rn(n, 0.5), 0.5);
	}

	// Tiles: direction d, tile (ta, tb), column (i, j) -> terrain[offset[d] + ((tb * nA + ta) * 1024) + j * 32 + i]
	//   Right/Left : A = world Y (up, local z), B = world Z (local y); i = local y, j = local z; noise(a = wy, b = wz)
	//   Up/Down    : A = world X (local x),     B = world Z (local y); i = local x, j = local y; noise(a = wx, b = wz)
	//   Back/Front : A = world X (local x),     B = world Y (local z); i = local x, j = local z; noise(a = wx, b = wy)
	// One launch for the six directions: the tiles of direction d are the blocks [terrainOffset[d] / 1024, terrainOffset[d + 1] / 1024).
	__global__ void __launch_bounds__(1024, 2) terrain_kernel(const TerrainArgs a)
	{
		// axis ids: 0 = world X, 1 = world Y (up), 2 = world Z
		// the tiles of direction d are the blocks [terrainOffset[d] / 1024, + tiles[d]); directions the call does not need have no tiles
		int dir = 0;
		size_t dirOffset = a.terrainOffset[0];
#pragma unroll
		for (int d = 1; d < 6; ++d)
			if (a.tiles[d] != 0u && size_t(blockIdx.x) >= a.terrainOffset[d] / SLAB)
			{
				dir = d;
				dirOffset = a.terrainOffset[d];
			}
		// selects, not indexing: a dynamically indexed kernel parameter would be copied to local memory
		const bool sideX = dir == D_RIGHT || dir == D_LEFT, sideY = dir == D_UP || dir == D_DOWN;
		const int nA = sideX ? a.tileCount[1] : a.tileCount[0];
		const int originA = sideX ? a.tileOrigin[1] : a.tileOrigin[0];
		const int originB = (sideX || sideY) ? a.tileOrigin[2] : a.tileOrigin[1];
		const int maxHAxis = sideX ? a.maxH[0] : (sideY ? a.maxH[1] : a.maxH[2]);

		__shared__ uint8_t perm[256];
		if (threadIdx.x < 256)
			perm[threadIdx.x] = a.perm[dir * 256 + threadIdx.x];
		__syncthreads();

		const int tile = int(size_t(blockIdx.x) - dirOffset / SLAB);
		const int ta = tile % nA, tb = tile / nA;
		const int i = threadIdx.x & 31, j = threadIdx.x >> 5;

		// world block coordinates of the column along the two table axes (GetBlockIndices: chunk*32 + local - 16)
		int wa, wb;
		if (dir == D_RIGHT || dir == D_LEFT)
		{
			wa = (originA + ta) * CS + j - CS / 2; // world Y from local z = j
			wb = (originB + tb) * CS + i - CS / 2; // world Z from local y = i
		}
		else
		{
			wa = (originA + ta) * CS + i - CS / 2; // world X from local x = i
			wb = (originB + tb) * CS + j - CS / 2; // world Z (local y) or world Y (local z) = j
		}

		const double scale = double(0.02f);
		const double heightScale = double(1.5f);
		const double height = __dmul_rn(pn_octave01(perm, __dmul_rn(double(wa), scale), __dmul_rn(double(wb), scale)), heightScale);

		// int terrainDepth = std::round(std::min<double>(height * (maxH / 2 - freeSpace) + freeSpace, maxH / 2));  freeSpace is size_t
		const int half = maxHAxis / 2;
		const unsigned long long span = static_cast<unsigned long long>(static_cast<long long>(half)) - 30ull;
		double t = __dadd_rn(__dmul_rn(height, __ull2double_rn(span)), 30.0);
		const double cap = double(half);
		if (cap < t)
			t = cap;
		const int td = int(round(t));
		a.terrain[dirOffset + size_t(tile) * SLAB + j * CS + i] = int16_t(td);

		// per-tile range: lets generate_kernel skip a pass for a whole chunk without touching the tile
		__shared__ int redMin[32], redMax[32];
		int mn = td, mx = td;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
			mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
		}
		if ((threadIdx.x & 31) == 0)
		{
			redMin[threadIdx.x >> 5] = mn;
			redMax[threadIdx.x >> 5] = mx;
		}
		__syncthreads();
		if (threadIdx.x < 32)
		{
			mn = redMin[threadIdx.x];
			mx = redMax[threadIdx.x];
#pragma unroll
			for (int o = 16; o > 0; o >>= 1)
			{
				mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
				mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
			}
			if (threadIdx.x == 0)
			{
				const size_t t2 = size_t(blockIdx.x) * 2;
				a.tileRange[t2] = int16_t(mn);
				a.tileRange[t2 + 1] = int16_t(mx);
			}
		}
	}

	void launch_terrain(const TerrainArgs& a, cudaStream_t st)
	{
		// tiles per direction (Direction order): Back / Front X x Y, Down / Up X x Z, Left / Right Y x Z — the layout terrainOffset describes
		size_t tiles = 0;
		for (int d = 0; d < 6; ++d)
			tiles += a.tiles[d];
		if (tiles)
			terrain_kernel<<<unsigned(tiles), 1024, 0, st>>>(a);
	}

	// ------------------------------------------------------------------------------------------------ std::minstd_rand skip-ahead
	constexpr uint32_t LCG_A = 48271u;
	constexpr uint32_t LCG_M = 2147483647u;

	// 32 x 32 -> 64 as ONE IMAD.WIDE: with a plain (u64)x * y the compiler hoists the zero-extension of loop-invariant operands out of
	// the block loop and then multiplies 64 x 32 (three instructions per product)
	__host__ __device__ __forceinline__ unsigned long long mulwide(uint32_t x, uint32_t y)
	{
#ifdef __CUDA_ARCH__
		unsigned long long p;
		asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x), "r"(y));
		return p;
#else
		return static_cast<unsigned long long>(x) * y;
#endif
	}

	__host__ __device__ __forceinline__ uint32_t mulmod31(uint32_t x, uint32_t y)
	{
		const unsigned long long p = mulwide(x, y);
		uint32_t s = uint32_t(p & LCG_M) + uint32_t(p >> 31);
		return s >= LCG_M ? s - LCG_M : s;
	}

	__host__ __device__ inline uint32_t powmod31(uint32_t base, unsigned long long e)
	{
		uint32_t r = 1u;
		while (e)
		{
			if (e & 1ull)
				r = mulmod31(r, base);
			base = mulmod31(base, base);
			e >>= 1;
		}
		return r;
	}

	// x * y mod (2^31 - 1) for y given DOUBLED (y2 = 2y < 2^32): with p = x * y2, hi = (x * y) >> 31 and lo >> 1 = (x * y) & (2^31 - 1),
	// so the folded sum is one IMAD.WIDE + one LEA.HI, and the final conditional subtraction is an unsigned min (sum - M wraps when
	// sum < M)
	__device__ __forceinline__ uint32_t mulmod31_doubled(uint32_t x, uint32_t y2)
	{
		const unsigned long long p = mulwide(x, y2);
		const uint32_t sum = uint32_t(p >> 32) + (uint32_t(p) >> 1);
		return min(sum, sum - LCG_M);
	}

	// 0xFFFF in the low half if a is negative, 0xFFFF0000 in the high half if b is negative: prmt with the sign-replicating selector
	// nibbles (msb set) — __byte_perm masks them off, hence the PTX
	__device__ __forceinline__ uint32_t sign_halves(uint32_t a, uint32_t b)
	{
		uint32_t d;
		asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(d) : "r"(a), "r"(b));
		return d;
	}

	// Eight consecutive drawing blocks of one row: packed ids in w[4]; returns true when one of the draws is a tie (u2 - 1 == bStar,
	// about once in 2^31 draws) and the caller has to run draw8_ties.  e = bStar + 1 - u2: > 0 stone, < 0 stone_mossy; the sign bytes of
	// two e's select a packed id pair with one PRMT + one LOP3.
	__device__ __forceinline__ bool draw8(uint32_t rowMul, const uint32_t (&px2)[8], uint32_t K, uint32_t stonePair, uint32_t deltaPair, uint32_t (&e)[8], uint32_t (&w)[4])
	{
#pragma unroll
		for (int j = 0; j < 8; ++j)
			e[j] = K - mulmod31_doubled(rowMul, px2[j]);
#pragma unroll
		for (int k = 0; k < 4; ++k)
			w[k] = stonePair ^ (deltaPair & sign_halves(e[2 * k], e[2 * k + 1]));
		return min(min(min(e[0], e[1]), min(e[2], e[3])), min(min(e[4], e[5]), min(e[6], e[7]))) == 0u;
	}

	// the first engine output of a tied draw decides: u1 = u2 / A = (bStar + 1) * A^-1 — the same for every tie
	__device__ __forceinline__ void draw8_ties(const uint32_t (&e)[8], uint32_t K, uint32_t aInv, uint32_t aStar, uint32_t stoneId, uint32_t mossyId, uint32_t (&w)[4])
	{
		const uint32_t v = (mulmod31(K, aInv) - 1u) < aStar ? stoneId : mossyId;
#pragma unroll
		for (int j = 0; j < 8; ++j)
			if (e[j] == 0u)
				w[j >> 1] = (j & 1) ? ((w[j >> 1] & 0xFFFFu) | (v << 16)) : ((w[j >> 1] & 0xFFFF0000u) | v);
	}

	constexpr int GEN_THREADS = 256;
	constexpr int ROWS_PER_CHUNK = CS * CS;

	struct GenSmem
	{
		uint8_t start[6][SLAB]; // per pass and column: start height (32 = pass does not touch the column)
		uint32_t powX[CS], powY[CS], powZ[CS];
		uint32_t above[ROWS_PER_CHUNK]; // per x-row (z, y): bit x = emptied by some pass (above a column start)
		uint32_t at[ROWS_PER_CHUNK];    // per x-row: bit x = sits on a column start (dirt -> grass)
		uint4 keepOf[256];              // byte of 8 block bits -> the four packed id pairs with the flagged 16-bit halves cleared
	};

	// 32x32 bit-matrix transpose across a warp: lane i holds row i; afterwards lane j holds column j (bit i = old M[i][j])
	__device__ __forceinline__ uint32_t warp_transpose32(uint32_t v)
	{
		const int lane = threadIdx.x & 31;
#pragma unroll
		for (int k = 16; k >= 1; k >>= 1)
		{
			const uint32_t lowMask = k == 16 ? 0x0000FFFFu : (k == 8 ? 0x00FF00FFu : (k == 4 ? 0x0F0F0F0Fu : (k == 2 ? 0x33333333u : 0x55555555u)));
			const uint32_t other = __shfl_xor_sync(0xffffffffu, v, k);
			if (lane & k)
				v = (v & ~lowMask) | ((other >> k) & lowMask);
			else
				v = (v & lowMask) | ((other << k) & ~lowMask);
		}
		return v;
	}

	// column of 32 cells along an axis with kept interval [lo, hi]: bits above (i > hi or i < lo) and at (i == hi or i == lo);
	// hi = 32 / lo = -1 mean "pass does not touch the column"
	__device__ __forceinline__ void column_masks(int hi, int lo, uint32_t& above, uint32_t& at)
	{
		above = (hi >= 31 ? 0u : ~((2u << hi) - 1u)) | (lo <= 0 ? 0u : ((1u << lo) - 1u));
		at = (hi < 32 ? (1u << hi) : 0u) | (lo >= 0 ? (1u << lo) : 0u);
	}

	// ids of 8 consecutive blocks of one row segment
	struct Ids8
	{
		unsigned v[8];
	};

	// Per pass (order +X Right, -X Left, +Y Up, -Y Down, +Z Back, -Z Front): blockDepth of the chunk's reference layer and the global tile
	// number of its 32 x 32 column tile (Planet.cpp:237,262,287,312,337,362)
	struct PassGeom
	{
		int blockDepth[6];
		size_t tileNo[6];
	};

	__device__ __forceinline__ PassGeom pass_geometry(const GenArgs& a, int cx, int cy, int cz)
	{
		const int wx0 = cx * CS - CS / 2, wy0 = cy * CS - CS / 2, wz0 = cz * CS - CS / 2;
		const int Hx = a.maxH[0], Hy = a.maxH[1], Hz = a.maxH[2];
		PassGeom g;
		g.blockDepth[0] = Hx - wx0 + 1;
		g.blockDepth[1] = Hx + (wx0 + CS - 1) + 1;
		g.blockDepth[2] = Hy - wy0 + 1;
		g.blockDepth[3] = Hy + (wy0 + CS - 1) + 1;
		g.blockDepth[4] = Hz - wz0 + 1;
		g.blockDepth[5] = Hz + (wz0 + CS - 1) + 1;
		const int tx = cx - a.tileOrigin[0], ty = cy - a.tileOrigin[1], tz = cz - a.tileOrigin[2];
		const size_t tYZ = size_t(tz) * a.tileCount[1] + ty; // (A = Y, B = Z)
		const size_t tXZ = size_t(tz) * a.tileCount[0] + tx; // (A = X, B = Z)
		const size_t tXY = size_t(ty) * a.tileCount[0] + tx; // (A = X, B = Y)
		g.tileNo[0] = a.terrainOffset[D_RIGHT] / SLAB + tYZ;
		g.tileNo[1] = a.terrainOffset[D_LEFT] / SLAB + tYZ;
		g.tileNo[2] = a.terrainOffset[D_UP] / SLAB + tXZ;
		g.tileNo[3] = a.terrainOffset[D_DOWN] / SLAB + tXZ;
		g.tileNo[4] = a.terrainOffset[D_BACK] / SLAB + tXY;
		g.tileNo[5] = a.terrainOffset[D_FRONT] / SLAB + tXY;
		return g;
	}

	// One thread per job, between terrain_kernel and generate_kernel: everything generate_kernel's CTA would otherwise find through a
	// chain of dependent global loads (job -> slot -> chunk index -> six tile ranges) in ONE 16-byte descriptor — chunk index, slot, and
	// which of the six terrain passes can touch the chunk (some column has blockDepth >= terrainDepth and blockDepth - terrainDepth < 32).
	constexpr uint32_t GEN_SLOT_MASK = 0x03FFFFFFu; // descriptor.w = slot | active passes << 26

	__global__ void __launch_bounds__(256) gen_classify_kernel(const GenArgs a, int4* __restrict__ desc)
	{
		const uint32_t job = blockIdx.x * 256u + threadIdx.x;
		if (job >= a.nJobs)
			return;
		const uint32_t slot = a.jobSlot[job];
		const int4 ci = a.chunkIdx[slot];
		const PassGeom g = pass_geometry(a, ci.x, ci.y, ci.z);
		const int passDir[6] = { D_RIGHT, D_LEFT, D_UP, D_DOWN, D_BACK, D_FRONT };
		uint32_t act = 0u;
#pragma unroll
		for (int p = 0; p < 6; ++p)
		{
			if (!((a.dirMask >> passDir[p]) & 1u))
				continue; // no tables: the host proved that no chunk of this call reaches the terrain band from this side
			const int tdMin = a.tileRange[g.tileNo[p] * 2], tdMax = a.tileRange[g.tileNo[p] * 2 + 1];
			if (g.blockDepth[p] >= tdMin && g.blockDepth[p] - tdMax < CS)
				act |= 1u << p;
		}
		desc[job] = make_int4(ci.x, ci.y, ci.z, int(slot | (act << 26)));
	}

	void launch_gen_classify(const GenArgs& a, int4* desc, cudaStream_t st)
	{
		if (a.nJobs)
			gen_classify_kernel<<<(a.nJobs + 255u) / 256u, 256, 0, st>>>(a, desc);
	}

#ifndef GEN_OCC
#define GEN_OCC 6 // 40 registers, 6 CTAs per SM: the per-chunk prologue is a chain of dependent global loads, more CTAs in flight hide it (47^3: 2.00 -> 1.87 ms)
#endif
	__global__ void __launch_bounds__(GEN_THREADS, GEN_OCC) generate_kernel(const GenArgs a)
	{
		__shared__ GenSmem sm;
		const int tid = threadIdx.x;
		const uint32_t job = blockIdx.x;
		if (job >= a.nJobs)
			return;
		const int4 ci = a.jobDesc[job]; // chunk index, slot | active passes << 26 (gen_classify_kernel)
		const uint32_t slot = uint32_t(ci.w) & GEN_SLOT_MASK;
		const uint32_t actBits = uint32_t(ci.w) >> 26;

		// world coordinates of local (0,0,0): GetBlockIndices = chunk * 32 + (x, z, y) - 16
		const int wx0 = ci.x * CS - CS / 2; // + local x
		const int wy0 = ci.y * CS - CS / 2; // + local z (up)
		const int wz0 = ci.z * CS - CS / 2; // + local y
		const int Hx = a.maxH[0], Hy = a.maxH[1], Hz = a.maxH[2];

		// ---- terrain passes: column start heights (Planet.cpp:241-248 and the five twins)
		// pass order in sm.start: 0 +X(Right) 1 -X(Left) 2 +Y(Up) 3 -Y(Down) 4 +Z(Back) 5 -Z(Front)
		bool act[6];
		{
			const PassGeom g = pass_geometry(a, ci.x, ci.y, ci.z);
#pragma unroll
			for (int p = 0; p < 6; ++p)
			{
				act[p] = ((actBits >> p) & 1u) != 0u;
				const size_t tileNo = g.tileNo[p];
				const int bd = g.blockDepth[p];
				if (act[p])
				{
					const int16_t* tile = a.terrain + tileNo * SLAB;
#pragma unroll
					for (int k = 0; k < SLAB / GEN_THREADS; ++k)
					{
						const int c = tid + k * GEN_THREADS;
						const int td = tile[c];
						int st = 32;
						if (bd >= td && bd - td < CS)
							st = bd - td;
						sm.start[p][c] = uint8_t(st);
					}
				}
			}
		}
		const bool actX = act[0] | act[1], actY = act[2] | act[3], actZ = act[4] | act[5];
		const bool anyAct = actX | actY | actZ;

		uint16_t* out = a.blocks + size_t(slot) * NB;
		uint16_t* xs = a.xslabs + size_t(slot) * 2 * SLAB;

		// ---- a chunk that lies wholly outside the base fill (depth < 30 for every block, Planet.cpp:205-206): zeros, nothing else.
		// depth = min(Hx - |wx|, Hy - |wz|, Hz - |wy|) (y / z crossed, :199-203); its maximum over the chunk is reached where |w| is smallest
		{
			auto minAbs = [](int w0) { return (w0 <= 0 && w0 + CS - 1 >= 0) ? 0 : min(abs(w0), abs(w0 + CS - 1)); };
			const int maxDepth = min(Hx - minAbs(wx0), min(Hy - minAbs(wz0), Hz - minAbs(wy0)));
			if (maxDepth < 30)
			{
				const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
				for (int k = 0; k < NB / 8 / GEN_THREADS; ++k)
					reinterpret_cast<uint4*>(out)[tid + k * GEN_THREADS] = z4;
				reinterpret_cast<uint4*>(xs)[tid] = z4; // 2 x 1024 ids = 256 x 16 B
				if (tid == 0)
					a.summary[slot] = SUM_EMPTY;
				return;
			}
		}

		// ---- RNG: the blocks that draw (depth - 30 > 18, Planet.cpp:218-219) form a box in local coordinates
		const int Lx = Hx - 49, Ly = Hy - 49, Lz = Hz - 49; // |wx| <= Lx, |wz| <= Ly, |wy| <= Lz (y/z crossed, :199-203)
		const int ax0 = max(0, -Lx - wx0), ax1 = min(CS - 1, Lx - wx0);
		const int ay0 = max(0, -Ly - wz0), ay1 = min(CS - 1, Ly - wz0); // local y <-> world Z
		const int az0 = max(0, -Lz - wy0), az1 = min(CS - 1, Lz - wy0); // local z <-> world Y
		const int NX = ax1 - ax0 + 1, NY = ay1 - ay0 + 1;
		const bool anyDraw = (NX > 0) && (NY > 0) && (az1 >= az0);
		const bool fullDraw = !anyAct && ax0 == 0 && ax1 == CS - 1 && ay0 == 0 && ay1 == CS - 1 && az0 == 0 && az1 == CS - 1;
		// the central shaft (|wx| <= 2 and |wz| <= 2, Planet.cpp:221-222) crosses the chunks with cx == 0 and cz == 0
		const bool shaftChunk = ci.x == 0 && ci.z == 0;

		// ---- every block draws and there is no shaft: before the terrain passes the chunk is a function of chunkSeed = seed + cx +
		// cy + cz alone (Planet.cpp:165-168) — its template chunk, which sits in L2 (a 47^3 planet has 139 of them for 90 k such
		// chunks).  Without an active pass the chunk IS the template; with one, the passes only empty blocks of it (there is no dirt to
		// turn into grass among stone / stone_mossy), see the masked copy below.
		const bool fromTemplate = ax0 == 0 && ax1 == CS - 1 && ay0 == 0 && ay1 == CS - 1 && az0 == 0 && az1 == CS - 1 && !shaftChunk && a.templates != nullptr &&
		                          uint32_t(ci.x + ci.y + ci.z - a.templateSumMin) < a.nTemplates;
		const uint4* tsrc = reinterpret_cast<const uint4*>(a.templates + size_t(uint32_t(ci.x + ci.y + ci.z - a.templateSumMin)) * TEMPLATE_STRIDE);
		if (fromTemplate && !anyAct)
		{
			uint4* dst = reinterpret_cast<uint4*>(out);
#pragma unroll 8
			for (int k = 0; k < NB / 8 / GEN_THREADS; ++k)
				dst[tid + k * GEN_THREADS] = __ldg(tsrc + tid + k * GEN_THREADS);
			reinterpret_cast<uint4*>(xs)[tid] = __ldg(tsrc + NB / 8 + tid);
			if (tid == 0)
				a.summary[slot] = a.fastSummary;
			return;
		}
		if (anyDraw && !fromTemplate && tid < 3 * CS)
		{
			const int j = tid & 31;
			const unsigned long long stride = tid < CS ? 2ull : (tid < 2 * CS ? 2ull * NX : 2ull * NX * NY);
			uint32_t* dst = tid < CS ? sm.powX : (tid < 2 * CS ? sm.powY : sm.powZ);
			dst[j] = powmod31(LCG_A, stride * j);
		}
		__syncthreads();

		// ---- the six passes as two bit masks per x-row: "Empty above any column start, grass if dirt at a column start"
		// (Planet.cpp:250-254 and twins).  X passes start per row; Y / Z passes start per (y, x) / (z, x) column: a lane builds its
		// column's 32-bit mask along the axis and a warp transpose turns 32 columns into 32 row masks.
		if (anyAct)
		{
			const int lane = tid & 31, warp = tid >> 5;
			{
				uint32_t m[4];
#pragma unroll
				for (int k = 0; k < 4; ++k)
					m[k] = ~((((unsigned(tid) >> (2 * k)) & 1u) * 0xFFFFu) | (((unsigned(tid) >> (2 * k + 1)) & 1u) * 0xFFFF0000u));
				sm.keepOf[tid] = make_uint4(m[0], m[1], m[2], m[3]);
			}
#pragma unroll
			for (int k = 0; k < ROWS_PER_CHUNK / GEN_THREADS; ++k)
			{
				const int row = tid + k * GEN_THREADS;
				uint32_t ab = 0u, at = 0u;
				if (actX)
					column_masks(act[0] ? int(sm.start[0][row]) : 32, CS - 1 - (act[1] ? int(sm.start[1][row]) : 32), ab, at);
				sm.above[row] = ab;
				sm.at[row] = at;
			}
			__syncthreads();
			if (actY)
			{
				// columns (y, x) along local z; warp owns y = 4 * warp .. 4 * warp + 3, lane = x -> after the transpose lane = z
#pragma unroll 1
				for (int y = 4 * warp; y < 4 * warp + 4; ++y)
				{
					uint32_t ab, at;
					column_masks(act[2] ? int(sm.start[2][y * CS + lane]) : 32, CS - 1 - (act[3] ? int(sm.start[3][y * CS + lane]) : 32), ab, at);
					ab = warp_transpose32(ab);
					at = warp_transpose32(at);
					sm.above[lane * CS + y] |= ab;
					sm.at[lane * CS + y] |= at;
				}
				__syncthreads();
			}
			if (actZ)
			{
				// columns (z, x) along local y; warp owns z = 4 * warp .. 4 * warp + 3, lane = x -> after the transpose lane = y
#pragma unroll 1
				for (int z = 4 * warp; z < 4 * warp + 4; ++z)
				{
					uint32_t ab, at;
					column_masks(act[4] ? int(sm.start[4][z * CS + lane]) : 32, CS - 1 - (act[5] ? int(sm.start[5][z * CS + lane]) : 32), ab, at);
					ab = warp_transpose32(ab);
					at = warp_transpose32(at);
					sm.above[z * CS + lane] |= ab;
					sm.at[z * CS + lane] |= at;
				}
				__syncthreads();
			}
		}

		// ---- template chunk with active passes: template AND NOT "above a column start" (thread = 8 blocks of a row, like below)
		if (fromTemplate)
		{
			const int xq = tid & 3, y = (tid >> 2) & 31, xb = xq * 8;
			uint32_t orAll = 0u, carved = 0u;
#pragma unroll 4
			for (int it = 0; it < CS / 2; ++it)
			{
				const int col = (2 * it + (tid >> 7)) * CS + y;
				uint4 w = __ldg(tsrc + col * (CS / 8) + xq);
				const uint32_t ab8 = (sm.above[col] >> xb) & 0xFFu;
				const uint4 m = sm.keepOf[ab8];
				w.x &= m.x;
				w.y &= m.y;
				w.z &= m.z;
				w.w &= m.w;
				*reinterpret_cast<uint4*>(out + size_t(col) * CS + xb) = w;
				if (xq == 0)
					xs[col] = uint16_t(w.x & 0xFFFFu);
				if (xq == 3)
					xs[SLAB + col] = uint16_t(w.w >> 16);
				orAll |= (w.x | w.y) | (w.z | w.w);
				carved |= ab8;
			}
			const int anyBlock = __syncthreads_or(orAll != 0u);
			const int anyCarved = __syncthreads_or(carved != 0u);
			if (tid == 0)
				a.summary[slot] = !anyBlock ? uint8_t(SUM_EMPTY) : (!anyCarved ? a.fastSummary : uint8_t(SUM_UNKNOWN));
			return;
		}

		uint32_t chunkSeed = a.seed + uint32_t(ci.x) + uint32_t(ci.y) + uint32_t(ci.z);
		uint32_t x0 = chunkSeed % LCG_M; // std::linear_congruential_engine::seed: 0 -> 1
		if (x0 == 0u)
			x0 = 1u;
		const uint32_t base2 = mulmod31(x0, mulmod31(LCG_A, LCG_A)); // second engine output of draw 0

		// ---- per-thread constants: this thread always produces x in [8*xq, 8*xq+8) of rows with local y = `y`
		const int xq = tid & 3;
		const int y = (tid >> 2) & 31;
		const int xb = xq * 8;
		const int wz = wz0 + y;
		const int absWz = abs(wz);
		const int dYZ_y = Hy - absWz; // the maxHeight.y - |blockPos.z| term

		int dX[8];
		uint32_t px2[8];
		int dXmin = INT_MAX, dXmax = INT_MIN;
		unsigned shaft = 0u;
#pragma unroll
		for (int j = 0; j < 8; ++j)
		{
			const int x = xb + j, wx = wx0 + x;
			dX[j] = Hx - abs(wx);
			dXmin = min(dXmin, dX[j]);
			dXmax = max(dXmax, dX[j]);
			px2[j] = (anyDraw && x >= ax0 && x <= ax1) ? sm.powX[x - ax0] << 1 : 0u; // doubled: see mulmod31_doubled
			if (abs(wx) <= 2 && absWz <= 2)
				shaft |= 1u << j;
		}
		const bool rowYDraws = anyDraw && y >= ay0 && y <= ay1;
		const uint32_t rowMulY = rowYDraws ? mulmod31(base2, sm.powY[y - ay0]) : 0u;
		const uint32_t stoneId = a.stone, mossyId = a.stoneMossy, dirtId = a.dirt, grassId = a.grass, snowId = a.snow;
		const uint32_t bStar = a.bStar;

		uint32_t keep[4]; // the central shaft: Empty after the draw (Planet.cpp:221-222); the same x for every row of this thread
#pragma unroll
		for (int k = 0; k < 4; ++k)
			keep[k] = ~((((shaft >> (2 * k)) & 1u) * 0xFFFFu) | (((shaft >> (2 * k + 1)) & 1u) * 0xFFFF0000u));
		const uint32_t K = bStar + 1u;
		const uint32_t stonePair = stoneId | (stoneId << 16), deltaPair = stonePair ^ (mossyId | (mossyId << 16));

		// ---- chunks in which every block draws and no terrain pass is active (two thirds of a large planet): a loop without the
		// depth classes and the pass masks.  Per block: x_{2n+2} = rowMul * A^(2x') mod (2^31 - 1) from ONE wide multiply by the doubled
		// power (hi = p >> 31, lo >> 1 = p & (2^31 - 1)), e = bStar + 1 - u2 (> 0 stone, < 0 stone_mossy, == 0 the first draw decides),
		// and the sign bytes of two e's select a packed id pair with one PRMT + one LOP3.
		if (fullDraw)
		{
			if (tid == 0)
				a.summary[slot] = shaftChunk ? uint8_t(SUM_UNKNOWN) : a.fastSummary;
			int col = (tid >> 7) * CS + y;
#pragma unroll 2
			for (int it = 0; it < CS / 2; ++it, col += 2 * CS)
			{
				const uint32_t rowMul = mulmod31(rowMulY, sm.powZ[2 * it + (tid >> 7)]);
				uint32_t e[8], w[4];
				if (draw8(rowMul, px2, K, stonePair, deltaPair, e, w))
					draw8_ties(e, K, a.aInv, a.aStar, stoneId, mossyId, w);
#pragma unroll
				for (int k = 0; k < 4; ++k)
					w[k] &= keep[k];
				*reinterpret_cast<uint4*>(out + size_t(col) * CS + xb) = make_uint4(w[0], w[1], w[2], w[3]);
				if (xq == 0)
					xs[col] = uint16_t(w[0] & 0xFFFFu);
				if (xq == 3)
					xs[SLAB + col] = uint16_t(w[3] >> 16);
			}
			return;
		}

		// depth class: 0 Empty (< 30), 1 snow (<= 36), 2 dirt (<= 48), 3 stone / stone_mossy by draw (Planet.cpp:205-219).  The class is
		// monotone in the depth, so class(min(dX, dYZ)) = min(class(dX), class(dYZ)): the x part is a per-thread constant
		auto depthClass = [](int d) { return (d >= 30) + (d >= 37) + (d >= 49); };
		const int cXmin = depthClass(dXmin), cXmax = depthClass(dXmax);
		uint32_t orAll = 0u, minPair = 0xFFFFFFFFu; // OR and per-half minimum of every id pair this thread stores (chunk summary)

#pragma unroll 1
		for (int it = 0; it < CS / 2; ++it)
		{
			const int z = 2 * it + (tid >> 7);
			const int wy = wy0 + z;
			const int mYZ = min(dYZ_y, Hz - abs(wy));
			const int col = z * CS + y;

			uint32_t ab8 = 0u, at8 = 0u;
			if (anyAct)
			{
				ab8 = (sm.above[col] >> xb) & 0xFFu;
				at8 = (sm.at[col] >> xb) & 0xFFu & ~ab8;
			}

			uint32_t w[4] = { 0u, 0u, 0u, 0u };
			if (mYZ >= 30 && ab8 != 0xFFu)
			{
				const int cYZ = depthClass(mYZ);
				const int cMin = min(cXmin, cYZ), cMax = min(cXmax, cYZ);
				if (cMin == 3)
				{
					// all eight blocks draw
					const uint32_t rowMul = mulmod31(rowMulY, sm.powZ[z - az0]);
					uint32_t e[8];
					if (draw8(rowMul, px2, K, stonePair, deltaPair, e, w))
						draw8_ties(e, K, a.aInv, a.aStar, stoneId, mossyId, w);
				}
				else if (cMin == cMax)
				{
					const unsigned v = cMin == 0 ? 0u : (cMin == 1 ? snowId : dirtId);
#pragma unroll
					for (int k = 0; k < 4; ++k)
						w[k] = v | (v << 16);
				}
				else
				{
					const bool rowDraws = rowYDraws && z >= az0 && z <= az1;
					const uint32_t rowMul = rowDraws ? mulmod31(rowMulY, sm.powZ[z - az0]) : 0u;
#pragma unroll
					for (int j = 0; j < 8; ++j)
					{
						const int depth = min(dX[j], mYZ);
						unsigned v;
						if (depth >= 49)
						{
							// bernoulli(0.9) of draw n == (u2-1, u1-1) lexicographically below (bStar, aStar), u2 = x_{2n+2}
							const uint32_t u2 = mulmod31_doubled(rowMul, px2[j]);
							bool stone = u2 <= bStar;
							if (u2 == K) // the first draw decides (about once in 2^31 draws)
								stone = (mulmod31(u2, a.aInv) - 1u) < a.aStar;
							v = stone ? stoneId : mossyId;
						}
						else
							v = depth < 30 ? 0u : (depth <= 36 ? snowId : dirtId);
						w[j >> 1] |= v << (16 * (j & 1));
					}
				}

				// the central shaft, then the terrain passes on the packed words: Empty above a column start, dirt -> grass on it
#pragma unroll
				for (int k = 0; k < 4; ++k)
					w[k] &= keep[k];
				if (ab8)
				{
					const uint4 m = sm.keepOf[ab8];
					w[0] &= m.x;
					w[1] &= m.y;
					w[2] &= m.z;
					w[3] &= m.w;
				}
				if (at8 && cMin <= 2 && cMax >= 2) // only a dirt block can turn into grass
				{
#pragma unroll
					for (int k = 0; k < 4; ++k)
					{
						if (((at8 >> (2 * k)) & 1u) && (w[k] & 0xFFFFu) == dirtId)
							w[k] = (w[k] & 0xFFFF0000u) | grassId;
						if (((at8 >> (2 * k + 1)) & 1u) && (w[k] >> 16) == dirtId)
							w[k] = (w[k] & 0xFFFFu) | (grassId << 16);
					}
				}
			}

			*reinterpret_cast<uint4*>(out + size_t(col) * CS + xb) = make_uint4(w[0], w[1], w[2], w[3]);
			if (xq == 0)
				xs[col] = uint16_t(w[0] & 0xFFFFu);
			if (xq == 3)
				xs[SLAB + col] = uint16_t(w[3] >> 16);
			orAll |= (w[0] | w[1]) | (w[2] | w[3]);
			minPair = __vminu2(__vminu2(minPair, __vminu2(w[0], w[1])), __vminu2(w[2], w[3]));
		}
		// ---- summary: all Empty, or (when the five generator ids are opaque types) no Empty block at all
		const int anyBlock = __syncthreads_or(orAll != 0u);
		const int noEmpty = __syncthreads_and((minPair & 0xFFFFu) != 0u && (minPair >> 16) != 0u);
		if (tid == 0)
			a.summary[slot] = !anyBlock ? uint8_t(SUM_EMPTY) : ((noEmpty && a.allOpaque) ? uint8_t(SUM_OPAQUE) : uint8_t(SUM_UNKNOWN));
	}

	// template chunks: every block draws (n = linear block index), no pass, no shaft; template t has chunkSeed = seed + templateSumMin + t
	__global__ void __launch_bounds__(GEN_THREADS) template_kernel(const GenArgs a, uint16_t* __restrict__ templates)
	{
		__shared__ uint32_t powX[CS], powY[CS], powZ[CS];
		const int tid = threadIdx.x;
		const uint32_t t = blockIdx.x;
		if (tid < 3 * CS)
		{
			const int j = tid & 31;
			const unsigned long long stride = tid < CS ? 2ull : (tid < 2 * CS ? 2ull * CS : 2ull * CS * CS);
			uint32_t* dst = tid < CS ? powX : (tid < 2 * CS ? powY : powZ);
			dst[j] = powmod31(LCG_A, stride * j);
		}
		__syncthreads();
		const uint32_t chunkSeed = a.seed + uint32_t(a.templateSumMin + int(t));
		uint32_t x0 = chunkSeed % LCG_M;
		if (x0 == 0u)
			x0 = 1u;
		const uint32_t base2 = mulmod31(x0, mulmod31(LCG_A, LCG_A));
		const int xq = tid & 3, y = (tid >> 2) & 31, xb = xq * 8;
		uint32_t px2[8];
#pragma unroll
		for (int j = 0; j < 8; ++j)
			px2[j] = powX[xb + j] << 1;
		const uint32_t rowMulY = mulmod31(base2, powY[y]);
		const uint32_t stoneId = a.stone, mossyId = a.stoneMossy;
		const uint32_t K = a.bStar + 1u;
		const uint32_t stonePair = stoneId | (stoneId << 16), deltaPair = stonePair ^ (mossyId | (mossyId << 16));
		uint16_t* out = templates + size_t(t) * TEMPLATE_STRIDE;
		uint16_t* xs = out + NB;
		int col = (tid >> 7) * CS + y;
#pragma unroll 2
		for (int it = 0; it < CS / 2; ++it, col += 2 * CS)
		{
			const uint32_t rowMul = mulmod31(rowMulY, powZ[2 * it + (tid >> 7)]);
			uint32_t e[8], w[4];
			if (draw8(rowMul, px2, K, stonePair, deltaPair, e, w))
				draw8_ties(e, K, a.aInv, a.aStar, stoneId, mossyId, w);
			*reinterpret_cast<uint4*>(out + size_t(col) * CS + xb) = make_uint4(w[0], w[1], w[2], w[3]);
			if (xq == 0)
				xs[col] = uint16_t(w[0] & 0xFFFFu);
			if (xq == 3)
				xs[SLAB + col] = uint16_t(w[3] >> 16);
		}
	}

	void launch_templates(const GenArgs& a, uint16_t* templates, cudaStream_t st)
	{
		if (a.nTemplates)
			template_kernel<<<a.nTemplates, GEN_THREADS, 0, st>>>(a, templates);
	}

	void launch_generate(const GenArgs& a, int grid, cudaStream_t st)
	{
		generate_kernel<<<grid, GEN_THREADS, 0, st>>>(a);
	}

	// ------------------------------------------------------------------------------------------------ x-slab packing
	// xslabs[slot][0][z][y] = ids[z][y][0], xslabs[slot][1][z][y] = ids[z][y][31]: what the Left/Right neighbours stage as halo
	__global__ void __launch_bounds__(256) build_xslabs_kernel(const uint16_t* __restrict__ blocks, uint16_t* __restrict__ xslabs, uint8_t* __restrict__ summary,
	                                                           const uint32_t* __restrict__ slots, uint32_t n)
	{
		const uint32_t i = blockIdx.x;
		if (i >= n)
			return;
		const uint32_t slot = slots ? slots[i] : i;
		const uint16_t* src = blocks + size_t(slot) * NB;
		uint16_t* dst = xslabs + size_t(slot) * 2 * SLAB;
		for (int c = threadIdx.x; c < SLAB; c += 256)
		{
			dst[c] = src[c * CS];
			dst[SLAB + c] = src[c * CS + CS - 1];
		}
		if (threadIdx.x == 0)
			summary[slot] = SUM_UNKNOWN; // new content of unknown make-up
	}

	void launch_build_xslabs(const uint16_t* blocks, uint16_t* xslabs, uint8_t* summary, const uint32_t* slots, uint32_t n, cudaStream_t st)
	{
		if (n)
			build_xslabs_kernel<<<n, 256, 0, st>>>(blocks, xslabs, summary, slots, n);
	}

	// ------------------------------------------------------------------------------------------------ edits
	// Chunk::UpdateBlock (Chunk.cpp:348-366) + the neighbour-mask rule of Planet.cpp:39-58, single-writer form (platform script)
	// neighbour mask of one edited block: Planet.cpp:39-58, or Ship.cpp:36-54 (Up / Down swapped) when `ship`
	__device__ __forceinline__ unsigned edit_mask(int x, int y, int z, bool ship)
	{
		unsigned mask = 0x80u; // bit 7: "is in m_invalidatedChunks" (a mask of 0 still invalidates the chunk itself)
		if (x == 0)
			mask |= 1u << D_LEFT;
		else if (x == CS - 1)
			mask |= 1u << D_RIGHT;
		if (y == 0)
			mask |= 1u << D_FRONT;
		else if (y == CS - 1)
			mask |= 1u << D_BACK;
		if (z == 0)
			mask |= 1u << (ship ? D_UP : D_DOWN);
		else if (z == CS - 1)
			mask |= 1u << (ship ? D_DOWN : D_UP);
		return mask;
	}

	// Batched edits: one thread per edit.  Only the order of edits to the SAME block matters (the later one wins), so a first
	// kernel records, per edited block, the index of its last edit in an open-addressing hash table (64-bit CAS on the key,
	// atomicMax on the value) and a second kernel lets exactly that edit write.  The neighbour masks are ORed atomically.
	constexpr int EDIT_THREADS = 256;

	__device__ __forceinline__ uint32_t edit_hash(unsigned long long k)
	{
		k ^= k >> 33;
		k *= 0xff51afd7ed558ccdull;
		k ^= k >> 33;
		return uint32_t(k);
	}

	__global__ void __launch_bounds__(EDIT_THREADS) edit_claim_kernel(const EditDev* __restrict__ edits, uint32_t n, unsigned long long* keys, uint32_t* vals, uint32_t capMask)
	{
		const uint32_t i = blockIdx.x * EDIT_THREADS + threadIdx.x;
		if (i >= n)
			return;
		const EditDev e = edits[i];
		if (e.slot == 0xFFFFFFFFu)
			return;
		const unsigned long long key = (static_cast<unsigned long long>(e.slot) << 15) | (e.packed & 0x7FFFu);
		uint32_t h = edit_hash(key) & capMask;
		for (;;)
		{
			const unsigned long long prev = atomicCAS(&keys[h], ~0ull, key);
			if (prev == ~0ull || prev == key)
			{
				atomicMax(&vals[h], i);
				return;
			}
			h = (h + 1u) & capMask;
		}
	}

	__global__ void __launch_bounds__(EDIT_THREADS) edit_apply_kernel(uint16_t* blocks, uint16_t* xslabs, uint8_t* invalidMask, uint8_t* summary, const EditDev* __restrict__ edits,
	                                                                  uint32_t n, const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t capMask, int shipMask)
	{
		const uint32_t i = blockIdx.x * EDIT_THREADS + threadIdx.x;
		if (i >= n)
			return;
		const EditDev e = edits[i];
		if (e.slot == 0xFFFFFFFFu)
			return;
		const unsigned long long key = (static_cast<unsigned long long>(e.slot) << 15) | (e.packed & 0x7FFFu);
		uint32_t h = edit_hash(key) & capMask;
		while (keys[h] != key)
			h = (h + 1u) & capMask;

		const int x = int(e.packed & 31u), y = int((e.packed >> 5) & 31u), z = int((e.packed >> 10) & 31u);
		const uint16_t id = uint16_t(e.packed >> 16);
		if (vals[h] == i)
		{
			blocks[size_t(e.slot) * NB + (z * CS + y) * CS + x] = id;
			if (x == 0)
				xslabs[size_t(e.slot) * 2 * SLAB + z * CS + y] = id;
			else if (x == CS - 1)
				xslabs[size_t(e.slot) * 2 * SLAB + SLAB + z * CS + y] = id;
		}
		summary[e.slot] = SUM_UNKNOWN;
		// every edit raises its mask, overwritten or not (each UpdateBlock call fires OnBlockUpdated, Chunk.cpp:365)
		atomicOr(reinterpret_cast<unsigned int*>(invalidMask) + (e.slot >> 2), edit_mask(x, y, z, shipMask != 0) << (8u * (e.slot & 3u)));
	}

	void launch_apply_edits(uint16_t* blocks, uint16_t* xslabs, uint8_t* invalidMask, uint8_t* summary, const EditDev* edits, uint32_t n, unsigned long long* keys, uint32_t* vals,
	                        uint32_t cap, int shipMask, cudaStream_t st)
	{
		if (!n)
			return;
		const int grid = int((n + EDIT_THREADS - 1) / EDIT_THREADS);
		edit_claim_kernel<<<grid, EDIT_THREADS, 0, st>>>(edits, n, keys, vals, cap - 1u);
		edit_apply_kernel<<<grid, EDIT_THREADS, 0, st>>>(blocks, xslabs, invalidMask, summary, edits, n, keys, vals, cap - 1u, shipMask);
	}

	// ------------------------------------------------------------------------------------------------ save format (Chunk::Serialize)
	// Device part of Chunk::Serialize / Deserialize (reference src/CommonLib/Chunk.cpp:252-346): the palette of block types that
	// occur in a chunk (ascending BlockIndex — the order the reference writes the names in, from m_blockTypeCount) and one
	// palette index per block: uint8 when the chunk has at most 8 distinct types, uint16 otherwise (Chunk.cpp:335-345).
	// One CTA per chunk: presence bitmap over the 65,536 possible ids in shared memory, popcount prefix -> rank of every id.
	constexpr int PAL_WORDS = 2048; // 65536 ids / 32

	__global__ void __launch_bounds__(256) palettize_kernel(const uint16_t* __restrict__ blocks, const uint32_t* __restrict__ slots, uint32_t n, uint16_t* __restrict__ palette,
	                                                        uint32_t stride, uint32_t* __restrict__ counts, uint8_t* __restrict__ payload, int bigEndian)
	{
		__shared__ uint32_t present[PAL_WORDS];
		__shared__ uint16_t rankBase[PAL_WORDS];
		__shared__ uint32_t warpTot[8];
		const uint32_t i = blockIdx.x;
		if (i >= n)
			return;
		const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
		const uint4* src = reinterpret_cast<const uint4*>(blocks + size_t(slots[i]) * NB);
		for (int w = tid; w < PAL_WORDS; w += 256)
			present[w] = 0u;
		__syncthreads();

		unsigned last = 0xFFFFFFFFu;
		for (int k = 0; k < NB / 8 / 256; ++k)
		{
			const uint4 v = src[tid + 256 * k];
			const uint32_t ws[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
			for (int j = 0; j < 8; ++j)
			{
				const unsigned id = (ws[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
				if (id != last)
				{
					atomicOr(&present[id >> 5], 1u << (id & 31u));
					last = id;
				}
			}
		}
		__syncthreads();

		// exclusive prefix of the popcounts: thread t owns words 8t .. 8t+7
		uint32_t pc[8], sum = 0;
#pragma unroll
		for (int j = 0; j < 8; ++j)
		{
			pc[j] = __popc(present[8 * tid + j]);
			sum += pc[j];
		}
		uint32_t incl = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o)
				incl += v;
		}
		if (lane == 31)
			warpTot[warp] = incl;
		__syncthreads();
		uint32_t base = 0, K = 0;
#pragma unroll
		for (int w = 0; w < 8; ++w)
		{
			if (w < warp)
				base += warpTot[w];
			K += warpTot[w];
		}
		uint32_t run = base + incl - sum;
#pragma unroll
		for (int j = 0; j < 8; ++j)
		{
			rankBase[8 * tid + j] = uint16_t(run);
			uint32_t bits = present[8 * tid + j];
			uint32_t r = run;
			while (bits)
			{
				const int b = __ffs(bits) - 1;
				bits &= bits - 1u;
				if (r < stride)
					palette[size_t(i) * stride + r] = uint16_t((8 * tid + j) * 32 + b);
				++r;
			}
			run += pc[j];
		}
		if (tid == 0)
			counts[i] = K;
		__syncthreads();

		const bool wide = K > 8u;
		uint8_t* dst = payload + size_t(i) * (NB * 2);
		for (int k = 0; k < NB / 8 / 256; ++k)
		{
			const uint4 v = src[tid + 256 * k];
			const uint32_t ws[4] = { v.x, v.y, v.z, v.w };
			uint32_t idx[8];
#pragma unroll
			for (int j = 0; j < 8; ++j)
			{
				const unsigned id = (ws[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
				idx[j] = rankBase[id >> 5] + __popc(present[id >> 5] & ((1u << (id & 31u)) - 1u));
			}
			if (wide)
			{
				uint32_t o[4];
#pragma unroll
				for (int j = 0; j < 4; ++j)
				{
					uint32_t lo = idx[2 * j], hi = idx[2 * j + 1];
					if (bigEndian)
					{
						lo = ((lo & 0xFFu) << 8) | (lo >> 8);
						hi = ((hi & 0xFFu) << 8) | (hi >> 8);
					}
					o[j] = lo | (hi << 16);
				}
				reinterpret_cast<uint4*>(dst)[tid + 256 * k] = make_uint4(o[0], o[1], o[2], o[3]);
			}
			else
			{
				const uint32_t lo = idx[0] | (idx[1] << 8) | (idx[2] << 16) | (idx[3] << 24);
				const uint32_t hi = idx[4] | (idx[5] << 8) | (idx[6] << 16) | (idx[7] << 24);
				reinterpret_cast<uint2*>(dst)[tid + 256 * k] = make_uint2(lo, hi);
			}
		}
	}

	__global__ void __launch_bounds__(256) depalettize_kernel(uint16_t* __restrict__ blocks, const uint32_t* __restrict__ slots, uint32_t n, const uint16_t* __restrict__ palette,
	                                                          uint32_t stride, const uint32_t* __restrict__ counts, const uint8_t* __restrict__ payload, int bigEndian, uint32_t* errFlag)
	{
		const uint32_t i = blockIdx.x;
		if (i >= n)
			return;
		const int tid = threadIdx.x;
		const uint32_t K = counts[i];
		const bool wide = K > 8u;
		const uint16_t* pal = palette + size_t(i) * stride;
		const uint8_t* src = payload + size_t(i) * (NB * 2);
		uint4* dst = reinterpret_cast<uint4*>(blocks + size_t(slots[i]) * NB);
		bool bad = false;
		for (int k = 0; k < NB / 8 / 256; ++k)
		{
			uint32_t idx[8];
			if (wide)
			{
				const uint4 v = reinterpret_cast<const uint4*>(src)[tid + 256 * k];
				const uint32_t ws[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
				for (int j = 0; j < 8; ++j)
				{
					uint32_t x = (ws[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
					if (bigEndian)
						x = ((x & 0xFFu) << 8) | (x >> 8);
					idx[j] = x;
				}
			}
			else
			{
				const uint2 v = reinterpret_cast<const uint2*>(src)[tid + 256 * k];
#pragma unroll
				for (int j = 0; j < 8; ++j)
					idx[j] = ((j < 4 ? v.x : v.y) >> (8 * (j & 3))) & 0xFFu;
			}
			uint32_t id[8];
#pragma unroll
			for (int j = 0; j < 8; ++j)
			{
				bad |= idx[j] >= K;
				id[j] = idx[j] < K ? uint32_t(__ldg(pal + idx[j])) : 0u;
			}
			dst[tid + 256 * k] = make_uint4(id[0] | (id[1] << 16), id[2] | (id[3] << 16), id[4] | (id[5] << 16), id[6] | (id[7] << 16));
		}
		if (bad)
			atomicExch(errFlag, 1u);
	}

	void launch_palettize(const uint16_t* blocks, const uint32_t* slots, uint32_t n, uint16_t* palette, uint32_t stride, uint32_t* counts, uint8_t* payload, int bigEndian, cudaStream_t st)
	{
		if (n)
			palettize_kernel<<<n, 256, 0, st>>>(blocks, slots, n, palette, stride, counts, payload, bigEndian);
	}

	void launch_depalettize(uint16_t* blocks, const uint32_t* slots, uint32_t n, const uint16_t* palette, uint32_t stride, const uint32_t* counts, const uint8_t* payload, int bigEndian,
	                        uint32_t* errFlag, cudaStream_t st)
	{
		if (n)
			depalettize_kernel<<<n, 256, 0, st>>>(blocks, slots, n, palette, stride, counts, payload, bigEndian, errFlag);
	}

	// ------------------------------------------------------------------------------------------------ Planet::GeneratePlatform
	// The edit script of Planet.cpp:405-512 on the device-resident blocks, one CTA, one thread per platform cell (15 x 15).
	// Why it parallelises: every cell of every layer names its own block (right / forward / up are three different axes), so no edit
	// of the script overwrites or reads the result of another one; the only sequential part is WHEN the pillar loop stops (the first
	// layer that places nothing), which is one CTA-wide OR per layer.
	struct PlatformAxis { int forwardAxis, rightAxis, upAxis, forwardDir, rightDir, upDir; };
	__constant__ PlatformAxis c_dirAxis[6] = {
		{ 1, 0, 2,  1,  1,  1 }, { 2, 0, 1,  1,  1, -1 }, { 1, 0, 2, -1,  1, -1 },
		{ 2, 1, 0, -1,  1, -1 }, { 2, 1, 0, -1, -1,  1 }, { 2, 0, 1, -1,  1,  1 },
	};

	// GetChunkIndicesByBlockIndices (ChunkContainer.inl:19-42) + slot lookup; returns -1 when the chunk is absent
	__device__ int platform_locate(const PlatformArgs& a, const int coord[3], int local[3])
	{
		int chunk[3], adj[3];
		for (int k = 0; k < 3; ++k)
		{
			adj[k] = coord[k] + CS / 2;
			int off = adj[k] < 0 ? adj[k] - CS + 1 : adj[k];
			chunk[k] = off / CS;
		}
		int l[3];
		for (int k = 0; k < 3; ++k)
		{
			l[k] = adj[k] - chunk[k] * CS;
			if (l[k] < 0)
				l[k] += CS;
		}
		local[0] = l[0];
		local[1] = l[2];
		local[2] = l[1]; // (x, z, y)
		for (int k = 0; k < 3; ++k)
		{
			int g = chunk[k] - a.gridOrigin[k];
			if (g < 0 || g >= a.gridDim[k])
				return -1;
		}
		return a.slotGrid[((chunk[2] - a.gridOrigin[2]) * a.gridDim[1] + (chunk[1] - a.gridOrigin[1])) * a.gridDim[0] + (chunk[0] - a.gridOrigin[0])];
	}

	// Chunk::UpdateBlock (Chunk.cpp:348-366) + the neighbour-mask rule of Planet.cpp:39-58 from many threads at once (distinct blocks)
	__device__ __forceinline__ void platform_update_block(const PlatformArgs& a, uint32_t slot, const int local[3], uint16_t id)
	{
		const int x = local[0], y = local[1], z = local[2];
		a.blocks[size_t(slot) * NB + (z * CS + y) * CS + x] = id;
		if (x == 0)
			a.xslabs[size_t(slot) * 2 * SLAB + z * CS + y] = id;
		else if (x == CS - 1)
			a.xslabs[size_t(slot) * 2 * SLAB + SLAB + z * CS + y] = id;
		a.summary[slot] = SUM_UNKNOWN;
		atomicOr(reinterpret_cast<unsigned int*>(a.invalidMask) + (slot >> 2), edit_mask(x, y, z, false) << (8u * (slot & 3u)));
	}

	__global__ void __launch_bounds__(256) platform_kernel(const PlatformArgs a)
	{
		constexpr int platformSize = 15;
		constexpr int freeHeight = 10;
		const PlatformAxis ax = c_dirAxis[a.upDirection];
		const int tid = threadIdx.x;
		// thread = cell (x, z): 16 lanes per row of the platform, lane 15 of each row idle
		const int x = tid & 15, z = tid >> 4;
		const bool cell = x < platformSize && z < platformSize;
		const bool border = x == 0 || x == platformSize - 1 || z == 0 || z == platformSize - 1;
		const bool corner = (x == 0 || x == platformSize - 1) && (z == 0 || z == platformSize - 1);

		int orig[3] = { a.center[0], a.center[1], a.center[2] };
		orig[ax.rightAxis] += -ax.rightDir * platformSize / 2;
		orig[ax.forwardAxis] += -ax.forwardDir * platformSize / 2;

		// ---- the deck and the free space above it (Planet.cpp:427-460): layer 0 copper border / stone bricks, layers 1..9 Empty; the
		// loop leaves after the layer whose up coordinate is 0 when up points to negative coordinates (:454-455)
		int layers = freeHeight;
		if (ax.upDir < 0)
			for (int y = 0; y < freeHeight; ++y)
				if (orig[ax.upAxis] + y * ax.upDir == 0)
				{
					layers = y + 1;
					break;
				}
		if (cell)
		{
			for (int y = 0; y < layers; ++y)
			{
				int coord[3] = { orig[0], orig[1], orig[2] };
				coord[ax.rightAxis] += x * ax.rightDir;
				coord[ax.forwardAxis] += z * ax.forwardDir;
				coord[ax.upAxis] += y * ax.upDir;
				const uint16_t id = y != 0 ? uint16_t(0) : (border ? a.copper : a.stoneBricks);
				int local[3];
				const int slot = platform_locate(a, coord, local);
				if (slot >= 0)
					platform_update_block(a, uint32_t(slot), local, id);
			}
		}

		// ---- the pillars below (Planet.cpp:462-511): layer y = 1, 2, ... until a layer places nothing.  Inside a row the reference
		// advances the right coordinate only behind a cell whose chunk exists (`continue` before the increment, :478-482): cell x sits at
		// orig + x * rightDir while every earlier cell of its row found its chunk, and the row is stuck on the first absent one.
		for (unsigned y = 1;; ++y)
		{
			int coord[3] = { orig[0], orig[1], orig[2] };
			coord[ax.rightAxis] += x * ax.rightDir;
			coord[ax.forwardAxis] += z * ax.forwardDir;
			coord[ax.upAxis] -= int(y) * ax.upDir;
			int local[3] = { 0, 0, 0 };
			const int slot = cell ? platform_locate(a, coord, local) : -1;
			const uint32_t have = __ballot_sync(0xffffffffu, slot >= 0 || x >= platformSize);
			const uint32_t rowHave = (have >> (16 * ((tid >> 4) & 1))) & 0x7FFFu;
			const int firstAbsent = (~rowHave & 0x7FFFu) ? __ffs(~rowHave & 0x7FFFu) - 1 : platformSize;
			bool place = false;
			if (cell && x < firstAbsent)
			{
				if (y % 3 == 0)
					place = border;
				else
					place = corner && a.blocks[size_t(slot) * NB + (local[2] * CS + local[1]) * CS + local[0]] == 0;
			}
			if (place)
				platform_update_block(a, uint32_t(slot), local, a.planks);
			if (!__syncthreads_or(place))
				break;
		}
	}

	void launch_platform(const PlatformArgs& a, cudaStream_t st) { platform_kernel<<<1, 256, 0, st>>>(a); }

	// ------------------------------------------------------------------------------------------------ FlatChunk::BuildCollider
	// Greedy box merge (FlatChunk.cpp:56-153) for a batch of chunks, one WARP per chunk.  The merge is order dependent (the next box
	// starts at the first cell still set in z, y, x order), so the boxes are found one after the other exactly like the reference
	// does; what the 32 lanes share is the work inside one step: the 32x32 rows of the collision mask are built with ballots
	// (Chunk::OnChunkReset / GetCollisionCellMask, Chunk.cpp:368-384), the Y extension tests 32 rows with one ballot, the Z extension
	// one layer per ballot, and the clear is one row per lane.  A box is one 30-bit word (x, y, z, sizes - 1) in the job's scratch.
	constexpr int BOXC_WARPS = 4;

	__global__ void __launch_bounds__(BOXC_WARPS * 32) box_colliders_kernel(const uint16_t* __restrict__ blocks, const uint32_t* __restrict__ jobSlot, uint32_t n,
	                                                                       const uint8_t* __restrict__ blkFlags, uint32_t nTypes, uint32_t* __restrict__ scratch, uint32_t* __restrict__ counts)
	{
		__shared__ uint32_t maskAll[BOXC_WARPS][ROWS_PER_CHUNK];
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		const uint32_t job = blockIdx.x * BOXC_WARPS + warp;
		if (job >= n)
			return;
		const uint32_t entry = jobSlot[job];
		if (!(entry & JOB_HAS_CONTENT))
		{
			if (lane == 0)
				counts[job] = 0u;
			return;
		}
		uint32_t* mask = maskAll[warp];
		const uint16_t* ids = blocks + size_t(entry & JOB_SLOT_MASK) * NB;
		constexpr uint32_t FULL = 0xffffffffu;
		// collision bits: lane = 8 consecutive ids of a row (16-byte load), 4 lanes per row, 8 rows per step
		for (int r0 = 0; r0 < ROWS_PER_CHUNK; r0 += 8)
		{
			const int row = r0 + (lane >> 2), seg = lane & 3;
			const uint4 w = *reinterpret_cast<const uint4*>(ids + size_t(row) * CS + seg * 8);
			const uint32_t ws[4] = { w.x, w.y, w.z, w.w };
			uint32_t bits = 0;
#pragma unroll
			for (int j = 0; j < 8; ++j)
			{
				const unsigned id = (ws[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
				if (__ldg(blkFlags + (id < nTypes ? id : nTypes - 1)) & BF_COLLIDES)
					bits |= 1u << j;
			}
			bits <<= seg * 8;
			bits |= __shfl_xor_sync(FULL, bits, 1);
			bits |= __shfl_xor_sync(FULL, bits, 2);
			if (seg == 0)
				mask[row] = bits;
		}
		__syncwarp();

		uint32_t* out = scratch + size_t(job) * BOX_SCRATCH_STRIDE;
		uint32_t nBoxes = 0;
		for (int z = 0; z < CS; ++z)
		{
			// rows of this layer that still hold a cell
			uint32_t live = __ballot_sync(FULL, mask[z * CS + lane] != 0u);
			while (live)
			{
				const int y = __ffs(live) - 1;
				uint32_t row = mask[z * CS + y];
				while (row)
				{
					const int sx = __ffs(row) - 1;
					const uint32_t run = row >> sx;
					const int len = (~run == 0u) ? CS - sx : __ffs(~run) - 1;
					const uint32_t rowBits = (len == 32 ? 0xffffffffu : ((1u << len) - 1u)) << sx;
					// grow on Y: rows y + 1 + lane of layer z
					const int yy = y + 1 + lane;
					const uint32_t okY = __ballot_sync(FULL, yy < CS && (mask[z * CS + (yy < CS ? yy : 0)] & rowBits) == rowBits);
					const int endY = y + ((~okY == 0u) ? 32 : __ffs(~okY) - 1);
					// grow on Z: layer by layer, lane = row y + lane of the box
					const int ny = endY - y + 1;
					int endZ = z;
					for (int zz = z + 1; zz < CS; ++zz)
					{
						const bool ok = lane >= ny || (mask[zz * CS + y + (lane < ny ? lane : 0)] & rowBits) == rowBits;
						if (!__all_sync(FULL, ok))
							break;
						endZ = zz;
					}
					for (int zz = z; zz <= endZ; ++zz)
						if (lane < ny)
							mask[zz * CS + y + lane] &= ~rowBits;
					__syncwarp();
					if (lane == 0 && nBoxes < BOX_SCRATCH_STRIDE)
						out[nBoxes] = uint32_t(sx) | (uint32_t(y) << 5) | (uint32_t(z) << 10) | (uint32_t(len - 1) << 15) | (uint32_t(endY - y) << 20) | (uint32_t(endZ - z) << 25);
					++nBoxes;
					row = mask[z * CS + y];
				}
				// the rows of this layer may have been cleared by the boxes just made
				live = __ballot_sync(FULL, mask[z * CS + lane] != 0u) & ~((2u << y) - 1u);
			}
		}
		if (lane == 0)
			counts[job] = nBoxes;
	}

	// decode the boxes of job j to their final place: arena[first[j] + k] = {pos.xyz, size.xyz} in the reference's axes (x, z, y):
	// pos = start - dims / 2, size = end + 1 - start (FlatChunk.cpp:121-127)
	__global__ void __launch_bounds__(256) box_pack_kernel(const uint32_t* __restrict__ scratch, const uint32_t* __restrict__ counts, const unsigned long long* __restrict__ first, uint32_t n,
	                                                       float* __restrict__ arena)
	{
		const uint32_t job = blockIdx.x;
		if (job >= n)
			return;
		const uint32_t cnt = min(counts[job], BOX_SCRATCH_STRIDE);
		const uint32_t* src = scratch + size_t(job) * BOX_SCRATCH_STRIDE;
		float* dst = arena + first[job] * 6ull;
		const float half = float(CS) * 0.5f;
		for (uint32_t k = threadIdx.x; k < cnt; k += blockDim.x)
		{
			const uint32_t b = src[k];
			const int sx = int(b & 31u), sy = int((b >> 5) & 31u), sz = int((b >> 10) & 31u);
			const int lx = int((b >> 15) & 31u) + 1, ly = int((b >> 20) & 31u) + 1, lz = int((b >> 25) & 31u) + 1;
			float* o = dst + size_t(k) * 6;
			o[0] = float(sx) - half;
			o[1] = float(sz) - half;
			o[2] = float(sy) - half;
			o[3] = float(sx + lx) - float(sx);
			o[4] = float(sz + lz) - float(sz);
			o[5] = float(sy + ly) - float(sy);
		}
	}

	void launch_box_colliders(const uint16_t* blocks, const uint32_t* jobSlot, uint32_t n, const uint8_t* blkFlags, uint32_t nTypes, uint32_t* scratch, uint32_t* counts, cudaStream_t st)
	{
		if (n)
			box_colliders_kernel<<<(n + BOXC_WARPS - 1) / BOXC_WARPS, BOXC_WARPS * 32, 0, st>>>(blocks, jobSlot, n, blkFlags, nTypes, scratch, counts);
	}

	void launch_box_pack(const uint32_t* scratch, const uint32_t* counts, const unsigned long long* first, uint32_t n, float* arena, cudaStream_t st)
	{
		if (n)
			box_pack_kernel<<<n, 256, 0, st>>>(scratch, counts, first, n, arena);
	}

	// ------------------------------------------------------------------------------------------------ P2P halo exchange
	// Ordering between GPUs lives on the device: every container owns epoch flags in its own HBM (blocks-final, slabs-pulled); the peers
	// poll them over NVLink.  A flag is raised by a one-thread kernel behind the producing kernels on the same stream (the kernel boundary
	// makes their writes visible device-wide, the system-scope fence + release store order them for the peers).
	__global__ void flag_publish_kernel(uint32_t* flag, uint32_t value)
	{
		if (threadIdx.x == 0 && blockIdx.x == 0)
		{
			__threadfence_system();
			asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
		}
	}

	__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
	{
		uint32_t v;
		asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
		return v;
	}

	__device__ __forceinline__ unsigned long long global_timer_ns()
	{
		unsigned long long t;
		asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
		return t;
	}

	// one warp: lane i waits until *flags[i] >= value (epochs only grow).  Gives up after timeoutNs and raises *err, so that a
	// peer that died shows up as an error on the host instead of a hung GPU.
	__global__ void flag_wait_kernel(const uint32_t* const* flags, uint32_t n, uint32_t value, unsigned long long timeoutNs, uint32_t* err)
	{
		const unsigned long long t0 = global_timer_ns();
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
		{
			const uint32_t* f = flags[i];
			while (int32_t(ld_acquire_sys(f) - value) < 0)
			{
				if (global_timer_ns() - t0 > timeoutNs)
				{
					atomicExch(err, 1u + i);
					break;
				}
				__nanosleep(200);
			}
		}
		__threadfence_system();
	}

	void launch_flag_publish(uint32_t* flag, uint32_t value, cudaStream_t st) { flag_publish_kernel<<<1, 32, 0, st>>>(flag, value); }

	void launch_flag_wait(const uint32_t* const* flags, uint32_t n, uint32_t value, unsigned long long timeoutNs, uint32_t* err, cudaStream_t st)
	{
		if (n)
			flag_wait_kernel<<<1, 32, 0, st>>>(flags, n, value, timeoutNs, err);
	}

	// One CTA per ghost slab: 2 KiB read straight from the owner GPU's HBM over NVLink (peer-mapped pointer, cache-volatile loads: the
	// owner rewrites these addresses every round), written locally.
	__global__ void __launch_bounds__(128) halo_pull_kernel(const GhostDesc* __restrict__ descs, uint32_t n, uint16_t* __restrict__ ghostSlabs)
	{
		const uint32_t i = blockIdx.x;
		if (i >= n)
			return;
		const GhostDesc d = descs[i];
		const int t = threadIdx.x; // 128 x 16 B = 2 KiB
		const uint4* src;
		if (d.contig)
			src = reinterpret_cast<const uint4*>(d.src) + t;
		else
			src = reinterpret_cast<const uint4*>(d.src + size_t(t >> 2) * SLAB) + (t & 3);
		reinterpret_cast<uint4*>(ghostSlabs + size_t(d.ghost) * SLAB)[t] = __ldcv(src);
	}

	void launch_halo_pull(const GhostDesc* descs, uint32_t n, uint16_t* ghostSlabs, cudaStream_t st)
	{
		if (n)
			halo_pull_kernel<<<n, 128, 0, st>>>(descs, n, ghostSlabs);
	}
}
