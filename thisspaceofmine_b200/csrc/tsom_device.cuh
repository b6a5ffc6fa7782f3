// tsom_device.cuh — device-side helpers shared by the sm_100a kernels (TMA bulk copy / mbarrier PTX, small math).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace tsomk
{
	constexpr int CS = 32;          // Planet::ChunkSize (reference include/CommonLib/ChunkContainer.hpp:42)
	constexpr int NB = CS * CS * CS; // blocks per chunk
	constexpr int SLAB = CS * CS;    // ids per boundary slab

	// Direction order (reference include/CommonLib/Direction.hpp:18-28)
	enum : int { D_BACK = 0, D_DOWN = 1, D_FRONT = 2, D_LEFT = 3, D_RIGHT = 4, D_UP = 5 };

	// block flag bits in the device block table
	enum : unsigned { BF_COLLIDES = 1u, BF_DOUBLE_SIDED = 2u, BF_TRANSPARENT = 4u };

	// halo pointer table encoding: device address | kind in the low bits (slabs are >= 16-byte aligned)
	enum : unsigned long long { HALO_CONTIG = 0ull, HALO_ROWS = 1ull, HALO_KIND_MASK = 15ull };

	// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) ----
	__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

	__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
	{
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
	}

	__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

	// order prior generic-proxy accesses to shared memory before subsequent async-proxy (TMA) writes
	__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

	__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
	{
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
	}

	__device__ __forceinline__ void tma_load_1d(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t* bar)
	{
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		             ::"r"(smem_u32(smemDst)), "l"(gmemSrc), "r"(bytes), "r"(smem_u32(bar))
		             : "memory");
	}

	__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
	{
		asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
	}

	// named barrier over a subset of the CTA's warps (id 1..15; `threads` must be a multiple of 32)
	__device__ __forceinline__ void named_bar_sync(int id, int threads)
	{
		asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
	}

	__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
	{
		uint32_t ok;
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(ok)
			: "r"(smem_u32(bar)), "r"(parity)
			: "memory");
		return ok != 0;
	}

	__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
	{
		while (!mbar_try_wait(bar, parity)) {}
	}

	// ---- float3 with the exact operation order of the reference's Nazara math (compiled with -fmad=false) ----
	struct F3
	{
		float x, y, z;
	};

	__device__ __forceinline__ F3 f3(float x, float y, float z) { return F3{ x, y, z }; }
	__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return F3{ a.x + b.x, a.y + b.y, a.z + b.z }; }
	__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return F3{ a.x - b.x, a.y - b.y, a.z - b.z }; }
	__device__ __forceinline__ F3 operator*(F3 a, float s) { return F3{ a.x * s, a.y * s, a.z * s }; }
	__device__ __forceinline__ F3 operator/(F3 a, float s) { return F3{ a.x / s, a.y / s, a.z / s }; }
	__device__ __forceinline__ F3 cross(F3 a, F3 b) { return F3{ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }

	// Nz::Vector3f::Normalize: multiply by 1/length when length > 0
	__device__ __forceinline__ F3 normalize(F3 v)
	{
		float n = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
		if (n > 0.f)
		{
			float inv = 1.f / n;
			v.x *= inv;
			v.y *= inv;
			v.z *= inv;
		}
		return v;
	}

	// DirectionFromNormal (reference include/CommonLib/Direction.inl:7-21): the dot with an axis normal is +-component;
	// ">=" keeps the last maximal direction; start at -1.
	__device__ __forceinline__ int direction_from_normal(F3 n)
	{
		int best = 0;
		float bestDot = -1.f;
		float dots[6] = { n.z, -n.y, -n.z, -n.x, n.x, n.y };
#pragma unroll
		for (int d = 0; d < 6; ++d)
		{
			if (dots[d] >= bestDot)
			{
				best = d;
				bestDot = dots[d];
			}
		}
		return best;
	}

	// Nz::Quaternionf * Vector3f
	__device__ __forceinline__ F3 quat_rotate(float qw, F3 qv, F3 v)
	{
		F3 uv = cross(qv, v);
		F3 uuv = cross(qv, uv);
		uv = uv * (2.f * qw);
		uuv = uuv * 2.f;
		return v + uv + uuv;
	}
}
