// lz4_kernels.cu — the LZ4 block format of `Packets::ChunkReset` on sm_100a.
//
// The reference ships a chunk to a client as tickIndex, entityId, chunkId, block count, compressed size and ONE LZ4 block holding the
// 65,536-byte block array (src/CommonLib/Protocol/Packets.cpp:172-212).  The block comes from BinaryCompressor::Compress =
// LZ4_compress_fast_extState(fresh state, acceleration 0 -> 1) and is read back with LZ4_decompress_safe
// (src/CommonLib/Utility/BinaryCompressor.cpp:17-47).  With the planet resident in HBM the block array never has to visit the host
// uncompressed: lz4_compress_kernel produces, per chunk, exactly the bytes the library produces (the greedy parse of lz4 1.9 / 1.10 with
// the 8192-entry 16-bit position table every input below 64 KiB + 11 bytes gets), lz4_decompress_kernel decodes a received block.
//
// Shape: the parse is one sequential chain per chunk (every table probe depends on the previous match), so parallelism comes from the
// chunks: one warp (= one CTA) per chunk, two CTAs per SM.  The chunk (64 KiB, one TMA bulk copy) and the position table (16 KiB) live in
// shared memory, so a step of the chain costs shared-memory latency instead of L2 latency.  All lanes run the scalar control flow in
// lock step on identical values (no divergence, table updates are same-value stores); the data-parallel parts — match length, literal
// copies, match copies of the decoder — use the 32 lanes.
#include "tsom_device.cuh"
#include "tsom_kernels.h"

namespace tsomk
{
	namespace
	{
		constexpr int LZ4_MINMATCH = 4;
		constexpr int LZ4_MFLIMIT = 12;     // no match starts within the last 12 bytes
		constexpr int LZ4_LASTLITERALS = 5; // the last 5 bytes are literals
		constexpr int LZ4_HASHLOG = 13;     // byU16 table: 8192 entries
		constexpr int LZ4_SKIP_TRIGGER = 6;
		constexpr uint32_t CHUNK_BYTES = NB * 2;
		constexpr uint32_t FULL = 0xffffffffu;

		struct alignas(128) Lz4Smem
		{
			uint8_t in[CHUNK_BYTES + 256]; // TMA destination; the pad keeps the 128-byte match-length probes inside the buffer
			uint16_t table[1 << LZ4_HASHLOG];
			alignas(8) uint64_t bar;
		};

		struct alignas(128) Lz4DecSmem
		{
			uint8_t out[CHUNK_BYTES];
		};

		// unaligned little-endian 32-bit read from shared memory: two aligned words + funnel shift
		__device__ __forceinline__ uint32_t lds32u(const uint8_t* base, uint32_t off)
		{
			const uint32_t* w = reinterpret_cast<const uint32_t*>(base + (off & ~3u));
			return __funnelshift_r(w[0], w[1], (off & 3u) * 8u);
		}

		__device__ __forceinline__ uint32_t lz4_hash(uint32_t sequence) { return (sequence * 2654435761u) >> (LZ4_MINMATCH * 8 - LZ4_HASHLOG); }

		// number of equal bytes of in[a..] and in[m..] (m < a), counting no byte at or past `limit`; 128 bytes per step over the warp
		__device__ __forceinline__ uint32_t common_length(const uint8_t* in, uint32_t a, uint32_t m, uint32_t limit, int lane)
		{
			uint32_t done = 0;
			for (;;)
			{
				const uint32_t pos = a + done + 4u * lane;
				const uint32_t diff = lds32u(in, pos) ^ lds32u(in, m + done + 4u * lane);
				const uint32_t ballot = __ballot_sync(FULL, diff != 0u);
				if (ballot)
				{
					const int first = __ffs(ballot) - 1;
					const uint32_t d = __shfl_sync(FULL, diff, first);
					done += 4u * first + ((__ffs(d) - 1) >> 3);
					break;
				}
				done += 128u;
				if (a + done >= limit)
					break;
			}
			return min(done, limit - a);
		}

		// The decoder's view of the compressed stream: tokens, offsets and length bytes are read one at a time and each decides where the
		// next one is, so a global load per byte would put L2 latency on every step.  The warp keeps 256 bytes of the stream in
		// registers (two words per lane, the second one prefetched) and fetches single bytes with a shuffle.
		struct StreamWindow
		{
			const uint32_t* words; // the stream's first word, rounded down to 4-byte alignment
			uint32_t shift;        // bytes between that word and the stream's first byte
			uint32_t base;         // window start (multiple of 128) in bytes from `words`
			uint32_t cur, next;    // this lane's word of [base, base + 128) and of [base + 128, base + 256)

			__device__ __forceinline__ void init(const uint8_t* src, int lane)
			{
				shift = uint32_t(reinterpret_cast<uintptr_t>(src) & 3u);
				words = reinterpret_cast<const uint32_t*>(src - shift);
				base = 0u;
				cur = words[lane];
				next = words[32 + lane];
			}
			__device__ __forceinline__ uint32_t byte(uint32_t ip, int lane)
			{
				const uint32_t a = ip + shift;
				if (a - base >= 128u)
				{
					if (a - base < 256u)
					{
						base += 128u;
						cur = next;
					}
					else
					{
						base = a & ~127u;
						cur = words[(base >> 2) + lane];
					}
					next = words[(base >> 2) + 32 + lane];
				}
				return (__shfl_sync(FULL, cur, (a - base) >> 2) >> ((a & 3u) * 8u)) & 255u;
			}
		};

		// length bytes of a sequence: 255, 255, ..., rest (all lanes return the advanced offset; lane 0 writes)
		__device__ __forceinline__ uint32_t put_length(uint8_t* dst, uint32_t op, uint32_t len, int lane)
		{
			const uint32_t n255 = len / 255u;
			for (uint32_t i = lane; i < n255; i += 32)
				dst[op + i] = 255;
			if (lane == 0)
				dst[op + n255] = uint8_t(len - n255 * 255u);
			return op + n255 + 1u;
		}
	}

	// One warp per chunk.  streams: n x stride bytes (stride >= LZ4_compressBound(65536) = 65,809); sizes[i] = compressed size.
	__global__ void __launch_bounds__(32) lz4_compress_kernel(const uint16_t* __restrict__ blocks, const uint32_t* __restrict__ slots, uint32_t n, uint8_t* __restrict__ streams,
	                                                          uint32_t stride, uint32_t* __restrict__ sizes)
	{
		extern __shared__ __align__(128) unsigned char smemRaw[];
		Lz4Smem& sm = *reinterpret_cast<Lz4Smem*>(smemRaw);
		const int lane = threadIdx.x;
		const uint32_t job = blockIdx.x;
		if (job >= n)
			return;
		if (lane == 0)
		{
			mbar_init(&sm.bar, 1);
			fence_mbar_init();
		}
		__syncwarp();
		if (lane == 0)
		{
			const uint8_t* src = reinterpret_cast<const uint8_t*>(blocks + size_t(slots[job]) * NB);
			mbar_arrive_expect_tx(&sm.bar, CHUNK_BYTES);
#pragma unroll
			for (int p = 0; p < 4; ++p)
				tma_load_1d(sm.in + p * (CHUNK_BYTES / 4), src + p * (CHUNK_BYTES / 4), CHUNK_BYTES / 4, &sm.bar);
		}
		// a fresh state (LZ4_compress_fast_extState re-initialises it on every call): every entry points at offset 0
		{
			uint4* t = reinterpret_cast<uint4*>(sm.table);
			for (int i = lane; i < int(sizeof(sm.table) / 16); i += 32)
				t[i] = make_uint4(0u, 0u, 0u, 0u);
			uint4* pad = reinterpret_cast<uint4*>(sm.in + CHUNK_BYTES);
			if (lane < 16)
				pad[lane] = make_uint4(0u, 0u, 0u, 0u);
		}
		__syncwarp();
		mbar_wait(&sm.bar, 0);

		const uint8_t* in = sm.in;
		uint8_t* dst = streams + size_t(job) * stride;
		constexpr uint32_t iend = CHUNK_BYTES;
		constexpr uint32_t mflimitPlusOne = iend - LZ4_MFLIMIT + 1;
		constexpr uint32_t matchlimit = iend - LZ4_LASTLITERALS;

		uint32_t ip = 0, anchor = 0, op = 0;
		sm.table[lz4_hash(lds32u(in, 0))] = 0; // the first position
		ip = 1;
		uint32_t forwardH = lz4_hash(lds32u(in, ip));
		for (;;)
		{
			// ---- find a match: probe the table at a stride that grows by one every 64 misses
			uint32_t match;
			{
				uint32_t forwardIp = ip;
				uint32_t step = 1, searchMatchNb = 1u << LZ4_SKIP_TRIGGER;
				bool found = false;
				for (;;)
				{
					const uint32_t h = forwardH;
					const uint32_t current = forwardIp;
					match = sm.table[h];
					ip = forwardIp;
					forwardIp += step;
					step = searchMatchNb++ >> LZ4_SKIP_TRIGGER;
					if (forwardIp > mflimitPlusOne)
						break;
					forwardH = lz4_hash(lds32u(in, forwardIp));
					__syncwarp(); // every lane has read the entry before any lane replaces it
					sm.table[h] = uint16_t(current);
					if (lds32u(in, match) == lds32u(in, ip))
					{
						found = true;
						break;
					}
				}
				if (!found)
					break; // -> last literals
			}
			// ---- catch up: extend the match backwards over the pending literals
			while (ip > anchor && match > 0u && in[ip - 1] == in[match - 1])
			{
				--ip;
				--match;
			}
			// ---- token + literals
			uint32_t token;
			{
				const uint32_t litLength = ip - anchor;
				token = op++;
				uint32_t tokenValue;
				if (litLength >= 15u)
				{
					tokenValue = 15u << 4;
					op = put_length(dst, op, litLength - 15u, lane);
				}
				else
					tokenValue = litLength << 4;
				for (uint32_t i = lane; i < litLength; i += 32)
					dst[op + i] = in[anchor + i];
				__syncwarp();
				op += litLength;
				// ---- sequences: a second, third ... one without literals while the position right after a match matches again
				for (;;)
				{
					const uint32_t offset = ip - match;
					if (lane == 0)
					{
						dst[op] = uint8_t(offset);
						dst[op + 1] = uint8_t(offset >> 8);
					}
					op += 2;
					uint32_t matchCode = common_length(in, ip + LZ4_MINMATCH, match + LZ4_MINMATCH, matchlimit, lane);
					ip += matchCode + LZ4_MINMATCH;
					if (matchCode >= 15u)
					{
						tokenValue += 15u;
						op = put_length(dst, op, matchCode - 15u, lane);
					}
					else
						tokenValue += matchCode;
					if (lane == 0)
						dst[token] = uint8_t(tokenValue);
					anchor = ip;
					if (ip >= mflimitPlusOne)
						break;
					// fill the table at ip - 2, then test the new position at once
					sm.table[lz4_hash(lds32u(in, ip - 2))] = uint16_t(ip - 2);
					const uint32_t h = lz4_hash(lds32u(in, ip));
					match = sm.table[h];
					__syncwarp();
					sm.table[h] = uint16_t(ip);
					if (lds32u(in, match) != lds32u(in, ip))
						break;
					token = op++;
					tokenValue = 0u;
				}
			}
			if (ip >= mflimitPlusOne)
				break;
			forwardH = lz4_hash(lds32u(in, ++ip));
		}
		// ---- last literals
		{
			const uint32_t lastRun = iend - anchor;
			if (lastRun >= 15u)
			{
				if (lane == 0)
					dst[op] = 15u << 4;
				op = put_length(dst, op + 1, lastRun - 15u, lane);
			}
			else
			{
				if (lane == 0)
					dst[op] = uint8_t(lastRun << 4);
				++op;
			}
			for (uint32_t i = lane; i < lastRun; i += 32)
				dst[op + i] = in[anchor + i];
			op += lastRun;
		}
		if (lane == 0)
			sizes[job] = op;
	}

	// streams scattered at `stride` -> one contiguous buffer at offsets[] (exclusive prefix sums of sizes)
	__global__ void __launch_bounds__(256) lz4_pack_kernel(const uint8_t* __restrict__ streams, uint32_t stride, const unsigned long long* __restrict__ offsets, uint32_t n,
	                                                       uint8_t* __restrict__ packed)
	{
		const uint32_t job = blockIdx.x;
		if (job >= n)
			return;
		const unsigned long long o = offsets[job];
		const uint32_t size = uint32_t(offsets[job + 1] - o);
		const uint8_t* src = streams + size_t(job) * stride;
		for (uint32_t i = threadIdx.x; i < size; i += blockDim.x)
			packed[o + i] = src[i];
	}

	// One warp per chunk: LZ4_decompress_safe of stream [offsets[i], offsets[i+1]) into shared memory (`packed` must be readable for 512
	// bytes past the last stream: the register window reads ahead); the 64 KiB land in out[i] only if the
	// stream is well formed and decodes to exactly 65,536 bytes.  status[i] = decoded size, or 0xFFFFFFFF for a malformed stream.
	__global__ void __launch_bounds__(32) lz4_decompress_kernel(const uint8_t* __restrict__ packed, const unsigned long long* __restrict__ offsets, uint32_t n, uint16_t* __restrict__ out,
	                                                            uint32_t* __restrict__ status)
	{
		extern __shared__ __align__(128) unsigned char smemRaw[];
		Lz4DecSmem& sm = *reinterpret_cast<Lz4DecSmem*>(smemRaw);
		const int lane = threadIdx.x;
		const uint32_t job = blockIdx.x;
		if (job >= n)
			return;
		const uint8_t* src = packed + offsets[job];
		const unsigned long long srcSize64 = offsets[job + 1] - offsets[job];
		constexpr uint32_t BAD = 0xFFFFFFFFu;
		uint32_t result = BAD;
		if (srcSize64 > 0ull && srcSize64 < (1ull << 31))
		{
			const uint32_t iend = uint32_t(srcSize64);
			constexpr uint32_t oend = CHUNK_BYTES;
			uint32_t ip = 0, op = 0;
			StreamWindow win;
			win.init(src, lane);
			for (;;)
			{
				if (ip >= iend)
					break;
				const uint32_t token = win.byte(ip++, lane);
				// literal length
				uint32_t length = token >> 4;
				bool bad = false;
				if (length == 15u)
				{
					uint32_t s;
					do
					{
						if (ip >= iend)
						{
							bad = true;
							break;
						}
						s = win.byte(ip++, lane);
						length += s;
					} while (s == 255u);
				}
				if (bad || length > iend - ip || length > oend - op)
					break;
				// literals that reach into the last 12 output bytes or the last 8 input bytes can only belong to the last sequence
				const bool mustBeLast = length + LZ4_MFLIMIT > oend - op || length + 8u > iend - ip;
				if (mustBeLast && length != iend - ip)
					break;
				for (uint32_t i = lane; i < length; i += 32)
					sm.out[op + i] = src[ip + i];
				ip += length;
				op += length;
				if (mustBeLast)
				{
					result = op; // the last sequence holds literals only
					break;
				}
				const uint32_t offset = win.byte(ip, lane) | (win.byte(ip + 1, lane) << 8);
				ip += 2;
				if (offset == 0u || offset > op)
					break;
				length = token & 15u;
				if (length == 15u)
				{
					uint32_t s;
					do
					{
						if (ip >= iend)
						{
							bad = true;
							break;
						}
						s = win.byte(ip++, lane);
						length += s;
					} while (s == 255u);
				}
				if (bad)
					break;
				length += LZ4_MINMATCH;
				if (length + LZ4_LASTLITERALS > oend - op)
					break; // the last 5 bytes of the output must be literals
				// out[op + i] = out[op - offset + i] byte by byte == out[op - offset + i mod offset]: every source byte precedes this match
				__syncwarp();
				const uint32_t from = op - offset;
				for (uint32_t i = lane; i < length; i += 32)
					sm.out[op + i] = sm.out[from + (offset >= length ? i : i % offset)];
				__syncwarp();
				op += length;
			}
		}
		__syncwarp();
		if (result == CHUNK_BYTES)
		{
			uint4* dst = reinterpret_cast<uint4*>(out + size_t(job) * NB);
			const uint4* s4 = reinterpret_cast<const uint4*>(sm.out);
			for (int i = lane; i < int(CHUNK_BYTES / 16); i += 32)
				dst[i] = s4[i];
		}
		if (lane == 0)
			status[job] = result;
	}

	uint32_t lz4_stream_stride() { return 65824u; } // LZ4_compressBound(65536) = 65536 + 65536 / 255 + 16 = 65,809, rounded up to 16

	cudaError_t lz4_kernels_init()
	{
		cudaError_t e = cudaFuncSetAttribute(lz4_compress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(Lz4Smem)));
		if (e != cudaSuccess)
			return e;
		return cudaFuncSetAttribute(lz4_decompress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(Lz4DecSmem)));
	}

	void launch_lz4_compress(const uint16_t* blocks, const uint32_t* slots, uint32_t n, uint8_t* streams, uint32_t* sizes, cudaStream_t st)
	{
		lz4_compress_kernel<<<n, 32, sizeof(Lz4Smem), st>>>(blocks, slots, n, streams, lz4_stream_stride(), sizes);
	}

	void launch_lz4_pack(const uint8_t* streams, const unsigned long long* offsets, uint32_t n, uint8_t* packed, cudaStream_t st)
	{
		lz4_pack_kernel<<<n, 256, 0, st>>>(streams, lz4_stream_stride(), offsets, n, packed);
	}

	void launch_lz4_decompress(const uint8_t* packed, const unsigned long long* offsets, uint32_t n, uint16_t* out, uint32_t* status, cudaStream_t st)
	{
		lz4_decompress_kernel<<<n, 32, sizeof(Lz4DecSmem), st>>>(packed, offsets, n, out, status);
	}
}
