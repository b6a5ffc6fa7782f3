// tsom_host.cu — host side of libtsom_b200.so: context / container bookkeeping and the C ABI of include/tsom_b200.h.
// Everything that computes runs in the sm_100a kernels of mesh_kernels.cu / gen_kernels.cu; there is no CPU fallback.
#include "../../include/tsom_b200.h"
#include "tsom_kernels.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <random>
#include <string>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>

// halo pointer encoding (mirrors HALO_CONTIG / HALO_ROWS of tsom_device.cuh, which is device-only)
static inline unsigned long long tsomk_halo_contig() { return 0ull; }
static inline unsigned long long tsomk_halo_rows() { return 1ull; }

namespace
{
	thread_local std::string g_lastError;

	int Fail(int code, const std::string& msg)
	{
		g_lastError = msg;
		return code;
	}

#define TSOM_CUDA(expr)                                                                                          \
	do                                                                                                           \
	{                                                                                                            \
		cudaError_t _e = (expr);                                                                                 \
		if (_e != cudaSuccess)                                                                                   \
			return Fail(TSOM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                      \
	} while (0)

#define TSOM_LOCK(ctxPtr) std::lock_guard<std::recursive_mutex> tsomLock_((ctxPtr)->mutex)

	constexpr int CS = TSOM_CHUNK_SIZE;
	constexpr size_t NB = TSOM_CHUNK_BLOCKS;
	constexpr size_t SLAB = 1024;

	// reference include/CommonLib/Direction.hpp:83-90 (world axes)
	const int kChunkDirOffset[6][3] = { { 0, 0, 1 }, { 0, -1, 0 }, { 0, 0, -1 }, { -1, 0, 0 }, { 1, 0, 0 }, { 0, 1, 0 } };

	struct Key
	{
		int32_t x, y, z;
		bool operator==(const Key& o) const { return x == o.x && y == o.y && z == o.z; }
		bool operator<(const Key& o) const { return std::tie(x, y, z) < std::tie(o.x, o.y, o.z); }
	};
	struct KeyHash
	{
		size_t operator()(const Key& k) const
		{
			uint64_t h = uint64_t(uint32_t(k.x)) * 0x9E3779B97F4A7C15ull;
			h ^= uint64_t(uint32_t(k.y)) * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
			h ^= uint64_t(uint32_t(k.z)) * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
			return size_t(h);
		}
	};

	// grow-only device scratch buffer
	template<typename T> struct DevBuf
	{
		T* ptr = nullptr;
		size_t cap = 0;
		cudaError_t Ensure(size_t n)
		{
			if (n <= cap)
				return cudaSuccess;
			size_t want = std::max(n, cap * 2);
			T* p = nullptr;
			cudaError_t e = cudaMalloc(&p, want * sizeof(T));
			if (e != cudaSuccess)
				return e;
			if (ptr)
				cudaFree(ptr);
			ptr = p;
			cap = want;
			return cudaSuccess;
		}
		// same, keeping the first `keep` elements (the copy is ordered on `st`, which is synchronised before the old buffer goes)
		cudaError_t EnsureKeep(size_t n, size_t keep, cudaStream_t st)
		{
			if (n <= cap)
				return cudaSuccess;
			size_t want = std::max(n, cap * 2);
			T* p = nullptr;
			cudaError_t e = cudaMalloc(&p, want * sizeof(T));
			if (e != cudaSuccess)
				return e;
			if (ptr && keep)
			{
				e = cudaMemcpyAsync(p, ptr, std::min(keep, cap) * sizeof(T), cudaMemcpyDeviceToDevice, st);
				if (e == cudaSuccess)
					e = cudaStreamSynchronize(st);
				if (e != cudaSuccess)
				{
					cudaFree(p);
					return e;
				}
			}
			else if (ptr)
				cudaStreamSynchronize(st);
			if (ptr)
				cudaFree(ptr);
			ptr = p;
			cap = want;
			return cudaSuccess;
		}
		void Free()
		{
			if (ptr)
				cudaFree(ptr);
			ptr = nullptr;
			cap = 0;
		}
	};

	// ---- product-side restatement of the small third-party pieces the host needs ----

	// siv::PerlinNoise::reseed (v3.0.0): iota then the library's own shuffle driven by std::mt19937{seed}
	void PerlinPermutation(uint32_t seed, uint8_t out[256])
	{
		std::mt19937 urbg{ seed };
		for (int i = 0; i < 256; ++i)
			out[i] = uint8_t(i);
		for (uint64_t i = 1; i < 256; ++i)
			std::swap(out[i], out[uint64_t(urbg()) % (i + 1)]);
	}

	// A URBG with std::minstd_rand's range that replays two chosen outputs, so the threshold below is derived from the very
	// std::generate_canonical the reference's std::bernoulli_distribution uses (libstdc++).
	struct ReplayEngine
	{
		using result_type = uint_fast32_t;
		static constexpr result_type min() { return 1u; }
		static constexpr result_type max() { return 2147483646u; }
		result_type v[2];
		int i = 0;
		result_type operator()() { return v[i++ & 1]; }
	};

	bool BernoulliFromDraws(uint32_t u1, uint32_t u2)
	{
		ReplayEngine e{ { u1, u2 } };
		return std::generate_canonical<double, std::numeric_limits<double>::digits>(e) < 0.9;
	}

	// Integer form of the decision: true <=> (u2-1, u1-1) < (bStar, aStar) lexicographically.  Verified at start-up.
	bool DeriveBernoulliThreshold(uint32_t& bStar, uint32_t& aStar)
	{
		const uint32_t R = 2147483646u; // outputs u in [1, R], a = u - 1 in [0, R)
		// smallest b for which not every a is accepted
		uint32_t lo = 0, hi = R; // invariant: all b < lo fully accepted
		while (lo < hi)
		{
			uint32_t mid = lo + (hi - lo) / 2;
			if (BernoulliFromDraws(R, mid + 1)) // a = R-1 (largest) still accepted -> every a accepted
				lo = mid + 1;
			else
				hi = mid;
		}
		bStar = lo;
		if (bStar >= R)
			return false;
		// number of accepted a at b = bStar (accepted set is a prefix)
		uint32_t alo = 0, ahi = R;
		while (alo < ahi)
		{
			uint32_t mid = alo + (ahi - alo) / 2;
			if (BernoulliFromDraws(mid + 1, bStar + 1))
				alo = mid + 1;
			else
				ahi = mid;
		}
		aStar = alo;
		// the next b must reject everything, otherwise the two-level form would not be exact
		if (bStar + 1 < R && BernoulliFromDraws(1, bStar + 2))
			return false;
		return true;
	}

	struct F3h
	{
		float x, y, z;
	};

	// Nz::Quaternionf::RotationBetween(from, Up) for an axis-aligned `from` (third-party arithmetic #3)
	void RotationBetween(F3h from, F3h to, float q[4])
	{
		auto dot = [](F3h a, F3h b) { return a.x * b.x + a.y * b.y + a.z * b.z; };
		auto cross = [](F3h a, F3h b) { return F3h{ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; };
		float d = dot(from, to);
		if (d < -0.999999f)
		{
			F3h axis = dot(from, F3h{ 1.f, 0.f, 0.f }) < 0.999999f ? cross(F3h{ 1.f, 0.f, 0.f }, from) : cross(F3h{ 0.f, 1.f, 0.f }, from);
			float n = std::sqrt(axis.x * axis.x + axis.y * axis.y + axis.z * axis.z);
			if (n > 0.f)
			{
				float inv = 1.f / n;
				axis.x *= inv;
				axis.y *= inv;
				axis.z *= inv;
			}
			float half = 3.14159265358979323846f / 2.f;
			float s = std::sin(half), c = std::cos(half);
			q[0] = c;
			q[1] = axis.x * s;
			q[2] = axis.y * s;
			q[3] = axis.z * s;
		}
		else if (d > 0.999999f)
		{
			q[0] = 1.f;
			q[1] = q[2] = q[3] = 0.f;
		}
		else
		{
			float norm = std::sqrt(dot(from, from) * dot(to, to));
			F3h c = cross(from, to);
			float w = norm + d;
			float n = std::sqrt(w * w + c.x * c.x + c.y * c.y + c.z * c.z);
			float inv = n > 0.f ? 1.f / n : 1.f;
			q[0] = w * inv;
			q[1] = c.x * inv;
			q[2] = c.y * inv;
			q[3] = c.z * inv;
		}
	}
}

// =====================================================================================================================
struct tsom_ctx
{
	int device = 0;
	int numSMs = 148;
	cudaStream_t stream = nullptr;
	uint64_t launches = 0;
	// every entry point holds this while it touches the context, its containers or its stream; recursive so that tsom_ctx_lock lets a
	// caller keep several calls together (mesh + copy out) and entry points may call each other
	std::recursive_mutex mutex;

	// block table
	uint8_t* dFlags = nullptr;
	uint32_t* dTex = nullptr;
	bool texFits8 = false; // every texture slice < 256: the mesher keeps the table as bytes in shared memory
	uint32_t nTypes = 0;
	std::vector<uint8_t> hostFlags; // BF_* per type, host copy (the generator derives the chunk summaries it may write from it)
	uint32_t tableEpoch = 0;        // bumped by tsom_set_block_table: chunk summaries made under another table are void

	// arenas (faces)
	float4* dVertices = nullptr;
	uint2* dIndices = nullptr;
	float4* dCollider = nullptr;
	uint64_t capVertices = 0, capIndices = 0, capCollider = 0;
	uint64_t totalFaces = 0;
	uint32_t lastMeshFlags = 0; // TSOM_MESH_RENDER / COLLIDER of the call that filled the arenas

	// per-call scratch
	DevBuf<uint32_t> jobSlot, faceCount;
	DevBuf<uint32_t> jobOrder; // ticket order of the unordered mesher: tsomk::JOB_BUCKETS lists of n jobs (job_classify_kernel)
	DevBuf<unsigned long long> firstFace, status;
	DevBuf<float> gravity;
	DevBuf<tsomk::EditDev> edits;
	DevBuf<uint8_t> ioPayload;   // save-format payloads (65536 B per chunk)
	DevBuf<uint16_t> ioPalette;
	DevBuf<uint32_t> ioCounts;
	DevBuf<uint8_t> ioStreams;            // LZ4 blocks: per-chunk scratch at a fixed stride, or the packed streams
	DevBuf<unsigned long long> ioOffsets; // n + 1 offsets of the packed LZ4 blocks
	DevBuf<unsigned long long> editKeys; // open-addressing table of the edited blocks of one tsom_apply_edits call
	DevBuf<uint32_t> editVals;           // ... -> index of the last edit that names the block
	uint32_t* dSmall = nullptr; // [1] job ticket [3] overflow [4..5] totalFaces(u64) [7] palette error [8] exchange time-out (peer + 1)
	uint32_t* hSmall = nullptr; // pinned mirror

	// box collider arena of the last tsom_box_colliders call: 6 floats per box
	float* dBoxes = nullptr;
	uint64_t capBoxes = 0, totalBoxes = 0;
	DevBuf<uint32_t> boxScratch, boxCounts;
	DevBuf<unsigned long long> boxFirst;

	// a mesh call in flight between MeshEnqueue and MeshFinish
	struct PendingMesh
	{
		bool active = false;
		uint32_t n = 0, flags = 0;
		bool wantViews = false;
		tsomk::MeshArgs args{};
	} pendingMesh;
	bool pendingGenerate = false;
	bool pendingHalo = false; // events of a tsom_halo_pull were recorded and not read yet
	void* hPinned = nullptr;    // pinned staging for small result read-backs
	size_t hPinnedBytes = 0;

	float quat[6][4];
	uint32_t bStar = 0, aStar = 0, aInv = 0;

	// flat fast path: face attribute table for block size `faceTableTile`
	tsomk::FaceTable* dFaceTable = nullptr;
	float faceTableTile = 0.f;

	// device-side timing of the last calls (CUDA events on `stream`)
	cudaEvent_t ev[12] = {}; // 0,1 terrain  2,3 generate  4,5 mesh  6,7 halo  8,9 classify / io / collider
	float timings[TSOM_TIMING_COUNT] = {};

	cudaError_t EnsurePinned(size_t bytes)
	{
		if (bytes <= hPinnedBytes)
			return cudaSuccess;
		if (hPinned)
			cudaFreeHost(hPinned);
		hPinned = nullptr;
		hPinnedBytes = 0;
		cudaError_t e = cudaMallocHost(&hPinned, bytes);
		if (e == cudaSuccess)
			hPinnedBytes = bytes;
		return e;
	}
};

struct tsom_container
{
	tsom_ctx* ctx = nullptr;
	uint32_t capacity = 0;
	float tile = 1.f;
	float center[3] = { 0.f, 0.f, 0.f };

	int kind = TSOM_CONTAINER_PLANET;

	uint16_t* dBlocks = nullptr;
	uint16_t* dXslabs = nullptr;
	unsigned long long* dHalo = nullptr;
	int4* dChunkIdx = nullptr;
	uint8_t* dInvalid = nullptr;
	uint8_t* dSummary = nullptr;  // [capacity] tsomk::SUM_*
	uint32_t* dLastFaces = nullptr; // [capacity] faces of the chunk's last mesh (orders the mesher's tickets)
	int32_t* dNbr = nullptr;      // [capacity][6] slot of the face-adjacent local chunk with content, -1 otherwise
	uint32_t* dSync = nullptr;    // [0] blocks-final epoch [1] slabs-pulled epoch (polled by the peers over NVLink)
	uint32_t summaryTableEpoch = 0;

	std::unordered_map<Key, uint32_t, KeyHash> slotOf;
	std::vector<Key> idxOf;
	std::vector<uint8_t> exists, hasContent, hostInvalid;
	std::vector<uint32_t> freeSlots;
	uint32_t highWater = 0;
	bool topoDirty = true;
	uint64_t topoVersion = 1; // bumped whenever the set of chunks or a content flag changes (keys of the job-list caches below)

	// the index list of the last generate / mesh call and its device-resident job list: a whole-planet regenerate + remesh names the
	// same chunks every time, so the per-chunk host work (slot lookup, job entry, upload) is one memcmp
	struct JobCache
	{
		std::vector<int32_t> idx;
		DevBuf<uint32_t> jobs;
		std::vector<uint32_t> hostSlots;
		bool coversAll = false; // the list names every chunk of the container exactly once
		uint64_t version = 0;
		uint32_t n = 0;
		bool Matches(const int32_t* chunk_indices, uint32_t count, uint64_t ver) const
		{
			return version == ver && n == count && count != 0 && std::memcmp(idx.data(), chunk_indices, size_t(count) * 3 * sizeof(int32_t)) == 0;
		}
	} genCache, meshCache;
	struct GenPlan // what depends on the chunk list only, cached with genCache
	{
		int lo[3], hi[3], sumMin, sumMax;
	} genPlan{};
	DevBuf<uint16_t> templates;
	DevBuf<int4> genDesc; // per-job descriptors of the last generate call (gen_classify_kernel)
	std::vector<int32_t> planetIdx;          // chunk list of the last tsom_planet_generate (GenerateChunks order)
	uint32_t planetIdxCount[3] = { 0, 0, 0 };

	// remote neighbours (other ranks) and the ghost slabs that mirror their boundary layers
	struct Remote
	{
		int rank;
		uint32_t slot;
	};
	std::unordered_map<Key, Remote, KeyHash> remotes;
	struct Peer
	{
		const uint16_t* blocks = nullptr;
		const uint16_t* xslabs = nullptr;
		const uint32_t* sync = nullptr;
		bool ipc = false;
		tsom_container* local = nullptr; // same-process peer: ordered with CUDA events instead of polling its flags
	};
	std::map<int, Peer> peers;
	DevBuf<uint16_t> ghostSlabs;
	DevBuf<tsomk::GhostDesc> ghostDescs;
	uint32_t nGhost = 0;
	// a (remote chunk, direction) keeps its ghost slab for the lifetime of the container: a topology change leaves pulled slabs valid
	struct GhostKey
	{
		Key k;
		int d;
		bool operator==(const GhostKey& o) const { return k == o.k && d == o.d; }
	};
	struct GhostKeyHash
	{
		size_t operator()(const GhostKey& g) const { return KeyHash()(g.k) * 6 + size_t(g.d); }
	};
	std::unordered_map<GhostKey, uint32_t, GhostKeyHash> ghostIndex;
	std::vector<tsomk::GhostDesc> ghostList;
	bool ghostsStale = false; // a ghost slab exists that was never pulled: tsom_mesh refuses until the next tsom_halo_pull

	// device-side ordering of the exchange (see tsom_halo_pull in tsom_b200.h)
	// Peers in other processes poll the epoch flags in dSync over NVLink.  Peers in THIS process wait on the events below instead
	// (cudaStreamWaitEvent): nothing then depends on two kernels of different streams being resident at the same time.
	cudaEvent_t evGen = nullptr, evPull = nullptr;       // recorded behind the blocks-final / slabs-pulled flag writes
	std::atomic<uint32_t> evGenEpoch{ 0 }, evPullEpoch{ 0 }; // round of the last record, as enqueued by the host
	uint32_t exchangeTimeoutMs = 5000;
	uint32_t epoch = 0, pulledEpoch = 0;
	bool published = false, dirtySincePublish = false, needWriteWait = false;
	DevBuf<const uint32_t*> peerGenFlags, peerPullFlags;
	bool peerFlagsDirty = true;

	// terrain tables of the last generate call
	DevBuf<int16_t> terrain, tileRange;
	uint8_t* dPerm = nullptr;
	uint8_t* hPerm = nullptr; // pinned staging of the six permutation tables
	DevBuf<int> slotGrid;
	uint64_t slotGridVersion = 0; // topoVersion the device slot grid of GeneratePlatform was built for

	// dense chunk-index -> slot grid over the bounding box of the existing chunks (O(1) lookups for per-tick batches: edits,
	// dirty sets, mesh jobs); rebuilt lazily after the set of chunks changed, skipped for very sparse containers
	mutable std::vector<uint32_t> hostGrid;
	mutable std::vector<uint32_t> gridOrderXYZ; // the slots of the grid in (x, y, z) order of their chunk indices (tsom_collect_dirty's output order)
	std::vector<uint8_t> dirtySeen;             // scratch of tsom_collect_dirty
	mutable int32_t gridOrigin[3] = { 0, 0, 0 }, gridDim[3] = { 0, 0, 0 };
	mutable bool hostGridValid = false, hostGridUsable = false;

	void EnsureHostGrid() const
	{
		if (hostGridValid)
			return;
		hostGridValid = true;
		hostGridUsable = false;
		hostGrid.clear();
		if (slotOf.empty())
			return;
		int32_t lo[3] = { INT32_MAX, INT32_MAX, INT32_MAX }, hi[3] = { INT32_MIN, INT32_MIN, INT32_MIN };
		for (auto& kv : slotOf)
		{
			const int32_t v[3] = { kv.first.x, kv.first.y, kv.first.z };
			for (int a = 0; a < 3; ++a)
			{
				lo[a] = std::min(lo[a], v[a]);
				hi[a] = std::max(hi[a], v[a]);
			}
		}
		double cells = 1.0;
		for (int a = 0; a < 3; ++a)
		{
			gridOrigin[a] = lo[a];
			gridDim[a] = hi[a] - lo[a] + 1;
			cells *= double(gridDim[a]);
		}
		if (cells > std::max(65536.0, 16.0 * double(slotOf.size())))
			return;
		hostGrid.assign(size_t(cells), 0xFFFFFFFFu);
		for (auto& kv : slotOf)
			hostGrid[(size_t(kv.first.z - lo[2]) * gridDim[1] + size_t(kv.first.y - lo[1])) * gridDim[0] + size_t(kv.first.x - lo[0])] = kv.second;
		gridOrderXYZ.clear();
		gridOrderXYZ.reserve(slotOf.size());
		for (size_t x = 0; x < size_t(gridDim[0]); ++x)
			for (size_t y = 0; y < size_t(gridDim[1]); ++y)
				for (size_t z = 0; z < size_t(gridDim[2]); ++z)
				{
					const uint32_t sl = hostGrid[(z * size_t(gridDim[1]) + y) * size_t(gridDim[0]) + x];
					if (sl != 0xFFFFFFFFu)
						gridOrderXYZ.push_back(sl);
				}
		hostGridUsable = true;
	}

	uint32_t SlotOf(const Key& k) const
	{
		EnsureHostGrid();
		if (hostGridUsable)
		{
			const int64_t gx = int64_t(k.x) - gridOrigin[0], gy = int64_t(k.y) - gridOrigin[1], gz = int64_t(k.z) - gridOrigin[2];
			if (gx < 0 || gy < 0 || gz < 0 || gx >= gridDim[0] || gy >= gridDim[1] || gz >= gridDim[2])
				return 0xFFFFFFFFu;
			return hostGrid[(size_t(gz) * gridDim[1] + size_t(gy)) * gridDim[0] + size_t(gx)];
		}
		auto it = slotOf.find(k);
		return it == slotOf.end() ? 0xFFFFFFFFu : it->second;
	}
};

namespace
{
	inline void TopoChanged(tsom_container* c)
	{
		c->topoDirty = true;
		++c->topoVersion;
	}

	int AddChunk(tsom_container* c, const Key& k, uint32_t* slotOut)
	{
		if (c->hostGridValid && c->hostGridUsable)
		{
			// the dense grid is current (nothing was added or removed since it was built): one array read instead of a hash probe —
			// regenerating or resetting a whole planet asks for every chunk again
			const uint32_t slot = c->SlotOf(k);
			if (slot != 0xFFFFFFFFu)
			{
				if (slotOut)
					*slotOut = slot;
				return TSOM_OK;
			}
		}
		else
		{
			auto it = c->slotOf.find(k);
			if (it != c->slotOf.end())
			{
				if (slotOut)
					*slotOut = it->second;
				return TSOM_OK;
			}
		}
		uint32_t slot;
		if (!c->freeSlots.empty())
		{
			slot = c->freeSlots.back();
			c->freeSlots.pop_back();
		}
		else
		{
			if (c->highWater >= c->capacity)
				return Fail(TSOM_ERR_CAPACITY, "container is full");
			slot = c->highWater++;
		}
		c->slotOf.emplace(k, slot);
		c->hostGridValid = false;
		c->idxOf[slot] = k;
		c->exists[slot] = 1;
		c->hasContent[slot] = 0;
		c->hostInvalid[slot] = 0;
		TopoChanged(c);
		if (slotOut)
			*slotOut = slot;
		return TSOM_OK;
	}

	// (re)build the per-chunk index, neighbour and halo pointer tables and the ghost list, then upload them
	int SyncTopology(tsom_container* c)
	{
		if (!c->topoDirty)
			return TSOM_OK;
		tsom_ctx* ctx = c->ctx;
		const uint32_t n = c->highWater;
		std::vector<int4> idx(n);
		std::vector<unsigned long long> halo(size_t(n) * 6, 0ull);
		std::vector<int32_t> nbr(size_t(n) * 6, -1);
		struct GhostUse
		{
			uint32_t s, d, ghost;
		};
		std::vector<GhostUse> ghostUses;
		bool ghostListChanged = false;

		for (uint32_t s = 0; s < n; ++s)
		{
			if (!c->exists[s])
			{
				idx[s] = make_int4(0, 0, 0, 0);
				continue;
			}
			const Key k = c->idxOf[s];
			idx[s] = make_int4(k.x, k.y, k.z, c->hasContent[s] ? 1 : 0);
			if (!c->hasContent[s])
				continue;
			for (int d = 0; d < 6; ++d)
			{
				const Key nk{ k.x + kChunkDirOffset[d][0], k.y + kChunkDirOffset[d][1], k.z + kChunkDirOffset[d][2] };
				// which boundary layer of the neighbour faces us
				const bool far = (d == TSOM_DOWN || d == TSOM_FRONT || d == TSOM_LEFT); // neighbour's layer 31
				const uint32_t ns = c->SlotOf(nk);
				if (ns != 0xFFFFFFFFu)
				{
					if (!c->hasContent[ns])
						continue; // !HasContent() => face drawn, same as no neighbour (Chunk.cpp:165-166)
					const uint16_t* nb = c->dBlocks + size_t(ns) * NB;
					const uint16_t* nx = c->dXslabs + size_t(ns) * 2 * SLAB;
					unsigned long long e;
					if (d == TSOM_UP || d == TSOM_DOWN)
						e = reinterpret_cast<unsigned long long>(nb + (far ? 31 * SLAB : 0)) | tsomk_halo_contig();
					else if (d == TSOM_BACK || d == TSOM_FRONT)
						e = reinterpret_cast<unsigned long long>(nb + (far ? 31 * CS : 0)) | tsomk_halo_rows();
					else
						e = reinterpret_cast<unsigned long long>(nx + (far ? SLAB : 0)) | tsomk_halo_contig();
					halo[size_t(s) * 6 + d] = e;
					nbr[size_t(s) * 6 + d] = int32_t(ns);
					continue;
				}
				if (c->remotes.empty())
					continue;
				auto rt = c->remotes.find(nk);
				if (rt != c->remotes.end())
				{
					auto pt = c->peers.find(rt->second.rank);
					if (pt == c->peers.end())
						return Fail(TSOM_ERR_INVALID, "remote neighbour's rank has no attached peer buffers");
					const uint16_t* nb = pt->second.blocks + size_t(rt->second.slot) * NB;
					const uint16_t* nx = pt->second.xslabs + size_t(rt->second.slot) * 2 * SLAB;
					tsomk::GhostDesc g;
					if (d == TSOM_UP || d == TSOM_DOWN)
					{
						g.src = nb + (far ? 31 * SLAB : 0);
						g.contig = 1;
					}
					else if (d == TSOM_BACK || d == TSOM_FRONT)
					{
						g.src = nb + (far ? 31 * CS : 0);
						g.contig = 0;
					}
					else
					{
						g.src = nx + (far ? SLAB : 0);
						g.contig = 1;
					}
					// stable ghost index per (remote chunk, direction): slabs that were pulled stay where they are
					const tsom_container::GhostKey gk{ nk, d };
					auto git = c->ghostIndex.find(gk);
					if (git == c->ghostIndex.end())
					{
						g.ghost = uint32_t(c->ghostList.size());
						c->ghostIndex.emplace(gk, g.ghost);
						c->ghostList.push_back(g);
						c->ghostsStale = true; // nothing has been pulled into this slab yet
						ghostListChanged = true;
					}
					else
					{
						g.ghost = git->second;
						if (c->ghostList[g.ghost].src != g.src || c->ghostList[g.ghost].contig != g.contig)
						{
							c->ghostList[g.ghost] = g; // the peer was re-attached or the chunk moved to another slot
							c->ghostsStale = true;
							ghostListChanged = true;
						}
					}
					ghostUses.push_back(GhostUse{ s, uint32_t(d), g.ghost });
				}
			}
		}

		c->nGhost = uint32_t(c->ghostList.size());
		if (c->nGhost)
		{
			TSOM_CUDA(c->ghostSlabs.EnsureKeep(size_t(c->nGhost) * SLAB, c->ghostSlabs.cap, ctx->stream));
			TSOM_CUDA(c->ghostDescs.Ensure(c->nGhost));
			if (ghostListChanged)
				TSOM_CUDA(cudaMemcpyAsync(c->ghostDescs.ptr, c->ghostList.data(), c->ghostList.size() * sizeof(tsomk::GhostDesc), cudaMemcpyHostToDevice, ctx->stream));
			for (const GhostUse& u : ghostUses)
				halo[size_t(u.s) * 6 + u.d] = reinterpret_cast<unsigned long long>(c->ghostSlabs.ptr + size_t(u.ghost) * SLAB) | tsomk_halo_contig();
		}

		if (n)
		{
			TSOM_CUDA(cudaMemcpyAsync(c->dChunkIdx, idx.data(), size_t(n) * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
			TSOM_CUDA(cudaMemcpyAsync(c->dHalo, halo.data(), halo.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
			TSOM_CUDA(cudaMemcpyAsync(c->dNbr, nbr.data(), nbr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
			TSOM_CUDA(cudaStreamSynchronize(ctx->stream)); // the host vectors die at scope exit
		}
		c->topoDirty = false;
		return TSOM_OK;
	}

	// ---- device-side ordering of the halo exchange (tsom_halo_publish / tsom_halo_pull, see tsom_b200.h)

	int SyncPeerFlags(tsom_container* c)
	{
		if (!c->peerFlagsDirty)
			return TSOM_OK;
		std::vector<const uint32_t*> gen, pull;
		for (auto& [rank, p] : c->peers)
		{
			gen.push_back(p.sync + 0);
			pull.push_back(p.sync + 1);
		}
		const size_t np = gen.size();
		if (np)
		{
			TSOM_CUDA(c->peerGenFlags.Ensure(np));
			TSOM_CUDA(c->peerPullFlags.Ensure(np));
			TSOM_CUDA(cudaMemcpyAsync(c->peerGenFlags.ptr, gen.data(), np * sizeof(const uint32_t*), cudaMemcpyHostToDevice, c->ctx->stream));
			TSOM_CUDA(cudaMemcpyAsync(c->peerPullFlags.ptr, pull.data(), np * sizeof(const uint32_t*), cudaMemcpyHostToDevice, c->ctx->stream));
			TSOM_CUDA(cudaStreamSynchronize(c->ctx->stream));
		}
		c->peerFlagsDirty = false;
		return TSOM_OK;
	}

	// Make this container's stream wait until every peer has reached `round` of the flag at `which` (0 blocks-final, 1 slabs-pulled).
	// A peer of this process whose host has already enqueued that round is waited for with its event; everybody else (peers in other
	// processes, or a local peer that is behind on the host) through the polling kernel, which gives up after the time-out.
	int WaitForPeers(tsom_container* c, int which, uint32_t round)
	{
		if (c->peers.empty())
			return TSOM_OK;
		tsom_ctx* ctx = c->ctx;
		bool poll = false;
		for (auto& [rank, p] : c->peers)
		{
			if (p.local)
			{
				const uint32_t done = which == 0 ? p.local->evGenEpoch.load() : p.local->evPullEpoch.load();
				if (int32_t(done - round) >= 0)
				{
					TSOM_CUDA(cudaStreamWaitEvent(ctx->stream, which == 0 ? p.local->evGen : p.local->evPull, 0));
					continue;
				}
			}
			poll = true;
		}
		if (!poll)
			return TSOM_OK;
		int rc = SyncPeerFlags(c);
		if (rc != TSOM_OK)
			return rc;
		// (a peer already waited for by event passes the poll at once: its flag write precedes its event)
		tsomk::launch_flag_wait(which == 0 ? c->peerGenFlags.ptr : c->peerPullFlags.ptr, uint32_t(c->peers.size()), round, 1000000ull * c->exchangeTimeoutMs, ctx->dSmall + 8, ctx->stream);
		ctx->launches += 1;
		TSOM_CUDA(cudaGetLastError());
		return TSOM_OK;
	}

	// Called by everything that writes block ids: once per round, before the first write, wait (on the device) until every peer
	// has pulled the slabs of the last published round — a fast GPU must not overwrite what a slower one is still reading.
	int BeforeBlockWrite(tsom_container* c)
	{
		c->dirtySincePublish = true;
		if (!c->needWriteWait)
			return TSOM_OK;
		c->needWriteWait = false;
		TSOM_CUDA(cudaSetDevice(c->ctx->device));
		return WaitForPeers(c, 1, c->pulledEpoch);
	}

	// after a stream synchronisation that brought dSmall[8] to the host: did a wait kernel give up?
	int CheckExchangeTimeout(tsom_ctx* ctx)
	{
		if (ctx->hSmall[8] == 0u)
			return TSOM_OK;
		const uint32_t who = ctx->hSmall[8] - 1u;
		ctx->hSmall[8] = 0u;
		cudaMemsetAsync(ctx->dSmall + 8, 0, sizeof(uint32_t), ctx->stream);
		return Fail(TSOM_ERR_INTERNAL, "halo exchange timed out: attached peer #" + std::to_string(who) + " did not publish / pull its round in time");
	}

	constexpr size_t ARENA_GUARD_BYTES = 256; // 0xA5 behind the last face of every arena, checked by tsom_debug_check_guards

	int EnsureArenas(tsom_ctx* ctx, uint64_t faces, uint32_t flags)
	{
		auto grow = [&](auto*& ptr, uint64_t& cap, size_t bytesPerFace) -> int {
			if (faces <= cap)
				return TSOM_OK;
			uint64_t want = std::max<uint64_t>(faces, 2 * cap); // geometric growth: a reallocation costs tens of ms
			if (ptr)
				cudaFree(ptr);
			ptr = nullptr;
			cap = 0;
			void* p = nullptr;
			cudaError_t e = cudaMalloc(&p, want * bytesPerFace + ARENA_GUARD_BYTES);
			if (e != cudaSuccess)
				return Fail(TSOM_ERR_NOMEM, std::string("mesh arena cudaMalloc: ") + cudaGetErrorString(e));
			if ((e = cudaMemsetAsync(static_cast<char*>(p) + want * bytesPerFace, 0xA5, ARENA_GUARD_BYTES, ctx->stream)) != cudaSuccess)
			{
				cudaFree(p);
				return Fail(TSOM_ERR_CUDA, std::string("mesh arena guard: ") + cudaGetErrorString(e));
			}
			ptr = reinterpret_cast<std::remove_reference_t<decltype(ptr)>>(p);
			cap = want;
			return TSOM_OK;
		};
		int rc;
		if ((rc = grow(ctx->dIndices, ctx->capIndices, TSOM_FACE_INDEX_BYTES)) != TSOM_OK)
			return rc;
		if ((flags & TSOM_MESH_RENDER) && (rc = grow(ctx->dVertices, ctx->capVertices, TSOM_FACE_VERTEX_BYTES)) != TSOM_OK)
			return rc;
		if ((flags & TSOM_MESH_COLLIDER) && (rc = grow(ctx->dCollider, ctx->capCollider, TSOM_FACE_COLLIDER_BYTES)) != TSOM_OK)
			return rc;
		return TSOM_OK;
	}

	uint64_t ArenaFaces(const tsom_ctx* ctx, uint32_t flags)
	{
		uint64_t cap = ctx->capIndices;
		if (flags & TSOM_MESH_RENDER)
			cap = std::min(cap, ctx->capVertices);
		if (flags & TSOM_MESH_COLLIDER)
			cap = std::min(cap, ctx->capCollider);
		return cap;
	}
}

// =====================================================================================================================
extern "C"
{
	const char* tsom_last_error(void) { return g_lastError.c_str(); }
	const char* tsom_version(void) { return "tsom_b200 0.1 (sm_100a)"; }

	int tsom_create(int cuda_device, tsom_ctx** out)
	{
		if (!out)
			return Fail(TSOM_ERR_INVALID, "out is null");
		*out = nullptr;
		int count = 0;
		cudaError_t e = cudaGetDeviceCount(&count);
		if (e != cudaSuccess || count == 0)
			return Fail(TSOM_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + " (libtsom_b200 has no CPU fallback)");
		if (cuda_device < 0 || cuda_device >= count)
			return Fail(TSOM_ERR_INVALID, "cuda_device out of range");
		TSOM_CUDA(cudaSetDevice(cuda_device));
		cudaDeviceProp prop;
		TSOM_CUDA(cudaGetDeviceProperties(&prop, cuda_device));
		if (prop.major < 10)
			return Fail(TSOM_ERR_CUDA, std::string("device ") + prop.name + " is not sm_100-class; this library ships sm_100a code only");

		// a failure below goes through tsom_destroy, which releases whatever was created so far
		tsom_ctx* ctx = new (std::nothrow) tsom_ctx();
		if (!ctx)
			return Fail(TSOM_ERR_NOMEM, "out of host memory");
		ctx->device = cuda_device;
		ctx->numSMs = prop.multiProcessorCount;
		auto init = [&]() -> int {
			TSOM_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
			TSOM_CUDA(tsomk::mesh_kernels_init());
			TSOM_CUDA(tsomk::lz4_kernels_init());
			TSOM_CUDA(cudaMalloc(&ctx->dSmall, 16 * sizeof(uint32_t)));
			TSOM_CUDA(cudaMemset(ctx->dSmall, 0, 16 * sizeof(uint32_t)));
			TSOM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&ctx->hSmall), 16 * sizeof(uint32_t)));
			std::memset(ctx->hSmall, 0, 16 * sizeof(uint32_t));
			for (auto& e : ctx->ev)
				TSOM_CUDA(cudaEventCreate(&e));
			TSOM_CUDA(cudaMalloc(&ctx->dFaceTable, sizeof(tsomk::FaceTable)));
			return TSOM_OK;
		};
		int rc = init();
		if (rc != TSOM_OK)
		{
			const std::string msg = g_lastError;
			tsom_destroy(ctx);
			return Fail(rc, msg);
		}

		const F3h normals[6] = { { 0.f, 0.f, 1.f }, { 0.f, -1.f, 0.f }, { 0.f, 0.f, -1.f }, { -1.f, 0.f, 0.f }, { 1.f, 0.f, 0.f }, { 0.f, 1.f, 0.f } };
		for (int d = 0; d < 6; ++d)
			RotationBetween(normals[d], F3h{ 0.f, 1.f, 0.f }, ctx->quat[d]);

		if (!DeriveBernoulliThreshold(ctx->bStar, ctx->aStar))
		{
			tsom_destroy(ctx);
			return Fail(TSOM_ERR_UNSUPPORTED, "std::bernoulli_distribution(0.9) over std::minstd_rand is not a two-level integer threshold on this standard library");
		}
		{
			// 48271^-1 mod (2^31 - 1) by Fermat
			auto mulmod = [](uint64_t a, uint64_t b) { return (a * b) % 2147483647ull; };
			uint64_t r = 1, b = 48271, ex = 2147483647ull - 2;
			while (ex)
			{
				if (ex & 1)
					r = mulmod(r, b);
				b = mulmod(b, b);
				ex >>= 1;
			}
			ctx->aInv = uint32_t(r);
		}
		*out = ctx;
		return TSOM_OK;
	}

	void tsom_destroy(tsom_ctx* ctx)
	{
		if (!ctx)
			return;
		cudaSetDevice(ctx->device);
		if (ctx->stream)
			cudaStreamSynchronize(ctx->stream);
		cudaFree(ctx->dFlags);
		cudaFree(ctx->dTex);
		cudaFree(ctx->dVertices);
		cudaFree(ctx->dIndices);
		cudaFree(ctx->dCollider);
		ctx->jobSlot.Free();
		ctx->faceCount.Free();
		ctx->jobOrder.Free();
		ctx->firstFace.Free();
		ctx->status.Free();
		ctx->editKeys.Free();
		ctx->ioPayload.Free();
		ctx->ioPalette.Free();
		ctx->ioCounts.Free();
		ctx->ioStreams.Free();
		ctx->ioOffsets.Free();
		ctx->editVals.Free();
		ctx->gravity.Free();
		ctx->edits.Free();
		cudaFree(ctx->dBoxes);
		ctx->boxScratch.Free();
		ctx->boxCounts.Free();
		ctx->boxFirst.Free();
		cudaFree(ctx->dSmall);
		cudaFree(ctx->dFaceTable);
		for (auto& e : ctx->ev)
			if (e)
				cudaEventDestroy(e);
		if (ctx->hSmall)
			cudaFreeHost(ctx->hSmall);
		if (ctx->hPinned)
			cudaFreeHost(ctx->hPinned);
		if (ctx->stream)
			cudaStreamDestroy(ctx->stream);
		delete ctx;
	}

	void* tsom_stream(tsom_ctx* ctx) { return ctx ? ctx->stream : nullptr; }

	int tsom_synchronize(tsom_ctx* ctx)
	{
		if (!ctx)
			return Fail(TSOM_ERR_INVALID, "ctx is null");
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
		if (ctx->pendingGenerate)
		{
			ctx->pendingGenerate = false;
			cudaEventElapsedTime(&ctx->timings[TSOM_TIMING_TERRAIN_MS], ctx->ev[0], ctx->ev[1]);
			cudaEventElapsedTime(&ctx->timings[TSOM_TIMING_GENERATE_MS], ctx->ev[2], ctx->ev[3]);
		}
		return TSOM_OK;
	}

	int tsom_ctx_lock(tsom_ctx* ctx)
	{
		if (!ctx)
			return Fail(TSOM_ERR_INVALID, "ctx is null");
		ctx->mutex.lock();
		return TSOM_OK;
	}

	int tsom_ctx_unlock(tsom_ctx* ctx)
	{
		if (!ctx)
			return Fail(TSOM_ERR_INVALID, "ctx is null");
		ctx->mutex.unlock();
		return TSOM_OK;
	}

	uint64_t tsom_launch_count(tsom_ctx* ctx) { return ctx ? ctx->launches : 0; }

	int tsom_get_timings(tsom_ctx* ctx, float* out, uint32_t n)
	{
		if (!ctx || !out)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(ctx);
		for (uint32_t i = 0; i < n && i < TSOM_TIMING_COUNT; ++i)
			out[i] = ctx->timings[i];
		return TSOM_OK;
	}

	int tsom_set_block_table(tsom_ctx* ctx, const tsom_block_desc* blocks, uint32_t n)
	{
		if (!ctx || !blocks || n == 0 || n > 65535)
			return Fail(TSOM_ERR_INVALID, "bad block table");
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		std::vector<uint8_t> flags(n);
		std::vector<uint32_t> tex(size_t(n) * 6);
		for (uint32_t i = 0; i < n; ++i)
		{
			flags[i] = uint8_t((blocks[i].has_collisions ? 1u : 0u) | (blocks[i].is_double_sided ? 2u : 0u) | (blocks[i].is_transparent ? 4u : 0u));
			for (int d = 0; d < 6; ++d)
				tex[size_t(i) * 6 + d] = blocks[i].tex[d];
		}
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->dFlags);
		cudaFree(ctx->dTex);
		ctx->dFlags = nullptr;
		ctx->dTex = nullptr;
		TSOM_CUDA(cudaMalloc(&ctx->dFlags, n));
		TSOM_CUDA(cudaMalloc(&ctx->dTex, size_t(n) * 6 * sizeof(uint32_t)));
		TSOM_CUDA(cudaMemcpy(ctx->dFlags, flags.data(), n, cudaMemcpyHostToDevice));
		TSOM_CUDA(cudaMemcpy(ctx->dTex, tex.data(), tex.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
		ctx->nTypes = n;
		ctx->texFits8 = std::all_of(tex.begin(), tex.end(), [](uint32_t t) { return t < 256u; });
		ctx->hostFlags = flags;
		++ctx->tableEpoch; // chunk summaries made under the previous table no longer hold
		return TSOM_OK;
	}

	int tsom_reserve_faces(tsom_ctx* ctx, uint64_t faces, uint32_t flags)
	{
		if (!ctx)
			return Fail(TSOM_ERR_INVALID, "ctx is null");
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
		return EnsureArenas(ctx, faces, flags);
	}

	int tsom_debug_check_guards(tsom_ctx* ctx)
	{
		if (!ctx)
			return Fail(TSOM_ERR_INVALID, "null context");
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
		struct { const void* ptr; uint64_t cap; size_t bytesPerFace; const char* name; } arenas[3] = {
			{ ctx->dIndices, ctx->capIndices, TSOM_FACE_INDEX_BYTES, "index" },
			{ ctx->dVertices, ctx->capVertices, TSOM_FACE_VERTEX_BYTES, "vertex" },
			{ ctx->dCollider, ctx->capCollider, TSOM_FACE_COLLIDER_BYTES, "collider" },
		};
		for (auto& a : arenas)
		{
			if (!a.ptr)
				continue;
			unsigned char guard[ARENA_GUARD_BYTES];
			TSOM_CUDA(cudaMemcpy(guard, static_cast<const char*>(a.ptr) + a.cap * a.bytesPerFace, ARENA_GUARD_BYTES, cudaMemcpyDeviceToHost));
			for (size_t i = 0; i < ARENA_GUARD_BYTES; ++i)
				if (guard[i] != 0xA5)
					return Fail(TSOM_ERR_INTERNAL, std::string("a kernel wrote behind the ") + a.name + " arena (guard byte " + std::to_string(i) + ")");
		}
		return TSOM_OK;
	}

	// ---------------------------------------------------------------------------------------------------- container
	int tsom_container_create(tsom_ctx* ctx, uint32_t capacity_chunks, float tile_size, tsom_container** out)
	{
		if (!ctx || !out || capacity_chunks == 0)
			return Fail(TSOM_ERR_INVALID, "bad container arguments");
		*out = nullptr;
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		auto c = std::make_unique<tsom_container>();
		c->ctx = ctx;
		c->capacity = capacity_chunks;
		c->tile = tile_size;
		cudaError_t e;
		const size_t maskBytes = (size_t(capacity_chunks) + 3u) & ~size_t(3); // edit_apply_kernel ORs whole 32-bit words
		if ((e = cudaMalloc(&c->dBlocks, size_t(capacity_chunks) * NB * sizeof(uint16_t))) != cudaSuccess ||
		    (e = cudaMalloc(&c->dXslabs, size_t(capacity_chunks) * 2 * SLAB * sizeof(uint16_t))) != cudaSuccess ||
		    (e = cudaMalloc(&c->dHalo, size_t(capacity_chunks) * 6 * sizeof(unsigned long long))) != cudaSuccess ||
		    (e = cudaMalloc(&c->dChunkIdx, size_t(capacity_chunks) * sizeof(int4))) != cudaSuccess ||
		    (e = cudaMalloc(&c->dInvalid, maskBytes)) != cudaSuccess ||
		    (e = cudaMalloc(&c->dSummary, maskBytes)) != cudaSuccess ||
		    (e = cudaMalloc(&c->dLastFaces, size_t(capacity_chunks) * sizeof(uint32_t))) != cudaSuccess ||
		    (e = cudaMalloc(&c->dNbr, size_t(capacity_chunks) * 6 * sizeof(int32_t))) != cudaSuccess ||
		    (e = cudaMalloc(&c->dSync, 64 * sizeof(uint32_t))) != cudaSuccess ||
		    (e = cudaMalloc(&c->dPerm, 6 * 256)) != cudaSuccess ||
		    (e = cudaMallocHost(reinterpret_cast<void**>(&c->hPerm), 6 * 256)) != cudaSuccess)
		{
			cudaFree(c->dBlocks);
			cudaFree(c->dXslabs);
			cudaFree(c->dHalo);
			cudaFree(c->dChunkIdx);
			cudaFree(c->dInvalid);
			cudaFree(c->dSummary);
			cudaFree(c->dLastFaces);
			cudaFree(c->dNbr);
			cudaFree(c->dSync);
			cudaFree(c->dPerm);
			return Fail(TSOM_ERR_NOMEM, std::string("container cudaMalloc: ") + cudaGetErrorString(e));
		}
		TSOM_CUDA(cudaMemset(c->dInvalid, 0, maskBytes));
		TSOM_CUDA(cudaMemset(c->dSummary, 0, maskBytes));
		TSOM_CUDA(cudaMemset(c->dLastFaces, 0, size_t(capacity_chunks) * sizeof(uint32_t)));
		TSOM_CUDA(cudaMemset(c->dSync, 0, 64 * sizeof(uint32_t)));
		TSOM_CUDA(cudaEventCreateWithFlags(&c->evGen, cudaEventDisableTiming));
		TSOM_CUDA(cudaEventCreateWithFlags(&c->evPull, cudaEventDisableTiming));
		c->summaryTableEpoch = ctx->tableEpoch;
		c->idxOf.resize(capacity_chunks);
		c->exists.assign(capacity_chunks, 0);
		c->hasContent.assign(capacity_chunks, 0);
		c->hostInvalid.assign(capacity_chunks, 0);
		*out = c.release();
		return TSOM_OK;
	}

	void tsom_container_destroy(tsom_container* c)
	{
		if (!c)
			return;
		TSOM_LOCK(c->ctx);
		cudaSetDevice(c->ctx->device);
		cudaStreamSynchronize(c->ctx->stream);
		for (auto& [rank, p] : c->peers)
		{
			if (p.ipc)
			{
				cudaIpcCloseMemHandle(const_cast<uint16_t*>(p.blocks));
				cudaIpcCloseMemHandle(const_cast<uint16_t*>(p.xslabs));
				cudaIpcCloseMemHandle(const_cast<uint32_t*>(p.sync));
			}
		}
		cudaFree(c->dBlocks);
		cudaFree(c->dXslabs);
		cudaFree(c->dHalo);
		cudaFree(c->dChunkIdx);
		cudaFree(c->dInvalid);
		cudaFree(c->dSummary);
		cudaFree(c->dLastFaces);
		cudaFree(c->dNbr);
		cudaFree(c->dSync);
		if (c->evGen)
			cudaEventDestroy(c->evGen);
		if (c->evPull)
			cudaEventDestroy(c->evPull);
		cudaFree(c->dPerm);
		if (c->hPerm)
			cudaFreeHost(c->hPerm);
		c->ghostSlabs.Free();
		c->ghostDescs.Free();
		c->peerGenFlags.Free();
		c->peerPullFlags.Free();
		c->genCache.jobs.Free();
		c->meshCache.jobs.Free();
		c->templates.Free();
		c->genDesc.Free();
		c->terrain.Free();
		c->tileRange.Free();
		c->slotGrid.Free();
		delete c;
	}

	int tsom_container_add_chunks(tsom_container* c, const int32_t* chunk_indices, uint32_t n)
	{
		if (!c || (!chunk_indices && n))
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(c->ctx);
		for (uint32_t i = 0; i < n; ++i)
		{
			int rc = AddChunk(c, Key{ chunk_indices[i * 3], chunk_indices[i * 3 + 1], chunk_indices[i * 3 + 2] }, nullptr);
			if (rc != TSOM_OK)
				return rc;
		}
		return TSOM_OK;
	}

	int tsom_container_remove_chunk(tsom_container* c, const int32_t chunk_index[3])
	{
		if (!c || !chunk_index)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(c->ctx);
		Key k{ chunk_index[0], chunk_index[1], chunk_index[2] };
		auto it = c->slotOf.find(k);
		if (it == c->slotOf.end())
			return Fail(TSOM_ERR_INVALID, "unknown chunk");
		TSOM_CUDA(cudaSetDevice(c->ctx->device));
		const uint32_t slot = it->second;
		c->exists[slot] = 0;
		c->hasContent[slot] = 0;
		c->hostInvalid[slot] = 0;
		TSOM_CUDA(cudaMemsetAsync(c->dInvalid + slot, 0, 1, c->ctx->stream));
		c->freeSlots.push_back(slot);
		c->slotOf.erase(it);
		c->hostGridValid = false;
		TopoChanged(c);
		return TSOM_OK;
	}

	int tsom_container_set_kind(tsom_container* c, int kind)
	{
		if (!c || (kind != TSOM_CONTAINER_PLANET && kind != TSOM_CONTAINER_SHIP))
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(c->ctx);
		c->kind = kind;
		return TSOM_OK;
	}

	uint32_t tsom_container_chunk_count(const tsom_container* c) { return c ? uint32_t(c->slotOf.size()) : 0; }

	// after new ids landed in dBlocks (ctx->jobSlot holds `slots` on the device): x-slabs, content flags, OnReset invalidation
	static int FinishReset(tsom_container* c, const std::vector<uint32_t>& slots)
	{
		tsom_ctx* ctx = c->ctx;
		const uint32_t n = uint32_t(slots.size());
		tsomk::launch_build_xslabs(c->dBlocks, c->dXslabs, c->dSummary, ctx->jobSlot.ptr, n, ctx->stream);
		ctx->launches += n ? 1 : 0;
		TSOM_CUDA(cudaGetLastError());
		for (uint32_t k = 0; k < n; ++k)
		{
			if (!c->hasContent[slots[k]])
				TopoChanged(c);
			c->hasContent[slots[k]] = 1;
			c->hostInvalid[slots[k]] = 0x80u | 0x3Fu; // OnReset -> OnChunkUpdated(DirectionMask_All) (Planet.cpp:34-37)
		}
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream)); // host staging (`slots`, possibly pageable sources) must outlive the copies
		return TSOM_OK;
	}

	static int ResetCommon(tsom_container* c, const int32_t* chunk_indices, uint32_t n, const void* src, cudaMemcpyKind kind)
	{
		if (!c || !chunk_indices || !src)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		tsom_ctx* ctx = c->ctx;
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		std::vector<uint32_t> slots(n);
		for (uint32_t i = 0; i < n; ++i)
		{
			int rc = AddChunk(c, Key{ chunk_indices[i * 3], chunk_indices[i * 3 + 1], chunk_indices[i * 3 + 2] }, &slots[i]);
			if (rc != TSOM_OK)
				return rc;
		}
		{
			int rc = BeforeBlockWrite(c);
			if (rc != TSOM_OK)
				return rc;
		}
		// coalesce runs of consecutive slots into single copies
		const uint16_t* s = static_cast<const uint16_t*>(src);
		uint32_t i = 0;
		while (i < n)
		{
			uint32_t j = i + 1;
			while (j < n && slots[j] == slots[j - 1] + 1)
				++j;
			TSOM_CUDA(cudaMemcpyAsync(c->dBlocks + size_t(slots[i]) * NB, s + size_t(i) * NB, size_t(j - i) * NB * sizeof(uint16_t), kind, ctx->stream));
			i = j;
		}
		TSOM_CUDA(ctx->jobSlot.Ensure(n));
		TSOM_CUDA(cudaMemcpyAsync(ctx->jobSlot.ptr, slots.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
		return FinishReset(c, slots);
	}

	int tsom_chunks_reset(tsom_container* c, const int32_t* chunk_indices, uint32_t n, const uint16_t* blocks_host)
	{
		return ResetCommon(c, chunk_indices, n, blocks_host, cudaMemcpyHostToDevice);
	}

	int tsom_chunks_reset_device(tsom_container* c, const int32_t* chunk_indices, uint32_t n, const void* blocks_device)
	{
		return ResetCommon(c, chunk_indices, n, blocks_device, cudaMemcpyDeviceToDevice);
	}

	int tsom_chunks_palettize(tsom_container* c, const int32_t* chunk_indices, uint32_t n, uint16_t* palette_out, uint32_t palette_stride, uint32_t* palette_counts_out,
	                          uint8_t* payload_out, int payload_big_endian)
	{
		if (!c || !chunk_indices || !palette_out || !palette_counts_out || !payload_out || palette_stride == 0)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		if (n == 0)
			return TSOM_OK;
		tsom_ctx* ctx = c->ctx;
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		std::vector<uint32_t> slots(n);
		for (uint32_t i = 0; i < n; ++i)
		{
			slots[i] = c->SlotOf(Key{ chunk_indices[i * 3], chunk_indices[i * 3 + 1], chunk_indices[i * 3 + 2] });
			if (slots[i] == 0xFFFFFFFFu || !c->hasContent[slots[i]])
				return Fail(TSOM_ERR_INVALID, "chunk is unknown or has no content");
		}
		TSOM_CUDA(ctx->jobSlot.Ensure(n));
		TSOM_CUDA(ctx->ioPayload.Ensure(size_t(n) * NB * 2));
		TSOM_CUDA(ctx->ioPalette.Ensure(size_t(n) * palette_stride));
		TSOM_CUDA(ctx->ioCounts.Ensure(n));
		TSOM_CUDA(cudaMemcpyAsync(ctx->jobSlot.ptr, slots.data(), size_t(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
		tsomk::launch_palettize(c->dBlocks, ctx->jobSlot.ptr, n, ctx->ioPalette.ptr, palette_stride, ctx->ioCounts.ptr, ctx->ioPayload.ptr, payload_big_endian, ctx->stream);
		ctx->launches += 1;
		TSOM_CUDA(cudaGetLastError());
		TSOM_CUDA(cudaMemcpyAsync(palette_counts_out, ctx->ioCounts.ptr, size_t(n) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		TSOM_CUDA(cudaMemcpyAsync(palette_out, ctx->ioPalette.ptr, size_t(n) * palette_stride * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
		TSOM_CUDA(cudaMemcpyAsync(payload_out, ctx->ioPayload.ptr, size_t(n) * NB * 2, cudaMemcpyDeviceToHost, ctx->stream));
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
		for (uint32_t i = 0; i < n; ++i)
			if (palette_counts_out[i] > palette_stride)
				return Fail(TSOM_ERR_CAPACITY, "a chunk has more distinct block types than palette_stride");
		return TSOM_OK;
	}

	int tsom_chunks_reset_palettized(tsom_container* c, const int32_t* chunk_indices, uint32_t n, const uint16_t* palettes, uint32_t palette_stride, const uint32_t* palette_counts,
	                                 const uint8_t* payload, int payload_big_endian)
	{
		if (!c || !chunk_indices || !palettes || !palette_counts || !payload || palette_stride == 0)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		if (n == 0)
			return TSOM_OK;
		tsom_ctx* ctx = c->ctx;
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		for (uint32_t i = 0; i < n; ++i)
			if (palette_counts[i] == 0 || palette_counts[i] > palette_stride)
				return Fail(TSOM_ERR_INVALID, "palette count must be in [1, palette_stride]");
		std::vector<uint32_t> slots(n);
		for (uint32_t i = 0; i < n; ++i)
		{
			int rc = AddChunk(c, Key{ chunk_indices[i * 3], chunk_indices[i * 3 + 1], chunk_indices[i * 3 + 2] }, &slots[i]);
			if (rc != TSOM_OK)
				return rc;
		}
		TSOM_CUDA(ctx->jobSlot.Ensure(n));
		TSOM_CUDA(ctx->ioPayload.Ensure(size_t(n) * NB * 2));
		TSOM_CUDA(ctx->ioPalette.Ensure(size_t(n) * palette_stride));
		TSOM_CUDA(ctx->ioCounts.Ensure(n));
		TSOM_CUDA(cudaMemcpyAsync(ctx->jobSlot.ptr, slots.data(), size_t(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
		TSOM_CUDA(cudaMemcpyAsync(ctx->ioCounts.ptr, palette_counts, size_t(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
		TSOM_CUDA(cudaMemcpyAsync(ctx->ioPalette.ptr, palettes, size_t(n) * palette_stride * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
		TSOM_CUDA(cudaMemcpyAsync(ctx->ioPayload.ptr, payload, size_t(n) * NB * 2, cudaMemcpyHostToDevice, ctx->stream));
		TSOM_CUDA(cudaMemsetAsync(ctx->dSmall + 7, 0, sizeof(uint32_t), ctx->stream));
		{
			int rc = BeforeBlockWrite(c);
			if (rc != TSOM_OK)
				return rc;
		}
		tsomk::launch_depalettize(c->dBlocks, ctx->jobSlot.ptr, n, ctx->ioPalette.ptr, palette_stride, ctx->ioCounts.ptr, ctx->ioPayload.ptr, payload_big_endian, ctx->dSmall + 7, ctx->stream);
		ctx->launches += 1;
		TSOM_CUDA(cudaGetLastError());
		TSOM_CUDA(cudaMemcpyAsync(ctx->hSmall + 7, ctx->dSmall + 7, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		int rc = FinishReset(c, slots);
		if (rc != TSOM_OK)
			return rc;
		if (ctx->hSmall[7] != 0)
			return Fail(TSOM_ERR_INVALID, "a payload index is out of its chunk's palette (malformed chunk data); such blocks were set to Empty");
		return TSOM_OK;
	}

	// Packets::ChunkReset payloads (Packets.cpp:172-212): one LZ4 block per chunk, the bytes BinaryCompressor::Compress produces.
	int tsom_chunks_lz4_compress(tsom_container* c, const int32_t* chunk_indices, uint32_t n, uint8_t* streams_out, uint64_t capacity, uint64_t* offsets_out)
	{
		if (!c || !chunk_indices || !offsets_out || (!streams_out && capacity))
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		offsets_out[0] = 0;
		if (n == 0)
			return TSOM_OK;
		tsom_ctx* ctx = c->ctx;
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		std::vector<uint32_t> slots(n);
		for (uint32_t i = 0; i < n; ++i)
		{
			slots[i] = c->SlotOf(Key{ chunk_indices[i * 3], chunk_indices[i * 3 + 1], chunk_indices[i * 3 + 2] });
			if (slots[i] == 0xFFFFFFFFu || !c->hasContent[slots[i]])
				return Fail(TSOM_ERR_INVALID, "chunk is unknown or has no content");
		}
		// sub-batches bound the scratch (65,824 B per chunk); each one: compress at a fixed stride -> sizes -> offsets -> pack -> D2H
		constexpr uint32_t BATCH = 8192;
		const uint32_t stride = tsomk::lz4_stream_stride();
		const uint32_t maxBatch = std::min(n, BATCH);
		TSOM_CUDA(ctx->jobSlot.Ensure(maxBatch));
		TSOM_CUDA(ctx->ioCounts.Ensure(maxBatch));
		TSOM_CUDA(ctx->ioOffsets.Ensure(size_t(maxBatch) + 1));
		TSOM_CUDA(ctx->ioStreams.Ensure(size_t(maxBatch) * stride));
		TSOM_CUDA(ctx->ioPayload.Ensure(size_t(maxBatch) * stride)); // the packed copy
		std::vector<uint32_t> sizes(maxBatch);
		std::vector<unsigned long long> offs(size_t(maxBatch) + 1);
		uint64_t total = 0;
		bool fits = true;
		float kernelMs = 0.f, ms = 0.f;
		for (uint32_t first = 0; first < n; first += BATCH)
		{
			const uint32_t m = std::min(BATCH, n - first);
			TSOM_CUDA(cudaMemcpyAsync(ctx->jobSlot.ptr, slots.data() + first, size_t(m) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
			TSOM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
			tsomk::launch_lz4_compress(c->dBlocks, ctx->jobSlot.ptr, m, ctx->ioStreams.ptr, ctx->ioCounts.ptr, ctx->stream);
			ctx->launches += 1;
			TSOM_CUDA(cudaGetLastError());
			TSOM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
			TSOM_CUDA(cudaMemcpyAsync(sizes.data(), ctx->ioCounts.ptr, size_t(m) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
			TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
			if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess)
				kernelMs += ms;
			offs[0] = 0;
			for (uint32_t i = 0; i < m; ++i)
			{
				offs[i + 1] = offs[i] + sizes[i];
				offsets_out[first + i + 1] = total + offs[i + 1];
			}
			const uint64_t batchBytes = offs[m];
			if (fits && total + batchBytes <= capacity)
			{
				TSOM_CUDA(cudaMemcpyAsync(ctx->ioOffsets.ptr, offs.data(), (size_t(m) + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
				TSOM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
				tsomk::launch_lz4_pack(ctx->ioStreams.ptr, ctx->ioOffsets.ptr, m, ctx->ioPayload.ptr, ctx->stream);
				ctx->launches += 1;
				TSOM_CUDA(cudaGetLastError());
				TSOM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
				TSOM_CUDA(cudaMemcpyAsync(streams_out + total, ctx->ioPayload.ptr, batchBytes, cudaMemcpyDeviceToHost, ctx->stream));
				TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
				if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess)
					kernelMs += ms;
			}
			else
				fits = false; // keep going for the sizes: offsets_out[n] tells the caller what to allocate
			total += batchBytes;
		}
		ctx->timings[TSOM_TIMING_IO_MS] = kernelMs;
		if (!fits)
			return Fail(TSOM_ERR_CAPACITY, "streams_out is too small (offsets_out[n] bytes are needed)");
		return TSOM_OK;
	}

	// the receiving side (ClientSessionHandler.cpp:156-176 after Packets.cpp:197-211): LZ4_decompress_safe + Chunk::Reset, all or nothing
	int tsom_chunks_reset_lz4(tsom_container* c, const int32_t* chunk_indices, uint32_t n, const uint8_t* streams, const uint64_t* offsets)
	{
		if (!c || !chunk_indices || !streams || !offsets)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		if (n == 0)
			return TSOM_OK;
		for (uint32_t i = 0; i < n; ++i)
			if (offsets[i + 1] < offsets[i])
				return Fail(TSOM_ERR_INVALID, "offsets must not decrease");
		tsom_ctx* ctx = c->ctx;
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		const uint64_t base = offsets[0], bytes = offsets[n] - offsets[0];
		std::vector<unsigned long long> offs(size_t(n) + 1);
		for (uint32_t i = 0; i <= n; ++i)
			offs[i] = offsets[i] - base;
		TSOM_CUDA(ctx->ioStreams.Ensure(bytes + 1024)); // the decoder's register window reads up to 512 bytes ahead
		TSOM_CUDA(ctx->ioOffsets.Ensure(size_t(n) + 1));
		TSOM_CUDA(ctx->ioCounts.Ensure(n));
		TSOM_CUDA(ctx->ioPayload.Ensure(size_t(n) * NB * 2));
		TSOM_CUDA(cudaMemcpyAsync(ctx->ioStreams.ptr, streams + base, bytes, cudaMemcpyHostToDevice, ctx->stream));
		TSOM_CUDA(cudaMemcpyAsync(ctx->ioOffsets.ptr, offs.data(), (size_t(n) + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
		TSOM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
		tsomk::launch_lz4_decompress(ctx->ioStreams.ptr, ctx->ioOffsets.ptr, n, reinterpret_cast<uint16_t*>(ctx->ioPayload.ptr), ctx->ioCounts.ptr, ctx->stream);
		ctx->launches += 1;
		TSOM_CUDA(cudaGetLastError());
		TSOM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
		std::vector<uint32_t> status(n);
		TSOM_CUDA(cudaMemcpyAsync(status.data(), ctx->ioCounts.ptr, size_t(n) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
		cudaEventElapsedTime(&ctx->timings[TSOM_TIMING_IO_MS], ctx->ev[0], ctx->ev[1]);
		for (uint32_t i = 0; i < n; ++i)
		{
			if (status[i] == 0xFFFFFFFFu)
				return Fail(TSOM_ERR_INVALID, "failed to decompress chunk (stream " + std::to_string(i) + "); no chunk was reset"); // Packets.cpp:204-205
			if (status[i] != NB * 2)
				return Fail(TSOM_ERR_INVALID, "malformed packet (decompressed size exceeds max packet size: " + std::to_string(status[i]) + " != " + std::to_string(NB * 2) +
				                                  "); no chunk was reset"); // Packets.cpp:207-208
		}
		return ResetCommon(c, chunk_indices, n, ctx->ioPayload.ptr, cudaMemcpyDeviceToDevice);
	}

	int tsom_chunks_get_content(tsom_container* c, const int32_t* chunk_indices, uint32_t n, uint16_t* blocks_host)
	{
		if (!c || !chunk_indices || !blocks_host)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(c->ctx);
		TSOM_CUDA(cudaSetDevice(c->ctx->device));
		for (uint32_t i = 0; i < n; ++i)
		{
			uint32_t slot = c->SlotOf(Key{ chunk_indices[i * 3], chunk_indices[i * 3 + 1], chunk_indices[i * 3 + 2] });
			if (slot == 0xFFFFFFFFu || !c->hasContent[slot])
				return Fail(TSOM_ERR_INVALID, "chunk is unknown or has no content");
			TSOM_CUDA(cudaMemcpyAsync(blocks_host + size_t(i) * NB, c->dBlocks + size_t(slot) * NB, NB * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->ctx->stream));
		}
		TSOM_CUDA(cudaStreamSynchronize(c->ctx->stream));
		return TSOM_OK;
	}

	const void* tsom_chunk_device_ptr(tsom_container* c, const int32_t chunk_index[3])
	{
		if (!c || !chunk_index)
			return nullptr;
		TSOM_LOCK(c->ctx);
		uint32_t slot = c->SlotOf(Key{ chunk_index[0], chunk_index[1], chunk_index[2] });
		if (slot == 0xFFFFFFFFu || !c->hasContent[slot])
			return nullptr;
		return c->dBlocks + size_t(slot) * NB;
	}

	int tsom_chunk_slot(tsom_container* c, const int32_t chunk_index[3], uint32_t* slot_out)
	{
		if (!c || !chunk_index || !slot_out)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(c->ctx);
		uint32_t slot = c->SlotOf(Key{ chunk_index[0], chunk_index[1], chunk_index[2] });
		if (slot == 0xFFFFFFFFu)
			return Fail(TSOM_ERR_INVALID, "unknown chunk");
		*slot_out = slot;
		return TSOM_OK;
	}

	// ---------------------------------------------------------------------------------------------------- generation
	// Everything of Planet::GenerateChunk(s) up to the last launch; GenerateFinish waits for it.  (Split so that one host thread can
	// keep several GPUs busy: tsom_sharded_generate enqueues on every part before it waits on any.)  Caller holds the context lock.
	static int GenerateEnqueue(tsom_container* c, uint32_t seed, const uint32_t chunk_count[3], const tsom_gen_blocks* ids, const int32_t* chunk_indices, uint32_t n)
	{
		tsom_ctx* ctx = c->ctx;
		TSOM_CUDA(cudaSetDevice(ctx->device));
		if (ctx->pendingGenerate)
			TSOM_CUDA(cudaStreamSynchronize(ctx->stream)); // an earlier asynchronous generate may not have consumed its pinned staging (permutation tables) yet

		int maxH[3];
		for (int k = 0; k < 3; ++k)
			maxH[k] = ((int(chunk_count[k]) + 1) / 2) * CS; // Planet.cpp:176-177

		auto& cache = c->genCache;
		if (!cache.Matches(chunk_indices, n, c->topoVersion))
		{
			// bounding box of the requested chunks + the reference's own precondition (Nz::SafeCaster at Planet.cpp:199: depth >= 0)
			tsom_container::GenPlan plan{ { INT32_MAX, INT32_MAX, INT32_MAX }, { INT32_MIN, INT32_MIN, INT32_MIN }, INT32_MAX, INT32_MIN };
			for (uint32_t i = 0; i < n; ++i)
			{
				const int32_t* ci = chunk_indices + size_t(i) * 3;
				for (int k = 0; k < 3; ++k)
				{
					plan.lo[k] = std::min(plan.lo[k], ci[k]);
					plan.hi[k] = std::max(plan.hi[k], ci[k]);
				}
				plan.sumMin = std::min(plan.sumMin, ci[0] + ci[1] + ci[2]);
				plan.sumMax = std::max(plan.sumMax, ci[0] + ci[1] + ci[2]);
				// world coordinate ranges: X from cx, Y(up) from cy, Z from cz; depth crosses Y/Z (Planet.cpp:199-203)
				auto maxAbs = [](int cidx) { int a = cidx * CS - CS / 2, b = a + CS - 1; return std::max(std::abs(a), std::abs(b)); };
				if (maxAbs(ci[0]) > maxH[0] || maxAbs(ci[2]) > maxH[1] || maxAbs(ci[1]) > maxH[2])
					return Fail(TSOM_ERR_UNSUPPORTED, "chunk lies outside +-maxHeight: the reference's depth computation asserts (Nz::SafeCaster, Planet.cpp:199)");
			}
			std::vector<uint32_t> slots(n);
			for (uint32_t i = 0; i < n; ++i)
			{
				int rc = AddChunk(c, Key{ chunk_indices[i * 3], chunk_indices[i * 3 + 1], chunk_indices[i * 3 + 2] }, &slots[i]);
				if (rc != TSOM_OK)
					return rc;
			}
			for (uint32_t i = 0; i < n; ++i)
			{
				if (!c->hasContent[slots[i]])
					TopoChanged(c);
				c->hasContent[slots[i]] = 1;
			}
			TSOM_CUDA(cache.jobs.Ensure(n));
			TSOM_CUDA(cudaMemcpyAsync(cache.jobs.ptr, slots.data(), size_t(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
			TSOM_CUDA(cudaStreamSynchronize(ctx->stream)); // `slots` dies at scope exit
			cache.idx.assign(chunk_indices, chunk_indices + size_t(n) * 3);
			cache.n = n;
			cache.version = c->topoVersion;
			c->genPlan = plan;
			cache.hostSlots = std::move(slots);
			// does the list name every chunk of the container exactly once?
			std::vector<uint8_t> seen(c->highWater, 0);
			cache.coversAll = n == c->slotOf.size();
			for (uint32_t s : cache.hostSlots)
			{
				cache.coversAll = cache.coversAll && !seen[s];
				seen[s] = 1;
			}
		}
		// OnReset -> OnChunkUpdated(DirectionMask_All) (Planet.cpp:34-37) for every generated chunk
		if (cache.coversAll && n == c->slotOf.size() && c->freeSlots.empty())
			std::fill(c->hostInvalid.begin(), c->hostInvalid.begin() + c->highWater, uint8_t(0x80u | 0x3Fu)); // the whole container: one fill
		else
			for (uint32_t s : cache.hostSlots)
				c->hostInvalid[s] = 0x80u | 0x3Fu;
		const tsom_container::GenPlan& plan = c->genPlan;

		// generation reads chunkIdx on the device: make sure new chunks are there
		int rc = SyncTopology(c);
		if (rc != TSOM_OK)
			return rc;
		if ((rc = BeforeBlockWrite(c)) != TSOM_OK)
			return rc;

		// terrain tables over the bounding box, only for the sides from which a chunk of this call can be carved: the pass of a
		// side touches a chunk only if blockDepth - terrainDepth < 32 for some column, and terrainDepth <= maxH / 2 (Planet.cpp:241-247)
		tsomk::TerrainArgs ta{};
		size_t total = 0;
		for (int k = 0; k < 3; ++k)
		{
			ta.tileOrigin[k] = plan.lo[k];
			ta.tileCount[k] = plan.hi[k] - plan.lo[k] + 1;
			ta.maxH[k] = maxH[k];
		}
		const size_t tilesX = size_t(ta.tileCount[1]) * ta.tileCount[2], tilesY = size_t(ta.tileCount[0]) * ta.tileCount[2], tilesZ = size_t(ta.tileCount[0]) * ta.tileCount[1];
		const size_t tiles[6] = { tilesZ, tilesY, tilesZ, tilesX, tilesX, tilesY }; // Direction order: Back Down Front Left Right Up
		uint32_t dirMask = 0;
		{
			// smallest blockDepth a chunk of the box can have per side (Planet.cpp:237,262,287,312,337,362)
			auto nearPlus = [&](int axis) { return maxH[axis] - (plan.hi[axis] * CS - CS / 2) + 1; };
			auto nearMinus = [&](int axis) { return maxH[axis] + (plan.lo[axis] * CS - CS / 2 + CS - 1) + 1; };
			const int minDepth[6] = { nearPlus(2), nearMinus(1), nearMinus(2), nearMinus(0), nearPlus(0), nearPlus(1) }; // Back +Z, Down -Y, Front -Z, Left -X, Right +X, Up +Y
			const int axisOf[6] = { 2, 1, 2, 0, 0, 1 };
			for (int d = 0; d < 6; ++d)
				if (minDepth[d] - maxH[axisOf[d]] / 2 < CS)
					dirMask |= 1u << d;
		}
		for (int d = 0; d < 6; ++d)
		{
			ta.terrainOffset[d] = total;
			ta.tiles[d] = (dirMask >> d) & 1u ? uint32_t(tiles[d]) : 0u;
			total += size_t(ta.tiles[d]) * SLAB;
		}
		TSOM_CUDA(c->terrain.Ensure(std::max<size_t>(total, SLAB)));
		TSOM_CUDA(c->tileRange.Ensure(std::max<size_t>(total / SLAB * 2, 2)));
		ta.terrain = c->terrain.ptr;
		ta.tileRange = c->tileRange.ptr;
		uint8_t* perm = c->hPerm; // pinned, owned by the container: the copy below may still be queued when this call returns
		for (int d = 0; d < 6; ++d)
			PerlinPermutation(seed + unsigned(d), perm + d * 256); // Planet.cpp:179-181
		TSOM_CUDA(cudaMemcpyAsync(c->dPerm, perm, 6 * 256, cudaMemcpyHostToDevice, ctx->stream));
		ta.perm = c->dPerm;
		TSOM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
		tsomk::launch_terrain(ta, ctx->stream);
		ctx->launches += total ? 1 : 0;
		TSOM_CUDA(cudaGetLastError());
		TSOM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));

		tsomk::GenArgs ga{};
		ga.blocks = c->dBlocks;
		ga.xslabs = c->dXslabs;
		ga.chunkIdx = c->dChunkIdx;
		ga.jobSlot = cache.jobs.ptr;
		ga.nJobs = n;
		ga.seed = seed;
		for (int k = 0; k < 3; ++k)
		{
			ga.maxH[k] = maxH[k];
			ga.tileOrigin[k] = ta.tileOrigin[k];
			ga.tileCount[k] = ta.tileCount[k];
		}
		ga.terrain = c->terrain.ptr;
		ga.tileRange = c->tileRange.ptr;
		for (int d = 0; d < 6; ++d)
			ga.terrainOffset[d] = ta.terrainOffset[d];
		ga.dirMask = dirMask;
		ga.dirt = ids->dirt;
		ga.grass = ids->grass;
		ga.stone = ids->stone;
		ga.stoneMossy = ids->stone_mossy;
		ga.snow = ids->snow;
		ga.bStar = ctx->bStar;
		ga.aStar = ctx->aStar;
		ga.aInv = ctx->aInv;
		// chunk summaries the generator may write, from the block table (none without a table: everything stays SUM_UNKNOWN)
		auto opaque = [&](uint16_t id) { return id != 0 && id < ctx->hostFlags.size() && !(ctx->hostFlags[id] & (4u | 2u) /* BF_TRANSPARENT | BF_DOUBLE_SIDED */); };
		ga.summary = c->dSummary;
		ga.fastSummary = (opaque(ids->stone) && opaque(ids->stone_mossy)) ? tsomk::SUM_OPAQUE : tsomk::SUM_UNKNOWN;
		ga.allOpaque = (opaque(ids->dirt) && opaque(ids->grass) && opaque(ids->stone) && opaque(ids->stone_mossy) && opaque(ids->snow)) ? 1 : 0;
		// template chunks: one per value of cx + cy + cz in the request; worth it when many chunks share each of them
		ga.templates = nullptr;
		ga.templateSumMin = plan.sumMin;
		ga.nTemplates = 0;
		const uint32_t nTemplates = uint32_t(plan.sumMax - plan.sumMin + 1);
		TSOM_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
		if (n >= 4u * nTemplates)
		{
			TSOM_CUDA(c->templates.Ensure(size_t(nTemplates) * tsomk::TEMPLATE_STRIDE));
			ga.nTemplates = nTemplates;
			tsomk::launch_templates(ga, c->templates.ptr, ctx->stream);
			ctx->launches += 1;
			TSOM_CUDA(cudaGetLastError());
			ga.templates = c->templates.ptr;
		}
		// one descriptor per job (chunk index, slot, active passes): the dependent loads a CTA would start with, done by n threads at once
		TSOM_CUDA(c->genDesc.Ensure(n));
		ga.jobDesc = c->genDesc.ptr;
		tsomk::launch_gen_classify(ga, c->genDesc.ptr, ctx->stream);
		tsomk::launch_generate(ga, int(n), ctx->stream);
		ctx->launches += 2;
		TSOM_CUDA(cudaGetLastError());
		TSOM_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
		TSOM_CUDA(cudaMemcpyAsync(ctx->hSmall + 8, ctx->dSmall + 8, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		ctx->pendingGenerate = true;
		return TSOM_OK;
	}

	static int GenerateFinish(tsom_container* c)
	{
		tsom_ctx* ctx = c->ctx;
		if (!ctx->pendingGenerate)
			return TSOM_OK;
		ctx->pendingGenerate = false;
		TSOM_CUDA(cudaSetDevice(ctx->device));
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream)); // blocking like Planet::GenerateChunks (WaitForTasks, Planet.cpp:402)
		cudaEventElapsedTime(&ctx->timings[TSOM_TIMING_TERRAIN_MS], ctx->ev[0], ctx->ev[1]);
		cudaEventElapsedTime(&ctx->timings[TSOM_TIMING_GENERATE_MS], ctx->ev[2], ctx->ev[3]);
		return CheckExchangeTimeout(ctx);
	}

	int tsom_planet_generate_chunks(tsom_container* c, uint32_t seed, const uint32_t chunk_count[3], const tsom_gen_blocks* ids, const int32_t* chunk_indices, uint32_t n)
	{
		if (!c || !chunk_count || !ids || (!chunk_indices && n))
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		if (n == 0)
			return TSOM_OK;
		TSOM_LOCK(c->ctx);
		int rc = GenerateEnqueue(c, seed, chunk_count, ids, chunk_indices, n);
		if (rc != TSOM_OK)
			return rc;
		return GenerateFinish(c);
	}

	int tsom_planet_generate_chunks_async(tsom_container* c, uint32_t seed, const uint32_t chunk_count[3], const tsom_gen_blocks* ids, const int32_t* chunk_indices, uint32_t n)
	{
		if (!c || !chunk_count || !ids || (!chunk_indices && n))
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		if (n == 0)
			return TSOM_OK;
		TSOM_LOCK(c->ctx);
		return GenerateEnqueue(c, seed, chunk_count, ids, chunk_indices, n); // completed by the next blocking call on the context
	}

	int tsom_planet_generate(tsom_container* c, uint32_t seed, const uint32_t chunk_count[3], const tsom_gen_blocks* ids)
	{
		if (!c || !chunk_count || !ids)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		const uint64_t n = uint64_t(chunk_count[0]) * chunk_count[1] * chunk_count[2];
		if (n == 0 || n > c->capacity)
			return Fail(TSOM_ERR_CAPACITY, "chunk_count exceeds the container capacity");
		// Planet.cpp:387-393: z, y, x loops; indices = (x, y, z) - chunkCount / 2 (the list is kept: a regenerate names the same chunks)
		std::vector<int32_t>& idx = c->planetIdx;
		{
			TSOM_LOCK(c->ctx);
			if (idx.size() != n * 3 || std::memcmp(c->planetIdxCount, chunk_count, sizeof(c->planetIdxCount)) != 0)
			{
				idx.clear();
				idx.reserve(n * 3);
				for (int cz = 0; cz < int(chunk_count[2]); ++cz)
					for (int cy = 0; cy < int(chunk_count[1]); ++cy)
						for (int cx = 0; cx < int(chunk_count[0]); ++cx)
						{
							idx.push_back(cx - int(chunk_count[0] / 2));
							idx.push_back(cy - int(chunk_count[1] / 2));
							idx.push_back(cz - int(chunk_count[2] / 2));
						}
				std::memcpy(c->planetIdxCount, chunk_count, sizeof(c->planetIdxCount));
			}
		}
		return tsom_planet_generate_chunks(c, seed, chunk_count, ids, idx.data(), uint32_t(n));
	}

	int tsom_planet_generate_platform(tsom_container* c, int up_direction, const int32_t platform_center[3], const tsom_platform_blocks* ids)
	{
		if (!c || !platform_center || !ids || up_direction < 0 || up_direction > 5)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		tsom_ctx* ctx = c->ctx;
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		// dense slot grid over the bounding box of existing chunks
		int lo[3] = { INT32_MAX, INT32_MAX, INT32_MAX }, hi[3] = { INT32_MIN, INT32_MIN, INT32_MIN };
		for (auto& [k, s] : c->slotOf)
		{
			const int v[3] = { k.x, k.y, k.z };
			for (int a = 0; a < 3; ++a)
			{
				lo[a] = std::min(lo[a], v[a]);
				hi[a] = std::max(hi[a], v[a]);
			}
		}
		if (c->slotOf.empty())
			return TSOM_OK;
		tsomk::PlatformArgs pa{};
		size_t cells = 1;
		for (int a = 0; a < 3; ++a)
		{
			pa.gridOrigin[a] = lo[a];
			pa.gridDim[a] = hi[a] - lo[a] + 1;
			cells *= size_t(pa.gridDim[a]);
		}
		if (cells > (size_t(1) << 28))
			return Fail(TSOM_ERR_UNSUPPORTED, "chunk bounding box too large for the platform slot grid");
		if (c->slotGridVersion != c->topoVersion)
		{
			std::vector<int> grid(cells, -1);
			for (auto& [k, s] : c->slotOf)
				if (c->hasContent[s])
					grid[(size_t(k.z - lo[2]) * pa.gridDim[1] + size_t(k.y - lo[1])) * pa.gridDim[0] + size_t(k.x - lo[0])] = int(s);
			TSOM_CUDA(c->slotGrid.Ensure(cells));
			TSOM_CUDA(cudaMemcpyAsync(c->slotGrid.ptr, grid.data(), cells * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
			TSOM_CUDA(cudaStreamSynchronize(ctx->stream)); // `grid` dies at scope exit
			c->slotGridVersion = c->topoVersion;
		}
		{
			int rc = BeforeBlockWrite(c);
			if (rc != TSOM_OK)
				return rc;
		}
		pa.blocks = c->dBlocks;
		pa.xslabs = c->dXslabs;
		pa.invalidMask = c->dInvalid;
		pa.summary = c->dSummary;
		pa.slotGrid = c->slotGrid.ptr;
		pa.upDirection = up_direction;
		for (int a = 0; a < 3; ++a)
			pa.center[a] = platform_center[a];
		pa.copper = ids->copper_block;
		pa.stoneBricks = ids->stone_bricks;
		pa.planks = ids->planks;
		tsomk::launch_platform(pa, ctx->stream);
		ctx->launches += 1;
		TSOM_CUDA(cudaGetLastError());
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
		return TSOM_OK;
	}

	// ---------------------------------------------------------------------------------------------------- edits / dirty
	int tsom_apply_edits(tsom_container* c, const tsom_edit* edits, uint32_t n)
	{
		if (!c || (!edits && n))
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		if (n == 0)
			return TSOM_OK;
		tsom_ctx* ctx = c->ctx;
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		TSOM_CUDA(ctx->EnsurePinned(size_t(n) * sizeof(tsomk::EditDev)));
		auto* staged = static_cast<tsomk::EditDev*>(ctx->hPinned);
		for (uint32_t i = 0; i < n; ++i)
		{
			const tsom_edit& e = edits[i];
			if (e.x >= CS || e.y >= CS || e.z >= CS)
				return Fail(TSOM_ERR_INVALID, "edit coordinates out of range");
			uint32_t slot = c->SlotOf(Key{ e.chunk[0], e.chunk[1], e.chunk[2] });
			if (slot != 0xFFFFFFFFu && !c->hasContent[slot])
				slot = 0xFFFFFFFFu;
			staged[i].slot = slot;
			staged[i].packed = uint32_t(e.x) | (uint32_t(e.y) << 5) | (uint32_t(e.z) << 10) | (uint32_t(e.block) << 16);
		}
		TSOM_CUDA(ctx->edits.Ensure(n));
		TSOM_CUDA(cudaMemcpyAsync(ctx->edits.ptr, staged, size_t(n) * sizeof(tsomk::EditDev), cudaMemcpyHostToDevice, ctx->stream));
		// last-writer-wins per block: a hash table of the edited blocks (capacity >= 2n, power of two) records the last edit index
		uint32_t cap = 1024;
		while (cap < 2u * n)
			cap <<= 1;
		TSOM_CUDA(ctx->editKeys.Ensure(cap));
		TSOM_CUDA(ctx->editVals.Ensure(cap));
		TSOM_CUDA(cudaMemsetAsync(ctx->editKeys.ptr, 0xFF, size_t(cap) * sizeof(unsigned long long), ctx->stream));
		TSOM_CUDA(cudaMemsetAsync(ctx->editVals.ptr, 0, size_t(cap) * sizeof(uint32_t), ctx->stream));
		{
			int rc = BeforeBlockWrite(c);
			if (rc != TSOM_OK)
				return rc;
		}
		tsomk::launch_apply_edits(c->dBlocks, c->dXslabs, c->dInvalid, c->dSummary, ctx->edits.ptr, n, ctx->editKeys.ptr, ctx->editVals.ptr, cap, c->kind == TSOM_CONTAINER_SHIP ? 1 : 0,
		                          ctx->stream);
		ctx->launches += 2;
		TSOM_CUDA(cudaGetLastError());
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream)); // the pinned staging buffer is reused by the next call
		return TSOM_OK;
	}

	int tsom_collect_dirty(tsom_container* c, int32_t* chunk_indices_out, uint32_t cap, uint32_t* n_out, int clear)
	{
		if (!c || !n_out)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		tsom_ctx* ctx = c->ctx;
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		const uint32_t n = c->highWater;
		const uint8_t* mask = nullptr; // the device's invalidation bytes (edits), read where the copy lands: pinned memory
		if (n)
		{
			TSOM_CUDA(ctx->EnsurePinned(n));
			TSOM_CUDA(cudaMemcpyAsync(ctx->hPinned, c->dInvalid, n, cudaMemcpyDeviceToHost, ctx->stream));
			TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
			mask = static_cast<const uint8_t*>(ctx->hPinned);
		}
		const bool wantClear = clear != 0;
		// ChunkEntities::Update's set: every invalidated chunk with content plus the neighbours its mask names (ChunkEntities.cpp:73-112)
		std::vector<uint8_t>& seen = c->dirtySeen;
		seen.assign(n, 0);
		c->EnsureHostGrid();
		uint32_t count = 0;
		for (uint32_t s = 0; s < n; ++s)
		{
			const uint8_t m = mask[s] | c->hostInvalid[s];
			if (!(m & 0x80u) || !c->exists[s] || !c->hasContent[s])
				continue;
			count += seen[s] ? 0u : 1u;
			seen[s] = 1;
			if (!(m & 0x3Fu))
				continue;
			const Key k = c->idxOf[s];
			for (int d = 0; d < 6; ++d)
			{
				if (!(m & (1u << d)))
					continue;
				const uint32_t ns = c->SlotOf(Key{ k.x + kChunkDirOffset[d][0], k.y + kChunkDirOffset[d][1], k.z + kChunkDirOffset[d][2] });
				if (ns == 0xFFFFFFFFu || !c->hasContent[ns] || seen[ns])
					continue;
				seen[ns] = 1;
				++count;
			}
		}
		*n_out = count;
		// a list that does not fit the caller's buffer must not be lost
		if (wantClear && chunk_indices_out && count > cap)
			return Fail(TSOM_ERR_CAPACITY, "chunk_indices_out holds " + std::to_string(cap) + " chunks, " + std::to_string(count) + " are invalidated; nothing was cleared");
		if (chunk_indices_out)
		{
			// sorted by (x, y, z): the slots of the dense grid are kept in that order; without a grid (scattered chunks) sort the set
			uint32_t i = 0;
			auto put = [&](const Key& k) {
				if (i < cap)
				{
					chunk_indices_out[i * 3 + 0] = k.x;
					chunk_indices_out[i * 3 + 1] = k.y;
					chunk_indices_out[i * 3 + 2] = k.z;
				}
				++i;
			};
			if (c->hostGridUsable)
			{
				for (uint32_t sl : c->gridOrderXYZ)
					if (sl < n && seen[sl])
						put(c->idxOf[sl]);
			}
			else
			{
				std::vector<Key> dirty;
				dirty.reserve(count);
				for (uint32_t s = 0; s < n; ++s)
					if (seen[s])
						dirty.push_back(c->idxOf[s]);
				std::sort(dirty.begin(), dirty.end());
				for (const Key& k : dirty)
					put(k);
			}
		}
		if (wantClear)
		{
			std::fill(c->hostInvalid.begin(), c->hostInvalid.begin() + n, uint8_t(0));
			if (n)
				TSOM_CUDA(cudaMemsetAsync(c->dInvalid, 0, n, ctx->stream));
		}
		return TSOM_OK;
	}

	// ---------------------------------------------------------------------------------------------------- meshing
	static cudaError_t LaunchFused(tsom_ctx* ctx, tsomk::MeshArgs& a, uint32_t n, uint32_t flags)
	{
		a.vertices = ctx->dVertices;
		a.indices = ctx->dIndices;
		a.collider = ctx->dCollider;
		a.arenaFaces = ArenaFaces(ctx, flags);
		a.jobCounter = ctx->dSmall + 1;
		a.status = ctx->status.ptr;
		a.firstFaceOut = ctx->firstFace.ptr;
		a.totalFacesOut = reinterpret_cast<unsigned long long*>(ctx->dSmall + 4);
		if (flags & TSOM_MESH_ORDERED)
		{
			cudaError_t e = cudaMemsetAsync(ctx->status.ptr, 0, size_t(n) * sizeof(unsigned long long), ctx->stream);
			if (e != cudaSuccess)
				return e;
		}
		tsomk::launch_mesh_fused(a, int(std::min<uint32_t>(n, uint32_t(2 * ctx->numSMs))), ctx->stream);
		ctx->launches += 1;
		return cudaGetLastError();
	}

	// Chunk::BuildMesh for a batch up to the launch and the queued read-back; MeshFinish waits and fills the views.  Caller holds the lock.
	static int MeshEnqueue(tsom_container* c, const int32_t* chunk_indices, uint32_t n, const tsom_mesh_params* params, bool wantViews, uint32_t* nOut)
	{
		tsom_ctx* ctx = c->ctx;
		const uint32_t flags = params->flags;
		if (!(flags & (TSOM_MESH_RENDER | TSOM_MESH_COLLIDER)))
			return Fail(TSOM_ERR_INVALID, "flags must request TSOM_MESH_RENDER and/or TSOM_MESH_COLLIDER");
		if ((flags & TSOM_MESH_TANGENTS) && !(flags & TSOM_MESH_RENDER))
			return Fail(TSOM_ERR_INVALID, "TSOM_MESH_TANGENTS needs TSOM_MESH_RENDER");
		if (!ctx->dFlags)
			return Fail(TSOM_ERR_INVALID, "no block table: call tsom_set_block_table first");
		TSOM_CUDA(cudaSetDevice(ctx->device));
		ctx->pendingMesh.active = false;

		// ---- job list: slot | has-content flag, one entry per chunk (cached on the device while the same list is meshed again)
		auto& cache = c->meshCache;
		if (chunk_indices)
		{
			if (!cache.Matches(chunk_indices, n, c->topoVersion))
			{
				std::vector<uint32_t> jobs(n);
				for (uint32_t i = 0; i < n; ++i)
				{
					const uint32_t slot = c->SlotOf(Key{ chunk_indices[i * 3], chunk_indices[i * 3 + 1], chunk_indices[i * 3 + 2] });
					if (slot == 0xFFFFFFFFu)
						return Fail(TSOM_ERR_INVALID, "tsom_mesh: unknown chunk");
					jobs[i] = slot | (c->hasContent[slot] ? tsomk::JOB_HAS_CONTENT : 0u);
				}
				cache.n = 0;
				if (n)
				{
					TSOM_CUDA(cache.jobs.Ensure(n));
					TSOM_CUDA(cudaMemcpyAsync(cache.jobs.ptr, jobs.data(), size_t(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
					TSOM_CUDA(cudaStreamSynchronize(ctx->stream)); // `jobs` dies at scope exit
					cache.idx.assign(chunk_indices, chunk_indices + size_t(n) * 3);
					cache.n = n;
					cache.version = c->topoVersion;
				}
			}
		}
		else
		{
			// every chunk with content, container order
			std::vector<uint32_t> jobs;
			for (uint32_t s = 0; s < c->highWater; ++s)
				if (c->exists[s] && c->hasContent[s])
					jobs.push_back(s | tsomk::JOB_HAS_CONTENT);
			if (wantViews && n != jobs.size())
				return Fail(TSOM_ERR_INVALID, "tsom_mesh: with chunk_indices == NULL, n must equal the number of chunks with content");
			n = uint32_t(jobs.size());
			cache.n = 0; // not keyed by an index list
			if (n)
			{
				TSOM_CUDA(cache.jobs.Ensure(n));
				TSOM_CUDA(cudaMemcpyAsync(cache.jobs.ptr, jobs.data(), size_t(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
				TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
			}
		}
		if (nOut)
			*nOut = n;
		ctx->totalFaces = 0;
		if (n == 0)
			return TSOM_OK;

		int rc = SyncTopology(c);
		if (rc != TSOM_OK)
			return rc;
		if (c->ghostsStale)
			return Fail(TSOM_ERR_INVALID, "tsom_mesh: a neighbour chunk on another GPU has never been pulled (the topology changed): call tsom_halo_pull first");
		if (c->summaryTableEpoch != ctx->tableEpoch)
		{
			// the block table changed since the summaries were written: opaque may no longer mean opaque
			TSOM_CUDA(cudaMemsetAsync(c->dSummary, 0, c->capacity, ctx->stream));
			c->summaryTableEpoch = ctx->tableEpoch;
		}

		TSOM_CUDA(ctx->jobSlot.Ensure(n));
		TSOM_CUDA(ctx->faceCount.Ensure(n));
		TSOM_CUDA(ctx->firstFace.Ensure(n));
		TSOM_CUDA(ctx->status.Ensure(n));
		TSOM_CUDA(ctx->EnsurePinned(std::max<size_t>(size_t(n) * (sizeof(uint32_t) + sizeof(unsigned long long)), 64)));
		if (params->gravity_centers)
		{
			TSOM_CUDA(ctx->gravity.Ensure(size_t(n) * 3));
			TSOM_CUDA(cudaMemcpyAsync(ctx->gravity.ptr, params->gravity_centers, size_t(n) * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
		}
		TSOM_CUDA(cudaMemsetAsync(ctx->dSmall, 0, 8 * sizeof(uint32_t), ctx->stream));
		if (ArenaFaces(ctx, flags) == 0 && (rc = EnsureArenas(ctx, 1u << 16, flags)) != TSOM_OK)
			return rc;

		// ---- chunks that provably have no face (summaries of the producers) are marked and skip their 64 KiB stage
		TSOM_CUDA(cudaEventRecord(ctx->ev[8], ctx->stream));
		// ... and, for the unordered mode, deals the tickets heavy chunks first (JOB_BUCKETS lists of n entries, counters in dSmall[9..12])
		const bool dealt = !(flags & TSOM_MESH_ORDERED);
		if (dealt)
		{
			TSOM_CUDA(ctx->jobOrder.Ensure(size_t(n) * tsomk::JOB_BUCKETS));
			TSOM_CUDA(cudaMemsetAsync(ctx->dSmall + 9, 0, tsomk::JOB_BUCKETS * sizeof(uint32_t), ctx->stream));
		}
		tsomk::launch_job_classify(cache.jobs.ptr, ctx->jobSlot.ptr, n, c->dSummary, c->dNbr, c->dLastFaces, dealt ? ctx->jobOrder.ptr : nullptr, ctx->dSmall + 9, ctx->stream);
		ctx->launches += 1;
		TSOM_CUDA(cudaGetLastError());
		TSOM_CUDA(cudaEventRecord(ctx->ev[9], ctx->stream));

		tsomk::MeshArgs& a = ctx->pendingMesh.args;
		a = tsomk::MeshArgs{};
		a.blocks = c->dBlocks;
		a.halo = c->dHalo;
		a.chunkIdx = c->dChunkIdx;
		a.jobSlot = ctx->jobSlot.ptr;
		a.gravity = params->gravity_centers ? ctx->gravity.ptr : nullptr;
		a.nJobs = n;
		a.blkFlags = ctx->dFlags;
		a.blkTex = ctx->dTex;
		a.texFits8 = ctx->texFits8 ? 1u : 0u;
		a.nTypes = ctx->nTypes;
		a.faceCount = ctx->faceCount.ptr;
		a.order = dealt ? ctx->jobOrder.ptr : nullptr;
		a.bucketCount = ctx->dSmall + 9;
		a.lastFaces = c->dLastFaces;
		a.overflow = ctx->dSmall + 3;
		a.flags = flags;
		a.tile = c->tile;
		a.deformRadius = params->deform_radius;
		std::memcpy(a.center, c->center, sizeof(a.center));
		std::memcpy(a.quat, ctx->quat, sizeof(a.quat));

		// table-driven path: FlatChunk corners and a power-of-two block size make every DrawFace intermediate exact
		int expo = 0;
		const bool pow2Tile = c->tile > 0.f && std::isfinite(c->tile) && std::frexp(c->tile, &expo) == 0.5f && expo > -100 && expo < 100;
		a.faceTable = nullptr;
		if (!(flags & TSOM_MESH_GENERIC_MATH) && pow2Tile)
		{
			if (ctx->faceTableTile != c->tile)
			{
				tsomk::launch_face_table(ctx->dFaceTable, c->tile, a, ctx->stream);
				ctx->launches += 1;
				ctx->faceTableTile = c->tile;
			}
			a.faceTable = ctx->dFaceTable;
		}

		// single pass: the count and the scan live inside mesh_fused_kernel
		TSOM_CUDA(cudaEventRecord(ctx->ev[4], ctx->stream));
		TSOM_CUDA(LaunchFused(ctx, a, n, flags));
		TSOM_CUDA(cudaEventRecord(ctx->ev[5], ctx->stream));

		// queue the read-back of the totals (+ views)
		TSOM_CUDA(cudaMemcpyAsync(ctx->hSmall, ctx->dSmall, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		auto* hFirst = static_cast<unsigned long long*>(ctx->hPinned);
		auto* hCount = reinterpret_cast<uint32_t*>(hFirst + n);
		if (wantViews)
		{
			TSOM_CUDA(cudaMemcpyAsync(hFirst, ctx->firstFace.ptr, size_t(n) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
			TSOM_CUDA(cudaMemcpyAsync(hCount, ctx->faceCount.ptr, size_t(n) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		}
		ctx->pendingMesh.active = true;
		ctx->pendingMesh.n = n;
		ctx->pendingMesh.flags = flags;
		ctx->pendingMesh.wantViews = wantViews;
		return TSOM_OK;
	}

	static int MeshFinish(tsom_container* c, tsom_mesh_view* views_out)
	{
		tsom_ctx* ctx = c->ctx;
		if (!ctx->pendingMesh.active)
			return TSOM_OK;
		ctx->pendingMesh.active = false;
		const uint32_t n = ctx->pendingMesh.n, flags = ctx->pendingMesh.flags;
		TSOM_CUDA(cudaSetDevice(ctx->device));
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream));

		cudaEventElapsedTime(&ctx->timings[TSOM_TIMING_CLASSIFY_MS], ctx->ev[8], ctx->ev[9]);
		cudaEventElapsedTime(&ctx->timings[TSOM_TIMING_MESH_EMIT_MS], ctx->ev[4], ctx->ev[5]);
		if (ctx->pendingGenerate)
		{
			// a tsom_planet_generate_chunks_async in front of this call has finished with it
			ctx->pendingGenerate = false;
			cudaEventElapsedTime(&ctx->timings[TSOM_TIMING_TERRAIN_MS], ctx->ev[0], ctx->ev[1]);
			cudaEventElapsedTime(&ctx->timings[TSOM_TIMING_GENERATE_MS], ctx->ev[2], ctx->ev[3]);
		}
		ctx->timings[TSOM_TIMING_MESH_REEMIT] = 0.f;
		if (ctx->pendingHalo)
		{
			cudaEventElapsedTime(&ctx->timings[TSOM_TIMING_HALO_MS], ctx->ev[6], ctx->ev[7]);
			ctx->pendingHalo = false;
		}
		int rc = CheckExchangeTimeout(ctx);
		if (rc != TSOM_OK)
			return rc;

		auto* hFirst = static_cast<unsigned long long*>(ctx->hPinned);
		auto* hCount = reinterpret_cast<uint32_t*>(hFirst + n);
		unsigned long long total;
		std::memcpy(&total, ctx->hSmall + 4, sizeof(total));
		if (ctx->hSmall[3] != 0)
		{
			ctx->timings[TSOM_TIMING_MESH_REEMIT] = 1.f;
			// arenas were too small: the first launch only counted.  Grow to the exact need and run again; the face counts are
			// the same, the first faces are not (unordered placement is first come, first served), so read the views again.
			if ((rc = EnsureArenas(ctx, total, flags)) != TSOM_OK)
				return rc;
			TSOM_CUDA(cudaMemsetAsync(ctx->dSmall + 1, 0, sizeof(uint32_t), ctx->stream));
			TSOM_CUDA(cudaMemsetAsync(ctx->dSmall + 3, 0, 3 * sizeof(uint32_t), ctx->stream)); // overflow flag + the 64-bit arena cursor / total
			TSOM_CUDA(LaunchFused(ctx, ctx->pendingMesh.args, n, flags));
			if (ctx->pendingMesh.wantViews)
				TSOM_CUDA(cudaMemcpyAsync(hFirst, ctx->firstFace.ptr, size_t(n) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
			TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
		}
		ctx->totalFaces = total;
		ctx->lastMeshFlags = flags;
		if (views_out && ctx->pendingMesh.wantViews)
		{
			for (uint32_t i = 0; i < n; ++i)
			{
				views_out[i].first_face = hFirst[i];
				views_out[i].face_count = hCount[i];
				views_out[i].reserved = 0;
			}
		}
		return TSOM_OK;
	}

	int tsom_mesh(tsom_container* c, const int32_t* chunk_indices, uint32_t n, const tsom_mesh_params* params, tsom_mesh_view* views_out)
	{
		if (!c || !params)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(c->ctx);
		int rc = MeshEnqueue(c, chunk_indices, n, params, views_out != nullptr, nullptr);
		if (rc != TSOM_OK)
			return rc;
		return MeshFinish(c, views_out);
	}

	int tsom_mesh_arenas(tsom_ctx* ctx, const void** vertices, const void** indices, const void** collider_positions, uint64_t* total_faces)
	{
		if (!ctx)
			return Fail(TSOM_ERR_INVALID, "ctx is null");
		TSOM_LOCK(ctx);
		if (vertices)
			*vertices = ctx->dVertices;
		if (indices)
			*indices = ctx->dIndices;
		if (collider_positions)
			*collider_positions = ctx->dCollider;
		if (total_faces)
			*total_faces = ctx->totalFaces;
		return TSOM_OK;
	}

	int tsom_mesh_copy_to_host(tsom_ctx* ctx, uint64_t first_face, uint64_t face_count, void* vertices, uint32_t* indices, float* collider_positions)
	{
		if (!ctx)
			return Fail(TSOM_ERR_INVALID, "ctx is null");
		TSOM_LOCK(ctx);
		if (first_face + face_count > ctx->totalFaces)
			return Fail(TSOM_ERR_INVALID, "face range exceeds the last mesh result");
		if (face_count == 0)
			return TSOM_OK;
		TSOM_CUDA(cudaSetDevice(ctx->device));
		if (vertices)
		{
			if (!ctx->dVertices || !(ctx->lastMeshFlags & TSOM_MESH_RENDER))
				return Fail(TSOM_ERR_INVALID, "last tsom_mesh call did not produce render vertices");
			TSOM_CUDA(cudaMemcpyAsync(vertices, reinterpret_cast<const char*>(ctx->dVertices) + first_face * TSOM_FACE_VERTEX_BYTES, face_count * TSOM_FACE_VERTEX_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
		}
		if (indices)
			TSOM_CUDA(cudaMemcpyAsync(indices, reinterpret_cast<const char*>(ctx->dIndices) + first_face * TSOM_FACE_INDEX_BYTES, face_count * TSOM_FACE_INDEX_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
		if (collider_positions)
		{
			if (!ctx->dCollider || !(ctx->lastMeshFlags & TSOM_MESH_COLLIDER))
				return Fail(TSOM_ERR_INVALID, "last tsom_mesh call did not produce collider positions");
			TSOM_CUDA(cudaMemcpyAsync(collider_positions, reinterpret_cast<const char*>(ctx->dCollider) + first_face * TSOM_FACE_COLLIDER_BYTES, face_count * TSOM_FACE_COLLIDER_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
		}
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
		return TSOM_OK;
	}

	int tsom_voxel_corners(tsom_container* c, const int32_t* chunk_indices, const uint32_t* local_indices, uint32_t n, const tsom_mesh_params* params, float* corners_out)
	{
		if (!c || !params || ((!chunk_indices || !local_indices || !corners_out) && n))
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		if (n == 0)
			return TSOM_OK;
		tsom_ctx* ctx = c->ctx;
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		// staging: n x int4 blocks | n x 3 float centres (pinned), device: status buffer as int4s, gravity buffer as centres + output
		TSOM_CUDA(ctx->EnsurePinned(size_t(n) * (sizeof(int4) + 3 * sizeof(float))));
		int4* hBlocks = static_cast<int4*>(ctx->hPinned);
		float* hCenters = reinterpret_cast<float*>(hBlocks + n);
		for (uint32_t i = 0; i < n; ++i)
		{
			const uint32_t* l = local_indices + size_t(i) * 3;
			if (l[0] >= uint32_t(CS) || l[1] >= uint32_t(CS) || l[2] >= uint32_t(CS))
				return Fail(TSOM_ERR_INVALID, "block indices out of range");
			hBlocks[i] = make_int4(int(l[0]), int(l[1]), int(l[2]), 0);
			const int32_t* ci = chunk_indices + size_t(i) * 3;
			for (int a = 0; a < 3; ++a) // DeformedChunk's deformation centre: the caller's, or container centre - GetChunkOffset (ChunkContainer.inl:53-56)
				hCenters[i * 3 + a] = params->gravity_centers ? params->gravity_centers[size_t(i) * 3 + a] : c->center[a] - float(ci[a]) * float(CS) * c->tile;
		}
		TSOM_CUDA(ctx->status.Ensure(size_t(n) * 2));         // n x int4
		TSOM_CUDA(ctx->gravity.Ensure(size_t(n) * (3 + 24))); // centres, then the corners
		TSOM_CUDA(cudaMemcpyAsync(ctx->status.ptr, hBlocks, size_t(n) * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
		TSOM_CUDA(cudaMemcpyAsync(ctx->gravity.ptr, hCenters, size_t(n) * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
		tsomk::launch_voxel_corners(reinterpret_cast<const int4*>(ctx->status.ptr), ctx->gravity.ptr, n, c->tile, params->deform_radius, (params->flags & TSOM_MESH_DEFORMED) != 0,
		                            ctx->gravity.ptr + size_t(n) * 3, ctx->stream);
		ctx->launches += 1;
		TSOM_CUDA(cudaGetLastError());
		TSOM_CUDA(cudaMemcpyAsync(corners_out, ctx->gravity.ptr + size_t(n) * 3, size_t(n) * 24 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
		return TSOM_OK;
	}

	int tsom_mesh_aabb(tsom_ctx* ctx, const tsom_mesh_view* views, uint32_t n, float* aabb_out)
	{
		if (!ctx || (!views && n) || (!aabb_out && n))
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		if (n == 0)
			return TSOM_OK;
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		const bool render = (ctx->lastMeshFlags & TSOM_MESH_RENDER) != 0;
		if (!render && !(ctx->lastMeshFlags & TSOM_MESH_COLLIDER))
			return Fail(TSOM_ERR_INVALID, "no mesh result on this context");
		std::vector<unsigned long long> first(n);
		std::vector<uint32_t> count(n);
		for (uint32_t i = 0; i < n; ++i)
		{
			if (views[i].first_face + views[i].face_count > ctx->totalFaces)
				return Fail(TSOM_ERR_INVALID, "a view exceeds the last mesh result");
			first[i] = views[i].first_face;
			count[i] = views[i].face_count;
		}
		TSOM_CUDA(ctx->firstFace.Ensure(n));
		TSOM_CUDA(ctx->faceCount.Ensure(n));
		TSOM_CUDA(ctx->gravity.Ensure(size_t(n) * 6));
		TSOM_CUDA(cudaMemcpyAsync(ctx->firstFace.ptr, first.data(), size_t(n) * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
		TSOM_CUDA(cudaMemcpyAsync(ctx->faceCount.ptr, count.data(), size_t(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
		tsomk::launch_mesh_aabb(render ? ctx->dVertices : nullptr, reinterpret_cast<const float*>(ctx->dCollider), ctx->firstFace.ptr, ctx->faceCount.ptr, n, ctx->gravity.ptr, ctx->stream);
		ctx->launches += 1;
		TSOM_CUDA(cudaGetLastError());
		TSOM_CUDA(cudaMemcpyAsync(aabb_out, ctx->gravity.ptr, size_t(n) * 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
		return TSOM_OK;
	}

	int tsom_mesh_checksums(tsom_ctx* ctx, const tsom_mesh_view* views, uint32_t n, uint64_t* out)
	{
		if (!ctx || (!views && n) || (!out && n))
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		if (n == 0)
			return TSOM_OK;
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		const bool render = (ctx->lastMeshFlags & TSOM_MESH_RENDER) != 0;
		if (!render && !(ctx->lastMeshFlags & TSOM_MESH_COLLIDER))
			return Fail(TSOM_ERR_INVALID, "no mesh result on this context");
		TSOM_CUDA(ctx->EnsurePinned(size_t(n) * (sizeof(unsigned long long) + sizeof(uint32_t))));
		auto* hFirst = static_cast<unsigned long long*>(ctx->hPinned);
		auto* hCount = reinterpret_cast<uint32_t*>(hFirst + n);
		for (uint32_t i = 0; i < n; ++i)
		{
			if (views[i].first_face + views[i].face_count > ctx->totalFaces)
				return Fail(TSOM_ERR_INVALID, "a view exceeds the last mesh result");
			hFirst[i] = views[i].first_face;
			hCount[i] = views[i].face_count;
		}
		TSOM_CUDA(ctx->firstFace.Ensure(n));
		TSOM_CUDA(ctx->faceCount.Ensure(n));
		TSOM_CUDA(ctx->status.Ensure(size_t(n) * 2));
		TSOM_CUDA(cudaMemcpyAsync(ctx->firstFace.ptr, hFirst, size_t(n) * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
		TSOM_CUDA(cudaMemcpyAsync(ctx->faceCount.ptr, hCount, size_t(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
		tsomk::launch_mesh_checksum(reinterpret_cast<const uint32_t*>(render ? static_cast<const void*>(ctx->dVertices) : static_cast<const void*>(ctx->dCollider)), render ? 48u : 12u,
		                            reinterpret_cast<const uint32_t*>(ctx->dIndices), ctx->firstFace.ptr, ctx->faceCount.ptr, n, ctx->status.ptr, ctx->stream);
		ctx->launches += 1;
		TSOM_CUDA(cudaGetLastError());
		TSOM_CUDA(cudaMemcpyAsync(out, ctx->status.ptr, size_t(n) * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
		return TSOM_OK;
	}

	int tsom_mesh_indices_for_faces(uint64_t face_count, uint32_t* indices_out)
	{
		if (!indices_out && face_count)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		if (face_count > 0xFFFFFFFFull / 4ull)
			return Fail(TSOM_ERR_INVALID, "a chunk's vertex ordinals do not fit 32 bits");
		auto fill = [indices_out](uint64_t f0, uint64_t f1) {
			for (uint64_t f = f0; f < f1; ++f)
			{
				uint32_t* o = indices_out + f * 6;
				const uint32_t b = uint32_t(f) * 4u;
				o[0] = b;
				o[1] = b + 2u;
				o[2] = b + 1u;
				o[3] = b + 1u;
				o[4] = b + 2u;
				o[5] = b + 3u;
			}
		};
		const unsigned workers = face_count < (1u << 20) ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
		if (workers == 1)
		{
			fill(0, face_count);
			return TSOM_OK;
		}
		std::vector<std::thread> pool;
		for (unsigned w = 0; w < workers; ++w)
			pool.emplace_back(fill, face_count * w / workers, face_count * (w + 1) / workers);
		for (auto& t : pool)
			t.join();
		return TSOM_OK;
	}

	// FlatChunk::BuildCollider for a batch: box_colliders_kernel (one warp per chunk) into a packed scratch, sizes to the host, offsets
	// back, box_pack_kernel into the dense arena.  Sub-batches bound the scratch (64 KiB per chunk); one synchronisation per sub-batch.
	int tsom_box_colliders(tsom_container* c, const int32_t* chunk_indices, uint32_t n, tsom_box_view* views_out, uint64_t* total_boxes_out)
	{
		if (!c || (!chunk_indices && n))
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		tsom_ctx* ctx = c->ctx;
		TSOM_LOCK(ctx);
		if (!ctx->dFlags)
			return Fail(TSOM_ERR_INVALID, "no block table");
		ctx->totalBoxes = 0;
		if (total_boxes_out)
			*total_boxes_out = 0;
		if (n == 0)
			return TSOM_OK;
		TSOM_CUDA(cudaSetDevice(ctx->device));
		std::vector<uint32_t> jobs(n);
		for (uint32_t i = 0; i < n; ++i)
		{
			const uint32_t slot = c->SlotOf(Key{ chunk_indices[i * 3], chunk_indices[i * 3 + 1], chunk_indices[i * 3 + 2] });
			if (slot == 0xFFFFFFFFu)
				return Fail(TSOM_ERR_INVALID, "tsom_box_colliders: unknown chunk");
			jobs[i] = slot | (c->hasContent[slot] ? tsomk::JOB_HAS_CONTENT : 0u);
		}
		constexpr uint32_t BATCH = 8192;
		const uint32_t maxBatch = std::min(n, BATCH);
		TSOM_CUDA(ctx->jobSlot.Ensure(maxBatch));
		TSOM_CUDA(ctx->boxCounts.Ensure(maxBatch));
		TSOM_CUDA(ctx->boxFirst.Ensure(maxBatch));
		TSOM_CUDA(ctx->boxScratch.Ensure(size_t(maxBatch) * tsomk::BOX_SCRATCH_STRIDE));
		TSOM_CUDA(ctx->EnsurePinned(size_t(maxBatch) * (sizeof(uint32_t) + sizeof(unsigned long long))));
		auto* hFirst = static_cast<unsigned long long*>(ctx->hPinned);
		auto* hCount = reinterpret_cast<uint32_t*>(hFirst + maxBatch);
		uint64_t total = 0;
		float kernelMs = 0.f, ms = 0.f;
		for (uint32_t first = 0; first < n; first += BATCH)
		{
			const uint32_t m = std::min(BATCH, n - first);
			TSOM_CUDA(cudaMemcpyAsync(ctx->jobSlot.ptr, jobs.data() + first, size_t(m) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
			TSOM_CUDA(cudaEventRecord(ctx->ev[8], ctx->stream));
			tsomk::launch_box_colliders(c->dBlocks, ctx->jobSlot.ptr, m, ctx->dFlags, ctx->nTypes, ctx->boxScratch.ptr, ctx->boxCounts.ptr, ctx->stream);
			ctx->launches += 1;
			TSOM_CUDA(cudaGetLastError());
			TSOM_CUDA(cudaEventRecord(ctx->ev[9], ctx->stream));
			TSOM_CUDA(cudaMemcpyAsync(hCount, ctx->boxCounts.ptr, size_t(m) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
			TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
			if (cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]) == cudaSuccess)
				kernelMs += ms;
			const uint64_t batchFirst = total;
			for (uint32_t i = 0; i < m; ++i)
			{
				hFirst[i] = total;
				if (views_out)
				{
					views_out[first + i].first_box = total;
					views_out[first + i].box_count = hCount[i];
					views_out[first + i].reserved = 0;
				}
				total += hCount[i];
			}
			if (total > batchFirst)
			{
				if (total > ctx->capBoxes)
				{
					// grow the arena, keeping the boxes of the earlier sub-batches
					const uint64_t want = std::max<uint64_t>(total, 2 * ctx->capBoxes);
					float* p = nullptr;
					cudaError_t e = cudaMalloc(&p, want * 6 * sizeof(float));
					if (e != cudaSuccess)
						return Fail(TSOM_ERR_NOMEM, std::string("box arena cudaMalloc: ") + cudaGetErrorString(e));
					if (ctx->dBoxes && batchFirst)
						cudaMemcpy(p, ctx->dBoxes, batchFirst * 6 * sizeof(float), cudaMemcpyDeviceToDevice);
					cudaFree(ctx->dBoxes);
					ctx->dBoxes = p;
					ctx->capBoxes = want;
				}
				TSOM_CUDA(cudaMemcpyAsync(ctx->boxFirst.ptr, hFirst, size_t(m) * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
				TSOM_CUDA(cudaEventRecord(ctx->ev[8], ctx->stream));
				tsomk::launch_box_pack(ctx->boxScratch.ptr, ctx->boxCounts.ptr, ctx->boxFirst.ptr, m, ctx->dBoxes, ctx->stream);
				ctx->launches += 1;
				TSOM_CUDA(cudaGetLastError());
				TSOM_CUDA(cudaEventRecord(ctx->ev[9], ctx->stream));
				TSOM_CUDA(cudaStreamSynchronize(ctx->stream)); // the pinned offsets are rewritten by the next sub-batch
				if (cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]) == cudaSuccess)
					kernelMs += ms;
			}
		}
		ctx->timings[TSOM_TIMING_COLLIDER_MS] = kernelMs;
		ctx->totalBoxes = total;
		if (total_boxes_out)
			*total_boxes_out = total;
		return TSOM_OK;
	}

	int tsom_box_colliders_copy_to_host(tsom_ctx* ctx, uint64_t first_box, uint64_t box_count, float* boxes_out)
	{
		if (!ctx || (!boxes_out && box_count))
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(ctx);
		if (first_box + box_count > ctx->totalBoxes)
			return Fail(TSOM_ERR_INVALID, "box range exceeds the last tsom_box_colliders result");
		if (box_count == 0)
			return TSOM_OK;
		TSOM_CUDA(cudaSetDevice(ctx->device));
		TSOM_CUDA(cudaMemcpyAsync(boxes_out, ctx->dBoxes + first_box * 6, box_count * 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
		TSOM_CUDA(cudaStreamSynchronize(ctx->stream));
		return TSOM_OK;
	}

	// one chunk, the reference's call shape (FlatChunk::BuildCollider): the batched path with n = 1
	int tsom_box_collider(tsom_container* c, const int32_t chunk_index[3], float* boxes_out, uint32_t cap, uint32_t* n_out)
	{
		if (!c || !chunk_index || !n_out)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(c->ctx);
		const uint32_t slot = c->SlotOf(Key{ chunk_index[0], chunk_index[1], chunk_index[2] });
		if (slot == 0xFFFFFFFFu || !c->hasContent[slot])
			return Fail(TSOM_ERR_INVALID, "chunk is unknown or has no content");
		tsom_box_view v{};
		int rc = tsom_box_colliders(c, chunk_index, 1, &v, nullptr);
		if (rc != TSOM_OK)
			return rc;
		*n_out = v.box_count;
		if (boxes_out && v.box_count)
			return tsom_box_colliders_copy_to_host(c->ctx, v.first_box, std::min<uint64_t>(v.box_count, cap), boxes_out);
		return TSOM_OK;
	}

	// Ship::Generate (Ship.cpp:115-152): chunk (0,0,0) reset to Empty, then the hull written block by block.  The final content is a
	// pure function of (small, ids); it is composed here and handed to the device in one copy.  What the reference leaves behind in
	// m_invalidatedChunks is kept too: Chunk::Reset() without a callback fires no OnReset (Chunk.inl:97-107), so the chunk is
	// invalidated by its UpdateBlock calls alone — the chunk itself plus the OR of their neighbour masks (Ship.cpp:36-54).
	int tsom_ship_generate(tsom_container* c, int small, uint16_t hull_block, uint16_t forcefield_block)
	{
		if (!c)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(c->ctx);
		std::vector<uint16_t> blocks(NB, 0);
		const unsigned boxSize = small ? 6 : 12, height = small ? 4 : 6;
		const unsigned sx = CS / 2 - boxSize / 2, sy = CS / 2 - boxSize / 2, sz = CS / 2 - height / 2;
		uint8_t mask = 0x80u;
		const bool ship = c->kind == TSOM_CONTAINER_SHIP;
		for (unsigned z = 0; z < height; ++z)
			for (unsigned y = 0; y < boxSize; ++y)
				for (unsigned x = 0; x < boxSize; ++x)
				{
					if (x != 0 && x != boxSize - 1 && y != 0 && y != boxSize - 1 && z != 0 && z != height - 1)
						continue;
					const bool door = x == 0 && y == boxSize / 2 && z > 0 && z < height - 1;
					const unsigned px = sx + x, py = sy + y, pz = sz + z;
					blocks[(size_t(pz) * CS + py) * CS + px] = door ? forcefield_block : hull_block;
					if (px == 0) mask |= 1u << TSOM_LEFT; else if (px == unsigned(CS) - 1) mask |= 1u << TSOM_RIGHT;
					if (py == 0) mask |= 1u << TSOM_FRONT; else if (py == unsigned(CS) - 1) mask |= 1u << TSOM_BACK;
					if (pz == 0) mask |= 1u << (ship ? TSOM_UP : TSOM_DOWN); else if (pz == unsigned(CS) - 1) mask |= 1u << (ship ? TSOM_DOWN : TSOM_UP);
				}
		const int32_t origin[3] = { 0, 0, 0 };
		int rc = tsom_chunks_reset(c, origin, 1, blocks.data());
		if (rc != TSOM_OK)
			return rc;
		const uint32_t slot = c->SlotOf(Key{ 0, 0, 0 });
		c->hostInvalid[slot] = mask;
		return TSOM_OK;
	}

	// ---------------------------------------------------------------------------------------------------- multi-GPU
	int tsom_container_add_remote_chunks(tsom_container* c, const int32_t* chunk_indices, const uint32_t* remote_slots, const int32_t* ranks, uint32_t n)
	{
		if (!c || !chunk_indices || !remote_slots || !ranks)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(c->ctx);
		for (uint32_t i = 0; i < n; ++i)
		{
			Key k{ chunk_indices[i * 3], chunk_indices[i * 3 + 1], chunk_indices[i * 3 + 2] };
			if (c->slotOf.count(k))
				return Fail(TSOM_ERR_INVALID, "remote chunk is already a local chunk");
			c->remotes[k] = tsom_container::Remote{ ranks[i], remote_slots[i] };
		}
		TopoChanged(c);
		return TSOM_OK;
	}

	int tsom_ipc_export(tsom_container* c, tsom_ipc_handle* out)
	{
		if (!c || !out)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
		static_assert(sizeof(tsom_ipc_handle) >= 3 * 64 + 4, "tsom_ipc_handle size");
		TSOM_LOCK(c->ctx);
		TSOM_CUDA(cudaSetDevice(c->ctx->device));
		cudaIpcMemHandle_t h0, h1, h2;
		TSOM_CUDA(cudaIpcGetMemHandle(&h0, c->dBlocks));
		TSOM_CUDA(cudaIpcGetMemHandle(&h1, c->dXslabs));
		TSOM_CUDA(cudaIpcGetMemHandle(&h2, c->dSync));
		std::memset(out, 0, sizeof(*out));
		std::memcpy(out->bytes, &h0, 64);
		std::memcpy(out->bytes + 64, &h1, 64);
		std::memcpy(out->bytes + 128, &h2, 64);
		std::memcpy(out->bytes + 192, &c->capacity, 4);
		return TSOM_OK;
	}

	static void ClosePeer(tsom_container::Peer& p)
	{
		if (!p.ipc)
			return;
		cudaIpcCloseMemHandle(const_cast<uint16_t*>(p.blocks));
		cudaIpcCloseMemHandle(const_cast<uint16_t*>(p.xslabs));
		cudaIpcCloseMemHandle(const_cast<uint32_t*>(p.sync));
		p = tsom_container::Peer{};
	}

	int tsom_ipc_attach(tsom_container* c, int rank, const tsom_ipc_handle* handle)
	{
		if (!c || !handle)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(c->ctx);
		TSOM_CUDA(cudaSetDevice(c->ctx->device));
		auto old = c->peers.find(rank);
		if (old != c->peers.end())
		{
			TSOM_CUDA(cudaStreamSynchronize(c->ctx->stream));
			ClosePeer(old->second); // re-attaching a rank: release the previous mappings first
		}
		cudaIpcMemHandle_t h0, h1, h2;
		std::memcpy(&h0, handle->bytes, 64);
		std::memcpy(&h1, handle->bytes + 64, 64);
		std::memcpy(&h2, handle->bytes + 128, 64);
		void *p0 = nullptr, *p1 = nullptr, *p2 = nullptr;
		TSOM_CUDA(cudaIpcOpenMemHandle(&p0, h0, cudaIpcMemLazyEnablePeerAccess));
		TSOM_CUDA(cudaIpcOpenMemHandle(&p1, h1, cudaIpcMemLazyEnablePeerAccess));
		TSOM_CUDA(cudaIpcOpenMemHandle(&p2, h2, cudaIpcMemLazyEnablePeerAccess));
		tsom_container::Peer p;
		p.blocks = static_cast<const uint16_t*>(p0);
		p.xslabs = static_cast<const uint16_t*>(p1);
		p.sync = static_cast<const uint32_t*>(p2);
		p.ipc = true;
		c->peers[rank] = p;
		c->peerFlagsDirty = true;
		TopoChanged(c);
		return TSOM_OK;
	}

	int tsom_peer_attach_local(tsom_container* c, int rank, tsom_container* peer)
	{
		if (!c || !peer)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(c->ctx);
		if (peer->ctx->device != c->ctx->device)
		{
			TSOM_CUDA(cudaSetDevice(c->ctx->device));
			cudaError_t e = cudaDeviceEnablePeerAccess(peer->ctx->device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
				return Fail(TSOM_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
			cudaGetLastError();
		}
		auto old = c->peers.find(rank);
		if (old != c->peers.end())
			ClosePeer(old->second);
		tsom_container::Peer p;
		p.blocks = peer->dBlocks;
		p.xslabs = peer->dXslabs;
		p.sync = peer->dSync;
		p.ipc = false;
		p.local = peer;
		c->peers[rank] = p;
		c->peerFlagsDirty = true;
		TopoChanged(c);
		return TSOM_OK;
	}

	// "my blocks are final for this round": raise the blocks-final flag to round pulledEpoch + 1 behind everything queued so far
	static int HaloPublish(tsom_container* c)
	{
		tsom_ctx* ctx = c->ctx;
		TSOM_CUDA(cudaSetDevice(ctx->device));
		c->epoch = c->pulledEpoch + 1u;
		tsomk::launch_flag_publish(c->dSync + 0, c->epoch, ctx->stream);
		ctx->launches += 1;
		TSOM_CUDA(cudaGetLastError());
		TSOM_CUDA(cudaEventRecord(c->evGen, ctx->stream));
		c->evGenEpoch.store(c->epoch);
		c->published = true;
		c->dirtySincePublish = false;
		return TSOM_OK;
	}

	int tsom_halo_set_timeout(tsom_container* c, uint32_t milliseconds)
	{
		if (!c || milliseconds == 0)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(c->ctx);
		c->exchangeTimeoutMs = milliseconds;
		return TSOM_OK;
	}

	int tsom_halo_publish(tsom_container* c)
	{
		if (!c)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		TSOM_LOCK(c->ctx);
		return HaloPublish(c);
	}

	int tsom_halo_pull(tsom_container* c)
	{
		if (!c)
			return Fail(TSOM_ERR_INVALID, "bad arguments");
		tsom_ctx* ctx = c->ctx;
		TSOM_LOCK(ctx);
		TSOM_CUDA(cudaSetDevice(ctx->device));
		int rc = SyncTopology(c);
		if (rc != TSOM_OK)
			return rc;
		if (!c->published || c->dirtySincePublish)
			if ((rc = HaloPublish(c)) != TSOM_OK)
				return rc;
		TSOM_CUDA(cudaEventRecord(ctx->ev[6], ctx->stream));
		// wait, on the device, until every peer's blocks are final for this round
		if ((rc = WaitForPeers(c, 0, c->epoch)) != TSOM_OK)
			return rc;
		if (c->nGhost)
		{
			tsomk::launch_halo_pull(c->ghostDescs.ptr, c->nGhost, c->ghostSlabs.ptr, ctx->stream);
			ctx->launches += 1;
			TSOM_CUDA(cudaGetLastError());
		}
		tsomk::launch_flag_publish(c->dSync + 1, c->epoch, ctx->stream);
		ctx->launches += 1;
		TSOM_CUDA(cudaGetLastError());
		TSOM_CUDA(cudaEventRecord(c->evPull, ctx->stream));
		c->evPullEpoch.store(c->epoch);
		TSOM_CUDA(cudaEventRecord(ctx->ev[7], ctx->stream));
		ctx->pendingHalo = true;
		c->pulledEpoch = c->epoch;
		c->published = false;
		c->needWriteWait = true;
		c->ghostsStale = false;
		return TSOM_OK; // nothing waited for on the host: the next blocking call on this context reports a time-out
	}

}

#include "tsom_sharded.inl"
