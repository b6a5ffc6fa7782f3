// tsom_kernels.h — host-visible declarations of the sm_100a kernels' argument blocks and launchers.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace tsomk
{
	enum : uint32_t { MESH_F_RENDER = 1u, MESH_F_COLLIDER = 2u, MESH_F_DEFORMED = 4u, MESH_F_GENERIC = 8u, MESH_F_ORDERED = 16u, MESH_F_TANGENTS = 32u };

	// per-chunk content summary (one byte per slot), written by the producers at no extra pass over the ids: lets the mesher skip the
	// 64 KiB stage of chunks that provably have no face
	enum : uint8_t
	{
		SUM_UNKNOWN = 0, // anything (every reset / edit puts a chunk back here)
		SUM_EMPTY = 1,   // every block is Empty
		SUM_OPAQUE = 2   // every block is non-Empty, not transparent and not double sided
	};
	// job entry of the mesher: slot | JOB_OPAQUE (summary SUM_OPAQUE: the row masks are known without reading the ids) | JOB_SKIP (no face,
	// proven from the summaries) | JOB_HAS_CONTENT
	constexpr uint32_t JOB_SLOT_MASK = 0x1FFFFFFFu, JOB_OPAQUE = 0x20000000u, JOB_SKIP = 0x40000000u, JOB_HAS_CONTENT = 0x80000000u;

	// per-(faceUp, slot, twin) face attributes of a FlatChunk with a power-of-two block size (see mesh_kernels.cu)
	struct FaceTable
	{
		float4 uv[6][6][2][2]; // (u0,v0,u1,v1), (u2,v2,u3,v3)
		float4 normal[6];      // per emission slot
		uint32_t texDir[6][6]; // [faceUp][slot]
	};

	struct MeshArgs
	{
		// container
		const uint16_t* blocks;          // [capacity][32768]
		const unsigned long long* halo;  // [capacity][6] encoded slab pointers (0 = no neighbour)
		const int4* chunkIdx;            // [capacity] (cx, cy, cz, hasContent)
		// jobs
		const uint32_t* jobSlot;         // [nJobs] slot | JOB_SKIP | JOB_HAS_CONTENT
		const float* gravity;            // [nJobs][3] or nullptr
		uint32_t nJobs;
		// block table
		const uint8_t* blkFlags;
		const uint32_t* blkTex;          // [nTypes][6]
		uint32_t nTypes;
		uint32_t texFits8;               // every texture slice of the table is below 256
		// results per job + the chained scan over the jobs (decoupled look-back)
		uint32_t* faceCount;                   // [nJobs]
		uint32_t* jobCounter;                  // work-stealing ticket, zeroed before each launch
		unsigned long long* status;            // [nJobs] (flag << 62) | faces, zeroed before each launch
		unsigned long long* firstFaceOut;      // [nJobs]
		unsigned long long* totalFacesOut;     // total faces; also the arena cursor of the unordered mode (zeroed before each launch)
		// arenas
		float4* vertices;
		uint2* indices;
		float4* collider;
		unsigned long long arenaFaces;
		uint32_t* overflow;
		// parameters
		uint32_t flags;
		float tile;
		float deformRadius;
		float center[3];  // ChunkContainer::GetCenter()
		float quat[6][4]; // RotationBetween(s_dirNormals[d], Up) as (w, x, y, z)
		const FaceTable* faceTable; // non-null: table-driven flat path
		// unordered mode: tickets are dealt heavy chunks first — order[b * nJobs + i] = i-th job of bucket b (by the face count of the
		// chunk's previous mesh), bucketCount[b] jobs in bucket b (job_classify_kernel); nullptr: ticket == job
		const uint32_t* order;
		const uint32_t* bucketCount;
		uint32_t* lastFaces; // [capacity] face count of the last mesh of every chunk slot
	};
	constexpr int JOB_BUCKETS = 4; // >= 4096 faces, >= 512, the rest, no face at all (skipped jobs)

	void launch_face_table(FaceTable* table, float tile, const MeshArgs& a, cudaStream_t st);

	cudaError_t mesh_kernels_init();
	size_t mesh_fused_smem_bytes();
	void launch_mesh_fused(const MeshArgs& a, int grid, cudaStream_t st);
	// dst[j] = src[j] | JOB_SKIP when the chunk of job j provably has no face: it is SUM_EMPTY, or SUM_OPAQUE with six existing
	// SUM_OPAQUE neighbours (nbrSlot: [capacity][6] slot of the face-adjacent chunk with content in this container, -1 otherwise)
	void launch_job_classify(const uint32_t* src, uint32_t* dst, uint32_t n, const uint8_t* summary, const int32_t* nbrSlot, const uint32_t* lastFaces, uint32_t* order,
	                         uint32_t* bucketCount, cudaStream_t st);

	// ---- generation ----
	struct GenArgs
	{
		uint16_t* blocks;        // [capacity][32768]
		uint16_t* xslabs;        // [capacity][2][1024]
		const int4* chunkIdx;    // [capacity]
		const uint32_t* jobSlot; // [nJobs]
		const int4* jobDesc;     // [nJobs] (cx, cy, cz, slot | active passes << 26), written by gen_classify_kernel
		uint32_t nJobs;
		uint32_t seed;
		int maxH[3];             // ((chunkCount + 1) / 2) * 32 per axis
		int tileOrigin[3];       // smallest chunk index covered by the terrain tables, per axis
		int tileCount[3];        // tiles per axis
		const int16_t* terrain;  // 6 tables of 32x32 tiles, see gen_kernels.cu
		const int16_t* tileRange; // (min, max) terrainDepth per tile, indexed by global tile number
		size_t terrainOffset[6]; // element offset of each direction's table
		uint16_t dirt, grass, stone, stoneMossy, snow;
		uint32_t bStar, aStar;   // integer form of `bernoulli(0.9)` over two minstd draws
		uint32_t aInv;           // 48271^-1 mod (2^31 - 1)
		uint32_t dirMask;        // bit d: the terrain tables of Direction d exist (else no chunk of this call can be carved from that side)
		uint8_t* summary;        // [capacity] SUM_*
		uint8_t fastSummary;     // summary of a chunk made of stone / stone_mossy only: SUM_OPAQUE when both are opaque types
		uint8_t allOpaque;       // 1 when dirt, grass, stone, stone_mossy and snow are all opaque types: no Empty block => SUM_OPAQUE
		// chunks in which every block draws and no pass is active depend on cx + cy + cz only (chunkSeed, Planet.cpp:165): they are
		// copies of template chunk (cx + cy + cz - templateSumMin), [nTemplates][32768 ids + 2 x 1024 x-slab ids]; nullptr = compute
		const uint16_t* templates;
		int templateSumMin;
		uint32_t nTemplates;
	};
	constexpr size_t TEMPLATE_STRIDE = 32768 + 2 * 1024; // ids per template chunk

	struct TerrainArgs
	{
		int16_t* terrain;
		int16_t* tileRange;
		size_t terrainOffset[6];
		uint32_t tiles[6];   // tiles per direction; 0 = direction not needed by this call
		int tileOrigin[3];
		int tileCount[3];
		int maxH[3];
		const uint8_t* perm; // [6][256] siv::PerlinNoise permutations, reseed(seed + dir)
	};

	void launch_terrain(const TerrainArgs& a, cudaStream_t st);
	void launch_generate(const GenArgs& a, int grid, cudaStream_t st);
	void launch_templates(const GenArgs& a, uint16_t* templates, cudaStream_t st);
	void launch_gen_classify(const GenArgs& a, int4* desc, cudaStream_t st);

	// ---- chunk maintenance ----
	void launch_build_xslabs(const uint16_t* blocks, uint16_t* xslabs, uint8_t* summary, const uint32_t* slots, uint32_t n, cudaStream_t st);

	struct EditDev
	{
		uint32_t slot; // 0xFFFFFFFF = skip
		uint32_t packed; // x | y<<5 | z<<10 | block<<16
	};
	// keys / vals: open-addressing table (capacity `cap`, a power of two >= 2n; keys preset to ~0, vals to 0)
	// shipMask: Ship::AddChunk's neighbour rule (z == 0 -> Up, z == 31 -> Down, Ship.cpp:48-51) instead of Planet's (Planet.cpp:50-53)
	void launch_apply_edits(uint16_t* blocks, uint16_t* xslabs, uint8_t* invalidMask, uint8_t* summary, const EditDev* edits, uint32_t n, unsigned long long* keys, uint32_t* vals,
	                        uint32_t cap, int shipMask, cudaStream_t st);

	struct PlatformArgs
	{
		uint16_t* blocks;
		uint16_t* xslabs;
		uint8_t* invalidMask;
		uint8_t* summary;
		const int* slotGrid; // dense grid of slots (-1 absent / no content) covering [gridOrigin, gridOrigin + gridDim)
		int gridOrigin[3];
		int gridDim[3];
		int upDirection;
		int center[3];
		uint16_t copper, stoneBricks, planks;
	};
	void launch_platform(const PlatformArgs& a, cudaStream_t st);

	// FlatChunk greedy box collider for n chunks (one CTA each): the boxes of job j go to scratch[j][0 .. count[j]) (one packed 30-bit word per box,
	// BOX_SCRATCH_STRIDE words per job); launch_box_pack then moves them to their final places first[j] in the dense arena
	constexpr uint32_t BOX_SCRATCH_STRIDE = 16384; // a 32^3 checkerboard: every other cell is its own box
	void launch_box_colliders(const uint16_t* blocks, const uint32_t* jobSlot, uint32_t n, const uint8_t* blkFlags, uint32_t nTypes, uint32_t* scratch, uint32_t* counts, cudaStream_t st);
	void launch_box_pack(const uint32_t* scratch, const uint32_t* counts, const unsigned long long* first, uint32_t n, float* arena, cudaStream_t st);

	// ---- save format: palette + per-block palette indices (payload stride 65536 bytes per chunk) ----
	void launch_palettize(const uint16_t* blocks, const uint32_t* slots, uint32_t n, uint16_t* palette, uint32_t stride, uint32_t* counts, uint8_t* payload, int bigEndian, cudaStream_t st);
	void launch_depalettize(uint16_t* blocks, const uint32_t* slots, uint32_t n, const uint16_t* palette, uint32_t stride, const uint32_t* counts, const uint8_t* payload, int bigEndian,
	                        uint32_t* errFlag, cudaStream_t st);

	// ---- Chunk::ComputeVoxelCorners as a query (block picking): n blocks -> n x 8 corners in Nz::BoxCorner order ----
	void launch_voxel_corners(const int4* blocks, const float* centers, uint32_t n, float tile, float radius, int deformed, float* out, cudaStream_t st);

	// ---- Nz::StaticMesh::GenerateAABB: min / max of each chunk's vertex positions (render arena, or the collider arena if vertices == null) ----
	void launch_mesh_aabb(const float4* vertices, const float* collider, const unsigned long long* firstFace, const uint32_t* faceCount, uint32_t n, float* out, cudaStream_t st);

	// ---- per-view checksums of a mesh result (vertex words or collider positions; indices), see mesh_checksum_kernel ----
	void launch_mesh_checksum(const uint32_t* vertexWords, uint32_t wordsPerFace, const uint32_t* indices, const unsigned long long* firstFace, const uint32_t* faceCount, uint32_t n,
	                          unsigned long long* out, cudaStream_t st);

	// ---- wire format: one LZ4 block per chunk (Packets::ChunkReset), lz4_kernels.cu ----
	cudaError_t lz4_kernels_init();
	uint32_t lz4_stream_stride(); // bytes reserved per chunk in the scratch of launch_lz4_compress (>= LZ4_compressBound(65536))
	void launch_lz4_compress(const uint16_t* blocks, const uint32_t* slots, uint32_t n, uint8_t* streams, uint32_t* sizes, cudaStream_t st);
	void launch_lz4_pack(const uint8_t* streams, const unsigned long long* offsets, uint32_t n, uint8_t* packed, cudaStream_t st);
	// out: n x 65536 bytes, written only for streams that decode to exactly 65536 bytes; status[i] = decoded size or 0xFFFFFFFF
	void launch_lz4_decompress(const uint8_t* packed, const unsigned long long* offsets, uint32_t n, uint16_t* out, uint32_t* status, cudaStream_t st);

	// ---- multi-GPU halo pull ----
	struct GhostDesc
	{
		const uint16_t* src; // peer address of the slab's first id
		uint32_t contig;     // 1 = contiguous 2 KiB, 0 = 32 rows of 64 B at stride 2 KiB
		uint32_t ghost;      // destination ghost slab index
	};
	void launch_halo_pull(const GhostDesc* descs, uint32_t n, uint16_t* ghostSlabs, cudaStream_t st);
	// epoch flags of the device-side exchange ordering (gen_kernels.cu)
	void launch_flag_publish(uint32_t* flag, uint32_t value, cudaStream_t st);
	void launch_flag_wait(const uint32_t* const* flags, uint32_t n, uint32_t value, unsigned long long timeoutNs, uint32_t* err, cudaStream_t st);
}
