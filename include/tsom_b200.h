/*
 * tsom_b200.h — C ABI of libtsom_b200.so: the B200-native (sm_100a) replacement for ThisSpaceOfMine's CPU voxel chunk
 * pipeline (chunk generation, face-culled meshing, collider triangles, block edits + dirty propagation).
 *
 * The reference has no FFI for this path — it sits behind C++ virtuals.  Every entry point below names
 * the reference interface it replaces (paths relative to the reference repository root).  Plain pointers and sizes only;
 * no C++ or torch types cross this boundary.  All functions return TSOM_OK (0) or an error code; tsom_last_error() gives
 * the message of the calling thread's last failure.  There is NO CPU fallback: without a CUDA device every compute
 * entry point fails with TSOM_ERR_CUDA.
 *
 * Conventions (reference: include/CommonLib/Chunk.inl:25-32, ChunkContainer.inl:14-17, Direction.hpp:18-28):
 *   - a chunk is 32x32x32 uint16 block ids, x fastest: idx = 32*(32*z + y) + x; local z is "up" (world Y);
 *   - chunk indices are world-axis int32 triples (cx, cy=up, cz);
 *   - Direction order: Back(0) Down(1) Front(2) Left(3) Right(4) Up(5); direction masks use bit (1 << Direction);
 *   - a render vertex is 48 bytes {position, normal, uvw, tangent} (include/ClientLib/ClientChunkEntities.hpp:27-33),
 *     tangent is written as zero exactly like Chunk::BuildMesh leaves it; indices are uint32, chunk-relative,
 *     4 vertices + 6 indices per face in the reference's emission order (src/CommonLib/Chunk.cpp:95-101,174-249).
 */
#ifndef TSOM_B200_H
#define TSOM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSOM_CHUNK_SIZE 32
#define TSOM_CHUNK_BLOCKS 32768 /* 32^3 */
#define TSOM_VERTEX_STRIDE 48   /* bytes */
#define TSOM_FACE_VERTEX_BYTES 192
#define TSOM_FACE_INDEX_BYTES 24
#define TSOM_FACE_COLLIDER_BYTES 48 /* 4 positions x 12 B */

enum tsom_status
{
	TSOM_OK = 0,
	TSOM_ERR_INVALID = 1,     /* bad argument / unknown chunk */
	TSOM_ERR_CUDA = 2,        /* CUDA runtime failure or no device */
	TSOM_ERR_UNSUPPORTED = 3, /* configuration on which the reference itself asserts (e.g. negative depth, Planet.cpp:199) */
	TSOM_ERR_NOMEM = 4,
	TSOM_ERR_CAPACITY = 5,    /* container or caller buffer too small */
	TSOM_ERR_INTERNAL = 6     /* a self-check of the library failed (tsom_debug_check_guards) */
};

enum tsom_direction { TSOM_BACK = 0, TSOM_DOWN = 1, TSOM_FRONT = 2, TSOM_LEFT = 3, TSOM_RIGHT = 4, TSOM_UP = 5 };

/* tsom_mesh flags */
enum
{
	TSOM_MESH_RENDER = 1,   /* 48-byte vertices + indices: ClientChunkEntities::BuildMesh (src/ClientLib/ClientChunkEntities.cpp:141-176) */
	TSOM_MESH_COLLIDER = 2, /* positions only + the same indices: DeformedChunk::BuildCollider (src/CommonLib/DeformedChunk.cpp:15-36) */
	TSOM_MESH_DEFORMED = 4, /* corner provider DeformedChunk::ComputeVoxelCorners (DeformedChunk.cpp:115-154) instead of FlatChunk (FlatChunk.cpp:48-54) */
	TSOM_MESH_GENERIC_MATH = 8, /* diagnostics: evaluate DrawFace's float math per face even where the exact table-driven flat path applies */
	TSOM_MESH_ORDERED = 16 /* place the chunks' face ranges in the arenas in job order (first_face[i] = sum of face_count[< i]) instead of
	                          first come, first served; costs a cross-chunk dependency (decoupled look-back) in the kernel */
};

typedef struct tsom_ctx tsom_ctx;             /* one per process and GPU: stream, block table, mesh arenas */
typedef struct tsom_container tsom_container; /* a ChunkContainer (Planet / Ship / isolated batch) resident in HBM */

/* BlockLibrary::BlockData (include/CommonLib/BlockLibrary.hpp:52-60), tex in Direction order */
typedef struct tsom_block_desc
{
	uint32_t tex[6];
	uint8_t has_collisions;
	uint8_t is_double_sided;
	uint8_t is_transparent;
	uint8_t reserved;
} tsom_block_desc;

/* the five ids Planet::GenerateChunk looks up by name (src/CommonLib/Planet.cpp:170-174) */
typedef struct tsom_gen_blocks
{
	uint16_t dirt, grass, stone, stone_mossy, snow;
} tsom_gen_blocks;

/* the three ids Planet::GeneratePlatform looks up (Planet.cpp:421-422,461) */
typedef struct tsom_platform_blocks
{
	uint16_t copper_block, stone_bricks, planks;
} tsom_platform_blocks;

/* one Chunk::UpdateBlock call (src/CommonLib/Chunk.cpp:348-366; wire form include/CommonLib/Protocol/Packets.hpp:283-300) */
typedef struct tsom_edit
{
	int32_t chunk[3];
	uint8_t x, y, z, reserved;
	uint16_t block;
	uint16_t reserved2;
} tsom_edit;

/* where one chunk's faces sit in the arenas of the last tsom_mesh call */
typedef struct tsom_mesh_view
{
	uint64_t first_face; /* face ordinal in the arenas; vertices at first_face*192 B, indices at first_face*24 B */
	uint32_t face_count; /* 0 <=> the reference would have returned a null mesh / collider */
	uint32_t reserved;
} tsom_mesh_view;

/* per-call meshing parameters */
typedef struct tsom_mesh_params
{
	uint32_t flags;               /* TSOM_MESH_* */
	float deform_radius;          /* DeformedChunk::m_deformationRadius (only with TSOM_MESH_DEFORMED) */
	const float* gravity_centers; /* optional host array n x 3: Chunk::BuildMesh `center` argument per chunk; NULL = container centre
	                                 minus chunk offset, as ClientChunkEntities.cpp:160 passes it.  Also used as the deformation centre. */
} tsom_mesh_params;

typedef struct tsom_ipc_handle
{
	unsigned char bytes[136]; /* 2 x cudaIpcMemHandle_t (blocks, x-slabs) + capacity */
} tsom_ipc_handle;

/* ---- library ---- */
const char* tsom_last_error(void);
const char* tsom_version(void);

/* ---- context ---- */
int tsom_create(int cuda_device, tsom_ctx** out);
void tsom_destroy(tsom_ctx* ctx);
/* the cudaStream_t all work of this context is enqueued on (for CUDA-event timing by the caller) */
void* tsom_stream(tsom_ctx* ctx);
int tsom_synchronize(tsom_ctx* ctx);
/* Threading (the reference calls BuildMesh / BuildCollider from many Nz::TaskScheduler workers, ClientChunkEntities.cpp:207-241): every entry
 * point takes the context's recursive lock, so calls on one context and its containers serialise and may come from any thread.  Results that
 * live in context-owned memory — the arenas of tsom_mesh — are replaced by the next tsom_mesh on the same context: a thread that needs
 * "mesh, then copy out" as one unit brackets the calls with tsom_ctx_lock / tsom_ctx_unlock (same thread, may nest).  Independent work
 * streams want one context each. */
int tsom_ctx_lock(tsom_ctx* ctx);
int tsom_ctx_unlock(tsom_ctx* ctx);
/* replaces BlockLibrary::GetBlockData (BlockLibrary.hpp:31) on the device: uploads the table once */
int tsom_set_block_table(tsom_ctx* ctx, const tsom_block_desc* blocks, uint32_t n);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t tsom_launch_count(tsom_ctx* ctx);
/* device-side durations (CUDA events on the context's stream) of the kernels of the most recent generate / mesh call */
enum tsom_timing
{
	TSOM_TIMING_TERRAIN_MS = 0,   /* terrain_kernel (the six directions in one launch) */
	TSOM_TIMING_GENERATE_MS = 1,  /* generate_kernel */
	TSOM_TIMING_MESH_COUNT_MS = 2, /* 0 on the single-pass path (count and scan live inside mesh_fused_kernel) */
	TSOM_TIMING_MESH_SCAN_MS = 3,
	TSOM_TIMING_MESH_EMIT_MS = 4,  /* mesh_fused_kernel: classify + count + chain + emit */
	TSOM_TIMING_MESH_REEMIT = 5,  /* 1.0 when the arenas had to grow and the kernel ran twice (the time is then of the first, count-only run) */
	TSOM_TIMING_IO_MS = 6,        /* kernels of the most recent wire-format call: lz4_compress (+ pack, summed over its sub-batches) or lz4_decompress */
	TSOM_TIMING_COUNT = 7
};
int tsom_get_timings(tsom_ctx* ctx, float* out, uint32_t n);
/* pre-size the mesh arenas (faces); they also grow on demand */
int tsom_reserve_faces(tsom_ctx* ctx, uint64_t faces, uint32_t flags);
/* every arena is allocated with a 256-byte guard filled with 0xA5 behind its last face; returns TSOM_ERR_INTERNAL if a kernel wrote into
   one (self-check for tests: compute-sanitizer is not always available) */
int tsom_debug_check_guards(tsom_ctx* ctx);

/* ---- container: Planet / ChunkContainer (include/CommonLib/ChunkContainer.hpp, Planet.hpp) ---- */
int tsom_container_create(tsom_ctx* ctx, uint32_t capacity_chunks, float tile_size, tsom_container** out);
void tsom_container_destroy(tsom_container* c);
/* Planet::AddChunk (Planet.cpp:28-68) without initCallback: the chunk exists but HasContent() is false */
int tsom_container_add_chunks(tsom_container* c, const int32_t* chunk_indices /* n x 3 */, uint32_t n);
/* Planet::RemoveChunk (Planet.cpp:514-521) */
int tsom_container_remove_chunk(tsom_container* c, const int32_t chunk_index[3]);
uint32_t tsom_container_chunk_count(const tsom_container* c);
/* Chunk::Reset(F) with F = copy of 32768 host ids (Chunk.inl:109-123; caller src/ClientLib/ClientSessionHandler.cpp:156-176).
 * Chunks that do not exist yet are added.  Marks each chunk invalidated with DirectionMask_All like the OnReset signal. */
int tsom_chunks_reset(tsom_container* c, const int32_t* chunk_indices, uint32_t n, const uint16_t* blocks_host);
/* same, source already in device memory (n x 32768 uint16, contiguous) */
int tsom_chunks_reset_device(tsom_container* c, const int32_t* chunk_indices, uint32_t n, const void* blocks_device);
/* Chunk::GetContent (Chunk.inl:70-74): copy n chunks' ids to host */
int tsom_chunks_get_content(tsom_container* c, const int32_t* chunk_indices, uint32_t n, uint16_t* blocks_host);
/* device address of one chunk's 32768 ids (NULL if unknown / no content) */
const void* tsom_chunk_device_ptr(tsom_container* c, const int32_t chunk_index[3]);

/* ---- save format: Chunk::Serialize / Chunk::Deserialize (src/CommonLib/Chunk.cpp:252-346; callers
 * src/ServerLib/ServerPlanetEnvironment.cpp:105-129,207-220), device part.  The header of a serialized chunk (version, size,
 * the NAMES of the block types that occur) is host work — names live in the caller's BlockLibrary (adapter/tsom_adapter.hpp
 * composes it); the payload, one palette index per block, is produced / consumed on the device.
 * tsom_chunks_palettize: for chunk i, palette_counts_out[i] = K = number of distinct block types, palette_out[i*palette_stride ..]
 * = their BlockIndex values ascending (the order the reference writes the names in), payload_out + i*65536 = 32768 palette
 * indices: uint8 when K <= 8, else uint16 (Chunk.cpp:335-345), big-endian when payload_big_endian != 0.  TSOM_ERR_CAPACITY if
 * some K exceeds palette_stride (the counts are still written). */
int tsom_chunks_palettize(tsom_container* c, const int32_t* chunk_indices, uint32_t n, uint16_t* palette_out, uint32_t palette_stride,
                          uint32_t* palette_counts_out, uint8_t* payload_out, int payload_big_endian);
/* Chunk::Deserialize after the host has parsed the header and mapped the names to BlockIndex values: expands the payloads
 * (same layout as above) into the chunks' ids like Chunk::Reset (chunks are added if missing, marked invalidated).
 * TSOM_ERR_INVALID if a payload index is >= its chunk's palette count (the reference indexes out of bounds there). */
int tsom_chunks_reset_palettized(tsom_container* c, const int32_t* chunk_indices, uint32_t n, const uint16_t* palettes, uint32_t palette_stride,
                                 const uint32_t* palette_counts, const uint8_t* payload, int payload_big_endian);

/* ---- wire format: Packets::ChunkReset (src/CommonLib/Protocol/Packets.cpp:172-212) ----
 * The reference sends a chunk as ONE LZ4 block of its 65,536-byte block array: BinaryCompressor::Compress =
 * LZ4_compress_fast_extState(fresh state, acceleration 0 -> 1) (src/CommonLib/Utility/BinaryCompressor.cpp:17-36).
 * tsom_chunks_lz4_compress produces exactly those bytes on the device, chunk i at streams_out[offsets_out[i] .. offsets_out[i+1]);
 * offsets_out has n + 1 entries.  TSOM_ERR_CAPACITY if capacity < offsets_out[n] (offsets_out is complete, nothing usable in
 * streams_out); 65,809 bytes per chunk (LZ4_compressBound) always suffice. */
int tsom_chunks_lz4_compress(tsom_container* c, const int32_t* chunk_indices, uint32_t n, uint8_t* streams_out, uint64_t capacity, uint64_t* offsets_out);
/* The receiving side (Packets.cpp:197-211 + Chunk::Reset, src/ClientLib/ClientSessionHandler.cpp:156-176): LZ4_decompress_safe of each
 * block on the device, then the chunks are reset (created if unknown).  All or nothing: TSOM_ERR_INVALID with the reference's message
 * ("failed to decompress chunk" / "malformed packet ...") if any block is malformed or does not decode to 65,536 bytes.  Deviation:
 * a match offset of 0, which LZ4_decompress_safe accepts with unspecified output, is rejected. */
int tsom_chunks_reset_lz4(tsom_container* c, const int32_t* chunk_indices, uint32_t n, const uint8_t* streams, const uint64_t* offsets);

/* ---- generation ---- */
/* Planet::GenerateChunks (Planet.cpp:385-403): adds chunk_count^3 chunks at (x,y,z) - chunk_count/2 in z,y,x order and
 * fills every one with Planet::GenerateChunk (Planet.cpp:160-383). */
int tsom_planet_generate(tsom_container* c, uint32_t seed, const uint32_t chunk_count[3], const tsom_gen_blocks* ids);
/* Planet::GenerateChunk on chosen (existing or new) chunks — the /regenchunk caller, src/ServerLib/Session/PlayerSessionHandler.cpp:290 */
int tsom_planet_generate_chunks(tsom_container* c, uint32_t seed, const uint32_t chunk_count[3], const tsom_gen_blocks* ids,
                                const int32_t* chunk_indices, uint32_t n);
/* Planet::GeneratePlatform (Planet.cpp:405-512), executed against the device-resident blocks */
int tsom_planet_generate_platform(tsom_container* c, int up_direction, const int32_t platform_center[3], const tsom_platform_blocks* ids);

/* ---- edits and dirty propagation ---- */
/* n x Chunk::UpdateBlock in order (later edits win), each raising the neighbour mask of Planet.cpp:39-58.
 * Edits that name a chunk that is absent or has no content are skipped, as the callers do. */
int tsom_apply_edits(tsom_container* c, const tsom_edit* edits, uint32_t n);
/* ChunkEntities::Update (src/CommonLib/ChunkEntities.cpp:108-111) + ClientChunkEntities::ProcessChunkUpdate
 * (src/ClientLib/ClientChunkEntities.cpp:243-256): the chunks to rebuild this tick = invalidated chunks with content plus
 * their masked neighbours that exist and have content.  Sorted by (x,y,z).  Clears the invalidated set when clear != 0.
 * *n_out receives the full count even if it exceeds cap. */
int tsom_collect_dirty(tsom_container* c, int32_t* chunk_indices_out, uint32_t cap, uint32_t* n_out, int clear);

/* ---- meshing ---- */
/* Chunk::BuildMesh (src/CommonLib/Chunk.cpp:20-250) for n chunks in ONE batch (chunk_indices NULL: every chunk with
 * content, container order).  Results land in the context's device arenas: every chunk gets one contiguous face range, its
 * faces in the reference's emission order; the ranges tile [0, total_faces) in no particular order unless TSOM_MESH_ORDERED is
 * set.  views_out (host, n entries, optional) tells where each chunk's faces are.  Chunks without content get 0 faces. */
int tsom_mesh(tsom_container* c, const int32_t* chunk_indices, uint32_t n, const tsom_mesh_params* params, tsom_mesh_view* views_out);
/* device arenas of the last tsom_mesh call on this context */
int tsom_mesh_arenas(tsom_ctx* ctx, const void** vertices, const void** indices, const void** collider_positions, uint64_t* total_faces);
/* copy faces [first_face, first_face + face_count) to host: the std::vector<VertexStruct> / std::vector<UInt32> /
 * std::vector<Vector3f> the reference's callers own.  Any destination may be NULL. */
int tsom_mesh_copy_to_host(tsom_ctx* ctx, uint64_t first_face, uint64_t face_count, void* vertices, uint32_t* indices, float* collider_positions);
/* Chunk::ComputeVoxelCorners (include/CommonLib/Chunk.hpp:54; FlatChunk.cpp:48-54, DeformedChunk.cpp:115-137) as a batched query: the pure
 * helper the game's block picking uses (src/Game/States/GameState.cpp:859, src/ServerLib/Session/PlayerSessionHandler.cpp:367,489),
 * answered by the device functions the mesher itself uses.  Block i = local_indices[3i..] of chunk chunk_indices[3i..] (the chunk need not
 * exist: the corners depend on its indices only).  params: TSOM_MESH_DEFORMED + deform_radius select DeformedChunk's provider,
 * gravity_centers (n x 3, optional) is its deformation centre (default: container centre - chunk offset).  corners_out: n x 8 x xyz in
 * Nz::BoxCorner enum order (LeftBottomFar, LeftTopFar, RightBottomFar, RightTopFar, LeftBottomNear, LeftTopNear, RightBottomNear,
 * RightTopNear). */
int tsom_voxel_corners(tsom_container* c, const int32_t* chunk_indices, const uint32_t* local_indices, uint32_t n, const tsom_mesh_params* params, float* corners_out);
/* Nz::StaticMesh::GenerateAABB as ClientChunkEntities::BuildMesh calls it on every chunk mesh (src/ClientLib/ClientChunkEntities.cpp:168):
 * the bounding box of the vertex positions of each view of the last tsom_mesh call, computed on the device from the floats in the arena
 * (render vertices, or the collider positions of a collider-only call).  aabb_out: n x {min.xyz, max.xyz}; zeros for a view without faces.
 * (The 12-byte tangent slot stays zero: Chunk::BuildMesh never writes it and StaticMesh::GenerateTangents is third-party.) */
int tsom_mesh_aabb(tsom_ctx* ctx, const tsom_mesh_view* views, uint32_t n, float* aabb_out);
/* FlatChunk::BuildCollider greedy box merge (src/CommonLib/FlatChunk.cpp:13-30,56-153): boxes_out = n_boxes x {pos.xyz, size.xyz}
 * in block units (multiply by the block size like FlatChunk.cpp:20-21 does). */
int tsom_box_collider(tsom_container* c, const int32_t chunk_index[3], float* boxes_out, uint32_t cap, uint32_t* n_out);

/* ---- multi-GPU: one process per GPU, chunks sharded by index, halo slabs pulled over NVLink P2P (no NCCL) ---- */
/* declare a neighbour chunk that lives on another rank; its facing slabs are refreshed by tsom_halo_pull */
int tsom_container_add_remote_chunks(tsom_container* c, const int32_t* chunk_indices, const uint32_t* remote_slots, const int32_t* ranks, uint32_t n);
/* slot of a local chunk (what a peer passes as remote_slots) */
int tsom_chunk_slot(tsom_container* c, const int32_t chunk_index[3], uint32_t* slot_out);
int tsom_ipc_export(tsom_container* c, tsom_ipc_handle* out);
int tsom_ipc_attach(tsom_container* c, int rank, const tsom_ipc_handle* handle);
/* same-process variant (tests, single-process multi-context): attach another container's buffers directly */
int tsom_peer_attach_local(tsom_container* c, int rank, tsom_container* peer);
/* copy the boundary slabs of all remote neighbours from their owners' HBM into this GPU's ghost slabs */
int tsom_halo_pull(tsom_container* c);

#ifdef __cplusplus
}
#endif

#endif /* TSOM_B200_H */
