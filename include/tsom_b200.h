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
	TSOM_MESH_ORDERED = 16, /* place the chunks' face ranges in the arenas in job order (first_face[i] = sum of face_count[< i]) instead of
	                           first come, first served; costs a cross-chunk dependency (decoupled look-back) in the kernel */
	TSOM_MESH_TANGENTS = 32 /* also fill the 12-byte tangent slot of every vertex the way Nz::StaticMesh::GenerateTangents does after
	                           ClientChunkEntities::BuildMesh (src/ClientLib/ClientChunkEntities.cpp:169); needs TSOM_MESH_RENDER */
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
	unsigned char bytes[200]; /* 3 x cudaIpcMemHandle_t (blocks, x-slabs, halo sync flags) + capacity */
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
	TSOM_TIMING_TERRAIN_MS = 0,  /* terrain_kernel (the six directions in one launch) */
	TSOM_TIMING_GENERATE_MS = 1, /* template_kernel + generate_kernel */
	TSOM_TIMING_HALO_MS = 2,     /* tsom_halo_pull: wait for the peers' publish flags + halo_pull_kernel */
	TSOM_TIMING_CLASSIFY_MS = 3, /* job_classify_kernel (marks the jobs whose chunk provably has no face) */
	TSOM_TIMING_MESH_EMIT_MS = 4, /* mesh_fused_kernel: classify + count + chain + emit */
	TSOM_TIMING_MESH_REEMIT = 5, /* 1.0 when the arenas had to grow and the kernel ran twice (the time is then of the first, count-only run) */
	TSOM_TIMING_IO_MS = 6,       /* kernels of the most recent wire-format call: lz4_compress (+ pack, summed over its sub-batches) or lz4_decompress */
	TSOM_TIMING_COLLIDER_MS = 7, /* box_colliders_kernel of the most recent tsom_box_colliders call */
	TSOM_TIMING_COUNT = 8
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
/* The same, returning as soon as the work is queued on the context's stream (Planet::GenerateChunks itself blocks in WaitForTasks; a caller
 * that goes straight on to tsom_halo_pull / tsom_mesh does not need the wait in between).  The next blocking call on the context —
 * tsom_mesh, tsom_synchronize, tsom_chunks_get_content ... — completes it; chunk_indices must stay valid only for the duration of the call. */
int tsom_planet_generate_chunks_async(tsom_container* c, uint32_t seed, const uint32_t chunk_count[3], const tsom_gen_blocks* ids,
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
/* Per-view checksums of the last tsom_mesh result, computed on the device: out[2i] over the 32-bit words of view i's vertices (the
 * collider positions after a collider-only call), out[2i + 1] over its indices.  Position-weighted 64-bit sums that do not depend on
 * where the faces sit in the arena: mesh(N GPUs) is compared with mesh(1 GPU) chunk by chunk without moving the arenas to the host
 * (bench.py's sharded_parity).  Zero for a view without faces. */
int tsom_mesh_checksums(tsom_ctx* ctx, const tsom_mesh_view* views, uint32_t n, uint64_t* out);
/* The index buffer of a chunk with face_count faces — a pure function of the face ordinal, 4k + {0,2,1,1,2,3} (src/CommonLib/Chunk.cpp:95-101)
 * — written by host threads: a caller that wants the std::vector<UInt32> of Chunk::BuildMesh can leave the 24 B per face on the device
 * (tsom_mesh_copy_to_host with indices == NULL) instead of carrying them across PCIe.  No device needed. */
int tsom_mesh_indices_for_faces(uint64_t face_count, uint32_t* indices_out);
/* FlatChunk::BuildCollider greedy box merge (src/CommonLib/FlatChunk.cpp:13-30,56-153): boxes_out = n_boxes x {pos.xyz, size.xyz}
 * in block units (multiply by the block size like FlatChunk.cpp:20-21 does). */
int tsom_box_collider(tsom_container* c, const int32_t chunk_index[3], float* boxes_out, uint32_t cap, uint32_t* n_out);

/* Batched form of the same (the server's physics path: ChunkEntities::UpdateChunkEntity calls BuildCollider for every invalidated chunk,
 * src/CommonLib/ChunkEntities.cpp:198): ONE launch for n chunks, one CTA per chunk, the boxes of chunk i at
 * [views_out[i].first_box, + box_count) of a device arena owned by the context, in the reference's order.  No per-chunk synchronisation.
 * Chunks without content or without a colliding block get 0 boxes (the reference returns a null collider, FlatChunk.cpp:25-26). */
typedef struct tsom_box_view
{
	uint64_t first_box;
	uint32_t box_count;
	uint32_t reserved;
} tsom_box_view;
int tsom_box_colliders(tsom_container* c, const int32_t* chunk_indices, uint32_t n, tsom_box_view* views_out, uint64_t* total_boxes_out);
/* boxes [first_box, first_box + box_count) of the last tsom_box_colliders call -> host, 6 floats each (pos.xyz, size.xyz, block units) */
int tsom_box_colliders_copy_to_host(tsom_ctx* ctx, uint64_t first_box, uint64_t box_count, float* boxes_out);

/* ---- the second ChunkContainer: Ship (include/CommonLib/Ship.hpp, src/CommonLib/Ship.cpp) ----
 * A container is a Planet by default.  TSOM_CONTAINER_SHIP switches the dirty-neighbour rule of tsom_apply_edits to Ship::AddChunk's
 * (Ship.cpp:36-54): local z == 0 raises Direction::Up and z == 31 Direction::Down — the opposite of Planet.cpp:50-53. */
enum tsom_container_kind { TSOM_CONTAINER_PLANET = 0, TSOM_CONTAINER_SHIP = 1 };
int tsom_container_set_kind(tsom_container* c, int kind);
/* Ship::Generate (Ship.cpp:115-152): chunk (0,0,0) is added, reset to Empty and the hull box (6x6x4 when small, else 12x12x6, one
 * forcefield door column) is written with Chunk::UpdateBlock semantics (the chunk ends up invalidated with the masks of those edits). */
int tsom_ship_generate(tsom_container* c, int small, uint16_t hull_block, uint16_t forcefield_block);

/* ---- multi-GPU: chunks sharded by chunk index, halo slabs pulled over NVLink P2P (no NCCL) ----
 * A planet is cut into axis-aligned boxes of chunks, one per GPU (tsom_shard_plan).  Each GPU generates and meshes its own box; between
 * the two phases it copies the 2 KiB facing slab of every neighbour chunk owned by another GPU straight out of the owner's HBM.  The
 * exchange is ordered on the DEVICE: every container owns two epoch flags in its HBM (blocks-final, slabs-pulled) that the peers poll
 * over NVLink from a one-warp wait kernel, so no host barrier sits between generation, exchange and meshing.
 * Two ways to run it: one process per GPU (tsom_ipc_export / tsom_ipc_attach; thisspaceofmine_b200/sharding.py under torchrun), or all
 * GPUs from ONE process — what a TSOM server is (src/Server/main.cpp:24-62) — through tsom_planet_create_sharded below. */

/* Cuts the chunk_count[0] x [1] x [2] chunk grid of Planet::GenerateChunks (Planet.cpp:385-403) into n_parts boxes by recursive
 * bisection: the part range is halved, the box is cut across its longest axis at the layer boundary that best balances the
 * summed chunk weights in proportion to the two part counts.  chunk_weights: optional, one float per chunk in GenerateChunks
 * order (z, y, x loops) — e.g. a + b * face count of a previous pass; NULL = a shell model (chunks that can hold a surface
 * weigh more).  boxes_out: n_parts x 6 int32 = lo.xyz, hi.xyz (inclusive) in chunk indices (already minus chunk_count / 2).
 * TSOM_ERR_INVALID when n_parts exceeds the number of chunks. */
int tsom_shard_plan(const uint32_t chunk_count[3], uint32_t n_parts, const float* chunk_weights, int32_t* boxes_out);

/* declare neighbour chunks that live on another GPU: rank = the peer's id as given to tsom_ipc_attach / tsom_peer_attach_local,
 * remote_slot = tsom_chunk_slot of the chunk in the peer's container.  Remote chunks are assumed to have content whenever their
 * owner has published (a sharded planet generates every chunk before the first exchange). */
int tsom_container_add_remote_chunks(tsom_container* c, const int32_t* chunk_indices, const uint32_t* remote_slots, const int32_t* ranks, uint32_t n);
/* slot of a local chunk (what a peer passes as remote_slots) */
int tsom_chunk_slot(tsom_container* c, const int32_t chunk_index[3], uint32_t* slot_out);
int tsom_ipc_export(tsom_container* c, tsom_ipc_handle* out);
int tsom_ipc_attach(tsom_container* c, int rank, const tsom_ipc_handle* handle);
/* same-process variant (single-process multi-GPU, tests): attach another container's buffers directly; enables peer access when
 * the two contexts sit on different devices */
int tsom_peer_attach_local(tsom_container* c, int rank, tsom_container* peer);
/* The exchange, one ROUND per tick, collective over the containers of a planet (every one of them runs every round, also when it
 * wrote nothing).  All of it is enqueued on the context's stream; neither call waits on the host.
 *   tsom_halo_publish : "my blocks are final for this round" — raises this container's blocks-final flag to round r.
 *   tsom_halo_pull    : publishes if this round has not been published yet, waits (on the device) until every attached peer has
 *                       published round r, copies the facing slabs of all remote neighbours into the local ghost slabs, raises
 *                       the slabs-pulled flag to r.
 * Any call that writes blocks (generate, reset, edits, platform ...) after round r first waits, on the device, until every peer has
 * pulled round r, so a fast GPU cannot overwrite slabs a slower one is still reading.  Nothing may write between a container's
 * publish and its pull.  Containers that share one context (one stream) must be driven publish-all, then pull-all.
 * tsom_mesh fails with TSOM_ERR_INVALID while a remote neighbour's slab has never been pulled (topology changed since the last pull).
 * Peers attached with tsom_peer_attach_local (same process) are waited for with CUDA events whenever their publish / pull has already
 * been enqueued by the host — publish-all, then pull-all guarantees it — so that path never polls; peers in other processes are polled.
 * A polled peer that does not arrive within the time-out (5 s unless tsom_halo_set_timeout changed it) makes the next blocking call on
 * the context fail with TSOM_ERR_INTERNAL instead of hanging the GPU. */
int tsom_halo_publish(tsom_container* c);
int tsom_halo_pull(tsom_container* c);
int tsom_halo_set_timeout(tsom_container* c, uint32_t milliseconds);

/* ---- one planet over several GPUs of ONE process (tsom_create(gpu_ids, n)) ----
 * Owns one context + container per part.  cuda_devices may name a device more than once (two parts on one GPU: what the
 * single-GPU parity test does).  Every call fans out to the parts on worker threads and returns when all are done. */
typedef struct tsom_sharded_planet tsom_sharded_planet;
int tsom_planet_create_sharded(const int* cuda_devices, uint32_t n_parts, const uint32_t chunk_count[3], float tile_size, const float* chunk_weights,
                               tsom_sharded_planet** out);
void tsom_sharded_destroy(tsom_sharded_planet* p);
uint32_t tsom_sharded_part_count(const tsom_sharded_planet* p);
tsom_ctx* tsom_sharded_ctx(tsom_sharded_planet* p, uint32_t part);
tsom_container* tsom_sharded_container(tsom_sharded_planet* p, uint32_t part);
/* lo.xyz, hi.xyz (inclusive chunk indices) of a part */
int tsom_sharded_part_box(const tsom_sharded_planet* p, uint32_t part, int32_t box_out[6]);
/* owner of a chunk (TSOM_ERR_INVALID outside the planet) */
int tsom_sharded_owner(const tsom_sharded_planet* p, const int32_t chunk_index[3], uint32_t* part_out);
int tsom_sharded_set_block_table(tsom_sharded_planet* p, const tsom_block_desc* blocks, uint32_t n);
/* Planet::GenerateChunks (Planet.cpp:385-403) over all parts + one exchange round */
int tsom_sharded_generate(tsom_sharded_planet* p, uint32_t seed, const tsom_gen_blocks* ids);
/* n x Chunk::UpdateBlock in order, each routed to the part that owns its chunk (edits outside the planet are skipped) */
int tsom_sharded_apply_edits(tsom_sharded_planet* p, const tsom_edit* edits, uint32_t n);
/* tsom_collect_dirty over the whole planet, including the neighbours across part boundaries (Planet.cpp:39-58); sorted by (x,y,z) */
int tsom_sharded_collect_dirty(tsom_sharded_planet* p, int32_t* chunk_indices_out, uint32_t cap, uint32_t* n_out, int clear);
/* one exchange round (if blocks changed since the last one), then Chunk::BuildMesh for the n chunks, each on its owner.  views_out[i] is
 * relative to the arenas of part parts_out[i] (tsom_mesh_arenas(tsom_sharded_ctx(p, part), ...)).  chunk_indices NULL: every chunk of the
 * planet in GenerateChunks order (n must be the chunk count when views are requested). */
int tsom_sharded_mesh(tsom_sharded_planet* p, const int32_t* chunk_indices, uint32_t n, const tsom_mesh_params* params, tsom_mesh_view* views_out, uint32_t* parts_out);
int tsom_sharded_mesh_copy_to_host(tsom_sharded_planet* p, uint32_t part, uint64_t first_face, uint64_t face_count, void* vertices, uint32_t* indices, float* collider_positions);
/* Chunk::GetContent for n chunks, wherever they live */
int tsom_sharded_get_content(tsom_sharded_planet* p, const int32_t* chunk_indices, uint32_t n, uint16_t* blocks_host);
/* out: part_count x TSOM_TIMING_COUNT floats (tsom_get_timings of every part) */
int tsom_sharded_get_timings(tsom_sharded_planet* p, float* out);

#ifdef __cplusplus
}
#endif

#endif /* TSOM_B200_H */
