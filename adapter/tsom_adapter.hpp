// tsom_adapter.hpp — C++17 host layer above the C ABI of libtsom_b200.so that keeps the call surface of the reference's
// CommonLib types for the chunk pipeline, so that ClientLib rendering and ServerLib physics consume the same vertex / index
// buffers they get from the CPU builders today.  Header-only; link with -ltsom_b200.
//
// Reference interfaces mirrored (paths relative to the reference repository root):
//   tsom::BlockLibrary        include/CommonLib/BlockLibrary.hpp:17-73, src/CommonLib/BlockLibrary.cpp:10-143
//   tsom::ChunkContainer      include/CommonLib/ChunkContainer.hpp:17-52, ChunkContainer.inl:14-61
//   tsom::Planet              include/CommonLib/Planet.hpp:25-75, src/CommonLib/Planet.cpp:28-68,160-521
//   tsom::Chunk               include/CommonLib/Chunk.hpp:41-115 (BuildMesh :51, BuildCollider :50, UpdateBlock :84, Reset :76-77)
//   ChunkEntities::Update     src/CommonLib/ChunkEntities.cpp:73-112, src/ClientLib/ClientChunkEntities.cpp:141-262
//
// What differs from the reference, on purpose:
//   * block ids live in HBM; Chunk is a light handle (container + indices), not an owner of a std::vector<BlockIndex>;
//   * Planet::GenerateChunks takes no Nz::TaskScheduler — the GPU is the scheduler (one launch for all chunks);
//   * Chunk::BuildMesh on one chunk works (same signature shape), but the intended entry is ChunkEntities-style batching:
//     ChunkContainer::BuildMeshes(indices) meshes all invalidated chunks of a tick in ONE tsom_mesh call;
//   * errors: the reference asserts; here every failing C call throws tsom_b200::Error carrying tsom_last_error().
//
// Nazara math types are replaced by the minimal PODs below (layout-compatible with Nz::Vector3<T>: three T's).
#pragma once

#include "../include/tsom_b200.h"

#include <algorithm>
#include <cmath>
#include <array>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace tsom_b200
{
	struct Error : std::runtime_error
	{
		int code;
		Error(int c, const char* where) : std::runtime_error(std::string(where) + ": " + tsom_last_error()), code(c) {}
	};

	inline void Check(int rc, const char* where)
	{
		if (rc != TSOM_OK)
			throw Error(rc, where);
	}

	template<typename T> struct Vector3
	{
		T x{}, y{}, z{};
		constexpr Vector3() = default;
		constexpr Vector3(T X, T Y, T Z) : x(X), y(Y), z(Z) {}
		constexpr explicit Vector3(T v) : x(v), y(v), z(v) {}
		constexpr bool operator==(const Vector3& o) const { return x == o.x && y == o.y && z == o.z; }
		constexpr bool operator!=(const Vector3& o) const { return !(*this == o); }
		constexpr Vector3 operator+(const Vector3& o) const { return { T(x + o.x), T(y + o.y), T(z + o.z) }; }
		constexpr Vector3 operator-(const Vector3& o) const { return { T(x - o.x), T(y - o.y), T(z - o.z) }; }
		constexpr Vector3 operator*(T s) const { return { T(x * s), T(y * s), T(z * s) }; }
		constexpr Vector3 operator/(T s) const { return { T(x / s), T(y / s), T(z / s) }; }
	};
	using Vector3f = Vector3<float>;
	using Vector3i = Vector3<std::int32_t>;
	using Vector3ui = Vector3<std::uint32_t>;

	using BlockIndex = std::uint16_t;                       // include/CommonLib/BlockIndex.hpp:15
	constexpr BlockIndex EmptyBlockIndex = 0;               // BlockIndex.hpp:17
	constexpr BlockIndex InvalidBlockIndex = 0xFFFF;        // BlockIndex.hpp:18
	using BlockIndices = Vector3i;                          // Chunk.hpp:36
	using ChunkIndices = Vector3i;                          // Chunk.hpp:37

	enum class Direction { Back = 0, Down = 1, Front = 2, Left = 3, Right = 4, Up = 5 }; // Direction.hpp:18-28

	// ------------------------------------------------------------------------------------------------ BlockLibrary
	class BlockLibrary
	{
		public:
			struct BlockInfo // BlockLibrary.hpp:37-50
			{
				std::string basePath, baseBackPath, baseDownPath, baseFrontPath, baseLeftPath, baseRightPath, baseSidePath, baseUpPath;
				bool hasCollisions = true;
				bool isDoubleSided = false;
				bool isTransparent = false;
				float permeability = 0.f;
			};

			struct BlockData // BlockLibrary.hpp:52-60
			{
				std::string name;
				unsigned int texIndices[6] = {};
				bool hasCollisions = true;
				bool isDoubleSided = false;
				bool isTransparent = false;
				float permeability = 0.f;
			};

			// the 14 built-in block types in the reference's registration order (BlockLibrary.cpp:10-113)
			BlockLibrary()
			{
				auto info = [](std::string base) { BlockInfo i; i.basePath = std::move(base); return i; };
				{ BlockInfo i; i.hasCollisions = false; i.isTransparent = true; i.permeability = 1.f; RegisterBlock("empty", i); }
				{ BlockInfo i; i.baseBackPath = "blocks/debug_back"; i.baseDownPath = "blocks/debug_down"; i.baseFrontPath = "blocks/debug_front";
				  i.baseLeftPath = "blocks/debug_left"; i.baseRightPath = "blocks/debug_right"; i.baseUpPath = "blocks/debug_up"; RegisterBlock("debug", i); }
				{ BlockInfo i = info("blocks/dirt"); i.permeability = 0.1f; RegisterBlock("dirt", i); }
				{ BlockInfo i = info("blocks/grass_top"); i.baseDownPath = "blocks/dirt"; i.baseSidePath = "blocks/grass_side"; i.permeability = 0.1f; RegisterBlock("grass", i); }
				RegisterBlock("hull", info("blocks/smooth_stone"));
				RegisterBlock("hull2", info("blocks/smooth_stone_slab_side"));
				{ BlockInfo i = info("blocks/snow"); i.permeability = 0.5f; RegisterBlock("snow", i); }
				RegisterBlock("stone", info("blocks/cobblestone"));
				RegisterBlock("stone_mossy", info("blocks/mossy_cobblestone"));
				{ BlockInfo i = info("blocks/forcefield"); i.hasCollisions = false; i.isTransparent = true; RegisterBlock("forcefield", i); }
				RegisterBlock("planks", info("blocks/planks"));
				RegisterBlock("stone_bricks", info("blocks/stone_bricks"));
				RegisterBlock("copper_block", info("blocks/copper_block"));
				{ BlockInfo i = info("blocks/glass"); i.isDoubleSided = true; i.isTransparent = true; RegisterBlock("glass", i); }
			}
			BlockLibrary(const BlockLibrary&) = delete;
			BlockLibrary& operator=(const BlockLibrary&) = delete;

			const BlockData& GetBlockData(BlockIndex blockIndex) const { return m_blocks.at(blockIndex); }

			BlockIndex GetBlockIndex(std::string_view blockName) const
			{
				auto it = m_blockIndices.find(std::string(blockName));
				return it == m_blockIndices.end() ? InvalidBlockIndex : it->second;
			}

			// BlockLibrary.cpp:115-134: per-direction path, else the side texture for the four sides, else the base texture
			BlockIndex RegisterBlock(std::string name, BlockInfo blockInfo)
			{
				BlockData data;
				data.name = name;
				data.hasCollisions = blockInfo.hasCollisions;
				data.isDoubleSided = blockInfo.isDoubleSided;
				data.isTransparent = blockInfo.isTransparent;
				data.permeability = blockInfo.permeability;
				const unsigned int base = blockInfo.basePath.empty() ? 0u : RegisterTexture(blockInfo.basePath);
				const unsigned int side = blockInfo.baseSidePath.empty() ? base : RegisterTexture(blockInfo.baseSidePath);
				const std::string* perDir[6] = { &blockInfo.baseBackPath, &blockInfo.baseDownPath, &blockInfo.baseFrontPath, &blockInfo.baseLeftPath, &blockInfo.baseRightPath, &blockInfo.baseUpPath };
				for (int d = 0; d < 6; ++d)
				{
					if (!perDir[d]->empty())
						data.texIndices[d] = RegisterTexture(*perDir[d]);
					else
						data.texIndices[d] = (d == int(Direction::Up) || d == int(Direction::Down)) ? base : side;
				}
				const BlockIndex index = BlockIndex(m_blocks.size());
				m_blockIndices.emplace(std::move(name), index);
				m_blocks.push_back(std::move(data));
				return index;
			}

			std::size_t GetBlockCount() const { return m_blocks.size(); }

			// device form of the table (tsom_set_block_table)
			std::vector<tsom_block_desc> ToDescs() const
			{
				std::vector<tsom_block_desc> out(m_blocks.size());
				for (std::size_t i = 0; i < m_blocks.size(); ++i)
				{
					for (int d = 0; d < 6; ++d)
						out[i].tex[d] = m_blocks[i].texIndices[d];
					out[i].has_collisions = m_blocks[i].hasCollisions;
					out[i].is_double_sided = m_blocks[i].isDoubleSided;
					out[i].is_transparent = m_blocks[i].isTransparent;
					out[i].reserved = 0;
				}
				return out;
			}

		private:
			unsigned int RegisterTexture(const std::string& path)
			{
				auto it = m_textureIndices.find(path);
				if (it != m_textureIndices.end())
					return it->second;
				const unsigned int tex = unsigned(m_textureIndices.size()) + 1u; // slice 0 stays empty (BlockLibrary.cpp:137-142)
				m_textureIndices.emplace(path, tex);
				return tex;
			}

			std::unordered_map<std::string, BlockIndex> m_blockIndices;
			std::unordered_map<std::string, unsigned int> m_textureIndices;
			std::vector<BlockData> m_blocks;
	};

	// ------------------------------------------------------------------------------------------------ device context
	// One per process and GPU.  Owns the stream, the device block table and the mesh arenas (tsom_ctx).
	class Device
	{
		public:
			explicit Device(const BlockLibrary& blockLibrary, int cudaDevice = 0)
			{
				Check(tsom_create(cudaDevice, &m_ctx), "tsom_create");
				auto descs = blockLibrary.ToDescs();
				Check(tsom_set_block_table(m_ctx, descs.data(), std::uint32_t(descs.size())), "tsom_set_block_table");
			}
			Device(const Device&) = delete;
			Device& operator=(const Device&) = delete;
			~Device() { tsom_destroy(m_ctx); }
			tsom_ctx* Handle() const { return m_ctx; }
			void Synchronize() const { Check(tsom_synchronize(m_ctx), "tsom_synchronize"); }

			// Keeps "mesh, then copy out of the arenas" together when several threads (the reference's TaskScheduler workers) share
			// the device: every C entry point locks the context by itself, but the arenas belong to the most recent BuildMeshes.
			class Lock
			{
				public:
					explicit Lock(const Device& device) : m_ctx(device.m_ctx) { Check(tsom_ctx_lock(m_ctx), "tsom_ctx_lock"); }
					Lock(const Lock&) = delete;
					Lock& operator=(const Lock&) = delete;
					~Lock() { tsom_ctx_unlock(m_ctx); }

				private:
					tsom_ctx* m_ctx;
			};

		private:
			tsom_ctx* m_ctx = nullptr;
	};

	// what ClientChunkEntities::BuildMesh hands to Nazara (48-byte VertexStruct, ClientChunkEntities.hpp:27-33) per chunk
	struct VertexStruct
	{
		Vector3f position, normal, uvw, tangent;
	};
	static_assert(sizeof(VertexStruct) == TSOM_VERTEX_STRIDE, "vertex layout");

	// result of one batched tsom_mesh call: per requested chunk its face range; CopyChunk() fills the caller's std::vectors
	class MeshBatch
	{
		public:
			std::vector<tsom_mesh_view> views;
			std::uint64_t totalFaces = 0;

			std::uint32_t FaceCount(std::size_t i) const { return views[i].face_count; }

			// vertices / indices / collider positions of chunk i (any may be null), the vectors the reference's callers own
			void CopyChunk(std::size_t i, std::vector<VertexStruct>* vertices, std::vector<std::uint32_t>* indices, std::vector<Vector3f>* colliderPositions) const
			{
				const tsom_mesh_view& v = views[i];
				if (vertices)
					vertices->resize(std::size_t(v.face_count) * 4);
				if (indices)
					indices->resize(std::size_t(v.face_count) * 6);
				if (colliderPositions)
					colliderPositions->resize(std::size_t(v.face_count) * 4);
				if (v.face_count == 0)
					return;
				Check(tsom_mesh_copy_to_host(m_ctx, v.first_face, v.face_count, vertices ? vertices->data() : nullptr, indices ? indices->data() : nullptr,
				                             colliderPositions ? &colliderPositions->data()->x : nullptr), "tsom_mesh_copy_to_host");
			}

			// The same without carrying the indices across PCIe: they are a pure function of the face ordinal (Chunk.cpp:95-101)
			void CopyChunkSynthesizeIndices(std::size_t i, std::vector<VertexStruct>* vertices, std::vector<std::uint32_t>& indices) const
			{
				CopyChunk(i, vertices, nullptr, nullptr);
				indices.resize(std::size_t(views[i].face_count) * 6);
				Check(tsom_mesh_indices_for_faces(views[i].face_count, indices.data()), "tsom_mesh_indices_for_faces");
			}

			// Nz::StaticMesh::GenerateAABB of every chunk (ClientChunkEntities.cpp:168), from the arena: {min, max} of the positions
			struct Aabb
			{
				Vector3f min, max;
			};
			std::vector<Aabb> ComputeAABBs() const
			{
				std::vector<Aabb> out(views.size());
				static_assert(sizeof(Aabb) == 6 * sizeof(float), "aabb layout");
				Check(tsom_mesh_aabb(m_ctx, views.data(), std::uint32_t(views.size()), out.empty() ? nullptr : &out[0].min.x), "tsom_mesh_aabb");
				return out;
			}

		private:
			friend class ChunkContainer;
			tsom_ctx* m_ctx = nullptr;
	};

	class ChunkContainer;

	enum class CornerMode { Flat, Deformed }; // FlatChunk.cpp:48-54 / DeformedChunk.cpp:115-154
	struct MeshOptions
	{
		bool render = true;                  // 48-byte vertices
		bool collider = false;               // positions only
		CornerMode corners = CornerMode::Flat;
		float deformRadius = 0.f;            // Planet::GetCornerRadius() for DeformedChunk
		bool ordered = false;                // arena ranges in job order (TSOM_MESH_ORDERED); default: first come, first served
		bool tangents = false;               // also what Nz::StaticMesh::GenerateTangents adds (ClientChunkEntities.cpp:169)
		const Vector3f* gravityCenters = nullptr; // per chunk `center` argument of Chunk::BuildMesh; null = GetCenter() - GetChunkOffset(chunk)
	};

	// ------------------------------------------------------------------------------------------------ Chunk (handle)
	class Chunk
	{
		public:
			// Chunk::VertexAttributes (Chunk.hpp:92-100): strided destinations; a null pointer means "do not write"
			struct VertexAttributes
			{
				std::uint32_t firstIndex = 0;
				Vector3f* position = nullptr;
				Vector3f* normal = nullptr;
				Vector3f* uv = nullptr;
				std::size_t stride = sizeof(VertexStruct); // bytes between consecutive vertices of each attribute
			};

			Chunk(ChunkContainer& owner, const ChunkIndices& indices) : m_owner(&owner), m_indices(indices) {}

			// Chunk::BuildMesh (Chunk.hpp:51, Chunk.cpp:20-250): appends this chunk's triangles to `indices` (offset by
			// firstIndex) and writes count = 4 * faces vertices through addVertices — one call, like the reference's single
			// allocation per face but batched.
			void BuildMesh(std::vector<std::uint32_t>& indices, const Vector3f& center, const std::function<VertexAttributes(std::uint32_t count)>& addVertices) const;

			// DeformedChunk::BuildCollider (DeformedChunk.cpp:15-36): positions + indices of the same face list; nullopt <=> nullptr
			struct TriangleCollider
			{
				std::vector<Vector3f> positions;
				std::vector<std::uint32_t> indices;
			};
			std::optional<TriangleCollider> BuildCollider() const;

			// FlatChunk::BuildCollider (FlatChunk.cpp:13-30,56-153): greedy boxes {position, size} already scaled by the block size
			struct Box
			{
				Vector3f position, size;
			};
			std::vector<Box> BuildBoxCollider() const;

			// Chunk::GetContent (Chunk.inl:70-74): a host copy of the 32768 ids
			std::vector<BlockIndex> GetContent() const;
			const ChunkIndices& GetIndices() const { return m_indices; }
			ChunkContainer& GetContainer() { return *m_owner; }
			bool HasContent() const;

			// Chunk::Reset(F) (Chunk.inl:109-123): func fills 32768 ids, zero-initialised as in the reference
			template<typename F> void Reset(F&& func);

			// Chunk::UpdateBlock (Chunk.cpp:348-366); batching many edits per tick: ChunkContainer::UpdateBlocks
			void UpdateBlock(const Vector3ui& indices, BlockIndex cellType);

			// Chunk::ComputeVoxelCorners (Chunk.hpp:54; FlatChunk.cpp:48-54, DeformedChunk.cpp:115-137), the helper block picking uses
			// (GameState.cpp:859, PlayerSessionHandler.cpp:367,489): the 8 corners of a block, indexed by BoxCorner; evaluated by the
			// device functions the mesher uses.  With CornerMode::Deformed the deformation centre is `deformCenter`, or by default the
			// container centre minus the chunk offset.
			enum class BoxCorner { LeftBottomFar, LeftTopFar, RightBottomFar, RightTopFar, LeftBottomNear, LeftTopNear, RightBottomNear, RightTopNear };
			std::array<Vector3f, 8> ComputeVoxelCorners(const Vector3ui& indices, CornerMode corners = CornerMode::Flat, float deformRadius = 0.f, const Vector3f* deformCenter = nullptr) const;

			// Chunk::Serialize / Deserialize (Chunk.cpp:252-346), the `<x>_<y>_<z>.chunk` save format: header composed / parsed
			// here (big-endian integers, UInt32-prefixed names — Nazara ByteStream conventions), palette + payload on the device.
			// Deserialize throws std::runtime_error with the reference's messages.
			std::vector<std::uint8_t> Serialize(const BlockLibrary& blockLibrary) const;
			void Deserialize(const BlockLibrary& blockLibrary, const std::uint8_t* data, std::size_t size);

			// The `content` of Packets::ChunkReset (Packets.cpp:172-212): the block array as one LZ4 block, byte for byte what
			// BinaryCompressor::Compress produces (LZ4_compress_fast_extState, acceleration 1), made on the device; and the receiving
			// side (LZ4_decompress_safe + Reset).  ResetCompressed throws with the reference's messages and leaves the chunk untouched
			// when the block is malformed.  Batches: tsom_chunks_lz4_compress / tsom_chunks_reset_lz4 take any number of chunks.
			std::vector<std::uint8_t> Compress() const;
			void ResetCompressed(const std::uint8_t* data, std::size_t size);

			static unsigned int GetBlockLocalIndex(const Vector3ui& i) { return 32u * (32u * i.z + i.y) + i.x; } // Chunk.inl:25-28

		private:
			ChunkContainer* m_owner;
			ChunkIndices m_indices;
	};

	// ------------------------------------------------------------------------------------------------ ChunkContainer
	class ChunkContainer
	{
		public:
			static constexpr unsigned int ChunkSize = 32; // ChunkContainer.hpp:42

			ChunkContainer(Device& device, float tileSize, std::uint32_t capacityChunks) : m_device(&device), m_tileSize(tileSize)
			{
				Check(tsom_container_create(device.Handle(), capacityChunks, tileSize, &m_container), "tsom_container_create");
			}
			ChunkContainer(const ChunkContainer&) = delete;
			ChunkContainer& operator=(const ChunkContainer&) = delete;
			virtual ~ChunkContainer() { tsom_container_destroy(m_container); }

			// ChunkContainer.inl:14-17 — note the y/z swap: local z is world Y
			BlockIndices GetBlockIndices(const ChunkIndices& chunkIndices, const Vector3ui& indices) const
			{
				return chunkIndices * std::int32_t(ChunkSize) + BlockIndices(std::int32_t(indices.x), std::int32_t(indices.z), std::int32_t(indices.y)) - BlockIndices(std::int32_t(ChunkSize)) / 2;
			}

			// ChunkContainer.inl:19-42
			ChunkIndices GetChunkIndicesByBlockIndices(const BlockIndices& indices, Vector3ui* localIndices = nullptr) const
			{
				const BlockIndices adjusted = indices + BlockIndices(std::int32_t(ChunkSize)) / 2;
				auto floorFix = [](int v) { return v < 0 ? v - int(ChunkSize) + 1 : v; };
				const ChunkIndices chunk(floorFix(adjusted.x) / int(ChunkSize), floorFix(adjusted.y) / int(ChunkSize), floorFix(adjusted.z) / int(ChunkSize));
				if (localIndices)
				{
					Vector3i l = adjusted - chunk * std::int32_t(ChunkSize);
					auto wrap = [](int v) { return v < 0 ? v + int(ChunkSize) : v; };
					*localIndices = Vector3ui(std::uint32_t(wrap(l.x)), std::uint32_t(wrap(l.z)), std::uint32_t(wrap(l.y)));
				}
				return chunk;
			}

			// ChunkContainer.inl:44-51
			ChunkIndices GetChunkIndicesByPosition(const Vector3f& position) const
			{
				const Vector3f rel = position - GetCenter();
				const float cs = float(ChunkSize) * m_tileSize;
				auto f = [cs](float v) { return std::roundf((v + cs * 0.5f) / cs - 0.5f); };
				return ChunkIndices(std::int32_t(f(rel.x)), std::int32_t(f(rel.y)), std::int32_t(f(rel.z)));
			}

			// ChunkContainer.inl:53-56
			Vector3f GetChunkOffset(const ChunkIndices& indices) const { return Vector3f(float(indices.x), float(indices.y), float(indices.z)) * float(ChunkSize) * m_tileSize; }

			virtual Vector3f GetCenter() const { return Vector3f(0.f, 0.f, 0.f); } // Planet.inl: Planet::GetCenter() == Zero
			float GetTileSize() const { return m_tileSize; }
			std::size_t GetChunkCount() const { return tsom_container_chunk_count(m_container); }

			Chunk GetChunk(const ChunkIndices& chunkIndices) { return Chunk(*this, chunkIndices); }

			void RemoveChunk(const ChunkIndices& indices) // Planet.cpp:514-521
			{
				const std::int32_t idx[3] = { indices.x, indices.y, indices.z };
				Check(tsom_container_remove_chunk(m_container, idx), "tsom_container_remove_chunk");
			}

			// ---- batched entry points (what ChunkEntities::Update would call once per tick) ----
			struct BlockUpdate
			{
				ChunkIndices chunk;
				Vector3ui indices;
				BlockIndex cellType;
			};
			// n x Chunk::UpdateBlock in order, with the neighbour invalidation of Planet.cpp:39-58
			void UpdateBlocks(const std::vector<BlockUpdate>& updates)
			{
				std::vector<tsom_edit> edits(updates.size());
				for (std::size_t i = 0; i < updates.size(); ++i)
				{
					edits[i] = tsom_edit{};
					edits[i].chunk[0] = updates[i].chunk.x;
					edits[i].chunk[1] = updates[i].chunk.y;
					edits[i].chunk[2] = updates[i].chunk.z;
					edits[i].x = std::uint8_t(updates[i].indices.x);
					edits[i].y = std::uint8_t(updates[i].indices.y);
					edits[i].z = std::uint8_t(updates[i].indices.z);
					edits[i].block = updates[i].cellType;
				}
				Check(tsom_apply_edits(m_container, edits.data(), std::uint32_t(edits.size())), "tsom_apply_edits");
			}

			// the chunks ChunkEntities::Update / ClientChunkEntities::ProcessChunkUpdate would rebuild this tick (sorted), and clears the set
			std::vector<ChunkIndices> CollectInvalidatedChunks()
			{
				std::uint32_t n = 0;
				Check(tsom_collect_dirty(m_container, nullptr, 0, &n, 0), "tsom_collect_dirty");
				std::vector<ChunkIndices> out(n);
				static_assert(sizeof(ChunkIndices) == 3 * sizeof(std::int32_t), "ChunkIndices layout");
				Check(tsom_collect_dirty(m_container, n ? &out[0].x : nullptr, n, &n, 1), "tsom_collect_dirty");
				return out;
			}

			using CornerMode = tsom_b200::CornerMode;
			using MeshOptions = tsom_b200::MeshOptions;

			// Chunk::BuildMesh for all `chunks` in ONE launch
			MeshBatch BuildMeshes(const std::vector<ChunkIndices>& chunks, const MeshOptions& opt = MeshOptions())
			{
				MeshBatch batch;
				batch.m_ctx = m_device->Handle();
				batch.views.resize(chunks.size());
				tsom_mesh_params p{};
				p.flags = (opt.render ? std::uint32_t(TSOM_MESH_RENDER) : 0u) | (opt.collider ? std::uint32_t(TSOM_MESH_COLLIDER) : 0u) | (opt.corners == CornerMode::Deformed ? std::uint32_t(TSOM_MESH_DEFORMED) : 0u) |
				          (opt.ordered ? std::uint32_t(TSOM_MESH_ORDERED) : 0u) | (opt.tangents ? std::uint32_t(TSOM_MESH_TANGENTS) : 0u);
				p.deform_radius = opt.deformRadius;
				p.gravity_centers = opt.gravityCenters ? &opt.gravityCenters->x : nullptr;
				if (!chunks.empty())
					Check(tsom_mesh(m_container, &chunks[0].x, std::uint32_t(chunks.size()), &p, batch.views.data()), "tsom_mesh");
				Check(tsom_mesh_arenas(m_device->Handle(), nullptr, nullptr, nullptr, &batch.totalFaces), "tsom_mesh_arenas");
				return batch;
			}

			// FlatChunk::BuildCollider for all `chunks` in ONE call (the server's physics path asks it per invalidated chunk,
			// ChunkEntities.cpp:198): boxes[i] = chunk i's greedy boxes {position, size}, scaled by the block size like FlatChunk.cpp:20-21;
			// empty <=> the reference returns a null collider
			std::vector<std::vector<Chunk::Box>> BuildBoxColliders(const std::vector<ChunkIndices>& chunks)
			{
				std::vector<std::vector<Chunk::Box>> out(chunks.size());
				if (chunks.empty())
					return out;
				std::vector<tsom_box_view> views(chunks.size());
				std::uint64_t total = 0;
				Device::Lock lock(*m_device);
				Check(tsom_box_colliders(m_container, &chunks[0].x, std::uint32_t(chunks.size()), views.data(), &total), "tsom_box_colliders");
				std::vector<float> raw(std::size_t(total) * 6);
				if (total)
					Check(tsom_box_colliders_copy_to_host(m_device->Handle(), 0, total, raw.data()), "tsom_box_colliders_copy_to_host");
				for (std::size_t i = 0; i < chunks.size(); ++i)
				{
					out[i].resize(views[i].box_count);
					for (std::uint32_t k = 0; k < views[i].box_count; ++k)
					{
						const float* b = raw.data() + (views[i].first_box + k) * 6;
						out[i][k].position = Vector3f(b[0], b[1], b[2]) * m_tileSize;
						out[i][k].size = Vector3f(b[3], b[4], b[5]) * m_tileSize;
					}
				}
				return out;
			}

			tsom_container* Handle() const { return m_container; }
			Device& GetDevice() const { return *m_device; }

		protected:
			Device* m_device;
			tsom_container* m_container = nullptr;
			float m_tileSize;
	};

	// ------------------------------------------------------------------------------------------------ Planet
	class Planet : public ChunkContainer
	{
		public:
			// Planet::Planet(tileSize, cornerRadius, gravity) (Planet.cpp:19-26) + the HBM capacity in chunks
			Planet(Device& device, float tileSize, float cornerRadius, float gravity, std::uint32_t capacityChunks) :
			ChunkContainer(device, tileSize, capacityChunks), m_cornerRadius(cornerRadius), m_gravity(gravity)
			{
			}

			// Planet::AddChunk (Planet.cpp:28-68): without initCallback the chunk exists but has no content
			Chunk AddChunk(const BlockLibrary& /*blockLibrary*/, const ChunkIndices& indices, const std::function<void(BlockIndex* blocks)>& initCallback = nullptr)
			{
				const std::int32_t idx[3] = { indices.x, indices.y, indices.z };
				if (initCallback)
				{
					std::vector<BlockIndex> blocks(TSOM_CHUNK_BLOCKS, EmptyBlockIndex);
					initCallback(blocks.data());
					Check(tsom_chunks_reset(m_container, idx, 1, blocks.data()), "tsom_chunks_reset");
				}
				else
					Check(tsom_container_add_chunks(m_container, idx, 1), "tsom_container_add_chunks");
				return Chunk(*this, indices);
			}

			// Planet::GenerateChunk (Planet.cpp:160-383) for one chunk — the /regenchunk path
			void GenerateChunk(const BlockLibrary& blockLibrary, Chunk& chunk, std::uint32_t seed, const Vector3ui& chunkCount)
			{
				const tsom_gen_blocks ids = GenBlocks(blockLibrary);
				const std::uint32_t cc[3] = { chunkCount.x, chunkCount.y, chunkCount.z };
				Check(tsom_planet_generate_chunks(m_container, seed, cc, &ids, &chunk.GetIndices().x, 1), "tsom_planet_generate_chunks");
			}

			// Planet::GenerateChunks (Planet.cpp:385-403); no TaskScheduler: 6 terrain launches + 1 generate launch for the whole planet
			void GenerateChunks(const BlockLibrary& blockLibrary, std::uint32_t seed, const Vector3ui& chunkCount)
			{
				const tsom_gen_blocks ids = GenBlocks(blockLibrary);
				const std::uint32_t cc[3] = { chunkCount.x, chunkCount.y, chunkCount.z };
				Check(tsom_planet_generate(m_container, seed, cc, &ids), "tsom_planet_generate");
			}

			// Planet::GeneratePlatform (Planet.cpp:405-512)
			void GeneratePlatform(const BlockLibrary& blockLibrary, Direction upDirection, const BlockIndices& platformCenter)
			{
				tsom_platform_blocks ids;
				ids.copper_block = blockLibrary.GetBlockIndex("copper_block");
				ids.stone_bricks = blockLibrary.GetBlockIndex("stone_bricks");
				ids.planks = blockLibrary.GetBlockIndex("planks");
				Check(tsom_planet_generate_platform(m_container, int(upDirection), &platformCenter.x, &ids), "tsom_planet_generate_platform");
			}

			float GetCornerRadius() const { return m_cornerRadius; }
			float GetGravity() const { return m_gravity; }
			void UpdateCornerRadius(float cornerRadius) { m_cornerRadius = cornerRadius; }

			// the chunk grid of GenerateChunks in its z, y, x order (Planet.cpp:387-393)
			static std::vector<ChunkIndices> ChunkGrid(const Vector3ui& chunkCount)
			{
				std::vector<ChunkIndices> out;
				out.reserve(std::size_t(chunkCount.x) * chunkCount.y * chunkCount.z);
				for (std::uint32_t z = 0; z < chunkCount.z; ++z)
					for (std::uint32_t y = 0; y < chunkCount.y; ++y)
						for (std::uint32_t x = 0; x < chunkCount.x; ++x)
							out.emplace_back(std::int32_t(x) - std::int32_t(chunkCount.x / 2), std::int32_t(y) - std::int32_t(chunkCount.y / 2), std::int32_t(z) - std::int32_t(chunkCount.z / 2));
				return out;
			}

		private:
			static tsom_gen_blocks GenBlocks(const BlockLibrary& lib)
			{
				tsom_gen_blocks ids;
				ids.dirt = lib.GetBlockIndex("dirt");
				ids.grass = lib.GetBlockIndex("grass");
				ids.stone = lib.GetBlockIndex("stone");
				ids.stone_mossy = lib.GetBlockIndex("stone_mossy");
				ids.snow = lib.GetBlockIndex("snow");
				return ids;
			}

			float m_cornerRadius;
			float m_gravity;
	};

	// ------------------------------------------------------------------------------------------------ Ship
	// Ship (include/CommonLib/Ship.hpp, src/CommonLib/Ship.cpp): the reference's second ChunkContainer.  Its block edits raise the
	// neighbour mask of Ship.cpp:36-54 (local z == 0 -> Direction::Up, z == 31 -> Direction::Down: the opposite of Planet's), its
	// physics body is the compound of its chunks' box colliders.
	class Ship : public ChunkContainer
	{
		public:
			Ship(Device& device, float tileSize, std::uint32_t capacityChunks = 8) : ChunkContainer(device, tileSize, capacityChunks)
			{
				Check(tsom_container_set_kind(m_container, TSOM_CONTAINER_SHIP), "tsom_container_set_kind");
			}

			// Ship::AddChunk (Ship.cpp:21-61)
			Chunk AddChunk(const BlockLibrary& /*blockLibrary*/, const ChunkIndices& indices, const std::function<void(BlockIndex* blocks)>& initCallback = nullptr)
			{
				const std::int32_t idx[3] = { indices.x, indices.y, indices.z };
				if (initCallback)
				{
					std::vector<BlockIndex> blocks(TSOM_CHUNK_BLOCKS, EmptyBlockIndex);
					initCallback(blocks.data());
					Check(tsom_chunks_reset(m_container, idx, 1, blocks.data()), "tsom_chunks_reset");
				}
				else
					Check(tsom_container_add_chunks(m_container, idx, 1), "tsom_container_add_chunks");
				if (std::find(m_chunks.begin(), m_chunks.end(), indices) == m_chunks.end())
					m_chunks.push_back(indices);
				return Chunk(*this, indices);
			}

			void RemoveChunk(const ChunkIndices& indices) // Ship.cpp:154-161
			{
				ChunkContainer::RemoveChunk(indices);
				m_chunks.erase(std::remove(m_chunks.begin(), m_chunks.end(), indices), m_chunks.end());
			}

			// Ship::Generate (Ship.cpp:115-152)
			void Generate(const BlockLibrary& blockLibrary, bool small)
			{
				if (std::find(m_chunks.begin(), m_chunks.end(), ChunkIndices(0, 0, 0)) == m_chunks.end())
					AddChunk(blockLibrary, ChunkIndices(0, 0, 0));
				const BlockIndex hull = blockLibrary.GetBlockIndex("hull");
				if (hull == InvalidBlockIndex)
					return;
				Check(tsom_ship_generate(m_container, small ? 1 : 0, hull, blockLibrary.GetBlockIndex("forcefield")), "tsom_ship_generate");
			}

			// Ship::BuildHullCollider (Ship.cpp:63-93): one child per chunk that has a collider — its boxes and GetChunkOffset — built by
			// ONE batched call; an empty result <=> nullptr.  (With a single chunk the reference returns that chunk's compound itself:
			// the one child here, offset zero.)
			struct HullChild
			{
				Vector3f offset;
				std::vector<Chunk::Box> boxes;
			};
			std::vector<HullChild> BuildHullCollider()
			{
				std::vector<HullChild> out;
				std::vector<std::vector<Chunk::Box>> boxes = BuildBoxColliders(m_chunks);
				for (std::size_t i = 0; i < m_chunks.size(); ++i)
					if (!boxes[i].empty())
						out.push_back(HullChild{ m_chunks.size() > 1 ? GetChunkOffset(m_chunks[i]) : Vector3f(0.f, 0.f, 0.f), std::move(boxes[i]) });
				return out;
			}

		private:
			std::vector<ChunkIndices> m_chunks; // insertion order (the reference iterates a hash map: no defined order)
	};

	// ------------------------------------------------------------------------------------------------ a planet over several GPUs
	// tsom_planet_create_sharded: ONE process — what a TSOM server is (src/Server/main.cpp:24-62) — drives all GPUs.  The planet is
	// cut into one box of chunks per GPU; generation, the halo exchange over NVLink and meshing run on all of them at once.  A chunk's
	// mesh lives in the arenas of the GPU that owns it: ShardedMeshBatch::CopyChunk fetches it from there.
	class ShardedPlanet;
	class ShardedMeshBatch
	{
		public:
			std::vector<tsom_mesh_view> views;
			std::vector<std::uint32_t> parts;

			std::uint32_t FaceCount(std::size_t i) const { return views[i].face_count; }
			void CopyChunk(std::size_t i, std::vector<VertexStruct>* vertices, std::vector<std::uint32_t>* indices, std::vector<Vector3f>* colliderPositions) const
			{
				const tsom_mesh_view& v = views[i];
				if (vertices)
					vertices->resize(std::size_t(v.face_count) * 4);
				if (indices)
					indices->resize(std::size_t(v.face_count) * 6);
				if (colliderPositions)
					colliderPositions->resize(std::size_t(v.face_count) * 4);
				if (v.face_count == 0)
					return;
				Check(tsom_sharded_mesh_copy_to_host(m_planet, parts[i], v.first_face, v.face_count, vertices ? vertices->data() : nullptr, indices ? indices->data() : nullptr,
				                                     colliderPositions ? &colliderPositions->data()->x : nullptr), "tsom_sharded_mesh_copy_to_host");
			}

		private:
			friend class ShardedPlanet;
			tsom_sharded_planet* m_planet = nullptr;
	};

	class ShardedPlanet
	{
		public:
			ShardedPlanet(const BlockLibrary& blockLibrary, const std::vector<int>& cudaDevices, const Vector3ui& chunkCount, float tileSize) : m_chunkCount(chunkCount), m_tileSize(tileSize)
			{
				const std::uint32_t cc[3] = { chunkCount.x, chunkCount.y, chunkCount.z };
				Check(tsom_planet_create_sharded(cudaDevices.data(), std::uint32_t(cudaDevices.size()), cc, tileSize, nullptr, &m_planet), "tsom_planet_create_sharded");
				auto descs = blockLibrary.ToDescs();
				Check(tsom_sharded_set_block_table(m_planet, descs.data(), std::uint32_t(descs.size())), "tsom_sharded_set_block_table");
			}
			ShardedPlanet(const ShardedPlanet&) = delete;
			ShardedPlanet& operator=(const ShardedPlanet&) = delete;
			~ShardedPlanet() { tsom_sharded_destroy(m_planet); }

			std::uint32_t GetPartCount() const { return tsom_sharded_part_count(m_planet); }
			std::uint32_t GetOwner(const ChunkIndices& indices) const
			{
				std::uint32_t part = 0;
				Check(tsom_sharded_owner(m_planet, &indices.x, &part), "tsom_sharded_owner");
				return part;
			}
			Vector3f GetChunkOffset(const ChunkIndices& indices) const { return Vector3f(float(indices.x), float(indices.y), float(indices.z)) * float(ChunkContainer::ChunkSize) * m_tileSize; }

			// Planet::GenerateChunks (Planet.cpp:385-403) on all GPUs + one halo exchange round
			void GenerateChunks(const BlockLibrary& lib, std::uint32_t seed)
			{
				tsom_gen_blocks ids;
				ids.dirt = lib.GetBlockIndex("dirt");
				ids.grass = lib.GetBlockIndex("grass");
				ids.stone = lib.GetBlockIndex("stone");
				ids.stone_mossy = lib.GetBlockIndex("stone_mossy");
				ids.snow = lib.GetBlockIndex("snow");
				Check(tsom_sharded_generate(m_planet, seed, &ids), "tsom_sharded_generate");
			}

			void UpdateBlocks(const std::vector<ChunkContainer::BlockUpdate>& updates)
			{
				std::vector<tsom_edit> edits(updates.size());
				for (std::size_t i = 0; i < updates.size(); ++i)
				{
					edits[i] = tsom_edit{};
					edits[i].chunk[0] = updates[i].chunk.x;
					edits[i].chunk[1] = updates[i].chunk.y;
					edits[i].chunk[2] = updates[i].chunk.z;
					edits[i].x = std::uint8_t(updates[i].indices.x);
					edits[i].y = std::uint8_t(updates[i].indices.y);
					edits[i].z = std::uint8_t(updates[i].indices.z);
					edits[i].block = updates[i].cellType;
				}
				Check(tsom_sharded_apply_edits(m_planet, edits.data(), std::uint32_t(edits.size())), "tsom_sharded_apply_edits");
			}

			std::vector<ChunkIndices> CollectInvalidatedChunks()
			{
				std::uint32_t n = 0;
				Check(tsom_sharded_collect_dirty(m_planet, nullptr, 0, &n, 0), "tsom_sharded_collect_dirty");
				std::vector<ChunkIndices> out(n);
				Check(tsom_sharded_collect_dirty(m_planet, n ? &out[0].x : nullptr, n, &n, 1), "tsom_sharded_collect_dirty");
				return out;
			}

			// Chunk::BuildMesh for `chunks`, each on the GPU that owns it (all GPUs at once)
			ShardedMeshBatch BuildMeshes(const std::vector<ChunkIndices>& chunks, const MeshOptions& opt = MeshOptions())
			{
				ShardedMeshBatch batch;
				batch.m_planet = m_planet;
				batch.views.resize(chunks.size());
				batch.parts.resize(chunks.size());
				tsom_mesh_params p{};
				p.flags = (opt.render ? std::uint32_t(TSOM_MESH_RENDER) : 0u) | (opt.collider ? std::uint32_t(TSOM_MESH_COLLIDER) : 0u) | (opt.corners == CornerMode::Deformed ? std::uint32_t(TSOM_MESH_DEFORMED) : 0u) |
				          (opt.tangents ? std::uint32_t(TSOM_MESH_TANGENTS) : 0u);
				p.deform_radius = opt.deformRadius;
				p.gravity_centers = opt.gravityCenters ? &opt.gravityCenters->x : nullptr;
				if (!chunks.empty())
					Check(tsom_sharded_mesh(m_planet, &chunks[0].x, std::uint32_t(chunks.size()), &p, batch.views.data(), batch.parts.data()), "tsom_sharded_mesh");
				return batch;
			}

			std::vector<BlockIndex> GetContent(const ChunkIndices& indices) const
			{
				std::vector<BlockIndex> out(TSOM_CHUNK_BLOCKS);
				Check(tsom_sharded_get_content(m_planet, &indices.x, 1, out.data()), "tsom_sharded_get_content");
				return out;
			}

			tsom_sharded_planet* Handle() const { return m_planet; }

		private:
			tsom_sharded_planet* m_planet = nullptr;
			Vector3ui m_chunkCount;
			float m_tileSize;
	};

	// ------------------------------------------------------------------------------------------------ ChunkEntities
	// ChunkEntities::Update (src/CommonLib/ChunkEntities.cpp:73-112) + ClientChunkEntities::ProcessChunkUpdate
	// (src/ClientLib/ClientChunkEntities.cpp:178-262).  The reference turns every invalidated chunk into collider / mesh tasks and
	// applies a finished job only when the jobs of its neighbour dependencies have finished too, so that a chunk and its
	// neighbours never show meshes of different block states.  Here all chunks of a tick are rebuilt from ONE snapshot of the
	// block store by one launch, so the dependency rule holds by construction: `apply` sees every chunk of the tick at once.
	class ChunkEntities
	{
		public:
			using ApplyFunc = std::function<void(const ChunkIndices& chunkIndices, const MeshBatch& batch, std::size_t indexInBatch)>;

			ChunkEntities(ChunkContainer& container, MeshOptions options, ApplyFunc apply) :
			m_container(&container), m_options(options), m_apply(std::move(apply))
			{
			}

			// returns the number of chunks rebuilt this tick
			std::size_t Update()
			{
				std::vector<ChunkIndices> dirty = m_container->CollectInvalidatedChunks();
				if (dirty.empty())
					return 0;
				MeshBatch batch = m_container->BuildMeshes(dirty, m_options);
				for (std::size_t i = 0; i < dirty.size(); ++i)
					m_apply(dirty[i], batch, i); // faces == 0: the reference destroys the mesh / collider (null mesh, ClientChunkEntities.cpp:161-162)
				return dirty.size();
			}

		private:
			ChunkContainer* m_container;
			MeshOptions m_options;
			ApplyFunc m_apply;
	};

	// ------------------------------------------------------------------------------------------------ Chunk members
	inline void Chunk::BuildMesh(std::vector<std::uint32_t>& indices, const Vector3f& center, const std::function<VertexAttributes(std::uint32_t count)>& addVertices) const
	{
		MeshOptions opt;
		opt.gravityCenters = &center;
		Device::Lock lock(m_owner->GetDevice()); // callable from many workers at once, like the reference's (ClientChunkEntities.cpp:207-241)
		MeshBatch batch = m_owner->BuildMeshes({ m_indices }, opt);
		const std::uint32_t faces = batch.FaceCount(0);
		if (faces == 0)
			return;
		std::vector<VertexStruct> vertices;
		std::vector<std::uint32_t> local;
		batch.CopyChunk(0, &vertices, &local, nullptr);
		VertexAttributes attr = addVertices(faces * 4);
		auto at = [&](Vector3f* base, std::size_t i) { return reinterpret_cast<Vector3f*>(reinterpret_cast<char*>(base) + i * attr.stride); };
		for (std::size_t i = 0; i < vertices.size(); ++i)
		{
			if (attr.position)
				*at(attr.position, i) = vertices[i].position;
			if (attr.normal)
				*at(attr.normal, i) = vertices[i].normal;
			if (attr.uv)
				*at(attr.uv, i) = vertices[i].uvw;
		}
		indices.reserve(indices.size() + local.size());
		for (std::uint32_t v : local)
			indices.push_back(attr.firstIndex + v);
	}

	inline std::optional<Chunk::TriangleCollider> Chunk::BuildCollider() const
	{
		MeshOptions opt;
		opt.render = false;
		opt.collider = true;
		Device::Lock lock(m_owner->GetDevice());
		MeshBatch batch = m_owner->BuildMeshes({ m_indices }, opt);
		if (batch.FaceCount(0) == 0)
			return std::nullopt; // the reference returns nullptr for an empty mesh (DeformedChunk.cpp:33-34)
		TriangleCollider c;
		batch.CopyChunk(0, nullptr, &c.indices, &c.positions);
		return c;
	}

	inline std::vector<Chunk::Box> Chunk::BuildBoxCollider() const
	{
		const std::int32_t idx[3] = { m_indices.x, m_indices.y, m_indices.z };
		std::uint32_t n = 0;
		std::vector<float> raw(16384 * 6);
		Check(tsom_box_collider(m_owner->Handle(), idx, raw.data(), 16384, &n), "tsom_box_collider");
		const float s = m_owner->GetTileSize();
		std::vector<Box> out(n);
		for (std::uint32_t i = 0; i < n; ++i)
		{
			out[i].position = Vector3f(raw[i * 6 + 0], raw[i * 6 + 1], raw[i * 6 + 2]) * s;
			out[i].size = Vector3f(raw[i * 6 + 3], raw[i * 6 + 4], raw[i * 6 + 5]) * s;
		}
		return out;
	}

	inline std::vector<BlockIndex> Chunk::GetContent() const
	{
		std::vector<BlockIndex> blocks(TSOM_CHUNK_BLOCKS);
		Check(tsom_chunks_get_content(m_owner->Handle(), &m_indices.x, 1, blocks.data()), "tsom_chunks_get_content");
		return blocks;
	}

	inline bool Chunk::HasContent() const
	{
		const std::int32_t idx[3] = { m_indices.x, m_indices.y, m_indices.z };
		return tsom_chunk_device_ptr(m_owner->Handle(), idx) != nullptr;
	}

	template<typename F> void Chunk::Reset(F&& func)
	{
		std::vector<BlockIndex> blocks(TSOM_CHUNK_BLOCKS, EmptyBlockIndex);
		func(blocks.data());
		Check(tsom_chunks_reset(m_owner->Handle(), &m_indices.x, 1, blocks.data()), "tsom_chunks_reset");
	}

	inline std::vector<std::uint8_t> Chunk::Serialize(const BlockLibrary& blockLibrary) const
	{
		const std::uint32_t stride = std::uint32_t(blockLibrary.GetBlockCount());
		std::vector<std::uint16_t> palette(stride);
		std::vector<std::uint8_t> payload(TSOM_CHUNK_BLOCKS * 2);
		std::uint32_t count = 0;
		Check(tsom_chunks_palettize(m_owner->Handle(), &m_indices.x, 1, palette.data(), stride, &count, payload.data(), 1), "tsom_chunks_palettize");
		std::vector<std::uint8_t> out;
		auto put32 = [&](std::uint32_t v) { for (int s = 24; s >= 0; s -= 8) out.push_back(std::uint8_t(v >> s)); };
		put32(1); // Constants::ChunkBinaryVersion (InternalConstants.hpp:24)
		put32(TSOM_CHUNK_SIZE); put32(TSOM_CHUNK_SIZE); put32(TSOM_CHUNK_SIZE);
		out.push_back(std::uint8_t(count >> 8));
		out.push_back(std::uint8_t(count));
		for (std::uint32_t i = 0; i < count; ++i)
		{
			const std::string& name = blockLibrary.GetBlockData(palette[i]).name;
			put32(std::uint32_t(name.size()));
			out.insert(out.end(), name.begin(), name.end());
		}
		out.insert(out.end(), payload.begin(), payload.begin() + TSOM_CHUNK_BLOCKS * (count > 8 ? 2 : 1));
		return out;
	}

	inline std::array<Vector3f, 8> Chunk::ComputeVoxelCorners(const Vector3ui& indices, CornerMode corners, float deformRadius, const Vector3f* deformCenter) const
	{
		tsom_mesh_params params{};
		params.flags = corners == CornerMode::Deformed ? std::uint32_t(TSOM_MESH_DEFORMED) : 0u;
		params.deform_radius = deformRadius;
		params.gravity_centers = deformCenter ? &deformCenter->x : nullptr;
		const std::uint32_t local[3] = { indices.x, indices.y, indices.z };
		std::array<Vector3f, 8> out;
		static_assert(sizeof(out) == 24 * sizeof(float), "corner layout");
		Check(tsom_voxel_corners(m_owner->Handle(), &m_indices.x, local, 1, &params, &out[0].x), "tsom_voxel_corners");
		return out;
	}

	inline std::vector<std::uint8_t> Chunk::Compress() const
	{
		std::vector<std::uint8_t> out(TSOM_CHUNK_BLOCKS * 2 + TSOM_CHUNK_BLOCKS * 2 / 255 + 16); // LZ4_compressBound
		std::uint64_t offsets[2] = { 0, 0 };
		Check(tsom_chunks_lz4_compress(m_owner->Handle(), &m_indices.x, 1, out.data(), out.size(), offsets), "tsom_chunks_lz4_compress");
		out.resize(std::size_t(offsets[1]));
		return out;
	}

	inline void Chunk::ResetCompressed(const std::uint8_t* data, std::size_t size)
	{
		const std::uint64_t offsets[2] = { 0, size };
		if (tsom_chunks_reset_lz4(m_owner->Handle(), &m_indices.x, 1, data, offsets) != TSOM_OK)
			throw std::runtime_error(tsom_last_error());
	}

	inline void Chunk::Deserialize(const BlockLibrary& blockLibrary, const std::uint8_t* data, std::size_t size)
	{
		std::size_t cur = 0;
		auto need = [&](std::size_t k) { if (cur + k > size) throw std::runtime_error("ByteStream: read past end"); };
		auto get32 = [&]() { need(4); std::uint32_t v = std::uint32_t(data[cur]) << 24 | std::uint32_t(data[cur + 1]) << 16 | std::uint32_t(data[cur + 2]) << 8 | data[cur + 3]; cur += 4; return v; };
		if (get32() != 1)
			throw std::runtime_error("incompatible chunk version");
		const std::uint32_t sx = get32(), sy = get32(), sz = get32();
		if (sx != TSOM_CHUNK_SIZE || sy != TSOM_CHUNK_SIZE || sz != TSOM_CHUNK_SIZE)
			throw std::runtime_error("incompatible chunk size");
		need(2);
		const std::uint32_t count = std::uint32_t(data[cur]) << 8 | data[cur + 1];
		cur += 2;
		std::vector<std::uint16_t> palette(std::max<std::uint32_t>(count, 1));
		for (std::uint32_t i = 0; i < count; ++i)
		{
			const std::uint32_t len = get32();
			need(len);
			std::string name(reinterpret_cast<const char*>(data + cur), len);
			cur += len;
			const BlockIndex idx = blockLibrary.GetBlockIndex(name);
			if (idx == InvalidBlockIndex)
				throw std::runtime_error("unknown block " + name);
			palette[i] = idx;
		}
		const std::size_t bytes = std::size_t(TSOM_CHUNK_BLOCKS) * (count > 8 ? 2 : 1);
		need(bytes);
		std::vector<std::uint8_t> payload(TSOM_CHUNK_BLOCKS * 2, 0);
		std::memcpy(payload.data(), data + cur, bytes);
		Check(tsom_chunks_reset_palettized(m_owner->Handle(), &m_indices.x, 1, palette.data(), std::uint32_t(palette.size()), &count, payload.data(), 1), "tsom_chunks_reset_palettized");
	}

	inline void Chunk::UpdateBlock(const Vector3ui& indices, BlockIndex cellType)
	{
		m_owner->UpdateBlocks({ ChunkContainer::BlockUpdate{ m_indices, indices, cellType } });
	}
}
