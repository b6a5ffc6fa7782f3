// oracle/ref_shim/ref_capi.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// C entry points (same names and meaning as oracle/capi.cpp) over THE REFERENCE'S OWN SOURCES: src/CommonLib/{Chunk,FlatChunk,
// DeformedChunk,Planet,Ship,BlockLibrary,ChunkContainer,GravityController}.cpp, compiled unmodified from /root/reference against the
// stand-in headers of oracle/ref_shim/include (NazaraEngine, siv::PerlinNoise, tsl, fmt are not on this box).
// What this pins: the restated oracle's STRUCTURE — loop order, neighbour rules, RNG consumption, pass order, platform script,
// greedy merge, mask rule — against the reference's real code.  What it cannot pin: the arithmetic the shim itself restates
// (Box corner naming, Quaternion::RotationBetween, Vector3::Normalize, PerlinNoise) — SURVEY.md §8c.
#include <CommonLib/BlockLibrary.hpp>
#include <CommonLib/Chunk.hpp>
#include <CommonLib/DeformedChunk.hpp>
#include <CommonLib/FlatChunk.hpp>
#include <CommonLib/Planet.hpp>
#include <CommonLib/Ship.hpp>
#include <Nazara/Physics3D/Collider3D.hpp>
#include <Nazara/Core/ByteStream.hpp>
#include <Nazara/Core/TaskScheduler.hpp>

#include "../nazara_tangents.hpp"
#include <atomic>
#include <chrono>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>

using namespace tsom;

namespace
{
	constexpr std::size_t ChunkBlockCount = 32 * 32 * 32;

	std::uint32_t g_registered = 0; // blocks added through orc_blocklib_register

	BlockLibrary& MutableLib()
	{
		static BlockLibrary lib;
		return lib;
	}
	const BlockLibrary& Lib() { return MutableLib(); }

	struct Key
	{
		int x, y, z;
		bool operator<(const Key& o) const { return std::tie(x, y, z) < std::tie(o.x, o.y, o.z); }
	};

	struct PlanetHandle
	{
		Planet planet;
		std::mutex mutex;
		std::map<Key, std::uint32_t> invalidated; // what ChunkEntities keeps in m_invalidatedChunks (ChunkEntities.cpp:43-48)
		NazaraSlot(ChunkContainer, OnChunkUpdated, onChunkUpdated);

		PlanetHandle(float tile, float radius) : planet(tile, radius, 9.81f)
		{
			onChunkUpdated.Connect(planet.OnChunkUpdated, [this](ChunkContainer*, Chunk* chunk, DirectionMask mask) {
				std::lock_guard<std::mutex> lock(mutex);
				const ChunkIndices& i = chunk->GetIndices();
				invalidated[Key{ i.x, i.y, i.z }] |= std::uint32_t(mask);
			});
		}
	};

	template<typename F> void ParallelFor(std::size_t n, int nthreads, F&& f)
	{
		Nz::TaskScheduler sched(nthreads <= 1 ? 1u : unsigned(nthreads));
		for (std::size_t i = 0; i < n; ++i)
			sched.AddTask([&f, i] { f(i); });
		sched.WaitForTasks();
	}

	double Now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

	struct MeshOut
	{
		std::vector<Nz::Vector3f> position, normal, uvw;
		std::vector<Nz::UInt32> indices;
	};

	// Chunk::BuildMesh through the reference's own VertexAttributes contract (null SparsePtr = "do not compute")
	void BuildMeshInto(const Chunk& chunk, const Nz::Vector3f& center, bool wantNormal, bool wantUv, MeshOut& out)
	{
		auto AddVertices = [&](Nz::UInt32 count) {
			Chunk::VertexAttributes a;
			a.firstIndex = Nz::UInt32(out.position.size());
			out.position.resize(out.position.size() + count);
			a.position = Nz::SparsePtr<Nz::Vector3f>(&out.position[a.firstIndex]);
			if (wantNormal)
			{
				out.normal.resize(out.position.size());
				a.normal = Nz::SparsePtr<Nz::Vector3f>(&out.normal[a.firstIndex]);
			}
			if (wantUv)
			{
				out.uvw.resize(out.position.size());
				a.uv = Nz::SparsePtr<Nz::Vector3f>(&out.uvw[a.firstIndex]);
			}
			return a;
		};
		chunk.BuildMesh(out.indices, center, AddVertices);
	}
}

extern "C"
{
	const char* orc_backend() { return "reference sources over oracle/ref_shim"; }

	void orc_get_block_indices(const std::int32_t chunk[3], const std::uint32_t local[3], std::int32_t out[3])
	{
		Planet planet(1.f, 0.f, 9.81f); // as tests/UnitTests/Common/ChunkTests.cpp:13 constructs it
		BlockIndices r = planet.GetBlockIndices({ chunk[0], chunk[1], chunk[2] }, { local[0], local[1], local[2] });
		out[0] = r.x; out[1] = r.y; out[2] = r.z;
	}

	void orc_get_chunk_indices_by_block_indices(const std::int32_t block[3], std::int32_t chunkOut[3], std::uint32_t localOut[3])
	{
		Planet planet(1.f, 0.f, 9.81f);
		Nz::Vector3ui l;
		ChunkIndices c = planet.GetChunkIndicesByBlockIndices({ block[0], block[1], block[2] }, &l);
		chunkOut[0] = c.x; chunkOut[1] = c.y; chunkOut[2] = c.z;
		localOut[0] = l.x; localOut[1] = l.y; localOut[2] = l.z;
	}

	int orc_direction_from_normal(const float n[3]) { return int(DirectionFromNormal({ n[0], n[1], n[2] })); }

	// BlockLibrary has no size accessor: count the built-in names (14 types, BlockLibrary.cpp:10-113) that resolve
	std::uint32_t orc_blocklib_count()
	{
		static const char* names[] = { "empty", "debug", "dirt", "grass", "hull", "hull2", "snow", "stone", "stone_mossy", "forcefield", "planks", "stone_bricks", "copper_block", "glass" };
		std::uint32_t n = 0;
		for (const char* nm : names)
			if (Lib().GetBlockIndex(nm) != InvalidBlockIndex)
				++n;
		return n + g_registered;
	}

	void orc_blocklib_get(std::uint32_t idx, std::uint32_t tex[6], std::uint8_t flags[3], char name[64])
	{
		const auto& d = Lib().GetBlockData(BlockIndex(idx));
		for (int i = 0; i < 6; ++i)
			tex[i] = d.texIndices[Direction(i)];
		flags[0] = d.hasCollisions; flags[1] = d.isDoubleSided; flags[2] = d.isTransparent;
		std::strncpy(name, d.name.c_str(), 63);
		name[63] = 0;
	}

	int orc_blocklib_index(const char* name) { return Lib().GetBlockIndex(name); }

	// BlockLibrary::RegisterBlock (BlockLibrary.cpp:85-131) on the process-wide library: appends one block type (tests of large
	// block tables run in their own process).  Empty strings mean "not given".
	int orc_blocklib_register(const char* name, const char* basePath, const char* sidePath, const char* upPath, int hasCollisions, int isDoubleSided, int isTransparent)
	{
		BlockLibrary::BlockInfo info;
		info.basePath = basePath;
		info.baseSidePath = sidePath;
		info.baseUpPath = upPath;
		info.hasCollisions = hasCollisions != 0;
		info.isDoubleSided = isDoubleSided != 0;
		info.isTransparent = isTransparent != 0;
		++g_registered;
		return MutableLib().RegisterBlock(name, std::move(info));
	}

	int orc_generate_chunk(std::uint32_t seed, const std::uint32_t count[3], const std::int32_t chunk[3], std::uint16_t* out)
	{
		Planet planet(1.f, 16.f, 9.81f);
		Chunk& c = planet.AddChunk(Lib(), { chunk[0], chunk[1], chunk[2] });
		try
		{
			planet.GenerateChunk(Lib(), c, seed, { count[0], count[1], count[2] });
		}
		catch (const Nz::SafeCastError&)
		{
			return 1;
		}
		std::memcpy(out, c.GetContent(), ChunkBlockCount * sizeof(BlockIndex));
		return 0;
	}

	void* orc_planet_create(float tile, float cornerRadius) { return new PlanetHandle(tile, cornerRadius); }
	void orc_planet_destroy(void* h) { delete static_cast<PlanetHandle*>(h); }

	int orc_planet_add_chunk(void* h, const std::int32_t idx[3], const std::uint16_t* blocks)
	{
		Planet& p = static_cast<PlanetHandle*>(h)->planet;
		ChunkIndices ci{ idx[0], idx[1], idx[2] };
		if (p.GetChunk(ci))
			return 1;
		Chunk& c = p.AddChunk(Lib(), ci);
		if (blocks)
			c.Reset([&](BlockIndex* dst) { std::memcpy(dst, blocks, ChunkBlockCount * sizeof(BlockIndex)); });
		return 0;
	}

	int orc_planet_reset_chunk(void* h, const std::int32_t idx[3], const std::uint16_t* blocks)
	{
		Chunk* c = static_cast<PlanetHandle*>(h)->planet.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c)
			return 1;
		c->Reset([&](BlockIndex* dst) { std::memcpy(dst, blocks, ChunkBlockCount * sizeof(BlockIndex)); });
		return 0;
	}

	// Planet::GenerateChunks itself (Planet.cpp:385-403), its tasks run by the shim TaskScheduler on nthreads workers
	int orc_planet_generate(void* h, std::uint32_t seed, const std::uint32_t count[3], int nthreads)
	{
		Planet& p = static_cast<PlanetHandle*>(h)->planet;
		for (int cz = 0; cz < int(count[2]); ++cz)
			for (int cy = 0; cy < int(count[1]); ++cy)
				for (int cx = 0; cx < int(count[0]); ++cx)
					if (p.GetChunk({ cx - int(count[0] / 2), cy - int(count[1] / 2), cz - int(count[2] / 2) }))
						return 1; // Planet::AddChunk asserts that the chunk does not exist yet (Planet.cpp:30)
		Nz::TaskScheduler sched(nthreads <= 1 ? 1u : unsigned(nthreads));
		try
		{
			p.GenerateChunks(Lib(), sched, seed, { count[0], count[1], count[2] });
		}
		catch (const Nz::SafeCastError&)
		{
			return 1;
		}
		return 0;
	}

	void orc_planet_generate_platform(void* h, int dir, const std::int32_t center[3])
	{
		static_cast<PlanetHandle*>(h)->planet.GeneratePlatform(Lib(), Direction(dir), { center[0], center[1], center[2] });
	}

	int orc_planet_get_blocks(void* h, const std::int32_t idx[3], std::uint16_t* out)
	{
		const Chunk* c = static_cast<PlanetHandle*>(h)->planet.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c || !c->HasContent())
			return 1;
		std::memcpy(out, c->GetContent(), ChunkBlockCount * sizeof(BlockIndex));
		return 0;
	}

	int orc_planet_update_block(void* h, const std::int32_t idx[3], const std::uint32_t local[3], std::uint16_t id)
	{
		Chunk* c = static_cast<PlanetHandle*>(h)->planet.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c || !c->HasContent())
			return 1;
		c->UpdateBlock({ local[0], local[1], local[2] }, id);
		return 0;
	}

	void orc_planet_apply_edits(void* h, const std::int32_t* edits, std::uint32_t n)
	{
		Planet& p = static_cast<PlanetHandle*>(h)->planet;
		for (std::uint32_t i = 0; i < n; ++i)
		{
			const std::int32_t* e = edits + std::size_t(i) * 7;
			Chunk* c = p.GetChunk({ e[0], e[1], e[2] });
			if (!c || !c->HasContent())
				continue;
			c->UpdateBlock({ unsigned(e[3]), unsigned(e[4]), unsigned(e[5]) }, BlockIndex(e[6]));
		}
	}

	void orc_planet_clear_invalidated(void* h) { static_cast<PlanetHandle*>(h)->invalidated.clear(); }

	// ChunkEntities::Update + ClientChunkEntities::ProcessChunkUpdate (harness logic: those files need entt/Jolt and are not
	// compiled here); the masks themselves come from the reference's OnBlockUpdated handler (Planet.cpp:39-58).
	std::uint32_t orc_planet_collect_dirty(void* hv, std::int32_t* out, std::uint32_t cap)
	{
		PlanetHandle* h = static_cast<PlanetHandle*>(hv);
		std::map<Key, bool> dirty;
		for (auto&& [idx, mask] : h->invalidated)
		{
			const Chunk* c = h->planet.GetChunk({ idx.x, idx.y, idx.z });
			if (!c || !c->HasContent())
				continue;
			dirty[idx] = true;
			for (int d = 0; d < 6; ++d)
			{
				if (!(mask & (1u << d)))
					continue;
				const Nz::Vector3i& o = s_chunkDirOffset[Direction(d)];
				const Chunk* nc = h->planet.GetChunk({ idx.x + o.x, idx.y + o.y, idx.z + o.z });
				if (!nc || !nc->HasContent())
					continue;
				dirty[Key{ idx.x + o.x, idx.y + o.y, idx.z + o.z }] = true;
			}
		}
		h->invalidated.clear();
		std::uint32_t i = 0;
		for (auto&& [idx, _] : dirty)
		{
			if (i < cap)
			{
				out[i * 3 + 0] = idx.x; out[i * 3 + 1] = idx.y; out[i * 3 + 2] = idx.z;
			}
			++i;
		}
		return i;
	}

	struct MeshHandle { MeshOut out; };

	void* orc_mesh_build(void* hv, const std::int32_t idx[3], int cornerMode, const float* gravityCenter, const float* deformCenter, float deformRadius, int wantNormal, int wantUv)
	{
		Planet& p = static_cast<PlanetHandle*>(hv)->planet;
		Chunk* c = p.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c || !c->HasContent())
			return nullptr;

		Nz::Vector3f center = p.GetCenter() - p.GetChunkOffset(c->GetIndices()); // ClientChunkEntities.cpp:160
		if (gravityCenter)
			center = Nz::Vector3f(gravityCenter[0], gravityCenter[1], gravityCenter[2]);

		auto* m = new MeshHandle;
		if (!cornerMode)
			BuildMeshInto(*c, center, wantNormal != 0, wantUv != 0, m->out);
		else
		{
			// Planet::AddChunk only instantiates FlatChunk in this snapshot (Planet.cpp:32); a DeformedChunk with the same
			// owner (hence the same neighbours) and the same content exercises DeformedChunk::ComputeVoxelCorners / DeformPosition.
			Nz::Vector3f dc = deformCenter ? Nz::Vector3f(deformCenter[0], deformCenter[1], deformCenter[2]) : center;
			auto deformed = std::make_shared<DeformedChunk>(Lib(), p, c->GetIndices(), Nz::Vector3ui{ Planet::ChunkSize }, p.GetTileSize(), dc, deformRadius);
			deformed->Reset([&](BlockIndex* dst) { std::memcpy(dst, c->GetContent(), ChunkBlockCount * sizeof(BlockIndex)); });
			BuildMeshInto(*deformed, center, wantNormal != 0, wantUv != 0, m->out);
		}
		return m;
	}

	// the reference's own FlatChunk::ComputeVoxelCorners / DeformedChunk::ComputeVoxelCorners: 8 corners x xyz in Nz::BoxCorner order
	int orc_voxel_corners(void* hv, const std::int32_t idx[3], const std::uint32_t local[3], int cornerMode, const float* deformCenter, float deformRadius, float out[24])
	{
		Planet& p = static_cast<PlanetHandle*>(hv)->planet;
		Chunk* c = p.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c)
			return 1;
		Nz::EnumArray<Nz::BoxCorner, Nz::Vector3f> corners;
		if (!cornerMode)
			corners = c->ComputeVoxelCorners(Nz::Vector3ui{ local[0], local[1], local[2] });
		else
		{
			const Nz::Vector3f center = p.GetCenter() - p.GetChunkOffset(c->GetIndices());
			Nz::Vector3f dc = deformCenter ? Nz::Vector3f(deformCenter[0], deformCenter[1], deformCenter[2]) : center;
			DeformedChunk deformed(Lib(), p, c->GetIndices(), Nz::Vector3ui{ Planet::ChunkSize }, p.GetTileSize(), dc, deformRadius);
			corners = deformed.ComputeVoxelCorners(Nz::Vector3ui{ local[0], local[1], local[2] });
		}
		int k = 0;
		for (const Nz::Vector3f& v : corners)
		{
			out[k++] = v.x;
			out[k++] = v.y;
			out[k++] = v.z;
		}
		return 0;
	}

	std::uint32_t orc_mesh_face_count(void* m) { return std::uint32_t(static_cast<MeshHandle*>(m)->out.indices.size() / 6); }

	void orc_mesh_copy(void* mh, float* pos, float* normal, float* uvw, std::uint32_t* indices)
	{
		const MeshOut& o = static_cast<MeshHandle*>(mh)->out;
		if (pos) std::memcpy(pos, o.position.data(), o.position.size() * sizeof(Nz::Vector3f));
		if (normal) std::memcpy(normal, o.normal.data(), o.normal.size() * sizeof(Nz::Vector3f));
		if (uvw) std::memcpy(uvw, o.uvw.data(), o.uvw.size() * sizeof(Nz::Vector3f));
		if (indices) std::memcpy(indices, o.indices.data(), o.indices.size() * sizeof(std::uint32_t));
	}

	void orc_mesh_free(void* m) { delete static_cast<MeshHandle*>(m); }

	// FlatChunk::BuildCollider(dims, mask, callback) (FlatChunk.cpp:56-153): out = n x {pos.xyz, size.xyz} in block units
	std::uint32_t orc_box_collider(void* hv, const std::int32_t idx[3], float* out, std::uint32_t cap)
	{
		const Chunk* c = static_cast<PlanetHandle*>(hv)->planet.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c || !c->HasContent())
			return 0;
		std::uint32_t n = 0;
		FlatChunk::BuildCollider(c->GetSize(), c->GetCollisionCellMask(), [&](const Nz::Boxf& box) {
			if (n < cap)
			{
				float* o = out + std::size_t(n) * 6;
				o[0] = box.x; o[1] = box.y; o[2] = box.z;
				o[3] = box.width; o[4] = box.height; o[5] = box.depth;
			}
			++n;
		});
		return n;
	}

	// Chunk::Serialize (Chunk.cpp:312-346) of one chunk -> bytes written (0 if the buffer is too small or the chunk is unknown)
	std::uint32_t orc_chunk_serialize(void* hv, const std::int32_t idx[3], std::uint8_t* out, std::uint32_t cap)
	{
		const Chunk* c = static_cast<PlanetHandle*>(hv)->planet.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c || !c->HasContent())
			return 0;
		Nz::ByteStream bs;
		c->Serialize(bs);
		if (bs.data.size() > cap)
			return 0;
		std::memcpy(out, bs.data.data(), bs.data.size());
		return std::uint32_t(bs.data.size());
	}

	// Chunk::Deserialize (Chunk.cpp:252-310) into an existing chunk: 0 ok, 1 unknown chunk, 2 rejected (message in err)
	int orc_chunk_deserialize(void* hv, const std::int32_t idx[3], const std::uint8_t* data, std::uint32_t size, char err[128])
	{
		Chunk* c = static_cast<PlanetHandle*>(hv)->planet.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c)
			return 1;
		try
		{
			Nz::ByteStream bs;
			bs.data.assign(data, data + size);
			c->Deserialize(bs);
		}
		catch (const std::exception& e)
		{
			std::strncpy(err, e.what(), 127);
			err[127] = 0;
			return 2;
		}
		return 0;
	}

	// ---- CPU baseline timers: the reference's own Chunk::BuildMesh / Planet::GenerateChunks, one task per chunk
	double orc_bench_mesh_isolated(const std::uint16_t* blocks, std::uint32_t n, int nthreads, std::uint64_t* facesOut)
	{
		std::vector<std::unique_ptr<PlanetHandle>> planets(n);
		for (std::uint32_t i = 0; i < n; ++i)
		{
			planets[i] = std::make_unique<PlanetHandle>(1.f, 16.f);
			std::int32_t idx[3] = { 0, 0, 0 };
			orc_planet_add_chunk(planets[i].get(), idx, blocks + std::size_t(i) * ChunkBlockCount);
		}
		std::atomic<std::uint64_t> faces{ 0 };
		double t0 = Now();
		ParallelFor(n, nthreads, [&](std::size_t i) {
			Planet& p = planets[i]->planet;
			Chunk* c = p.GetChunk({ 0, 0, 0 });
			MeshOut out;
			BuildMeshInto(*c, p.GetCenter() - p.GetChunkOffset(c->GetIndices()), true, true, out);
			faces += out.indices.size() / 6;
		});
		double t1 = Now();
		if (facesOut)
			*facesOut = faces;
		return t1 - t0;
	}

	int orc_bench_planet(std::uint32_t seed, const std::uint32_t count[3], int nthreads, std::uint32_t stride, double times[2], std::uint64_t* facesOut, std::uint32_t* meshedOut)
	{
		PlanetHandle h(1.f, 16.f);
		double t0 = Now();
		int bad = orc_planet_generate(&h, seed, count, nthreads);
		double t1 = Now();
		std::vector<Chunk*> chunks;
		std::size_t k = 0;
		for (int cz = 0; cz < int(count[2]); ++cz)
			for (int cy = 0; cy < int(count[1]); ++cy)
				for (int cx = 0; cx < int(count[0]); ++cx, ++k)
					if (k % stride == 0)
						chunks.push_back(h.planet.GetChunk({ cx - int(count[0] / 2), cy - int(count[1] / 2), cz - int(count[2] / 2) }));
		std::atomic<std::uint64_t> faces{ 0 };
		double t2 = Now();
		ParallelFor(chunks.size(), nthreads, [&](std::size_t i) {
			MeshOut out;
			BuildMeshInto(*chunks[i], h.planet.GetCenter() - h.planet.GetChunkOffset(chunks[i]->GetIndices()), true, true, out);
			faces += out.indices.size() / 6;
		});
		double t3 = Now();
		times[0] = t1 - t0;
		times[1] = t3 - t2;
		if (facesOut)
			*facesOut = faces;
		if (meshedOut)
			*meshedOut = std::uint32_t(chunks.size());
		return bad;
	}

	// ---- a bounded, planet-wide sample of "generate + mesh" for the CPU arm of bench.py -----------------------------------------
	// Every stride-th chunk of the planet (GenerateChunks order, starting at `offset`) is the sample; its face-adjacent neighbours
	// are generated once at set-up so that the sample meshes with real halos.  A step regenerates the sample chunks with the
	// reference's own Planet::GenerateChunk and meshes them with Chunk::BuildMesh, one task per chunk like the reference's callers.
	struct SampleHandle
	{
		PlanetHandle h{ 1.f, 16.f };
		std::vector<Chunk*> sample;
		std::uint32_t seed = 0;
		Nz::Vector3ui count;
	};

	void* orc_sample_create(std::uint32_t seed, const std::uint32_t count[3], std::uint32_t stride, std::uint32_t offset, int nthreads)
	{
		auto s = new SampleHandle;
		s->seed = seed;
		s->count = { count[0], count[1], count[2] };
		Planet& p = s->h.planet;
		auto inside = [&](const ChunkIndices& c) {
			return c.x >= -int(count[0] / 2) && c.x < int(count[0]) - int(count[0] / 2) && c.y >= -int(count[1] / 2) && c.y < int(count[1]) - int(count[1] / 2) &&
			       c.z >= -int(count[2] / 2) && c.z < int(count[2]) - int(count[2] / 2);
		};
		std::vector<Chunk*> all;
		std::size_t k = 0;
		for (int cz = 0; cz < int(count[2]); ++cz)
			for (int cy = 0; cy < int(count[1]); ++cy)
				for (int cx = 0; cx < int(count[0]); ++cx, ++k)
				{
					if (k % stride != offset % stride)
						continue;
					const ChunkIndices ci{ cx - int(count[0] / 2), cy - int(count[1] / 2), cz - int(count[2] / 2) };
					Chunk* c = p.GetChunk(ci);
					if (!c)
					{
						c = &p.AddChunk(Lib(), ci);
						all.push_back(c);
					}
					s->sample.push_back(c);
					for (int d = 0; d < 6; ++d)
					{
						const ChunkIndices ni = ci + s_chunkDirOffset[Direction(d)];
						if (inside(ni) && !p.GetChunk(ni))
							all.push_back(&p.AddChunk(Lib(), ni));
					}
				}
		try
		{
			ParallelFor(all.size(), nthreads, [&](std::size_t i) { p.GenerateChunk(Lib(), *all[i], seed, s->count); });
		}
		catch (const Nz::SafeCastError&)
		{
			delete s;
			return nullptr;
		}
		return s;
	}

	std::uint32_t orc_sample_size(void* hv) { return std::uint32_t(static_cast<SampleHandle*>(hv)->sample.size()); }

	// times[0] = generate seconds, times[1] = mesh seconds of the sample; returns its face count
	std::uint64_t orc_sample_step(void* hv, int nthreads, double times[2])
	{
		auto* s = static_cast<SampleHandle*>(hv);
		Planet& p = s->h.planet;
		const double t0 = Now();
		ParallelFor(s->sample.size(), nthreads, [&](std::size_t i) { p.GenerateChunk(Lib(), *s->sample[i], s->seed, s->count); });
		const double t1 = Now();
		std::atomic<std::uint64_t> faces{ 0 };
		ParallelFor(s->sample.size(), nthreads, [&](std::size_t i) {
			MeshOut out;
			BuildMeshInto(*s->sample[i], p.GetCenter() - p.GetChunkOffset(s->sample[i]->GetIndices()), true, true, out);
			faces += out.indices.size() / 6;
		});
		const double t2 = Now();
		times[0] = t1 - t0;
		times[1] = t2 - t1;
		return faces;
	}

	void orc_sample_destroy(void* hv) { delete static_cast<SampleHandle*>(hv); }

	// FlatChunk::BuildCollider for every stride-th chunk of a generated planet, one task per chunk: seconds (boxes counted)
	double orc_bench_box_colliders(void* hv, const std::uint32_t count[3], std::uint32_t stride, int nthreads, std::uint64_t* boxesOut, std::uint32_t* chunksOut)
	{
		Planet& p = static_cast<PlanetHandle*>(hv)->planet;
		std::vector<const Chunk*> chunks;
		std::size_t k = 0;
		for (int cz = 0; cz < int(count[2]); ++cz)
			for (int cy = 0; cy < int(count[1]); ++cy)
				for (int cx = 0; cx < int(count[0]); ++cx, ++k)
					if (k % stride == 0)
						if (const Chunk* c = p.GetChunk({ cx - int(count[0] / 2), cy - int(count[1] / 2), cz - int(count[2] / 2) }))
							chunks.push_back(c);
		std::atomic<std::uint64_t> boxes{ 0 };
		const double t0 = Now();
		ParallelFor(chunks.size(), nthreads, [&](std::size_t i) {
			auto col = chunks[i]->BuildCollider();
			if (auto* cc = dynamic_cast<Nz::CompoundCollider3D*>(col.get()))
				boxes += cc->GetGeoms().size();
		});
		const double t1 = Now();
		if (boxesOut)
			*boxesOut = boxes;
		if (chunksOut)
			*chunksOut = std::uint32_t(chunks.size());
		return t1 - t0;
	}

	// ---- Ship (include/CommonLib/Ship.hpp, src/CommonLib/Ship.cpp): the second ChunkContainer ------------------------------------
	struct ShipHandle
	{
		Ship ship;
		std::map<Key, std::uint32_t> invalidated;
		NazaraSlot(ChunkContainer, OnChunkUpdated, onChunkUpdated);

		explicit ShipHandle(float tile) : ship(tile)
		{
			onChunkUpdated.Connect(ship.OnChunkUpdated, [this](ChunkContainer*, Chunk* chunk, DirectionMask mask) {
				const ChunkIndices& i = chunk->GetIndices();
				invalidated[Key{ i.x, i.y, i.z }] |= std::uint32_t(mask) | 0x80u; // bit 7: the chunk itself is in the set even with an empty mask
			});
		}
	};

	void* orc_ship_create(float tile) { return new ShipHandle(tile); }
	void orc_ship_destroy(void* h) { delete static_cast<ShipHandle*>(h); }
	void orc_ship_generate(void* h, int small) { static_cast<ShipHandle*>(h)->ship.Generate(Lib(), small != 0); }

	int orc_ship_add_chunk(void* h, const std::int32_t idx[3], const std::uint16_t* blocks)
	{
		Ship& s = static_cast<ShipHandle*>(h)->ship;
		ChunkIndices ci{ idx[0], idx[1], idx[2] };
		if (s.GetChunk(ci))
			return 1;
		if (blocks)
			s.AddChunk(Lib(), ci, [&](BlockIndex* dst) { std::memcpy(dst, blocks, ChunkBlockCount * sizeof(BlockIndex)); });
		else
			s.AddChunk(Lib(), ci);
		return 0;
	}

	int orc_ship_get_blocks(void* h, const std::int32_t idx[3], std::uint16_t* out)
	{
		const Chunk* c = static_cast<ShipHandle*>(h)->ship.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c || !c->HasContent())
			return 1;
		std::memcpy(out, c->GetContent(), ChunkBlockCount * sizeof(BlockIndex));
		return 0;
	}

	int orc_ship_update_block(void* h, const std::int32_t idx[3], const std::uint32_t local[3], std::uint16_t id)
	{
		Chunk* c = static_cast<ShipHandle*>(h)->ship.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c || !c->HasContent())
			return 1;
		c->UpdateBlock({ local[0], local[1], local[2] }, id);
		return 0;
	}

	// the (chunk, mask | 0x80) pairs collected since the last call, sorted by chunk; cleared
	std::uint32_t orc_ship_take_invalidated(void* h, std::int32_t* idxOut, std::uint32_t* maskOut, std::uint32_t cap)
	{
		auto* s = static_cast<ShipHandle*>(h);
		std::uint32_t n = 0;
		for (auto&& [k, m] : s->invalidated)
		{
			if (n < cap)
			{
				idxOut[n * 3 + 0] = k.x;
				idxOut[n * 3 + 1] = k.y;
				idxOut[n * 3 + 2] = k.z;
				maskOut[n] = m;
			}
			++n;
		}
		s->invalidated.clear();
		return n;
	}

	// Ship::BuildHullCollider (Ship.cpp:63-93), flattened: per box 9 floats = child offset (GetChunkOffset, zero for the single-chunk form),
	// box centre * blockSize, box lengths * blockSize.  Returns the number of boxes (may exceed cap); 0 = null collider.
	std::uint32_t orc_ship_hull_collider(void* h, float* out, std::uint32_t cap)
	{
		std::shared_ptr<Nz::Collider3D> hull = static_cast<ShipHandle*>(h)->ship.BuildHullCollider();
		auto* top = dynamic_cast<Nz::CompoundCollider3D*>(hull.get());
		if (!top)
			return 0;
		std::uint32_t n = 0;
		auto emit = [&](const Nz::Vector3f& chunkOffset, const Nz::CompoundCollider3D::ChildCollider& box) {
			auto* b = dynamic_cast<Nz::BoxCollider3D*>(box.collider.get());
			if (n < cap && b)
			{
				float* o = out + std::size_t(n) * 9;
				o[0] = chunkOffset.x; o[1] = chunkOffset.y; o[2] = chunkOffset.z;
				o[3] = box.offset.x; o[4] = box.offset.y; o[5] = box.offset.z;
				o[6] = b->GetLengths().x; o[7] = b->GetLengths().y; o[8] = b->GetLengths().z;
			}
			++n;
		};
		for (const auto& child : top->GetGeoms())
		{
			if (auto* inner = dynamic_cast<Nz::CompoundCollider3D*>(child.collider.get()))
				for (const auto& box : inner->GetGeoms())
					emit(child.offset, box);
			else
				emit(Nz::Vector3f::Zero(), child);
		}
		return n;
	}

	// Nz::SubMesh::GenerateTangents (ClientChunkEntities.cpp:169): restated in oracle/nazara_tangents.hpp, unpinned third-party arithmetic
	void orc_generate_tangents(const float* pos, const float* uvw, const std::uint32_t* indices, std::uint32_t vertexCount, std::uint32_t indexCount, float* tangentsOut)
	{
		nazara_tangents::Generate(pos, uvw, indices, vertexCount, indexCount, tangentsOut);
	}
}
