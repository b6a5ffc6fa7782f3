// oracle/capi.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  C entry points (ctypes) over tsom_oracle.hpp.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load this library.
#include "tsom_oracle.hpp"
#include "lz4_block.hpp"

#include "nazara_tangents.hpp"
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

using namespace orc;

namespace
{
	BlockLibrary& MutableLib()
	{
		static BlockLibrary lib;
		return lib;
	}
	const BlockLibrary& Lib() { return MutableLib(); }

	struct PlanetHandle
	{
		Planet planet;
		PlanetHandle(float tile, float radius) : planet(tile, radius, 9.81f) {}
	};

	// One task per item on nthreads workers (the shape of Nz::TaskScheduler fan-out, Planet.cpp:394-402)
	template<typename F> void ParallelFor(std::size_t n, int nthreads, F&& f)
	{
		if (nthreads <= 1)
		{
			for (std::size_t i = 0; i < n; ++i)
				f(i);
			return;
		}
		std::atomic<std::size_t> next{ 0 };
		std::vector<std::thread> workers;
		for (int t = 0; t < nthreads; ++t)
			workers.emplace_back([&] {
				for (;;)
				{
					std::size_t i = next.fetch_add(1);
					if (i >= n)
						return;
					f(i);
				}
			});
		for (auto& w : workers)
			w.join();
	}

	double Now()
	{
		return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
	}
}

extern "C"
{
	// ---- index math (KAT surface: tests/UnitTests/Common/ChunkTests.cpp:17-43)
	void orc_get_block_indices(const std::int32_t chunk[3], const std::uint32_t local[3], std::int32_t out[3])
	{
		Vec3i r = ChunkContainer::GetBlockIndices({ chunk[0], chunk[1], chunk[2] }, { local[0], local[1], local[2] });
		out[0] = r.x; out[1] = r.y; out[2] = r.z;
	}

	void orc_get_chunk_indices_by_block_indices(const std::int32_t block[3], std::int32_t chunkOut[3], std::uint32_t localOut[3])
	{
		Vec3ui l;
		Vec3i c = ChunkContainer::GetChunkIndicesByBlockIndices({ block[0], block[1], block[2] }, &l);
		chunkOut[0] = c.x; chunkOut[1] = c.y; chunkOut[2] = c.z;
		localOut[0] = l.x; localOut[1] = l.y; localOut[2] = l.z;
	}

	int orc_direction_from_normal(const float n[3]) { return DirectionFromNormal({ n[0], n[1], n[2] }); }

	// ---- block library
	std::uint32_t orc_blocklib_count() { return std::uint32_t(Lib().GetBlockCount()); }

	void orc_blocklib_get(std::uint32_t idx, std::uint32_t tex[6], std::uint8_t flags[3], char name[64])
	{
		const BlockData& d = Lib().GetBlockData(BlockIndex(idx));
		for (int i = 0; i < 6; ++i)
			tex[i] = d.texIndices[i];
		flags[0] = d.hasCollisions; flags[1] = d.isDoubleSided; flags[2] = d.isTransparent;
		std::strncpy(name, d.name.c_str(), 63);
		name[63] = 0;
	}

	int orc_blocklib_index(const char* name) { return Lib().GetBlockIndex(name); }

	// BlockLibrary::RegisterBlock (BlockLibrary.cpp:85-131) on the process-wide library: appends one block type (tests of large
	// block tables run in their own process).  Empty strings mean "not given".
	int orc_blocklib_register(const char* name, const char* basePath, const char* sidePath, const char* upPath, int hasCollisions, int isDoubleSided, int isTransparent)
	{
		BlockInfo info;
		info.basePath = basePath;
		info.baseSidePath = sidePath;
		info.baseUpPath = upPath;
		info.hasCollisions = hasCollisions != 0;
		info.isDoubleSided = isDoubleSided != 0;
		info.isTransparent = isTransparent != 0;
		return MutableLib().RegisterBlock(name, info);
	}

	// ---- perlin / rng probes
	void orc_perlin_perm(std::uint32_t seed, std::uint8_t out[256])
	{
		PerlinNoise p;
		p.reseed(seed);
		std::memcpy(out, p.permutation(), 256);
	}

	// noise3D over a caller-supplied permutation (siv::PerlinNoise::deserialize): lets the tests check the noise core against Ken Perlin's
	// published improved-noise reference value
	double orc_perlin_noise3d_perm(const std::uint8_t perm[256], double x, double y, double z)
	{
		PerlinNoise p;
		p.deserialize(perm);
		return p.noise3D(x, y, z);
	}

	double orc_perlin_octave01(std::uint32_t seed, double x, double y, int octaves)
	{
		PerlinNoise p;
		p.reseed(seed);
		return p.normalizedOctave2D_01(x, y, octaves);
	}

	// the first n draws of std::bernoulli_distribution(0.9) over std::minstd_rand(seed)
	void orc_bernoulli_draws(std::uint32_t seed, std::uint32_t n, std::uint8_t* out)
	{
		std::minstd_rand rand(seed);
		std::bernoulli_distribution dis(0.9);
		for (std::uint32_t i = 0; i < n; ++i)
			out[i] = dis(rand);
	}

	// ---- generation of one chunk without a container
	int orc_generate_chunk(std::uint32_t seed, const std::uint32_t count[3], const std::int32_t chunk[3], std::uint16_t* out)
	{
		return Planet::GenerateChunkBlocks(Lib(), { chunk[0], chunk[1], chunk[2] }, seed, { count[0], count[1], count[2] }, out) ? 0 : 1;
	}

	// ---- planet handle
	void* orc_planet_create(float tile, float cornerRadius) { return new PlanetHandle(tile, cornerRadius); }
	void orc_planet_destroy(void* h) { delete static_cast<PlanetHandle*>(h); }

	// blocks == nullptr: chunk exists but has no content (Chunk::HasContent() == false)
	int orc_planet_add_chunk(void* h, const std::int32_t idx[3], const std::uint16_t* blocks)
	{
		Planet& p = static_cast<PlanetHandle*>(h)->planet;
		Vec3i ci{ idx[0], idx[1], idx[2] };
		if (p.GetChunk(ci))
			return 1;
		Chunk& c = p.AddChunk(Lib(), ci);
		if (blocks)
			c.Reset([&](BlockIndex* dst) { std::memcpy(dst, blocks, ChunkBlockCount * sizeof(BlockIndex)); });
		return 0;
	}

	int orc_planet_reset_chunk(void* h, const std::int32_t idx[3], const std::uint16_t* blocks)
	{
		Planet& p = static_cast<PlanetHandle*>(h)->planet;
		Chunk* c = p.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c)
			return 1;
		c->Reset([&](BlockIndex* dst) { std::memcpy(dst, blocks, ChunkBlockCount * sizeof(BlockIndex)); });
		return 0;
	}

	// Planet::GenerateChunks with one task per chunk on nthreads workers; returns 0 ok, 1 if SafeCaster would have asserted
	int orc_planet_generate(void* h, std::uint32_t seed, const std::uint32_t count[3], int nthreads)
	{
		Planet& p = static_cast<PlanetHandle*>(h)->planet;
		Vec3ui cc{ count[0], count[1], count[2] };
		std::vector<Chunk*> chunks;
		for (int cz = 0; cz < int(cc.z); ++cz)
			for (int cy = 0; cy < int(cc.y); ++cy)
				for (int cx = 0; cx < int(cc.x); ++cx)
					chunks.push_back(&p.AddChunk(Lib(), { cx - int(cc.x / 2), cy - int(cc.y / 2), cz - int(cc.z / 2) }));

		// OnChunkUpdated is not thread-safe here (the reference guards it with a mutex): detach, generate, then signal serially
		std::atomic<int> bad{ 0 };
		std::vector<std::vector<BlockIndex>> tmp(chunks.size());
		ParallelFor(chunks.size(), nthreads, [&](std::size_t i) {
			tmp[i].resize(ChunkBlockCount);
			if (!Planet::GenerateChunkBlocks(Lib(), chunks[i]->GetIndices(), seed, cc, tmp[i].data()))
				bad = 1;
		});
		for (std::size_t i = 0; i < chunks.size(); ++i)
			chunks[i]->Reset([&](BlockIndex* dst) { std::memcpy(dst, tmp[i].data(), ChunkBlockCount * sizeof(BlockIndex)); });
		return bad;
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
		std::memcpy(out, c->Blocks().data(), ChunkBlockCount * sizeof(BlockIndex));
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

	// edits: n x {cx,cy,cz,x,y,z,id} int32; applied in order (Chunk::UpdateBlock); edits to missing chunks are skipped
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

	void orc_planet_clear_invalidated(void* h) { static_cast<PlanetHandle*>(h)->planet.invalidated.clear(); }

	// dirty set of the tick (sorted by (x,y,z) chunk index); returns the count, writes up to cap triples
	std::uint32_t orc_planet_collect_dirty(void* h, std::int32_t* out, std::uint32_t cap)
	{
		std::vector<Vec3i> d = static_cast<PlanetHandle*>(h)->planet.CollectDirtyChunks(true);
		for (std::uint32_t i = 0; i < d.size() && i < cap; ++i)
		{
			out[i * 3 + 0] = d[i].x; out[i * 3 + 1] = d[i].y; out[i * 3 + 2] = d[i].z;
		}
		return std::uint32_t(d.size());
	}

	// ---- meshing
	struct MeshHandle { MeshOut out; };

	// cornerMode 0 flat / 1 deformed.  gravityCenter == nullptr: ClientChunkEntities.cpp:160 (container centre - chunk offset).
	// deformCenter == nullptr: same point.
	void* orc_mesh_build(void* h, const std::int32_t idx[3], int cornerMode, const float* gravityCenter, const float* deformCenter, float deformRadius, int wantNormal, int wantUv)
	{
		Planet& p = static_cast<PlanetHandle*>(h)->planet;
		Chunk* c = p.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c || !c->HasContent())
			return nullptr;

		Vec3f center = p.GetCenter() - p.GetChunkOffset(c->GetIndices());
		if (gravityCenter)
			center = Vec3f{ gravityCenter[0], gravityCenter[1], gravityCenter[2] };

		c->cornerMode = cornerMode ? CornerMode::Deformed : CornerMode::Flat;
		c->deformationCenter = deformCenter ? Vec3f{ deformCenter[0], deformCenter[1], deformCenter[2] } : center;
		c->deformationRadius = deformRadius;

		auto* m = new MeshHandle;
		c->BuildMesh(m->out, center, wantNormal != 0, wantUv != 0);
		return m;
	}

	// Chunk::ComputeVoxelCorners (Chunk.hpp:54; FlatChunk.cpp:48-54 / DeformedChunk.cpp:115-137) of one block: out = 8 corners x xyz in
	// Nz::BoxCorner enum order.  deformCenter == nullptr: container centre - chunk offset.  Returns 0, or 1 for an unknown chunk.
	int orc_voxel_corners(void* h, const std::int32_t idx[3], const std::uint32_t local[3], int cornerMode, const float* deformCenter, float deformRadius, float out[24])
	{
		Planet& p = static_cast<PlanetHandle*>(h)->planet;
		Chunk* c = p.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c)
			return 1;
		const Vec3f center = p.GetCenter() - p.GetChunkOffset(c->GetIndices());
		c->cornerMode = cornerMode ? CornerMode::Deformed : CornerMode::Flat;
		c->deformationCenter = deformCenter ? Vec3f{ deformCenter[0], deformCenter[1], deformCenter[2] } : center;
		c->deformationRadius = deformRadius;
		const Corners corners = c->ComputeVoxelCorners({ local[0], local[1], local[2] });
		int k = 0;
		for (const Vec3f& v : corners)
		{
			out[k++] = v.x;
			out[k++] = v.y;
			out[k++] = v.z;
		}
		return 0;
	}

	std::uint32_t orc_mesh_face_count(void* m) { return std::uint32_t(static_cast<MeshHandle*>(m)->out.indices.size() / 6); }

	// any pointer may be null
	void orc_mesh_copy(void* mh, float* pos, float* normal, float* uvw, std::uint32_t* indices)
	{
		const MeshOut& o = static_cast<MeshHandle*>(mh)->out;
		if (pos) std::memcpy(pos, o.position.data(), o.position.size() * sizeof(Vec3f));
		if (normal) std::memcpy(normal, o.normal.data(), o.normal.size() * sizeof(Vec3f));
		if (uvw) std::memcpy(uvw, o.uvw.data(), o.uvw.size() * sizeof(Vec3f));
		if (indices) std::memcpy(indices, o.indices.data(), o.indices.size() * sizeof(std::uint32_t));
	}

	void orc_mesh_free(void* m) { delete static_cast<MeshHandle*>(m); }

	// FlatChunk greedy box collider: out = n x {pos.xyz, size.xyz}
	std::uint32_t orc_box_collider(void* h, const std::int32_t idx[3], float* out, std::uint32_t cap)
	{
		const Chunk* c = static_cast<PlanetHandle*>(h)->planet.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c || !c->HasContent())
			return 0;
		auto boxes = c->BuildBoxCollider();
		for (std::uint32_t i = 0; i < boxes.size() && i < cap; ++i)
		{
			float* o = out + std::size_t(i) * 6;
			o[0] = boxes[i].pos.x; o[1] = boxes[i].pos.y; o[2] = boxes[i].pos.z;
			o[3] = boxes[i].size.x; o[4] = boxes[i].size.y; o[5] = boxes[i].size.z;
		}
		return std::uint32_t(boxes.size());
	}

	// Chunk::Serialize (Chunk.cpp:312-346) -> bytes written (0: unknown chunk / no content / buffer too small)
	std::uint32_t orc_chunk_serialize(void* h, const std::int32_t idx[3], std::uint8_t* out, std::uint32_t cap)
	{
		const Chunk* c = static_cast<PlanetHandle*>(h)->planet.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c || !c->HasContent())
			return 0;
		std::vector<std::uint8_t> bytes;
		c->Serialize(bytes);
		if (bytes.size() > cap)
			return 0;
		std::memcpy(out, bytes.data(), bytes.size());
		return std::uint32_t(bytes.size());
	}

	// Chunk::Deserialize (Chunk.cpp:252-310) into an existing chunk: 0 ok, 1 unknown chunk, 2 rejected (message in err)
	int orc_chunk_deserialize(void* h, const std::int32_t idx[3], const std::uint8_t* data, std::uint32_t size, char err[128])
	{
		Chunk* c = static_cast<PlanetHandle*>(h)->planet.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c)
			return 1;
		try
		{
			c->Deserialize(data, size);
		}
		catch (const std::exception& e)
		{
			std::strncpy(err, e.what(), 127);
			err[127] = 0;
			return 2;
		}
		return 0;
	}

	// ---- CPU baseline timers (bench.py cpu_baseline / --impl reference).  Return seconds; *facesOut = total faces.

	// mesh (render attributes) n isolated chunks, one task per chunk — BASELINE config 2
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
			c->BuildMesh(out, p.GetCenter() - p.GetChunkOffset(c->GetIndices()), true, true);
			faces += out.indices.size() / 6;
		});
		double t1 = Now();
		if (facesOut)
			*facesOut = faces;
		return t1 - t0;
	}

	// generate a planet then mesh every chunk (render attributes, real neighbours) — BASELINE configs 1/3.
	// times[0] = gen seconds, times[1] = mesh seconds.  stride > 1 meshes only every stride-th chunk (bounded sample).
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
			chunks[i]->BuildMesh(out, h.planet.GetCenter() - h.planet.GetChunkOffset(chunks[i]->GetIndices()), true, true);
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
	// ---- LZ4 block format as used by Packets::ChunkReset (BinaryCompressor.cpp:17-50): see oracle/lz4_block.hpp
	int orc_lz4_bound(int n) { return tsom_oracle_lz4::compress_bound(n); }
	int orc_lz4_compress(const std::uint8_t* src, int n, std::uint8_t* dst, int cap) { return tsom_oracle_lz4::compress_block(src, n, dst, cap); }
	int orc_lz4_decompress(const std::uint8_t* src, int n, std::uint8_t* dst, int cap) { return tsom_oracle_lz4::decompress_block(src, n, dst, cap); }

	// ---- a bounded, planet-wide sample of "generate + mesh" for the CPU arm of bench.py (same entry points as oracle/ref_shim) ----
	struct SampleHandle
	{
		PlanetHandle h{ 1.f, 16.f };
		std::vector<Chunk*> sample;
		std::uint32_t seed = 0;
		Vec3ui count{};
	};

	void* orc_sample_create(std::uint32_t seed, const std::uint32_t count[3], std::uint32_t stride, std::uint32_t offset, int nthreads)
	{
		auto s = new SampleHandle;
		s->seed = seed;
		s->count = Vec3ui{ count[0], count[1], count[2] };
		Planet& p = s->h.planet;
		auto inside = [&](const Vec3i& c) {
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
					const Vec3i ci{ cx - int(count[0] / 2), cy - int(count[1] / 2), cz - int(count[2] / 2) };
					Chunk* c = p.GetChunk(ci);
					if (!c)
					{
						c = &p.AddChunk(Lib(), ci);
						all.push_back(c);
					}
					s->sample.push_back(c);
					for (int d = 0; d < 6; ++d)
					{
						const Vec3i ni{ ci.x + s_chunkDirOffset[d].x, ci.y + s_chunkDirOffset[d].y, ci.z + s_chunkDirOffset[d].z };
						if (inside(ni) && !p.GetChunk(ni))
							all.push_back(&p.AddChunk(Lib(), ni));
					}
				}
		std::atomic<int> bad{ 0 };
		std::vector<std::vector<BlockIndex>> tmp(all.size());
		ParallelFor(all.size(), nthreads, [&](std::size_t i) {
			tmp[i].resize(ChunkBlockCount);
			if (!Planet::GenerateChunkBlocks(Lib(), all[i]->GetIndices(), seed, s->count, tmp[i].data()))
				bad = 1;
		});
		for (std::size_t i = 0; i < all.size(); ++i)
			all[i]->Reset([&](BlockIndex* dst) { std::memcpy(dst, tmp[i].data(), ChunkBlockCount * sizeof(BlockIndex)); });
		if (bad)
		{
			delete s;
			return nullptr;
		}
		return s;
	}

	std::uint32_t orc_sample_size(void* hv) { return std::uint32_t(static_cast<SampleHandle*>(hv)->sample.size()); }

	std::uint64_t orc_sample_step(void* hv, int nthreads, double times[2])
	{
		auto* s = static_cast<SampleHandle*>(hv);
		Planet& p = s->h.planet;
		const double t0 = Now();
		std::vector<std::vector<BlockIndex>> tmp(s->sample.size());
		ParallelFor(s->sample.size(), nthreads, [&](std::size_t i) {
			tmp[i].resize(ChunkBlockCount);
			Planet::GenerateChunkBlocks(Lib(), s->sample[i]->GetIndices(), s->seed, s->count, tmp[i].data());
		});
		for (std::size_t i = 0; i < s->sample.size(); ++i)
			s->sample[i]->Reset([&](BlockIndex* dst) { std::memcpy(dst, tmp[i].data(), ChunkBlockCount * sizeof(BlockIndex)); });
		const double t1 = Now();
		std::atomic<std::uint64_t> faces{ 0 };
		ParallelFor(s->sample.size(), nthreads, [&](std::size_t i) {
			MeshOut out;
			s->sample[i]->BuildMesh(out, p.GetCenter() - p.GetChunkOffset(s->sample[i]->GetIndices()), true, true);
			faces += out.indices.size() / 6;
		});
		const double t2 = Now();
		times[0] = t1 - t0;
		times[1] = t2 - t1;
		return faces;
	}

	void orc_sample_destroy(void* hv) { delete static_cast<SampleHandle*>(hv); }

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
		ParallelFor(chunks.size(), nthreads, [&](std::size_t i) { boxes += chunks[i]->BuildBoxCollider().size(); });
		const double t1 = Now();
		if (boxesOut)
			*boxesOut = boxes;
		if (chunksOut)
			*chunksOut = std::uint32_t(chunks.size());
		return t1 - t0;
	}

	// ---- Ship (same entry points as oracle/ref_shim) ----
	struct ShipHandle
	{
		Ship ship;
		explicit ShipHandle(float tile) : ship(tile) {}
	};

	void* orc_ship_create(float tile) { return new ShipHandle(tile); }
	void orc_ship_destroy(void* h) { delete static_cast<ShipHandle*>(h); }
	void orc_ship_generate(void* h, int small) { static_cast<ShipHandle*>(h)->ship.Generate(Lib(), small != 0); }

	int orc_ship_add_chunk(void* h, const std::int32_t idx[3], const std::uint16_t* blocks)
	{
		Ship& s = static_cast<ShipHandle*>(h)->ship;
		Vec3i ci{ idx[0], idx[1], idx[2] };
		if (s.GetChunk(ci))
			return 1;
		Chunk& c = s.AddChunk(Lib(), ci);
		if (blocks)
			c.Reset([&](BlockIndex* dst) { std::memcpy(dst, blocks, ChunkBlockCount * sizeof(BlockIndex)); });
		return 0;
	}

	int orc_ship_get_blocks(void* h, const std::int32_t idx[3], std::uint16_t* out)
	{
		const Chunk* c = static_cast<ShipHandle*>(h)->ship.GetChunk({ idx[0], idx[1], idx[2] });
		if (!c || !c->HasContent())
			return 1;
		std::memcpy(out, c->Blocks().data(), ChunkBlockCount * sizeof(BlockIndex));
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

	std::uint32_t orc_ship_take_invalidated(void* h, std::int32_t* idxOut, std::uint32_t* maskOut, std::uint32_t cap)
	{
		Ship& s = static_cast<ShipHandle*>(h)->ship;
		std::uint32_t n = 0;
		for (auto&& [k, m] : s.invalidated)
		{
			if (n < cap)
			{
				idxOut[n * 3 + 0] = k.x;
				idxOut[n * 3 + 1] = k.y;
				idxOut[n * 3 + 2] = k.z;
				maskOut[n] = std::uint32_t(m) | 0x80u;
			}
			++n;
		}
		s.invalidated.clear();
		return n;
	}

	std::uint32_t orc_ship_hull_collider(void* h, float* out, std::uint32_t cap)
	{
		auto boxes = static_cast<ShipHandle*>(h)->ship.BuildHullCollider();
		for (std::uint32_t i = 0; i < boxes.size() && i < cap; ++i)
		{
			float* o = out + std::size_t(i) * 9;
			o[0] = boxes[i].chunkOffset.x; o[1] = boxes[i].chunkOffset.y; o[2] = boxes[i].chunkOffset.z;
			o[3] = boxes[i].center.x; o[4] = boxes[i].center.y; o[5] = boxes[i].center.z;
			o[6] = boxes[i].lengths.x; o[7] = boxes[i].lengths.y; o[8] = boxes[i].lengths.z;
		}
		return std::uint32_t(boxes.size());
	}

	// Nz::SubMesh::GenerateTangents (ClientChunkEntities.cpp:169): restated in oracle/nazara_tangents.hpp, unpinned third-party arithmetic
	void orc_generate_tangents(const float* pos, const float* uvw, const std::uint32_t* indices, std::uint32_t vertexCount, std::uint32_t indexCount, float* tangentsOut)
	{
		nazara_tangents::Generate(pos, uvw, indices, vertexCount, indexCount, tangentsOut);
	}
}
