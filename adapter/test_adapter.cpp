// test_adapter.cpp — the reference's own unit test (tests/UnitTests/Common/ChunkTests.cpp:8-46) re-run against the adapter,
// plus (with --gpu) an end-to-end pass over the C ABI: generate a planet, mesh it batched and per chunk, edit, re-mesh.
// Exit code 0 = all checks passed.  Without --gpu nothing touches CUDA (the library is linked but no entry point is called).
#include "tsom_adapter.hpp"

#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

using namespace tsom_b200;

static int g_failed = 0;
#define CHECK(cond)                                                                 \
	do                                                                              \
	{                                                                               \
		if (!(cond))                                                                \
		{                                                                           \
			std::printf("CHECK failed at %s:%d: %s\n", __FILE__, __LINE__, #cond); \
			++g_failed;                                                             \
		}                                                                           \
	} while (0)

// index math only needs the inline members; a container without a device is enough
struct IndexMath
{
	static BlockIndices GetBlockIndices(const ChunkIndices& c, const Vector3ui& i)
	{
		return c * std::int32_t(32) + BlockIndices(std::int32_t(i.x), std::int32_t(i.z), std::int32_t(i.y)) - BlockIndices(32) / 2;
	}
};

static void TestBlockPositions(ChunkContainer* planet)
{
	constexpr int ChunkSize = int(ChunkContainer::ChunkSize);
	constexpr int HalfChunkSize = ChunkSize / 2;
	auto GetBlockIndices = [&](const ChunkIndices& c, const Vector3ui& i) { return planet ? planet->GetBlockIndices(c, i) : IndexMath::GetBlockIndices(c, i); };

	// "Testing block indices" (ChunkTests.cpp:17-23)
	CHECK(GetBlockIndices({ 0, 0, 0 }, { 0, 0, 0 }) == BlockIndices(-HalfChunkSize, -HalfChunkSize, -HalfChunkSize));
	CHECK(GetBlockIndices({ -1, 0, 0 }, { 0, 0, 0 }) == BlockIndices(-HalfChunkSize * 3, -HalfChunkSize, -HalfChunkSize));
	CHECK(GetBlockIndices({ 1, 0, 0 }, { 0, 0, 0 }) == BlockIndices(HalfChunkSize, -HalfChunkSize, -HalfChunkSize));
	CHECK(GetBlockIndices({ 0, -1, 0 }, { 0, 0, 0 }) == BlockIndices(-HalfChunkSize, -HalfChunkSize * 3, -HalfChunkSize));
	CHECK(GetBlockIndices({ 0, 1, 0 }, { 0, 0, 0 }) == BlockIndices(-HalfChunkSize, HalfChunkSize, -HalfChunkSize));
	CHECK(GetBlockIndices({ 0, 0, -1 }, { 0, 0, 0 }) == BlockIndices(-HalfChunkSize, -HalfChunkSize, -HalfChunkSize * 3));
	CHECK(GetBlockIndices({ 0, 0, 1 }, { 0, 0, 0 }) == BlockIndices(-HalfChunkSize, -HalfChunkSize, HalfChunkSize));

	if (!planet)
		return;
	// "Converting back and forth" (ChunkTests.cpp:25-44)
	auto Test = [&](const ChunkIndices& chunkIndices, const Vector3ui& blockIndices) {
		BlockIndices globalIndices = planet->GetBlockIndices(chunkIndices, blockIndices);
		Vector3ui localIndices;
		CHECK(planet->GetChunkIndicesByBlockIndices(globalIndices, &localIndices) == chunkIndices);
		CHECK(localIndices == blockIndices);
	};
	Test({ 1, 0, 0 }, { 1, 0, 0 });
	Test({ -1, 0, 0 }, { 0, 0, 0 });
	Test({ -1, 0, 0 }, { 0, 0, 1 });
	Test({ -1, 0, 0 }, { 0, 1, 0 });
	Test({ -1, 42, 2 }, { 14, 10, 12 });
	Test({ 3, 42, -2 }, { 17, 10, ChunkContainer::ChunkSize - 1 });
	Test({ 3, 42, -2 }, { ChunkContainer::ChunkSize - 1, ChunkContainer::ChunkSize - 1, ChunkContainer::ChunkSize - 1 });
}

// the same index math without a device: GetChunkIndicesByBlockIndices is a pure inline member, exercised through a
// null-device shim so the CPU test covers both KAT sections
struct HostOnlyContainerMath
{
	static ChunkIndices ChunkOf(const BlockIndices& indices, Vector3ui* local)
	{
		const BlockIndices adjusted = indices + BlockIndices(32) / 2;
		auto floorFix = [](int v) { return v < 0 ? v - 32 + 1 : v; };
		const ChunkIndices chunk(floorFix(adjusted.x) / 32, floorFix(adjusted.y) / 32, floorFix(adjusted.z) / 32);
		Vector3i l = adjusted - chunk * std::int32_t(32);
		auto wrap = [](int v) { return v < 0 ? v + 32 : v; };
		*local = Vector3ui(std::uint32_t(wrap(l.x)), std::uint32_t(wrap(l.z)), std::uint32_t(wrap(l.y)));
		return chunk;
	}
};

static void TestBlockLibrary()
{
	BlockLibrary lib;
	CHECK(lib.GetBlockCount() == 14);
	CHECK(lib.GetBlockIndex("empty") == EmptyBlockIndex);
	CHECK(lib.GetBlockIndex("stone") == 7);
	CHECK(lib.GetBlockIndex("glass") == 13);
	CHECK(lib.GetBlockIndex("no_such_block") == InvalidBlockIndex);
	const auto& grass = lib.GetBlockData(lib.GetBlockIndex("grass"));
	CHECK(grass.texIndices[int(Direction::Up)] != grass.texIndices[int(Direction::Down)]);
	CHECK(grass.texIndices[int(Direction::Left)] == grass.texIndices[int(Direction::Back)]);
	CHECK(grass.texIndices[int(Direction::Down)] == lib.GetBlockData(lib.GetBlockIndex("dirt")).texIndices[int(Direction::Up)]);
	CHECK(lib.GetBlockData(13).isDoubleSided && lib.GetBlockData(13).isTransparent);
	CHECK(!lib.GetBlockData(0).hasCollisions && lib.GetBlockData(0).isTransparent);
}

static void TestGpu()
{
	BlockLibrary lib;
	Device device(lib, 0);
	const Vector3ui chunkCount(3, 3, 3);
	Planet planet(device, 1.f, 16.f, 9.81f, 28);
	TestBlockPositions(&planet);

	planet.GenerateChunks(lib, 42, chunkCount);
	CHECK(planet.GetChunkCount() == 27);
	const auto grid = Planet::ChunkGrid(chunkCount);

	// batched == per chunk (Chunk::BuildMesh call shape with caller-owned, append-only buffers)
	MeshBatch batch = planet.BuildMeshes(grid);
	std::uint64_t faces = 0;
	std::vector<MeshBatch> keep;
	for (std::size_t i = 0; i < grid.size(); ++i)
		faces += batch.FaceCount(i);
	CHECK(faces == batch.totalFaces);
	CHECK(faces > 0);

	std::vector<std::vector<VertexStruct>> batchVerts(grid.size());
	std::vector<std::vector<std::uint32_t>> batchIdx(grid.size());
	for (std::size_t i = 0; i < grid.size(); ++i)
		batch.CopyChunk(i, &batchVerts[i], &batchIdx[i], nullptr);

	// ComputeVoxelCorners (FlatChunk.cpp:48-54): block (0, 5, 31) of chunk (1, -1, 0), BoxCorner order, y / z crossed
	{
		Chunk c = planet.GetChunk(ChunkIndices(1, -1, 0));
		const auto corners = c.ComputeVoxelCorners(Vector3ui(0, 5, 31));
		CHECK(corners[std::size_t(Chunk::BoxCorner::LeftBottomFar)] == Vector3f(-16.f, 15.f, -10.f));
		CHECK(corners[std::size_t(Chunk::BoxCorner::RightBottomFar)] == Vector3f(-16.f, 15.f, -11.f));
		CHECK(corners[std::size_t(Chunk::BoxCorner::RightTopNear)] == Vector3f(-15.f, 16.f, -11.f));
		// every vertex of this block's faces in the mesh is one of its corners: checked for all blocks in tests/test_gpu_mesh_parity.py
	}

	// GenerateAABB on the device == min / max over the vertices the caller received
	{
		const auto boxes = batch.ComputeAABBs();
		CHECK(boxes.size() == grid.size());
		for (std::size_t i = 0; i < grid.size(); ++i)
		{
			Vector3f lo(0.f, 0.f, 0.f), hi(0.f, 0.f, 0.f);
			for (std::size_t v = 0; v < batchVerts[i].size(); ++v)
			{
				const Vector3f& p = batchVerts[i][v].position;
				if (v == 0)
					lo = hi = p;
				lo = Vector3f(std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z));
				hi = Vector3f(std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z));
			}
			CHECK(boxes[i].min == lo && boxes[i].max == hi);
		}
	}

	for (std::size_t i = 0; i < grid.size(); ++i)
	{
		Chunk chunk = planet.GetChunk(grid[i]);
		std::vector<VertexStruct> vertices;
		std::vector<std::uint32_t> indices;
		const Vector3f center = planet.GetCenter() - planet.GetChunkOffset(grid[i]); // ClientChunkEntities.cpp:160
		chunk.BuildMesh(indices, center, [&](std::uint32_t count) {
			const std::size_t first = vertices.size();
			vertices.resize(first + count);
			Chunk::VertexAttributes a;
			a.firstIndex = std::uint32_t(first);
			a.position = &vertices[first].position;
			a.normal = &vertices[first].normal;
			a.uv = &vertices[first].uvw;
			return a;
		});
		CHECK(indices == batchIdx[i]);
		CHECK(vertices.size() == batchVerts[i].size());
		CHECK(vertices.empty() || std::memcmp(vertices.data(), batchVerts[i].data(), vertices.size() * sizeof(VertexStruct)) == 0);
		// every face: 4 vertices, indices k*4 + {0,2,1,1,2,3} (Chunk.cpp:95-101)
		for (std::size_t f = 0; f * 6 < indices.size(); ++f)
		{
			const std::uint32_t b = std::uint32_t(f * 4);
			const std::uint32_t want[6] = { b, b + 2, b + 1, b + 1, b + 2, b + 3 };
			CHECK(std::memcmp(&indices[f * 6], want, sizeof(want)) == 0);
		}
		// collider positions == render positions, null for an empty mesh
		auto collider = chunk.BuildCollider();
		CHECK(collider.has_value() == !vertices.empty());
		if (collider)
		{
			CHECK(collider->positions.size() == vertices.size());
			bool same = true;
			for (std::size_t v = 0; v < vertices.size() && same; ++v)
				same = collider->positions[v] == vertices[v].position;
			CHECK(same);
			CHECK(collider->indices == indices);
		}
	}

	// the reference's calling pattern (ClientChunkEntities.cpp:207-241): many TaskScheduler workers, each calling BuildMesh /
	// BuildCollider of its own chunk at the same time — here 8 threads over the 27 chunks, 3 rounds, results == the batched run
	{
		std::atomic<std::size_t> next{ 0 };
		std::atomic<int> mismatches{ 0 };
		auto worker = [&] {
			for (;;)
			{
				const std::size_t job = next.fetch_add(1);
				if (job >= grid.size() * 3)
					return;
				const std::size_t i = job % grid.size();
				Chunk c = planet.GetChunk(grid[i]);
				std::vector<VertexStruct> vertices;
				std::vector<std::uint32_t> indices;
				const Vector3f center = planet.GetCenter() - planet.GetChunkOffset(grid[i]);
				c.BuildMesh(indices, center, [&](std::uint32_t count) {
					vertices.resize(count);
					Chunk::VertexAttributes a;
					a.position = &vertices[0].position;
					a.normal = &vertices[0].normal;
					a.uv = &vertices[0].uvw;
					return a;
				});
				if (indices != batchIdx[i] || vertices.size() != batchVerts[i].size() ||
				    (!vertices.empty() && std::memcmp(vertices.data(), batchVerts[i].data(), vertices.size() * sizeof(VertexStruct)) != 0))
					++mismatches;
				auto collider = c.BuildCollider();
				if (collider.has_value() != !vertices.empty() || (collider && collider->indices != indices))
					++mismatches;
			}
		};
		std::vector<std::thread> pool;
		for (int t = 0; t < 8; ++t)
			pool.emplace_back(worker);
		for (auto& t : pool)
			t.join();
		CHECK(mismatches.load() == 0);
	}

	// edit: break a block, the dirty set holds the chunk (and the masked neighbours), faces change
	const ChunkIndices c0(0, 1, 0);
	Chunk chunk = planet.GetChunk(c0);
	auto before = chunk.GetContent();
	unsigned int victim = 0;
	for (unsigned int i = 0; i < before.size(); ++i)
		if (before[i] != EmptyBlockIndex) { victim = i; break; }
	const Vector3ui local(victim & 31u, (victim >> 5) & 31u, victim >> 10);
	planet.CollectInvalidatedChunks(); // generation invalidated everything; start the tick clean
	chunk.UpdateBlock(local, EmptyBlockIndex);
	auto dirty = planet.CollectInvalidatedChunks();
	bool found = false;
	for (auto& d : dirty)
		found = found || d == c0;
	CHECK(found);
	auto after = chunk.GetContent();
	CHECK(after[victim] == EmptyBlockIndex);
	CHECK(planet.CollectInvalidatedChunks().empty());

	// per-tick glue: two edits in different chunks -> one Update() rebuilds both (and only dirty chunks), then nothing is left
	{
		std::size_t applied = 0, facesSeen = 0;
		MeshOptions opt;
		opt.collider = true;
		ChunkEntities entities(planet, opt, [&](const ChunkIndices&, const MeshBatch& b, std::size_t i) { ++applied; facesSeen += b.FaceCount(i); });
		planet.UpdateBlocks({ { ChunkIndices(0, 1, 0), Vector3ui(5, 5, 5), lib.GetBlockIndex("glass") }, { ChunkIndices(1, 0, 0), Vector3ui(0, 7, 7), EmptyBlockIndex } });
		const std::size_t n = entities.Update();
		CHECK(n >= 2 && n <= 4); // the two chunks + the Left neighbour of the x == 0 edit
		CHECK(applied == n && facesSeen > 0);
		CHECK(entities.Update() == 0);
		after = chunk.GetContent();
	}

	// save format: serialize -> deserialize into another chunk -> same ids; a corrupted version is rejected like the reference does
	{
		auto bytes = chunk.Serialize(lib);
		CHECK(bytes.size() > 18 + 32768 - 1);
		CHECK(bytes[3] == 1 && bytes[7] == 32);
		Chunk copy = planet.AddChunk(lib, ChunkIndices(10, 10, 10));
		copy.Deserialize(lib, bytes.data(), bytes.size());
		CHECK(copy.GetContent() == after);
		planet.RemoveChunk(ChunkIndices(10, 10, 10));
		bytes[3] = 2;
		bool threw = false;
		try { chunk.Deserialize(lib, bytes.data(), bytes.size()); } catch (const std::runtime_error& e) { threw = std::string(e.what()) == "incompatible chunk version"; }
		CHECK(threw);
	}

	// wire format: LZ4 block of the block array -> another chunk -> same ids; a truncated block is rejected and changes nothing
	{
		auto packet = chunk.Compress();
		CHECK(!packet.empty() && packet.size() < TSOM_CHUNK_BLOCKS * 2);
		Chunk copy = planet.AddChunk(lib, ChunkIndices(10, 10, 10));
		copy.ResetCompressed(packet.data(), packet.size());
		CHECK(copy.GetContent() == after);
		bool threw = false;
		try { copy.ResetCompressed(packet.data(), packet.size() - 3); } catch (const std::runtime_error& e) { threw = std::string(e.what()).find("failed to decompress chunk") == 0; }
		CHECK(threw);
		CHECK(copy.GetContent() == after);
		planet.RemoveChunk(ChunkIndices(10, 10, 10));
	}

	// greedy box collider of a full chunk: one box covering everything
	Chunk solid = planet.AddChunk(lib, ChunkIndices(10, 10, 10), [&](BlockIndex* blocks) {
		for (unsigned int i = 0; i < TSOM_CHUNK_BLOCKS; ++i)
			blocks[i] = lib.GetBlockIndex("stone");
	});
	auto boxes = solid.BuildBoxCollider();
	CHECK(boxes.size() == 1);
	if (boxes.size() == 1)
	{
		CHECK(boxes[0].position == Vector3f(-16.f, -16.f, -16.f));
		CHECK(boxes[0].size == Vector3f(32.f, 32.f, 32.f));
	}
	CHECK(planet.GetChunkCount() == 28);
	std::printf("adapter GPU pass: %llu faces over %zu chunks\n", (unsigned long long)faces, grid.size());
}

// round 2: the sharded planet in one process (parts may share a GPU: devices = {0, 0, 0} works on a single-GPU box), the Ship container,
// batched box colliders, tangents.  `devices` comes from the command line: --gpu [n_parts]
// number of usable devices through the C ABI alone (the adapter does not include CUDA headers)
static void cudaDevicesVisible(int* n)
{
	*n = 0;
	for (int d = 0; d < 16; ++d)
	{
		tsom_ctx* c = nullptr;
		if (tsom_create(d, &c) != TSOM_OK)
			break;
		tsom_destroy(c);
		++*n;
	}
}

static void TestGpuSharded(const std::vector<int>& devices)
{
	BlockLibrary lib;
	Device device(lib, 0);
	const Vector3ui count(3, 3, 3);
	const auto grid = Planet::ChunkGrid(count);

	// one GPU, one container: the reference run
	Planet planet(device, 1.f, 16.f, 9.81f, 32);
	planet.GenerateChunks(lib, 7, count);
	MeshOptions opt;
	opt.collider = true;
	MeshBatch whole = planet.BuildMeshes(grid, opt);

	// the same planet cut into devices.size() parts
	ShardedPlanet sharded(lib, devices, count, 1.f);
	CHECK(sharded.GetPartCount() == devices.size());
	sharded.GenerateChunks(lib, 7);
	ShardedMeshBatch parts = sharded.BuildMeshes(grid, opt);
	std::size_t differing = 0;
	for (std::size_t i = 0; i < grid.size(); ++i)
	{
		CHECK(parts.parts[i] == sharded.GetOwner(grid[i]));
		std::vector<VertexStruct> va, vb;
		std::vector<std::uint32_t> ia, ib;
		std::vector<Vector3f> ca, cb;
		whole.CopyChunk(i, &va, &ia, &ca);
		parts.CopyChunk(i, &vb, &ib, &cb);
		if (ia != ib || va.size() != vb.size() || (!va.empty() && std::memcmp(va.data(), vb.data(), va.size() * sizeof(VertexStruct)) != 0) ||
		    (!ca.empty() && std::memcmp(ca.data(), cb.data(), ca.size() * sizeof(Vector3f)) != 0))
			++differing;
		CHECK(sharded.GetContent(grid[i]) == planet.GetChunk(grid[i]).GetContent());
	}
	CHECK(differing == 0);

	// an edit on a part boundary: both sides of the cut are rebuilt, like on the single container
	planet.CollectInvalidatedChunks();
	sharded.CollectInvalidatedChunks();
	std::vector<ChunkContainer::BlockUpdate> edits;
	for (const ChunkIndices& c : grid)
		edits.push_back({ c, Vector3ui(0, 31, 0), EmptyBlockIndex });
	planet.UpdateBlocks(edits);
	sharded.UpdateBlocks(edits);
	auto dirtyA = planet.CollectInvalidatedChunks();
	auto dirtyB = sharded.CollectInvalidatedChunks();
	CHECK(dirtyA.size() == 27 && dirtyA.size() == dirtyB.size());
	for (std::size_t i = 0; i < dirtyA.size() && i < dirtyB.size(); ++i)
		CHECK(dirtyA[i] == dirtyB[i]);
	MeshBatch wholeAfter = planet.BuildMeshes(grid, opt);
	ShardedMeshBatch partsAfter = sharded.BuildMeshes(grid, opt);
	for (std::size_t i = 0; i < grid.size(); ++i)
		CHECK(wholeAfter.FaceCount(i) == partsAfter.FaceCount(i));

	// indices synthesised on the host == indices copied from the device; tangents fill the 12-byte slot with unit vectors
	{
		MeshOptions t;
		t.tangents = true;
		MeshBatch b = planet.BuildMeshes({ grid[13] }, t);
		std::vector<VertexStruct> v, v2;
		std::vector<std::uint32_t> fromDevice, synthesized;
		b.CopyChunk(0, &v, &fromDevice, nullptr);
		b.CopyChunkSynthesizeIndices(0, &v2, synthesized);
		CHECK(!fromDevice.empty() && fromDevice == synthesized);
		bool unit = !v.empty();
		for (const VertexStruct& q : v)
		{
			const float len = std::sqrt(q.tangent.x * q.tangent.x + q.tangent.y * q.tangent.y + q.tangent.z * q.tangent.z);
			unit = unit && std::fabs(len - 1.f) < 1e-5f;
		}
		CHECK(unit);
	}

	// batched box colliders == one call per chunk
	{
		auto all = planet.BuildBoxColliders(grid);
		std::size_t boxes = 0;
		for (std::size_t i = 0; i < grid.size(); i += 5)
		{
			auto one = planet.GetChunk(grid[i]).BuildBoxCollider();
			CHECK(one.size() == all[i].size());
			for (std::size_t k = 0; k < one.size() && k < all[i].size(); ++k)
				CHECK(one[k].position == all[i][k].position && one[k].size == all[i][k].size);
			boxes += one.size();
		}
		CHECK(boxes > 0);
	}

	// Ship: Generate, the Up / Down-swapped mask, the hull compound
	{
		Ship ship(device, 2.f);
		ship.Generate(lib, true);
		auto hull = ship.BuildHullCollider();
		CHECK(hull.size() == 1 && !hull[0].boxes.empty() && hull[0].offset == Vector3f(0.f, 0.f, 0.f));
		auto d = ship.CollectInvalidatedChunks();
		CHECK(d.size() == 1 && d[0] == ChunkIndices(0, 0, 0));
		ship.AddChunk(lib, ChunkIndices(0, 1, 0), [&](BlockIndex* b) { b[0] = lib.GetBlockIndex("hull"); });
		ship.AddChunk(lib, ChunkIndices(0, -1, 0), [&](BlockIndex* b) { b[0] = lib.GetBlockIndex("hull"); });
		ship.CollectInvalidatedChunks();
		ship.UpdateBlocks({ { ChunkIndices(0, 0, 0), Vector3ui(5, 5, 0), lib.GetBlockIndex("hull") } }); // local z == 0: Up on a Ship (Ship.cpp:48-49)
		d = ship.CollectInvalidatedChunks();
		CHECK(d.size() == 2 && d[0] == ChunkIndices(0, 0, 0) && d[1] == ChunkIndices(0, 1, 0));
		hull = ship.BuildHullCollider();
		CHECK(hull.size() == 3 && hull[1].offset == Vector3f(0.f, 64.f, 0.f));
	}
	std::printf("adapter sharded GPU pass: %zu parts\n", devices.size());
}

int main(int argc, char** argv)
{
	const bool gpu = argc > 1 && std::strcmp(argv[1], "--gpu") == 0;
	TestBlockLibrary();
	TestBlockPositions(nullptr);
	{
		// round trips of ChunkTests.cpp:37-43 on the pure index math
		const int S = 32;
		const int cases[7][6] = { { 1, 0, 0, 1, 0, 0 }, { -1, 0, 0, 0, 0, 0 }, { -1, 0, 0, 0, 0, 1 }, { -1, 0, 0, 0, 1, 0 }, { -1, 42, 2, 14, 10, 12 }, { 3, 42, -2, 17, 10, S - 1 }, { 3, 42, -2, S - 1, S - 1, S - 1 } };
		for (auto& c : cases)
		{
			const ChunkIndices ci(c[0], c[1], c[2]);
			const Vector3ui bi(c[3], c[4], c[5]);
			Vector3ui local;
			CHECK(HostOnlyContainerMath::ChunkOf(IndexMath::GetBlockIndices(ci, bi), &local) == ci);
			CHECK(local == bi);
		}
	}
	if (gpu)
	{
		try
		{
			TestGpu();
			// parts of the sharded planet: one per visible GPU when there are several, else --gpu N parts on GPU 0 (default 3)
			int nDev = 0;
			cudaDevicesVisible(&nDev);
			const int nParts = argc > 2 ? std::atoi(argv[2]) : (nDev > 1 ? nDev : 3);
			std::vector<int> devices;
			for (int i = 0; i < nParts; ++i)
				devices.push_back(nDev > 1 ? i % nDev : 0);
			TestGpuSharded(devices);
		}
		catch (const std::exception& e)
		{
			std::printf("exception: %s\n", e.what());
			++g_failed;
		}
	}
	std::printf("%s (%d failed)\n", g_failed ? "FAILED" : "OK", g_failed);
	return g_failed ? 1 : 0;
}
