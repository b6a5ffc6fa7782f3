// tsom_sharded.inl — included at the end of tsom_host.cu (inside extern "C"): the shard plan and the single-process multi-GPU planet.
//
// A TSOM server is ONE process (reference src/Server/main.cpp:24-62), so "this planet over these GPUs" has to be a call of the
// library, not a launcher: tsom_planet_create_sharded owns one context + container per part and fans every call out to them.
// The parts talk to each other exactly like the one-process-per-GPU deployment does (peer-mapped HBM, epoch flags polled on the
// device, tsom_halo_pull), so both deployments run the same kernels and the same ordering protocol.

namespace
{
	// weight of a chunk when the caller gives none: every chunk costs its 64 KiB stage; chunks that reach into the band where the
	// terrain surface can lie (|w| > maxH / 2 - 32 on some axis, Planet.cpp:241-247) also emit faces
	float ShellWeight(const uint32_t cc[3], int cx, int cy, int cz)
	{
		const int c[3] = { cx, cy, cz };
		bool shell = false;
		for (int k = 0; k < 3; ++k)
		{
			const int maxH = ((int(cc[k]) + 1) / 2) * CS;
			const int far = std::max(std::abs(c[k] * CS - CS / 2), std::abs(c[k] * CS + CS / 2 - 1));
			shell = shell || far > maxH / 2 - CS;
		}
		return shell ? 4.f : 1.f;
	}

	struct PlanCtx
	{
		const uint32_t* cc;
		const float* weights; // GenerateChunks order, or null
		int32_t* boxes;
		float Weight(int cx, int cy, int cz) const
		{
			if (!weights)
				return ShellWeight(cc, cx, cy, cz);
			const size_t x = size_t(cx + int(cc[0] / 2)), y = size_t(cy + int(cc[1] / 2)), z = size_t(cz + int(cc[2] / 2));
			return weights[(z * cc[1] + y) * cc[0] + x];
		}
	};

	int PlanSplit(const PlanCtx& pc, const int lo[3], const int hi[3], uint32_t r0, uint32_t r1)
	{
		const uint32_t nr = r1 - r0;
		if (nr == 1)
		{
			int32_t* b = pc.boxes + size_t(r0) * 6;
			for (int k = 0; k < 3; ++k)
			{
				b[k] = lo[k];
				b[3 + k] = hi[k];
			}
			return TSOM_OK;
		}
		const uint32_t nL = nr / 2, nR = nr - nL;
		const int ext[3] = { hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, hi[2] - lo[2] + 1 };
		// longest axis; ties go to z, then y (the slower axes of the linear chunk id: the parts stay contiguous in it as far as possible)
		int axis = 2;
		if (ext[1] > ext[axis])
			axis = 1;
		if (ext[0] > ext[axis])
			axis = 0;
		if (ext[axis] < 2)
			return Fail(TSOM_ERR_INVALID, "tsom_shard_plan: more parts than chunks");
		std::vector<double> layer(size_t(ext[axis]), 0.0);
		for (int z = lo[2]; z <= hi[2]; ++z)
			for (int y = lo[1]; y <= hi[1]; ++y)
				for (int x = lo[0]; x <= hi[0]; ++x)
				{
					const int c[3] = { x, y, z };
					layer[size_t(c[axis] - lo[axis])] += pc.Weight(x, y, z);
				}
		double total = 0.0;
		for (double v : layer)
			total += v;
		const double target = total * double(nL) / double(nr);
		const uint64_t layerChunks = uint64_t(ext[(axis + 1) % 3]) * uint64_t(ext[(axis + 2) % 3]);
		int best = -1;
		double bestErr = 0.0, prefix = 0.0;
		for (int k = 1; k < ext[axis]; ++k) // cut between layers k - 1 and k
		{
			prefix += layer[size_t(k - 1)];
			if (layerChunks * uint64_t(k) < nL || layerChunks * uint64_t(ext[axis] - k) < nR)
				continue; // a side must hold at least one chunk per part
			const double err = std::fabs(prefix - target);
			if (best < 0 || err < bestErr)
			{
				best = k;
				bestErr = err;
			}
		}
		if (best < 0)
			return Fail(TSOM_ERR_INVALID, "tsom_shard_plan: more parts than chunks");
		int loR[3] = { lo[0], lo[1], lo[2] }, hiL[3] = { hi[0], hi[1], hi[2] };
		hiL[axis] = lo[axis] + best - 1;
		loR[axis] = lo[axis] + best;
		int rc = PlanSplit(pc, lo, hiL, r0, r0 + nL);
		if (rc != TSOM_OK)
			return rc;
		return PlanSplit(pc, loR, hi, r0 + nL, r1);
	}

	// runs fn(part) for every part on its own thread and reports the first failure with its message (tsom_last_error is per thread)
	template<class F> int ForEachPart(uint32_t n, F&& fn)
	{
		std::vector<int> rcs(n, TSOM_OK);
		std::vector<std::string> msgs(n);
		auto run = [&](uint32_t i) {
			rcs[i] = fn(i);
			if (rcs[i] != TSOM_OK)
				msgs[i] = g_lastError;
		};
		std::vector<std::thread> workers;
		for (uint32_t i = 1; i < n; ++i)
			workers.emplace_back(run, i);
		if (n)
			run(0);
		for (auto& t : workers)
			t.join();
		for (uint32_t i = 0; i < n; ++i)
			if (rcs[i] != TSOM_OK)
				return Fail(rcs[i], "part " + std::to_string(i) + ": " + msgs[i]);
		return TSOM_OK;
	}
}

struct tsom_sharded_planet
{
	uint32_t cc[3] = { 0, 0, 0 };
	float tile = 1.f;
	struct Part
	{
		tsom_ctx* ctx = nullptr;
		tsom_container* c = nullptr;
		int32_t box[6] = {};
		std::vector<int32_t> idx;     // the part's chunks, z, y, x loops inside its box (the order they were added in: slot == position)
		std::vector<uint32_t> linear; // their positions in Planet::GenerateChunks order
		std::vector<Key> extraDirty;  // chunks of this part that face an edited boundary layer of another part's chunk
	};
	std::vector<Part> parts;
	std::vector<uint8_t> owner; // per chunk in GenerateChunks order
	bool needExchange = false;  // blocks changed since the last exchange round
	std::mutex mutex;

	bool Inside(const int32_t ci[3]) const
	{
		for (int k = 0; k < 3; ++k)
		{
			const int v = ci[k] + int(cc[k] / 2);
			if (v < 0 || v >= int(cc[k]))
				return false;
		}
		return true;
	}
	size_t Linear(const int32_t ci[3]) const
	{
		const size_t x = size_t(ci[0] + int(cc[0] / 2)), y = size_t(ci[1] + int(cc[1] / 2)), z = size_t(ci[2] + int(cc[2] / 2));
		return (z * cc[1] + y) * cc[0] + x;
	}
};

// one exchange round over all parts: every publish is queued before any pull is (parts that share a GPU, or a host call that
// synchronises a whole device, can then never keep a peer's flag from being raised)
static int ShardedExchange(tsom_sharded_planet* p)
{
	const uint32_t np = uint32_t(p->parts.size());
	int rc = ForEachPart(np, [&](uint32_t i) { return tsom_halo_publish(p->parts[i].c); });
	if (rc != TSOM_OK)
		return rc;
	rc = ForEachPart(np, [&](uint32_t i) { return tsom_halo_pull(p->parts[i].c); });
	p->needExchange = false;
	return rc;
}

extern "C"
{
int tsom_shard_plan(const uint32_t chunk_count[3], uint32_t n_parts, const float* chunk_weights, int32_t* boxes_out)
{
	if (!chunk_count || !boxes_out || n_parts == 0 || chunk_count[0] == 0 || chunk_count[1] == 0 || chunk_count[2] == 0)
		return Fail(TSOM_ERR_INVALID, "bad arguments");
	if (uint64_t(chunk_count[0]) * chunk_count[1] * chunk_count[2] < n_parts)
		return Fail(TSOM_ERR_INVALID, "tsom_shard_plan: more parts than chunks");
	int lo[3], hi[3];
	for (int k = 0; k < 3; ++k)
	{
		lo[k] = -int(chunk_count[k] / 2); // Planet.cpp:392-394: indices = (x, y, z) - chunkCount / 2
		hi[k] = lo[k] + int(chunk_count[k]) - 1;
	}
	PlanCtx pc{ chunk_count, chunk_weights, boxes_out };
	return PlanSplit(pc, lo, hi, 0, n_parts);
}

void tsom_sharded_destroy(tsom_sharded_planet* p)
{
	if (!p)
		return;
	// containers first (a peer's buffers must outlive nobody's kernels: every stream is drained before anything is freed)
	for (auto& part : p->parts)
		if (part.ctx)
			tsom_synchronize(part.ctx);
	for (auto& part : p->parts)
		if (part.c)
			tsom_container_destroy(part.c);
	for (auto& part : p->parts)
		if (part.ctx)
			tsom_destroy(part.ctx);
	delete p;
}

int tsom_planet_create_sharded(const int* cuda_devices, uint32_t n_parts, const uint32_t chunk_count[3], float tile_size, const float* chunk_weights, tsom_sharded_planet** out)
{
	if (!out)
		return Fail(TSOM_ERR_INVALID, "out is null");
	*out = nullptr;
	if (!cuda_devices || !chunk_count || n_parts == 0 || n_parts > 255)
		return Fail(TSOM_ERR_INVALID, "bad arguments");
	std::vector<int32_t> boxes(size_t(n_parts) * 6);
	int rc = tsom_shard_plan(chunk_count, n_parts, chunk_weights, boxes.data());
	if (rc != TSOM_OK)
		return rc;
	auto p = new (std::nothrow) tsom_sharded_planet();
	if (!p)
		return Fail(TSOM_ERR_NOMEM, "out of host memory");
	auto fail = [&](int code) {
		const std::string msg = g_lastError;
		tsom_sharded_destroy(p);
		return Fail(code, msg);
	};
	for (int k = 0; k < 3; ++k)
		p->cc[k] = chunk_count[k];
	p->tile = tile_size;
	p->parts.resize(n_parts);
	p->owner.assign(size_t(chunk_count[0]) * chunk_count[1] * chunk_count[2], 0);
	for (uint32_t i = 0; i < n_parts; ++i)
	{
		auto& part = p->parts[i];
		std::memcpy(part.box, boxes.data() + size_t(i) * 6, sizeof(part.box));
		for (int z = part.box[2]; z <= part.box[5]; ++z)
			for (int y = part.box[1]; y <= part.box[4]; ++y)
				for (int x = part.box[0]; x <= part.box[3]; ++x)
				{
					const int32_t ci[3] = { x, y, z };
					part.idx.insert(part.idx.end(), ci, ci + 3);
					const size_t lin = p->Linear(ci);
					part.linear.push_back(uint32_t(lin));
					p->owner[lin] = uint8_t(i);
				}
		if ((rc = tsom_create(cuda_devices[i], &part.ctx)) != TSOM_OK)
			return fail(rc);
		const uint32_t n = uint32_t(part.idx.size() / 3);
		if ((rc = tsom_container_create(part.ctx, n, tile_size, &part.c)) != TSOM_OK)
			return fail(rc);
		if ((rc = tsom_container_add_chunks(part.c, part.idx.data(), n)) != TSOM_OK)
			return fail(rc);
	}
	// neighbours across part boundaries: the owner's slot of a chunk is its position in the owner's box (chunks were added in order)
	for (uint32_t i = 0; i < n_parts; ++i)
	{
		auto& part = p->parts[i];
		std::vector<int32_t> ridx;
		std::vector<uint32_t> rslot;
		std::vector<int32_t> rrank;
		std::vector<uint8_t> attached(n_parts, 0);
		const uint32_t n = uint32_t(part.idx.size() / 3);
		std::unordered_map<Key, uint8_t, KeyHash> seen;
		for (uint32_t k = 0; k < n; ++k)
		{
			const int32_t* ci = part.idx.data() + size_t(k) * 3;
			bool onFace = false;
			for (int a = 0; a < 3; ++a)
				onFace = onFace || ci[a] == part.box[a] || ci[a] == part.box[3 + a];
			if (!onFace)
				continue;
			for (int d = 0; d < 6; ++d)
			{
				const int32_t nb[3] = { ci[0] + kChunkDirOffset[d][0], ci[1] + kChunkDirOffset[d][1], ci[2] + kChunkDirOffset[d][2] };
				if (!p->Inside(nb))
					continue;
				const uint32_t o = p->owner[p->Linear(nb)];
				if (o == i)
					continue;
				const Key nk{ nb[0], nb[1], nb[2] };
				if (!seen.emplace(nk, 1).second)
					continue;
				const int32_t* ob = p->parts[o].box;
				const uint32_t slot = uint32_t(((nb[2] - ob[2]) * (ob[4] - ob[1] + 1) + (nb[1] - ob[1])) * (ob[3] - ob[0] + 1) + (nb[0] - ob[0]));
				ridx.insert(ridx.end(), nb, nb + 3);
				rslot.push_back(slot);
				rrank.push_back(int32_t(o));
				attached[o] = 1;
			}
		}
		for (uint32_t o = 0; o < n_parts; ++o)
			if (attached[o] && (rc = tsom_peer_attach_local(part.c, int(o), p->parts[o].c)) != TSOM_OK)
				return fail(rc);
		if (!rslot.empty() && (rc = tsom_container_add_remote_chunks(part.c, ridx.data(), rslot.data(), rrank.data(), uint32_t(rslot.size()))) != TSOM_OK)
			return fail(rc);
	}
	*out = p;
	return TSOM_OK;
}

uint32_t tsom_sharded_part_count(const tsom_sharded_planet* p) { return p ? uint32_t(p->parts.size()) : 0; }
tsom_ctx* tsom_sharded_ctx(tsom_sharded_planet* p, uint32_t part) { return (p && part < p->parts.size()) ? p->parts[part].ctx : nullptr; }
tsom_container* tsom_sharded_container(tsom_sharded_planet* p, uint32_t part) { return (p && part < p->parts.size()) ? p->parts[part].c : nullptr; }

int tsom_sharded_part_box(const tsom_sharded_planet* p, uint32_t part, int32_t box_out[6])
{
	if (!p || !box_out || part >= p->parts.size())
		return Fail(TSOM_ERR_INVALID, "bad arguments");
	std::memcpy(box_out, p->parts[part].box, sizeof(p->parts[part].box));
	return TSOM_OK;
}

int tsom_sharded_owner(const tsom_sharded_planet* p, const int32_t chunk_index[3], uint32_t* part_out)
{
	if (!p || !chunk_index || !part_out)
		return Fail(TSOM_ERR_INVALID, "bad arguments");
	if (!p->Inside(chunk_index))
		return Fail(TSOM_ERR_INVALID, "chunk lies outside the planet");
	*part_out = p->owner[p->Linear(chunk_index)];
	return TSOM_OK;
}

int tsom_sharded_set_block_table(tsom_sharded_planet* p, const tsom_block_desc* blocks, uint32_t n)
{
	if (!p)
		return Fail(TSOM_ERR_INVALID, "bad arguments");
	std::lock_guard<std::mutex> lock(p->mutex);
	for (auto& part : p->parts)
	{
		int rc = tsom_set_block_table(part.ctx, blocks, n);
		if (rc != TSOM_OK)
			return rc;
	}
	return TSOM_OK;
}

int tsom_sharded_generate(tsom_sharded_planet* p, uint32_t seed, const tsom_gen_blocks* ids)
{
	if (!p || !ids)
		return Fail(TSOM_ERR_INVALID, "bad arguments");
	std::lock_guard<std::mutex> lock(p->mutex);
	const uint32_t np = uint32_t(p->parts.size());
	// every part generates its own box (terrain tables only for the planet faces its box can reach) ...
	int rc = ForEachPart(np, [&](uint32_t i) {
		auto& part = p->parts[i];
		TSOM_LOCK(part.ctx);
		return GenerateEnqueue(part.c, seed, p->cc, ids, part.idx.data(), uint32_t(part.idx.size() / 3));
	});
	// ... then one exchange round, ordered on the devices, and only then does the host wait (Planet::GenerateChunks is blocking)
	if (rc == TSOM_OK)
		rc = ShardedExchange(p);
	const int rcFinish = ForEachPart(np, [&](uint32_t i) {
		auto& part = p->parts[i];
		TSOM_LOCK(part.ctx);
		return GenerateFinish(part.c);
	});
	return rc != TSOM_OK ? rc : rcFinish;
}

int tsom_sharded_apply_edits(tsom_sharded_planet* p, const tsom_edit* edits, uint32_t n)
{
	if (!p || (!edits && n))
		return Fail(TSOM_ERR_INVALID, "bad arguments");
	std::lock_guard<std::mutex> lock(p->mutex);
	const uint32_t np = uint32_t(p->parts.size());
	std::vector<std::vector<tsom_edit>> routed(np);
	for (uint32_t i = 0; i < n; ++i)
	{
		const tsom_edit& e = edits[i];
		if (e.x >= CS || e.y >= CS || e.z >= CS)
			return Fail(TSOM_ERR_INVALID, "edit coordinates out of range");
		if (!p->Inside(e.chunk))
			continue; // no such chunk: skipped like the callers of Chunk::UpdateBlock do
		const uint32_t o = p->owner[p->Linear(e.chunk)];
		routed[o].push_back(e);
		// the neighbour-mask rule of Planet.cpp:39-58 across a part boundary: the part that owns the facing chunk rebuilds it
		const int dirs[3] = { e.x == 0 ? TSOM_LEFT : (e.x == CS - 1 ? TSOM_RIGHT : -1), e.y == 0 ? TSOM_FRONT : (e.y == CS - 1 ? TSOM_BACK : -1),
			                  e.z == 0 ? TSOM_DOWN : (e.z == CS - 1 ? TSOM_UP : -1) };
		for (int d : dirs)
		{
			if (d < 0)
				continue;
			const int32_t nb[3] = { e.chunk[0] + kChunkDirOffset[d][0], e.chunk[1] + kChunkDirOffset[d][1], e.chunk[2] + kChunkDirOffset[d][2] };
			if (!p->Inside(nb))
				continue;
			const uint32_t no = p->owner[p->Linear(nb)];
			if (no != o)
				p->parts[no].extraDirty.push_back(Key{ nb[0], nb[1], nb[2] });
		}
	}
	p->needExchange = true;
	return ForEachPart(np, [&](uint32_t i) { return routed[i].empty() ? int(TSOM_OK) : tsom_apply_edits(p->parts[i].c, routed[i].data(), uint32_t(routed[i].size())); });
}

int tsom_sharded_collect_dirty(tsom_sharded_planet* p, int32_t* chunk_indices_out, uint32_t cap, uint32_t* n_out, int clear)
{
	if (!p || !n_out)
		return Fail(TSOM_ERR_INVALID, "bad arguments");
	std::lock_guard<std::mutex> lock(p->mutex);
	std::vector<Key> dirty;
	for (auto& part : p->parts)
	{
		const uint32_t n = uint32_t(part.idx.size() / 3);
		std::vector<int32_t> local(size_t(n) * 3);
		uint32_t got = 0;
		int rc = tsom_collect_dirty(part.c, local.data(), n, &got, 0);
		if (rc != TSOM_OK)
			return rc;
		for (uint32_t i = 0; i < got; ++i)
			dirty.push_back(Key{ local[i * 3], local[i * 3 + 1], local[i * 3 + 2] });
		// a neighbour across a part boundary is rebuilt only if it has content (ChunkEntities.cpp:108-111), like a local one
		for (const Key& k : part.extraDirty)
		{
			const uint32_t s = part.c->SlotOf(k);
			if (s != 0xFFFFFFFFu && part.c->hasContent[s])
				dirty.push_back(k);
		}
	}
	std::sort(dirty.begin(), dirty.end());
	dirty.erase(std::unique(dirty.begin(), dirty.end()), dirty.end());
	*n_out = uint32_t(dirty.size());
	if (clear)
	{
		if (chunk_indices_out && dirty.size() > cap)
			return Fail(TSOM_ERR_CAPACITY, "chunk_indices_out is too small; nothing was cleared");
		for (auto& part : p->parts)
		{
			uint32_t got = 0;
			int rc = tsom_collect_dirty(part.c, nullptr, 0, &got, 1);
			if (rc != TSOM_OK)
				return rc;
			part.extraDirty.clear();
		}
	}
	if (chunk_indices_out)
		for (uint32_t i = 0; i < dirty.size() && i < cap; ++i)
		{
			chunk_indices_out[i * 3 + 0] = dirty[i].x;
			chunk_indices_out[i * 3 + 1] = dirty[i].y;
			chunk_indices_out[i * 3 + 2] = dirty[i].z;
		}
	return TSOM_OK;
}

int tsom_sharded_mesh(tsom_sharded_planet* p, const int32_t* chunk_indices, uint32_t n, const tsom_mesh_params* params, tsom_mesh_view* views_out, uint32_t* parts_out)
{
	if (!p || !params)
		return Fail(TSOM_ERR_INVALID, "bad arguments");
	std::lock_guard<std::mutex> lock(p->mutex);
	const uint32_t np = uint32_t(p->parts.size());
	const size_t nAll = p->owner.size();
	if (!chunk_indices && (views_out || parts_out) && n != nAll)
		return Fail(TSOM_ERR_INVALID, "tsom_sharded_mesh: with chunk_indices == NULL, n must be the number of chunks of the planet");
	int rc = TSOM_OK;
	if (p->needExchange && (rc = ShardedExchange(p)) != TSOM_OK)
		return rc;

	// route the chunks to their owners (whole planet: the parts' own lists, already in their containers' order)
	struct Routed
	{
		std::vector<int32_t> idx;
		std::vector<uint32_t> pos; // position in the caller's list
		std::vector<float> gravity;
		std::vector<tsom_mesh_view> views;
	};
	std::vector<Routed> routed(np);
	if (chunk_indices)
	{
		for (uint32_t i = 0; i < n; ++i)
		{
			const int32_t* ci = chunk_indices + size_t(i) * 3;
			if (!p->Inside(ci))
				return Fail(TSOM_ERR_INVALID, "tsom_sharded_mesh: unknown chunk");
			Routed& r = routed[p->owner[p->Linear(ci)]];
			r.idx.insert(r.idx.end(), ci, ci + 3);
			r.pos.push_back(i);
			if (params->gravity_centers)
				r.gravity.insert(r.gravity.end(), params->gravity_centers + size_t(i) * 3, params->gravity_centers + size_t(i) * 3 + 3);
		}
	}
	else if (params->gravity_centers)
	{
		for (uint32_t i = 0; i < np; ++i)
			for (uint32_t lin : p->parts[i].linear)
				routed[i].gravity.insert(routed[i].gravity.end(), params->gravity_centers + size_t(lin) * 3, params->gravity_centers + size_t(lin) * 3 + 3);
	}
	auto listOf = [&](uint32_t i) -> const std::vector<int32_t>& { return chunk_indices ? routed[i].idx : p->parts[i].idx; };
	const bool wantViews = views_out != nullptr;
	rc = ForEachPart(np, [&](uint32_t i) {
		auto& part = p->parts[i];
		const auto& list = listOf(i);
		const uint32_t m = uint32_t(list.size() / 3);
		if (wantViews)
			routed[i].views.resize(m);
		if (m == 0)
			return int(TSOM_OK);
		tsom_mesh_params pp = *params;
		pp.gravity_centers = params->gravity_centers ? routed[i].gravity.data() : nullptr;
		TSOM_LOCK(part.ctx);
		return MeshEnqueue(part.c, list.data(), m, &pp, wantViews, nullptr);
	});
	const int rcFinish = ForEachPart(np, [&](uint32_t i) {
		auto& part = p->parts[i];
		TSOM_LOCK(part.ctx);
		return MeshFinish(part.c, wantViews ? routed[i].views.data() : nullptr);
	});
	if (rc != TSOM_OK)
		return rc;
	if (rcFinish != TSOM_OK)
		return rcFinish;
	for (uint32_t i = 0; i < np; ++i)
	{
		const uint32_t m = uint32_t(listOf(i).size() / 3);
		for (uint32_t k = 0; k < m; ++k)
		{
			const size_t pos = chunk_indices ? routed[i].pos[k] : p->parts[i].linear[k];
			if (views_out)
				views_out[pos] = routed[i].views[k];
			if (parts_out)
				parts_out[pos] = i;
		}
	}
	return TSOM_OK;
}

int tsom_sharded_mesh_copy_to_host(tsom_sharded_planet* p, uint32_t part, uint64_t first_face, uint64_t face_count, void* vertices, uint32_t* indices, float* collider_positions)
{
	if (!p || part >= p->parts.size())
		return Fail(TSOM_ERR_INVALID, "bad arguments");
	return tsom_mesh_copy_to_host(p->parts[part].ctx, first_face, face_count, vertices, indices, collider_positions);
}

int tsom_sharded_get_content(tsom_sharded_planet* p, const int32_t* chunk_indices, uint32_t n, uint16_t* blocks_host)
{
	if (!p || (!chunk_indices && n) || (!blocks_host && n))
		return Fail(TSOM_ERR_INVALID, "bad arguments");
	std::lock_guard<std::mutex> lock(p->mutex);
	for (uint32_t i = 0; i < n; ++i)
	{
		const int32_t* ci = chunk_indices + size_t(i) * 3;
		if (!p->Inside(ci))
			return Fail(TSOM_ERR_INVALID, "chunk is unknown or has no content");
		int rc = tsom_chunks_get_content(p->parts[p->owner[p->Linear(ci)]].c, ci, 1, blocks_host + size_t(i) * NB);
		if (rc != TSOM_OK)
			return rc;
	}
	return TSOM_OK;
}

int tsom_sharded_get_timings(tsom_sharded_planet* p, float* out)
{
	if (!p || !out)
		return Fail(TSOM_ERR_INVALID, "bad arguments");
	for (size_t i = 0; i < p->parts.size(); ++i)
	{
		int rc = tsom_get_timings(p->parts[i].ctx, out + i * TSOM_TIMING_COUNT, TSOM_TIMING_COUNT);
		if (rc != TSOM_OK)
			return rc;
	}
	return TSOM_OK;
}
} // extern "C"
