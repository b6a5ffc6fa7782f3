"""Whole-planet sharding across GPUs, one process per GPU: the planet is cut into one axis-aligned box of chunks per rank
(``tsom_shard_plan``: recursive bisection balanced by chunk weights), halo slabs are pulled over NVLink P2P.

Chunks are independent given the boundary layers of their six face-adjacent neighbours (reference src/CommonLib/Chunk.cpp:104-172).
Each rank generates and meshes its own box (chunks in the z, y, x order Planet::GenerateChunks walks, reference
src/CommonLib/Planet.cpp:387-393); between the two phases it copies the 2 KiB facing slab of every neighbour chunk owned by another
rank straight out of that rank's HBM (cudaIpc mapping + a copy kernel, ``tsom_halo_pull``).  The exchange is ordered ON THE DEVICE:
every container owns two epoch flags in its HBM that the peers poll over NVLink (see ``tsom_halo_pull`` in include/tsom_b200.h), so
there is neither a collective nor a host barrier on the data path; ``torch.distributed`` only carries the 200-byte IPC handles once.

``ShardPlan`` is integer host logic over the boxes the C library plans (tested on CPU with gloo, world_size 2); ``ShardedPlanet``
drives the C ABI.  The same plan and the same exchange run inside ONE process through ``tsom_planet_create_sharded``
(``thisspaceofmine_b200.MultiGpuPlanet``).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# Direction.hpp:83-90 (world axes): Back, Down, Front, Left, Right, Up
CHUNK_DIR_OFFSET = np.array([[0, 0, 1], [0, -1, 0], [0, 0, -1], [-1, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.int32)


def chunk_grid(chunkCount: Sequence[int]) -> np.ndarray:
    """Chunk indices in GenerateChunks order (z, y, x loops; (x,y,z) - chunkCount/2, Planet.cpp:387-393) -> [n, 3] int32."""
    nx, ny, nz = (int(v) for v in chunkCount)
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    return np.stack([x.ravel() - nx // 2, y.ravel() - ny // 2, z.ravel() - nz // 2], axis=1).astype(np.int32)


def plan_boxes(chunkCount: Sequence[int], world: int, weights=None) -> np.ndarray:
    """tsom_shard_plan: [world, 6] int32 = lo.xyz, hi.xyz (inclusive chunk indices) of every rank's box.  Pure host code of the
    library (no GPU needed)."""
    from . import _lib as L

    cc = np.ascontiguousarray(chunkCount, dtype=np.uint32)
    boxes = np.zeros((int(world), 6), np.int32)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32).reshape(-1)
    L.check(L.lib().tsom_shard_plan(cc.ctypes.data_as(C.POINTER(C.c_uint32)), int(world), None if w is None else w.ctypes.data_as(C.POINTER(C.c_float)),
                                    boxes.ctypes.data_as(C.POINTER(C.c_int32))))
    return boxes


class ShardPlan:
    """Who owns which chunk (one box per rank), and which neighbour chunks each rank must mirror."""

    def __init__(self, chunkCount: Sequence[int], world: int, weights=None):
        if world < 1:
            raise ValueError("world must be >= 1")
        self.chunkCount = tuple(int(v) for v in chunkCount)
        self.world = int(world)
        self.grid = chunk_grid(self.chunkCount)
        self.n = len(self.grid)
        self.boxes = plan_boxes(self.chunkCount, self.world, weights)
        self.owner = np.full(self.n, -1, np.int32)
        self._ids: List[np.ndarray] = []
        for r, b in enumerate(self.boxes):
            inside = np.all((self.grid >= b[0:3]) & (self.grid <= b[3:6]), axis=1)
            self.owner[inside] = r
            self._ids.append(np.nonzero(inside)[0].astype(np.int64))  # ascending linear id == z, y, x loops inside the box
        assert np.all(self.owner >= 0), "the boxes tile the planet"

    def linear_id(self, idx) -> np.ndarray:
        """Linear chunk id of chunk indices [k, 3]; -1 outside the planet."""
        a = np.asarray(idx, np.int64).reshape(-1, 3)
        nx, ny, nz = self.chunkCount
        x, y, z = a[:, 0] + nx // 2, a[:, 1] + ny // 2, a[:, 2] + nz // 2
        ok = (x >= 0) & (x < nx) & (y >= 0) & (y < ny) & (z >= 0) & (z < nz)
        return np.where(ok, (z * ny + y) * nx + x, -1)

    def local_ids(self, rank: int) -> np.ndarray:
        """Linear ids (GenerateChunks order) of the chunks of `rank`, in the order it adds them to its container."""
        return self._ids[rank]

    def local(self, rank: int) -> np.ndarray:
        return self.grid[self._ids[rank]]

    def slot_of(self, linear: np.ndarray) -> np.ndarray:
        """Slot of a chunk inside its owner's container: chunks are added in box order (z, y, x loops) into an empty container."""
        lin = np.asarray(linear, np.int64).reshape(-1)
        idx = self.grid[lin].astype(np.int64)
        b = self.boxes[self.owner[lin]].astype(np.int64)
        ex, ey = b[:, 3] - b[:, 0] + 1, b[:, 4] - b[:, 1] + 1
        return (((idx[:, 2] - b[:, 2]) * ey + (idx[:, 1] - b[:, 1])) * ex + (idx[:, 0] - b[:, 0])).astype(np.uint32)

    def remote_neighbours(self, rank: int) -> Dict[int, np.ndarray]:
        """peer rank -> sorted unique linear ids of the chunks it owns that are face-adjacent to a chunk of `rank`."""
        mine = self.local(rank)
        out: Dict[int, set] = {}
        for d in range(6):
            lid = self.linear_id(mine + CHUNK_DIR_OFFSET[d])
            lid = lid[lid >= 0]
            own = self.owner[lid]
            for peer in np.unique(own):
                if peer != rank:
                    out.setdefault(int(peer), set()).update(int(v) for v in lid[own == peer])
        return {p: np.array(sorted(v), np.int64) for p, v in sorted(out.items())}

    def route_edits(self, rank: int, edits) -> Tuple[np.ndarray, np.ndarray]:
        """A tick's edits, replicated on every rank ([n, 7] int: cx, cy, cz, x, y, z, block) -> what `rank` has to do with them:
        (mask of the edits whose chunk it owns — it applies those, in order —, sorted linear ids of ITS chunks that must be rebuilt
        because a chunk owned by someone else was edited on the boundary layer facing them).  The boundary rule is the reference's
        neighbour mask (Planet.cpp:39-58): x = 0 Left, x = 31 Right, y = 0 Front, y = 31 Back, z = 0 Down, z = 31 Up (local axes);
        the owner of an edited chunk finds its own masked neighbours through tsom_collect_dirty, so only the cross-rank pairs are
        listed here.  Edits outside the planet are owned by nobody."""
        e = np.asarray(edits, np.int64).reshape(-1, 7)
        lid = self.linear_id(e[:, 0:3])
        inside = lid >= 0
        own = np.zeros(len(e), bool)
        own[inside] = self.owner[lid[inside]] == rank
        foreign = inside & ~own
        touched = []
        # (local coordinate column, value on the boundary, direction of the neighbour that sees that layer)
        for col, edge, d in ((3, 0, 3), (3, 31, 4), (4, 0, 2), (4, 31, 0), (5, 0, 1), (5, 31, 5)):
            sel = foreign & (e[:, col] == edge)
            if sel.any():
                n = self.linear_id(e[sel, 0:3] + CHUNK_DIR_OFFSET[d])
                n = n[n >= 0]
                touched.append(n[self.owner[n] == rank])
        extra = np.unique(np.concatenate(touched)) if touched else np.zeros(0, np.int64)
        return own, extra.astype(np.int64)

    def halo_bytes(self, rank: int) -> int:
        """Bytes `rank` pulls per exchange: one 2 KiB slab per (local chunk, direction) whose neighbour is remote."""
        mine = self.local(rank)
        total = 0
        for d in range(6):
            lid = self.linear_id(mine + CHUNK_DIR_OFFSET[d])
            lid = lid[lid >= 0]
            total += int((self.owner[lid] != rank).sum()) * 2048
        return total


class ShardedPlanet:
    """One rank's share of a Planet (Planet::GenerateChunks + Chunk::BuildMesh over the rank's chunk range)."""

    def __init__(self, ctx, chunkCount: Sequence[int], rank: int, world: int, tileSize: float = 1.0, group=None, weights=None):
        from .host import Planet

        self.ctx, self.rank, self.world, self.group = ctx, int(rank), int(world), group
        self.plan = ShardPlan(chunkCount, world, weights)
        self.local_grid = np.ascontiguousarray(self.plan.local(rank))
        self.planet = Planet(ctx, tileSize, capacity=max(len(self.local_grid), 1))
        self._connected = False

    def close(self):
        self.planet.close()

    # ---- phase 1: generation (no communication)
    def GenerateChunks(self, blockLibrary, seed: int, wait: bool = True):
        """wait=False: only queue the work; PullHalos + BuildMeshes behind it are ordered on the stream and the mesh call completes all three."""
        if len(self.local_grid):
            self.planet.GenerateChunk(blockLibrary, self.local_grid, seed, self.plan.chunkCount, wait=wait)

    # ---- phase 1.5: halo exchange over P2P
    def _dist(self):
        import torch.distributed as dist

        return dist

    def ConnectPeers(self):
        """Once per planet: swap IPC handles of the block stores, map the peers, declare the remote neighbour chunks."""
        from . import _lib as L

        if self._connected or self.world == 1:
            self._connected = True
            return
        dist = self._dist()
        h = L.IpcHandle()
        L.check(L.lib().tsom_ipc_export(self.planet._h, C.byref(h)))
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, bytes(h.bytes), group=self.group)
        rem = self.plan.remote_neighbours(self.rank)
        for peer, lids in rem.items():
            ph = L.IpcHandle()
            C.memmove(ph.bytes, handles[peer], len(handles[peer]))
            L.check(L.lib().tsom_ipc_attach(self.planet._h, peer, C.byref(ph)))
            idx = np.ascontiguousarray(self.plan.grid[lids])
            slots = np.ascontiguousarray(self.plan.slot_of(lids))
            ranks = np.full(len(lids), peer, np.int32)
            L.check(L.lib().tsom_container_add_remote_chunks(self.planet._h, idx.ctypes.data_as(C.POINTER(C.c_int32)), slots.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                             ranks.ctypes.data_as(C.POINTER(C.c_int32)), len(lids)))
        self._connected = True

    def PullHalos(self):
        """One exchange round (collective: every rank runs every round).  Queued on the context's stream and ordered on the device —
        wait for the peers' blocks-final flags, copy the remote boundary slabs into the local ghost slabs, raise the slabs-pulled
        flag; no host barrier, no stream synchronisation."""
        from . import _lib as L

        if self.world == 1:
            return
        self.ConnectPeers()
        L.check(L.lib().tsom_halo_pull(self.planet._h))

    # ---- phase 2: meshing (no communication)
    def BuildMeshes(self, indices=None, **kw):
        return self.planet.BuildMeshes(self.local_grid if indices is None else indices, **kw)

    # ---- edits: every rank sees the tick's edit list, applies the edits of its own chunks and remeshes its own dirty chunks
    def UpdateBlocks(self, edits):
        """n x Chunk::UpdateBlock on a sharded planet.  `edits` ([n, 7] int) must be the same on every rank (it is a few hundred KB per
        tick; the server broadcasts it).  The owner of a chunk applies its edits in order; chunks of this rank that face an edited
        boundary layer of another rank's chunk are remembered for the next CollectInvalidated."""
        own, extra = self.plan.route_edits(self.rank, edits)
        e = np.asarray(edits).reshape(-1, 7)
        if own.any():
            self.planet.UpdateBlocks(e[own])
        self._extra_dirty = np.union1d(getattr(self, "_extra_dirty", np.zeros(0, np.int64)), extra)

    def CollectInvalidated(self, clear: bool = True) -> np.ndarray:
        """This rank's share of what ChunkEntities::Update would rebuild: its edited chunks, their masked neighbours on this rank, and
        its chunks next to boundary edits made on other ranks -> [k, 3] int32 sorted by (x, y, z)."""
        local = self.planet.CollectInvalidated(clear=clear)
        lids = self.plan.linear_id(local) if len(local) else np.zeros(0, np.int64)
        lids = lids[(lids >= 0)]
        lids = lids[self.plan.owner[lids] == self.rank]  # ghost chunks of other ranks are rebuilt by their owners
        lids = np.union1d(lids, getattr(self, "_extra_dirty", np.zeros(0, np.int64))).astype(np.int64)
        if clear:
            self._extra_dirty = np.zeros(0, np.int64)
        idx = self.plan.grid[lids]
        order = np.lexsort((idx[:, 2], idx[:, 1], idx[:, 0])) if len(idx) else np.zeros(0, np.int64)
        return np.ascontiguousarray(idx[order], np.int32)

    def Update(self, **kw):
        """One tick after UpdateBlocks (collective: every rank calls it): refresh the ghost slabs, rebuild this rank's dirty chunks.
        Returns (dirty chunk indices, MeshBatch or None)."""
        dirty = self.CollectInvalidated(clear=True)
        self.PullHalos()
        return dirty, (self.planet.BuildMeshes(dirty, **kw) if len(dirty) else None)
