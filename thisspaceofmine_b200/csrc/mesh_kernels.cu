// mesh_kernels.cu — face-culled chunk meshing on sm_100a (replaces Chunk::BuildMesh, reference src/CommonLib/Chunk.cpp:20-250).
//
// One kernel, mesh_fused_kernel (see its header comment below): stage the ids by TMA, classify, count, chain the chunk into
// the arena by decoupled look-back, enumerate the faces in the reference's order and stream them out with coalesced stores.
// A face is drawn iff self != Empty and (neighbour chunk absent, or neighbour id != self and neighbour is transparent)
// (Chunk.cpp:190-201).  An absent neighbour chunk reads as a slab of Empty ids, which is equivalent.
//
// Compiled with -fmad=false: the float math reproduces the reference's operation order without FMA contraction.
#include "tsom_kernels.h"
#include "tsom_device.cuh"

namespace tsomk
{
	constexpr int ROWS = CS * CS; // 1024 x-rows per chunk

	// staging: one face = 12 float4 (4 vertices x 48 B) padded to 13 so lane-strided 16-byte stores are conflict free
	constexpr int STAGE_FACE_F4 = 13;

	// emission order of the reference (Chunk.cpp:200-246) as Direction values
	__constant__ int c_slotDir[6] = { D_UP, D_DOWN, D_FRONT, D_BACK, D_LEFT, D_RIGHT };

	// Per emission slot, the 4 quad corners as bit codes dX | dY<<1 | dZ<<2 in world-local axes (X = local x, Y = up = local z,
	// Z = local y) — the effective Nz::BoxCorner semantics derived in SURVEY.md §8.1.  The double-sided twin swaps v0<->v1, v2<->v3.
	__constant__ unsigned char c_faceCorners[6][4] = {
		{ 0b011, 0b111, 0b010, 0b110 }, // Up    : (1,1,0) (1,1,1) (0,1,0) (0,1,1)
		{ 0b101, 0b001, 0b100, 0b000 }, // Down  : (1,0,1) (1,0,0) (0,0,1) (0,0,0)
		{ 0b001, 0b011, 0b000, 0b010 }, // Front : (1,0,0) (1,1,0) (0,0,0) (0,1,0)
		{ 0b111, 0b101, 0b110, 0b100 }, // Back  : (1,1,1) (1,0,1) (0,1,1) (0,0,1)
		{ 0b010, 0b110, 0b000, 0b100 }, // Left  : (0,1,0) (0,1,1) (0,0,0) (0,0,1)
		{ 0b111, 0b011, 0b101, 0b001 }, // Right : (1,1,1) (1,1,0) (1,0,1) (1,0,0)
	};

	// order in which the 8 corners are summed for the block centre (enum order of Nz::BoxCorner as listed at
	// DeformedChunk.cpp:123-130 under the §8.1 mapping)
	// (= codes 0b100, 0b101, 0b000, 0b001, 0b110, 0b111, 0b010, 0b011; spelled out in face_geometry)

	__device__ __forceinline__ uint32_t row_faces(const uint32_t v[6], uint32_t d)
	{
		uint32_t c = 0;
#pragma unroll
		for (int s = 0; s < 6; ++s)
			c += __popc(v[s]) + __popc(v[s] & d);
		return c;
	}

	// =====================================================================================================================
	struct FaceCtx
	{
		float tile;
		F3 gravity; // Chunk::BuildMesh `center` argument; also the deformation centre
		float radius;
		bool deformed;
		bool wantAttr; // normals + uv (render) — collider-only skips them like null SparsePtrs do (Chunk.cpp:33,39)
		// Table-driven flat path only: faceCenter - gravityCenter is EXACT (power-of-two block size, container centre 0, chunk indices
		// below 2^14: every coordinate is a multiple of tile / 2 far below 2^24), so DirectionFromNormal gives the same answer on the
		// raw vector as on the normalised one — scaling by 1 / length keeps equal components equal and distinct ones (at least
		// 2^-21 apart, relatively) distinct, and the largest component is >= 0 > the -1 the search starts from either way.
		bool rawUp;
		// chunk-local index that maps to coordinate 0: size / 2 for a FlatChunk (FlatChunk.cpp:50-53), 0 for the corners of a
		// DeformedChunk (DeformedChunk.cpp:117-119) when the table-driven path serves a chunk the deformation leaves untouched
		float flatOrigin;
		bool exactDeform; // deform_is_exact: regions may be classified as undeformed without evaluating DeformPosition
		const uint8_t* tex8; // shared-memory copy of the texture-slice table ([TEX8N][6] bytes) or nullptr
	};
	constexpr int TEX8N = 128;

	// DeformedChunk::DeformPosition (reference src/CommonLib/DeformedChunk.cpp:139-154)
	__device__ __forceinline__ F3 deform_position(const FaceCtx& cx, F3 p)
	{
		const F3 c = cx.gravity;
		float dist = fmaxf(fmaxf(fabsf(p.x - c.x), fabsf(p.y - c.y)), fabsf(p.z - c.z));
		float inner = fmaxf(dist - cx.radius, 0.f);
		F3 bmin = c - f3(inner, inner, inner);
		float len = inner * 2.f;
		F3 bmax = f3(bmin.x + len, bmin.y + len, bmin.z + len);
		F3 ip = f3(fmaxf(fminf(p.x, bmax.x), bmin.x), fmaxf(fminf(p.y, bmax.y), bmin.y), fmaxf(fminf(p.z, bmax.z), bmin.z));
		F3 n = normalize(p - ip);
		return ip + n * fminf(cx.radius, dist);
	}

	// one voxel corner, code = dX | dY<<1 | dZ<<2 (FlatChunk.cpp:48-54 / DeformedChunk.cpp:115-137)
	__device__ __forceinline__ F3 voxel_corner(const FaceCtx& cx, int x, int y, int z, unsigned code)
	{
		const float s = cx.tile;
		F3 p;
		if (!cx.deformed)
		{
			float bx = (float(x) - float(CS) * 0.5f) * s;
			float by = (float(y) - float(CS) * 0.5f) * s;
			float bz = (float(z) - float(CS) * 0.5f) * s;
			// Boxf(blockPos.x, blockPos.z, blockPos.y, s, s, s)
			p = f3((code & 1u) ? bx + s : bx, (code & 2u) ? bz + s : bz, (code & 4u) ? by + s : by);
		}
		else
		{
			float fx = float(unsigned(x)) * s;
			float fy = float(unsigned(y)) * s;
			float fz = float(unsigned(z)) * s;
			p = f3((code & 1u) ? fx + s : fx, (code & 2u) ? fz + s : fz, (code & 4u) ? fy + s : fy);
			p = deform_position(cx, p);
		}
		return p;
	}

	// DeformedChunk corner plus whether the deformation left it where it was (bit for bit up to the sign of a zero): true for every
	// corner further than the deformation radius from two of the planet's faces, i.e. for most of a planet's surface
	__device__ __forceinline__ F3 deformed_corner(const FaceCtx& cx, int x, int y, int z, unsigned code, bool& same)
	{
		const float s = cx.tile;
		const float fx = float(unsigned(x)) * s;
		const float fy = float(unsigned(y)) * s;
		const float fz = float(unsigned(z)) * s;
		const F3 p = f3((code & 1u) ? fx + s : fx, (code & 2u) ? fz + s : fz, (code & 4u) ? fy + s : fy);
		const F3 d = deform_position(cx, p);
		same = d.x == p.x && d.y == p.y && d.z == p.z;
		return d;
	}

	// DeformPosition returns every point of an axis-aligned region unchanged, bit for bit, when
	//   * one axis m dominates the whole region and the other two stay inside the inner box: max |p_j - c_j| <= min |p_m - c_m| - radius.
	//     Then the clamp moves a point by exactly (+-radius, 0, 0) along m, the normalised offset is (+-1, 0, 0) and the point comes
	//     back to where it was — most of a planet: everything further than the radius from two of its faces (region_is_undeformed);
	//   * and the arithmetic on the way is exact (deform_is_exact): centre and radius are multiples of 2^-10 below 2^13 like every
	//     corner coordinate (power-of-two block size: the caller checked), so no sum or difference rounds; sqrt(radius^2) == radius
	//     in binary floating point, and radius * (1 / radius) == 1 is checked (it fails for a few radii).
	// Conservative: whatever fails takes the per-corner path, which compares every deformed corner with the undeformed one.
	__device__ __forceinline__ bool deform_is_exact(const FaceCtx& cx)
	{
		const float r = cx.radius, s = cx.tile;
		auto gridded = [](float v) { return fabsf(v) < 8192.f && v * 1024.f == truncf(v * 1024.f); };
		if (!(r >= 0.f) || !gridded(r) || !gridded(cx.gravity.x) || !gridded(cx.gravity.y) || !gridded(cx.gravity.z) || !(s >= 1.f / 1024.f) || !(float(CS) * s < 8192.f))
			return false;
		return r == 0.f || r * (1.f / r) == 1.f;
	}

	// region = [lo, lo + ext] on each axis, in the chunk's own coordinates
	__device__ __forceinline__ bool region_is_undeformed(const FaceCtx& cx, float lox, float loy, float loz, float ext)
	{
		const float lo[3] = { lox - cx.gravity.x, loy - cx.gravity.y, loz - cx.gravity.z };
		float qmin[3], qmax[3];
#pragma unroll
		for (int k = 0; k < 3; ++k)
		{
			const float hi = lo[k] + ext;
			qmax[k] = fmaxf(fabsf(lo[k]), fabsf(hi));
			qmin[k] = (lo[k] <= 0.f && hi >= 0.f) ? 0.f : fminf(fabsf(lo[k]), fabsf(hi));
		}
		const float lead = fmaxf(fmaxf(qmin[0], qmin[1]), qmin[2]);
		// the two non-leading axes: the largest qmax among the axes whose qmin is not the lead (ties: any one leads, the rest must fit)
		const int m = qmin[0] == lead ? 0 : (qmin[1] == lead ? 1 : 2);
		const float other = m == 0 ? fmaxf(qmax[1], qmax[2]) : (m == 1 ? fmaxf(qmax[0], qmax[2]) : fmaxf(qmax[0], qmax[1]));
		return other <= lead - cx.radius;
	}

	// geometry of one face: block centre (mean of the 8 corners, summed in enum order, Chunk.cpp:186-188), the 4 quad corners
	// and their mean (Chunk.cpp:30).  The 8 corners are evaluated once (the deformed ones cost a normalisation each) and the quad
	// corners are picked from them by their 3-bit codes.
	__device__ __forceinline__ void face_geometry(const FaceCtx& cx, int x, int y, int z, int slot, bool twin, F3& blockCenter, F3 pos[4], F3& fc)
	{
		F3 c[8]; // indexed by corner code dX | dY << 1 | dZ << 2
#pragma unroll
		for (int code = 0; code < 8; ++code)
			c[code] = voxel_corner(cx, x, y, z, unsigned(code));
		// enum order of Nz::BoxCorner (c_cornerSumOrder): 0b100, 0b101, 0b000, 0b001, 0b110, 0b111, 0b010, 0b011
		F3 sum = f3(0.f, 0.f, 0.f);
		sum = sum + c[4];
		sum = sum + c[5];
		sum = sum + c[0];
		sum = sum + c[1];
		sum = sum + c[6];
		sum = sum + c[7];
		sum = sum + c[2];
		sum = sum + c[3];
		blockCenter = sum / 8.f;
#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			const unsigned code = c_faceCorners[slot][twin ? (i ^ 1) : i];
			const bool b0 = (code & 1u) != 0u, b1 = (code & 2u) != 0u, b2 = (code & 4u) != 0u;
			auto pick = [&](float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7) {
				const float lo = b1 ? (b0 ? v3 : v2) : (b0 ? v1 : v0);
				const float hi = b1 ? (b0 ? v7 : v6) : (b0 ? v5 : v4);
				return b2 ? hi : lo;
			};
			pos[i] = f3(pick(c[0].x, c[1].x, c[2].x, c[3].x, c[4].x, c[5].x, c[6].x, c[7].x), pick(c[0].y, c[1].y, c[2].y, c[3].y, c[4].y, c[5].y, c[6].y, c[7].y),
			            pick(c[0].z, c[1].z, c[2].z, c[3].z, c[4].z, c[5].z, c[6].z, c[7].z));
		}
		fc = f3(0.f, 0.f, 0.f);
#pragma unroll
		for (int i = 0; i < 4; ++i)
			fc = fc + pos[i];
		fc = fc / 4.f;
	}

	// normal + texture direction + per-vertex uv of one face for a given faceUp (DrawFace, Chunk.cpp:31-92)
	__device__ __forceinline__ void face_attrs(const float (*quat)[4], int faceUp, const F3& blockCenter, const F3 pos[4], const F3& fc, F3& normal, int& texDir, float u[4], float v[4])
	{
		normal = normalize(fc - blockCenter);
		const float qw = quat[faceUp][0];
		const F3 qv = f3(quat[faceUp][1], quat[faceUp][2], quat[faceUp][3]);
		texDir = direction_from_normal(quat_rotate(qw, qv, normal));
#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			F3 dir = quat_rotate(qw, qv, pos[i] - blockCenter);
			// the three cases of Chunk.cpp:52-88 as selects (texDir differs from lane to lane: no divergent branches)
			const bool zFace = texDir == D_BACK || texDir == D_FRONT, yFace = texDir == D_DOWN || texDir == D_UP;
			const float major = zFace ? dir.x : (yFace ? dir.y : dir.z);
			const float mag = 0.5f / fabsf(major);
			const float uz = dir.x < 0.f ? -dir.z : dir.z;  // Back / Front
			const float ux = dir.z < 0.f ? dir.x : -dir.x;  // Left / Right
			const float vy = dir.y < 0.f ? -dir.z : dir.z;  // Down / Up
			const float uu = zFace ? uz : (yFace ? dir.x : ux);
			const float vv = yFace ? vy : -dir.y;
			u[i] = uu * mag + 0.5f;
			v[i] = vv * mag + 0.5f;
		}
	}

	// out of line so that the shared-memory path below is a real branch around the global load (inlined, both loads were issued)
	__device__ __noinline__ unsigned tex_slice_global(const uint32_t* blkTex, uint32_t nTypes, unsigned id, int texDir)
	{
		return __ldg(blkTex + size_t(id < nTypes ? id : nTypes - 1) * 6 + texDir);
	}

	__device__ __forceinline__ unsigned tex_slice(const MeshArgs& a, const FaceCtx& cx, unsigned id, int texDir)
	{
		// the texture slices of the first TEX8N block types as bytes in shared memory (when they all fit a byte)
		if (cx.tex8 != nullptr && id < unsigned(TEX8N))
			return cx.tex8[id * 6u + unsigned(texDir)];
		return tex_slice_global(a.blkTex, a.nTypes, id, texDir);
	}

	// one face held in registers: 4 positions, the shared normal, 4 uv pairs and the texture slice
	struct FaceOut
	{
		F3 pos[4];
		F3 n;
		float u[4], v[4];
		float slice;
	};

	// Nz::SubMesh::GenerateTangents as ClientChunkEntities::BuildMesh runs it after Chunk::BuildMesh (ClientChunkEntities.cpp:169), for
	// one face.  NazaraEngine is not in the reference tree: this restates the engine's algorithm (UNPINNED third-party arithmetic, single
	// CPU statement in oracle/nazara_tangents.hpp): every triangle adds coef * (dv0 * duv1.y - dv1 * duv0.y), coef = 1 / (duv0.x * duv1.y -
	// duv1.x * duv0.y), to its three vertices, then every vertex tangent is normalised.  A face owns its 4 vertices and its two triangles
	// (0, 2, 1) and (1, 2, 3) (Chunk.cpp:95-101), so the sums close inside the face: v0 <- A, v1 and v2 <- A + B, v3 <- B.
	__device__ __forceinline__ F3 triangle_tangent(const F3& p0, const F3& p1, const F3& p2, float u0, float v0, float u1, float v1, float u2, float v2)
	{
		const F3 dv0 = p1 - p0, dv1 = p2 - p0;
		const float du0 = u1 - u0, dw0 = v1 - v0, du1 = u2 - u0, dw1 = v2 - v0;
		const float coef = 1.f / (du0 * dw1 - du1 * dw0);
		return f3(coef * (dv0.x * dw1 + dv1.x * (-dw0)), coef * (dv0.y * dw1 + dv1.y * (-dw0)), coef * (dv0.z * dw1 + dv1.z * (-dw0)));
	}

	__device__ __forceinline__ void face_tangents(const FaceOut& f, F3 t[4])
	{
		const F3 tA = triangle_tangent(f.pos[0], f.pos[2], f.pos[1], f.u[0], f.v[0], f.u[2], f.v[2], f.u[1], f.v[1]);
		const F3 tB = triangle_tangent(f.pos[1], f.pos[2], f.pos[3], f.u[1], f.v[1], f.u[2], f.v[2], f.u[3], f.v[3]);
		const F3 zero = f3(0.f, 0.f, 0.f);
		t[0] = normalize(zero + tA);
		t[1] = normalize((zero + tA) + tB);
		t[2] = normalize((zero + tA) + tB);
		t[3] = normalize(zero + tB);
	}

	// 4 vertices as 12 float4 {px py pz nx | ny nz u v | w tx ty tz}; the tangent slot stays zero (Chunk::BuildMesh never writes it)
	// unless the caller asked for what GenerateTangents adds afterwards
	__device__ __forceinline__ void store_face(float4* out, const FaceOut& f, bool wantTangents)
	{
		F3 t[4];
#pragma unroll
		for (int i = 0; i < 4; ++i)
			t[i] = f3(0.f, 0.f, 0.f);
		if (wantTangents)
			face_tangents(f, t);
#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			out[i * 3 + 0] = make_float4(f.pos[i].x, f.pos[i].y, f.pos[i].z, f.n.x);
			out[i * 3 + 1] = make_float4(f.n.y, f.n.z, f.u[i], f.v[i]);
			out[i * 3 + 2] = make_float4(f.slice, t[i].x, t[i].y, t[i].z);
		}
	}

	// the same face without tangents in 9 float4: the third float4 {w 0 0 0} of the four vertices is staged once
	__device__ __forceinline__ void store_face_compact(float4* out, const FaceOut& f)
	{
#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			out[i] = make_float4(f.pos[i].x, f.pos[i].y, f.pos[i].z, f.n.x);
			out[4 + i] = make_float4(f.n.y, f.n.z, f.u[i], f.v[i]);
		}
		out[8] = make_float4(f.slice, 0.f, 0.f, 0.f);
	}

	// DrawFace (Chunk.cpp:22-102), generic arithmetic
	__device__ __forceinline__ void draw_face_r(const MeshArgs& a, const FaceCtx& cx, int x, int y, int z, int slot, bool twin, unsigned id, FaceOut& o)
	{
		F3 blockCenter, fc;
		face_geometry(cx, x, y, z, slot, twin, blockCenter, o.pos, fc);

		o.n = f3(0.f, 0.f, 0.f);
#pragma unroll
		for (int i = 0; i < 4; ++i)
			o.u[i] = o.v[i] = 0.f;
		o.slice = 0.f;
		if (cx.wantAttr)
		{
			const int faceUp = direction_from_normal(normalize(fc - cx.gravity));
			int texDir;
			face_attrs(a.quat, faceUp, blockCenter, o.pos, fc, o.n, texDir, o.u, o.v);
			o.slice = float(tex_slice(a, cx, id, texDir));
		}
	}

	// DrawFace with the block's geometry already known: the four quad corners (in emission order) and the block centre
	__device__ __forceinline__ void draw_face_cached(const MeshArgs& a, const FaceCtx& cx, const F3& blockCenter, const F3 pos[4], unsigned id, FaceOut& o)
	{
		F3 fc = f3(0.f, 0.f, 0.f);
#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			o.pos[i] = pos[i];
			fc = fc + pos[i];
		}
		fc = fc / 4.f;
		o.n = f3(0.f, 0.f, 0.f);
#pragma unroll
		for (int i = 0; i < 4; ++i)
			o.u[i] = o.v[i] = 0.f;
		o.slice = 0.f;
		if (cx.wantAttr)
		{
			const int faceUp = direction_from_normal(normalize(fc - cx.gravity));
			int texDir;
			face_attrs(a.quat, faceUp, blockCenter, o.pos, fc, o.n, texDir, o.u, o.v);
			o.slice = float(tex_slice(a, cx, id, texDir));
		}
	}

	// ---- table-driven flat path ----
	// For a FlatChunk whose block size is a power of two every intermediate of DrawFace is exact in fp32, so the normal, the
	// texture direction and the four uv pairs of a face depend only on (faceUp, slot, twin) — not on the block.  They are
	// evaluated once by face_table_kernel with the generic arithmetic above (so both paths are bit-identical) and then looked up.
	__global__ void face_table_kernel(FaceTable* table, float tile, MeshArgs a)
	{
		const int t = threadIdx.x;
		if (t >= 72)
			return;
		const int faceUp = t / 12, slot = (t >> 1) % 6;
		const bool twin = t & 1;
		FaceCtx cx;
		cx.tile = tile;
		cx.gravity = f3(0.f, 0.f, 0.f);
		cx.radius = 0.f;
		cx.deformed = false;
		cx.wantAttr = true;
		F3 blockCenter, pos[4], fc, normal;
		face_geometry(cx, 0, 0, 0, slot, twin, blockCenter, pos, fc);
		int texDir;
		float u[4], v[4];
		face_attrs(a.quat, faceUp, blockCenter, pos, fc, normal, texDir, u, v);
		table->uv[faceUp][slot][twin][0] = make_float4(u[0], v[0], u[1], v[1]);
		table->uv[faceUp][slot][twin][1] = make_float4(u[2], v[2], u[3], v[3]);
		table->texDir[faceUp][slot] = uint32_t(texDir);
		if (faceUp == 0 && !twin)
		{
			// .w: the corner codes of the slot's four vertices, 3 bits each, in emission order (bits 0-11) and in the twin's order (bits 16-27)
			uint32_t packed = 0u;
			for (int i = 0; i < 4; ++i)
				packed |= (uint32_t(c_faceCorners[slot][i]) << (3 * i)) | (uint32_t(c_faceCorners[slot][i ^ 1]) << (16 + 3 * i));
			table->normal[slot] = make_float4(normal.x, normal.y, normal.z, __uint_as_float(packed));
		}
	}

	void launch_face_table(FaceTable* table, float tile, const MeshArgs& a, cudaStream_t st) { face_table_kernel<<<1, 96, 0, st>>>(table, tile, a); }

	__device__ __forceinline__ void draw_face_flat_r(const MeshArgs& a, const FaceCtx& cx, const FaceTable& tb, float origin, int x, int y, int z, int slot, bool twin, unsigned id, FaceOut& o)
	{
		const float s = cx.tile;
		// Boxf(blockPos.x, blockPos.z, blockPos.y, s, s, s) with blockPos = (indices - size / 2) * s (FlatChunk.cpp:50-53)
		const float ox = (float(x) - origin) * s;
		const float oy = (float(z) - origin) * s;
		const float oz = (float(y) - origin) * s;
		const float ex = ox + s, ey = oy + s, ez = oz + s;
		// one shared-memory load instead of four constant-bank loads at a per-lane index (the slot differs from lane to lane)
		const float4 nc = tb.normal[slot];
		const uint32_t codes = twin ? __float_as_uint(nc.w) >> 16 : __float_as_uint(nc.w);
#pragma unroll
		for (int i = 0; i < 4; ++i)
			o.pos[i] = f3((codes >> (3 * i)) & 1u ? ex : ox, (codes >> (3 * i + 1)) & 1u ? ey : oy, (codes >> (3 * i + 2)) & 1u ? ez : oz);
		float4 uv0 = make_float4(0.f, 0.f, 0.f, 0.f), uv1 = uv0, n = uv0;
		o.slice = 0.f;
		if (cx.wantAttr)
		{
			F3 fc = f3(0.f, 0.f, 0.f);
#pragma unroll
			for (int i = 0; i < 4; ++i)
				fc = fc + o.pos[i];
			fc = fc / 4.f;
			const int faceUp = cx.rawUp ? direction_from_normal(fc - cx.gravity) : direction_from_normal(normalize(fc - cx.gravity));
			uv0 = tb.uv[faceUp][slot][twin][0];
			uv1 = tb.uv[faceUp][slot][twin][1];
			n = nc;
			o.slice = float(tex_slice(a, cx, id, int(tb.texDir[faceUp][slot])));
		}
		o.n = f3(n.x, n.y, n.z);
		o.u[0] = uv0.x; o.v[0] = uv0.y;
		o.u[1] = uv0.z; o.v[1] = uv0.w;
		o.u[2] = uv1.x; o.v[2] = uv1.y;
		o.u[3] = uv1.z; o.v[3] = uv1.w;
	}

	// DrawFace of a DeformedChunk block whose 8 corners the deformation left in place (deformed_corner): with a power-of-two block
	// size its corners are exact multiples of the block size, so normal, texture direction and uvs are the table's — the generic
	// arithmetic on these inputs produces the very same bits (it is what filled the table)
	__device__ __forceinline__ void draw_face_cached_table(const MeshArgs& a, const FaceCtx& cx, const FaceTable& tb, const F3 pos[4], int slot, bool twin, unsigned id, FaceOut& o)
	{
		F3 fc = f3(0.f, 0.f, 0.f);
#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			o.pos[i] = pos[i];
			fc = fc + pos[i];
		}
		fc = fc / 4.f;
		float4 uv0 = make_float4(0.f, 0.f, 0.f, 0.f), uv1 = uv0, n = uv0;
		o.slice = 0.f;
		if (cx.wantAttr)
		{
			const int faceUp = direction_from_normal(normalize(fc - cx.gravity));
			uv0 = tb.uv[faceUp][slot][twin][0];
			uv1 = tb.uv[faceUp][slot][twin][1];
			n = tb.normal[slot];
			o.slice = float(tex_slice(a, cx, id, int(tb.texDir[faceUp][slot])));
		}
		o.n = f3(n.x, n.y, n.z);
		o.u[0] = uv0.x; o.v[0] = uv0.y;
		o.u[1] = uv0.z; o.v[1] = uv0.w;
		o.u[2] = uv1.x; o.v[2] = uv1.y;
		o.u[3] = uv1.z; o.v[3] = uv1.w;
	}

	// =====================================================================================================================
	// mesh_fused_kernel — the whole of Chunk::BuildMesh for a batch of chunks in ONE pass over the ids.
	//
	//   * 256 threads, 2 CTAs per SM (113 KB of shared memory each): while one CTA streams faces out, the other stages and
	//     classifies its next chunk, so the front end hides behind the HBM writes instead of serialising with them.
	//   * ids arrive by TMA bulk copies (4 x 16 KiB on one mbarrier); the six halo slabs are read straight from global memory
	//     (16-byte loads) while the TMA is in flight, and only their 32x32 transparency bits are kept.
	//   * row masks: each thread classifies 8 ids per 16-byte shared-memory load through a packed flag table
	//     (transparent | non-empty << 8 | double-sided << 16), with a one-lookup shortcut when the 8 ids are equal.
	//   * the per-chunk face count is chained across chunks by a decoupled look-back over `status` (job order == ticket order),
	//     so every chunk learns its first face in the arena without a separate count kernel; `firstFace`/`faceCount` are
	//     identical to an exclusive scan in job order.  If the arena is too small the kernel only counts; the host grows the
	//     arena to the exact total and runs it again.
	//   * faces are enumerated from a compact list of the rows that have faces (`ne`): a warp takes a window of 128 consecutive
	//     faces, finds the first row by a two-level ballot search, and lanes walk one row each, writing 16-bit face descriptors
	//     in the reference's emission order.  Faces are then built one per lane in registers and streamed out through a
	//     16-face staging tile with fully coalesced 16-byte stores.
	// The arenas are written once and read by someone else much later (a copy to the host, a renderer): streaming stores
	// (st.global.cs, evict-first in L2) keep 40 GB of output from pushing the ids the mesher still has to read out of L2 (47^3: 8.44 -> 8.34 ms)
#define ST_OUT(ptr, val) __stcs((ptr), (val))
	constexpr int MODE_LEAN = 0, MODE_GENERIC = 1, MODE_TANGENTS = 2;
	constexpr int FUSED_THREADS = 256;
	constexpr int FUSED_WARPS = FUSED_THREADS / 32;
	constexpr int WIN = 128;        // faces per descriptor window
	constexpr int HALF_FACES = 16;  // faces per staging round
	constexpr int LUTN = 128;       // block types below this id are classified from shared memory
	constexpr uint32_t NE_OFF_MASK = 0x7FFFFu; // ne[k] = first face of the row (19 bits; a chunk has at most 393,216 faces) | row << 19
	constexpr int NE_ROW_SHIFT = 19;

	struct FusedSmem
	{
		uint16_t ids[NB];                 // 65536 B, TMA destination
		uint32_t mT[ROWS], mE[ROWS], mD[ROWS];
		uint32_t ne[ROWS + 4];            // compact list of rows with faces, + sentinel ne[K] = F
		uint32_t hT[6][CS];               // transparency bits of the six halo slabs (Left/Right: bit = y, row = z)
		float4 stage[FUSED_WARPS][HALF_FACES * STAGE_FACE_F4];
		uint16_t desc[FUSED_WARPS][WIN];  // x | slot << 5 | twin << 8 | (compact row - first compact row of the window) << 9
		uint32_t lut32[LUTN];
		uint8_t tex8[TEX8N * 6];          // texture slice per (block type, direction) when MeshArgs::texFits8
		alignas(16) FaceTable table;
		uint32_t warpSum[FUSED_WARPS];
		uint32_t job, slot, hasContent;   // the ticket and what it names (written by the lane that took the ticket)
		uint32_t warpsDone;               // warps that have finished the current chunk; the last one takes the next ticket
		uint32_t ffSeq;                   // == job + 1 once firstFace of the current job is known
		unsigned long long firstFace;
		alignas(8) uint64_t bar;
	};
	static_assert(sizeof(FusedSmem) <= 113 * 1024, "two CTAs of mesh_fused_kernel must fit one SM");

	__device__ __forceinline__ uint32_t pack_flags(unsigned f, unsigned id)
	{
		return ((f & BF_TRANSPARENT) ? 1u : 0u) | (id != 0u ? 0x100u : 0u) | ((f & BF_DOUBLE_SIDED) ? 0x10000u : 0u);
	}

	__device__ __forceinline__ uint32_t flags32(const MeshArgs& a, const uint32_t* lut, unsigned id)
	{
		if (id < LUTN)
			return lut[id];
		return pack_flags(__ldg(a.blkFlags + (id < a.nTypes ? id : a.nTypes - 1)), id);
	}

	// 8 ids -> t8 | e8 << 8 | d8 << 16 (bit j of each byte = id j); out of line: only taken when an id is >= LUTN
	__device__ __noinline__ uint32_t flags8_slow(const MeshArgs& a, const uint32_t* lut, uint4 w)
	{
		const uint32_t ws[4] = { w.x, w.y, w.z, w.w };
		uint32_t acc = 0;
#pragma unroll
		for (int j = 0; j < 4; ++j)
		{
			acc |= flags32(a, lut, ws[j] & 0xffffu) << (2 * j);
			acc |= flags32(a, lut, ws[j] >> 16) << (2 * j + 1);
		}
		return acc;
	}

	// branch-free classification of 8 ids that are all < LUTN (callers mask; ids >= LUTN are fixed up separately):
	// the three flag bits of an entry sit 8 bits apart, so the 8 shifted entries add up without carries
	__device__ __forceinline__ uint32_t flags8_fast(const uint32_t* lut, uint4 w)
	{
		const uint32_t ws[4] = { w.x, w.y, w.z, w.w };
		uint32_t acc = 0;
#pragma unroll
		for (int j = 0; j < 4; ++j)
		{
			acc += lut[ws[j] & uint32_t(LUTN - 1)] << (2 * j);
			acc += lut[(ws[j] >> 16) & uint32_t(LUTN - 1)] << (2 * j + 1);
		}
		return acc;
	}

	__device__ __forceinline__ bool has_big_id(uint4 w) { return ((w.x | w.y | w.z | w.w) & ~(uint32_t(LUTN - 1) * 0x10001u)) != 0u; }

	// transparency bits of the six facing slabs, read from global memory; an absent neighbour reads as a slab of Empty.
	// All loads of a thread are issued before any is used (no branch or vote in between), so their latencies overlap.
	template<int NT>
	__device__ __forceinline__ void build_halo_masks(const MeshArgs& a, const uint32_t* lut, uint32_t slot, uint32_t (*hT)[CS], int tid)
	{
		constexpr int N = 6 * (SLAB / 8) / NT;
		const unsigned long long* hp = a.halo + size_t(slot) * 6;
		const uint32_t emptyT = (lut[0] & 1u) ? 0xFFu : 0u;
		unsigned long long e[N];
		uint4 w[N];
#pragma unroll
		for (int j = 0; j < N; ++j)
			e[j] = hp[(tid + j * NT) >> 7];
#pragma unroll
		for (int j = 0; j < N; ++j)
		{
			const int q = (tid + j * NT) & 127, row = q >> 2, seg = q & 3;
			// an absent neighbour loads from the chunk's own ids instead (any valid address; the result is discarded)
			const uint16_t* src = e[j] != 0ull ? reinterpret_cast<const uint16_t*>(e[j] & ~HALO_KIND_MASK) : a.blocks + size_t(slot) * NB;
			const size_t rowStride = (e[j] & HALO_KIND_MASK) == HALO_ROWS ? SLAB : CS;
			w[j] = *reinterpret_cast<const uint4*>(src + size_t(row) * rowStride + seg * 8);
		}
#pragma unroll
		for (int j = 0; j < N; ++j)
		{
			const int u = tid + j * NT;
			const int d = u >> 7, q = u & 127, row = q >> 2, seg = q & 3;
			uint32_t t8 = flags8_fast(lut, w[j]) & 0xFFu;
			if (has_big_id(w[j]))
				t8 = flags8_slow(a, lut, w[j]) & 0xFFu;
			if (e[j] == 0ull)
				t8 = emptyT;
			uint32_t t = t8 << (seg * 8);
			t |= __shfl_xor_sync(0xffffffffu, t, 1);
			t |= __shfl_xor_sync(0xffffffffu, t, 2);
			if (seg == 0)
				hT[d][row] = t;
		}
	}

	// row masks of the staged chunk, one x-row (32 ids = four 16-byte loads) per thread and step: no shuffles, 32 independent
	// table lookups in flight per thread.  Lane i of a quarter-warp starts with 16-byte segment (i >> 1) & 3 of its row and rotates
	// from there, which makes the 64-byte-strided 16-byte loads bank-conflict free.  Block ids beyond the in-shared-memory table
	// (never seen with the reference's 14 block types) are fixed up by a second, slow pass.
	template<int NT, class B>
	__device__ __forceinline__ void build_masks_fused(const MeshArgs& a, const uint32_t* lut, B& sm, int tid)
	{
		const int rot = (tid >> 1) & 3;
		bool big = false;
#pragma unroll 2
		for (int k = 0; k < ROWS / NT; ++k)
		{
			const int row = k * NT + tid;
			uint32_t t = 0, e = 0, d = 0;
#pragma unroll
			for (int j = 0; j < 4; ++j)
			{
				const int seg = (j + rot) & 3;
				const uint4 w = *reinterpret_cast<const uint4*>(&sm.ids[row * CS + seg * 8]);
				big |= has_big_id(w);
				const uint32_t acc = flags8_fast(lut, w);
				t |= (acc & 0xFFu) << (seg * 8);
				e |= ((acc >> 8) & 0xFFu) << (seg * 8);
				d |= ((acc >> 16) & 0xFFu) << (seg * 8);
			}
			sm.mT[row] = t;
			sm.mE[row] = e;
			sm.mD[row] = d;
		}
		if (__any_sync(0xffffffffu, big))
		{
#pragma unroll 1
			for (int k = 0; k < ROWS / NT; ++k)
			{
				const int row = k * NT + tid;
				uint32_t t = 0, e = 0, d = 0;
				for (int seg = 0; seg < 4; ++seg)
				{
					const uint4 w = *reinterpret_cast<const uint4*>(&sm.ids[row * CS + seg * 8]);
					const uint32_t acc = flags8_slow(a, lut, w);
					t |= (acc & 0xFFu) << (seg * 8);
					e |= ((acc >> 8) & 0xFFu) << (seg * 8);
					d |= ((acc >> 16) & 0xFFu) << (seg * 8);
				}
				sm.mT[row] = t;
				sm.mE[row] = e;
				sm.mD[row] = d;
			}
		}
	}

	// id of a block of the facing slab, from global memory (rare path only)
	__device__ __forceinline__ unsigned halo_id(const MeshArgs& a, uint32_t slot, int d, int slow, int fast)
	{
		const unsigned long long e = a.halo[size_t(slot) * 6 + d];
		if (e == 0ull)
			return 0u;
		const uint16_t* src = reinterpret_cast<const uint16_t*>(e & ~HALO_KIND_MASK);
		const size_t rowStride = (e & HALO_KIND_MASK) == HALO_ROWS ? SLAB : CS;
		return src[size_t(slow) * rowStride + fast];
	}

	// visibility masks of row r in emission-slot order; returns the row's double-sided mask.  Thread-local (no warp collectives).
	// The same-type suppression of transparent blocks is applied to the blocks of `fixMask` only (a lane that shares a row with others
	// fixes its own segment); `hadNeed` says whether the row has transparent non-empty blocks at all.
	template<class B>
	__device__ __forceinline__ uint32_t row_vis(const MeshArgs& a, const B& sm, uint32_t slot, int r, uint32_t v[6], uint32_t fixMask, bool& hadNeed)
	{
		const int z = r >> 5, y = r & 31;
		const uint32_t t = sm.mT[r], e = sm.mE[r];
		const uint32_t tUp = z < CS - 1 ? sm.mT[r + CS] : sm.hT[D_UP][y];
		const uint32_t tDown = z > 0 ? sm.mT[r - CS] : sm.hT[D_DOWN][y];
		const uint32_t tFront = y > 0 ? sm.mT[r - 1] : sm.hT[D_FRONT][z];
		const uint32_t tBack = y < CS - 1 ? sm.mT[r + 1] : sm.hT[D_BACK][z];
		const uint32_t tLeft = (t << 1) | ((sm.hT[D_LEFT][z] >> y) & 1u);
		const uint32_t tRight = (t >> 1) | (((sm.hT[D_RIGHT][z] >> y) & 1u) << 31);
		v[0] = e & tUp;
		v[1] = e & tDown;
		v[2] = e & tFront;
		v[3] = e & tBack;
		v[4] = e & tLeft;
		v[5] = e & tRight;

		// transparent non-empty blocks (glass, forcefield): no face towards a neighbour of the same type (Chunk.cpp:192-194)
		uint32_t need = e & t;
		hadNeed = need != 0u;
		need &= fixMask;
		while (need)
		{
			const int x = __ffs(need) - 1;
			need &= need - 1;
			const uint32_t bit = 1u << x;
			const unsigned self = sm.ids[r * CS + x];
			if (v[0] & bit)
				if ((z < CS - 1 ? sm.ids[(r + CS) * CS + x] : halo_id(a, slot, D_UP, y, x)) == self)
					v[0] &= ~bit;
			if (v[1] & bit)
				if ((z > 0 ? sm.ids[(r - CS) * CS + x] : halo_id(a, slot, D_DOWN, y, x)) == self)
					v[1] &= ~bit;
			if (v[2] & bit)
				if ((y > 0 ? sm.ids[(r - 1) * CS + x] : halo_id(a, slot, D_FRONT, z, x)) == self)
					v[2] &= ~bit;
			if (v[3] & bit)
				if ((y < CS - 1 ? sm.ids[(r + 1) * CS + x] : halo_id(a, slot, D_BACK, z, x)) == self)
					v[3] &= ~bit;
			if (v[4] & bit)
				if ((x > 0 ? sm.ids[r * CS + x - 1] : halo_id(a, slot, D_LEFT, z, y)) == self)
					v[4] &= ~bit;
			if (v[5] & bit)
				if ((x < CS - 1 ? sm.ids[r * CS + x + 1] : halo_id(a, slot, D_RIGHT, z, y)) == self)
					v[5] &= ~bit;
		}
		return sm.mD[r];
	}

	template<class B>
	__device__ __forceinline__ uint32_t row_vis(const MeshArgs& a, const B& sm, uint32_t slot, int r, uint32_t v[6])
	{
		bool hadNeed;
		return row_vis(a, sm, slot, r, v, 0xFFFFFFFFu, hadNeed);
	}

	// per-lane write-out constants: a staging round is 16 faces = 192 float4; an index tile is 32 faces = 96 uint2
	struct LaneConst
	{
		uint32_t stageOff[6];
		uint2 idxPat[3];
	};

	// `compact`: the staging slot of a face without tangents holds 9 float4 — 4 x {px py pz nx}, 4 x {ny nz u v} and ONE {w 0 0 0}
	// that all four vertices read (store_face_compact)
	__device__ __forceinline__ LaneConst lane_constants(int lane, bool compact)
	{
		LaneConst lc;
#pragma unroll
		for (int k = 0; k < 6; ++k)
		{
			const uint32_t i = uint32_t(lane) + 32u * k;
			const uint32_t e = i % 12u, v = e / 3u, part = e % 3u;
			lc.stageOff[k] = (i / 12u) * STAGE_FACE_F4 + (compact ? (part == 0u ? v : (part == 1u ? 4u + v : 8u)) : e);
		}
#pragma unroll
		for (int k = 0; k < 3; ++k)
		{
			const uint32_t i = uint32_t(lane) + 32u * k;
			const uint32_t b = (i / 3u) * 4u, j = i % 3u;
			lc.idxPat[k] = j == 0 ? make_uint2(b, b + 2u) : (j == 1 ? make_uint2(b + 1u, b + 1u) : make_uint2(b + 2u, b + 3u));
		}
		return lc;
	}

	// Descriptors of the window of faces [w0, w0 + nw) (nw <= 128) of the chunk held by `sm`, written by one warp; returns the
	// first compact row of the window (the descriptors count rows from it).
	template<class B>
	__device__ __forceinline__ uint32_t expand_window(const MeshArgs& a, const B& sm, uint32_t slot, uint32_t K, uint32_t w0, uint32_t nw, uint16_t* desc, int lane)
	{
		constexpr uint32_t FULL = 0xffffffffu;

		// ---- compact rows [kA, kB] of the window: largest k with ne[k].off <= first / last face (offsets increase strictly)
		uint32_t kA, kB;
		{
			const uint32_t wLast = w0 + nw - 1u;
			const uint32_t i1 = uint32_t(lane) * 32u;
			const uint32_t o1 = i1 < K ? (sm.ne[i1] & NE_OFF_MASK) : 0xFFFFFFFFu;
			const uint32_t topA = __popc(__ballot_sync(FULL, o1 <= w0)) - 1u;
			const uint32_t topB = __popc(__ballot_sync(FULL, o1 <= wLast)) - 1u;
			const uint32_t iA = topA * 32u + lane, iB = topB * 32u + lane;
			const uint32_t oA = iA < K ? (sm.ne[iA] & NE_OFF_MASK) : 0xFFFFFFFFu;
			const uint32_t oB = iB < K ? (sm.ne[iB] & NE_OFF_MASK) : 0xFFFFFFFFu;
			kA = topA * 32u + __popc(__ballot_sync(FULL, oA <= w0)) - 1u;
			kB = topB * 32u + __popc(__ballot_sync(FULL, oB <= wLast)) - 1u;
		}

		// ---- descriptors, faces in the reference's order (x, then Up/Down/Front/Back/Left/Right, twin).  A window of R
		// rows is walked by 32 lanes: S = 32 / R (power of two) lanes share a row, each taking 32 / S consecutive blocks.
		const uint32_t R = kB - kA + 1u;
		const uint32_t lgS = R > 16u ? 0u : (R > 8u ? 1u : (R > 4u ? 2u : (R > 2u ? 3u : (R > 1u ? 4u : 5u))));
		const uint32_t x0 = (uint32_t(lane) & ((1u << lgS) - 1u)) << (5u - lgS);
		const uint32_t segMask = lgS == 0u ? 0xFFFFFFFFu : (((1u << (32u >> lgS)) - 1u) << x0);
		const uint32_t below = (1u << x0) - 1u;
		for (uint32_t k0 = kA; k0 <= kB; k0 += 32u >> lgS)
		{
			const uint32_t k = k0 + (uint32_t(lane) >> lgS);
			uint32_t v[6] = { 0u, 0u, 0u, 0u, 0u, 0u };
			uint32_t d = 0u, p = 0u, base = 0u, any = 0u;
			bool hadNeed = false;
			if (k <= kB)
			{
				const uint32_t en = sm.ne[k];
				d = row_vis(a, sm, slot, int(en >> NE_ROW_SHIFT), v, segMask, hadNeed);
				p = (en & NE_OFF_MASK) - w0; // wraps for the part of the first row that precedes the window
				base = (k - kA) << 9;
				any = (v[0] | v[1] | v[2] | v[3] | v[4] | v[5]) & segMask;
			}
			if (!__any_sync(FULL, hadNeed))
			{
				// nothing to suppress in these rows: the masks are final everywhere, the faces in front of this lane's segment are counted directly
#pragma unroll
				for (int s = 0; s < 6; ++s)
					p += __popc(v[s] & below) + __popc(v[s] & d & below);
			}
			else
			{
				// rows with transparent non-empty blocks (glass next to glass shows no face): every lane has run the suppression for its own
				// segment only, so the faces in front of it are the sum over the lanes to its left in the row's group
				uint32_t cnt = 0u;
#pragma unroll
				for (int s = 0; s < 6; ++s)
					cnt += __popc(v[s] & segMask) + __popc(v[s] & d & segMask);
				uint32_t incl = cnt;
				const uint32_t sub = uint32_t(lane) & ((1u << lgS) - 1u);
				for (uint32_t o = 1u; o < (1u << lgS); o <<= 1)
				{
					const uint32_t n = __shfl_up_sync(FULL, incl, o);
					if (sub >= o)
						incl += n;
				}
				p += incl - cnt;
			}
			// no double-sided block in any of these rows (a planet has none): half the loop body
			if (!__any_sync(FULL, d != 0u))
			{
				while (any && int(p) < int(nw))
				{
					const uint32_t x = uint32_t(__ffs(any) - 1);
					any &= any - 1u;
					const uint32_t bit = 1u << x;
#pragma unroll
					for (uint32_t s = 0; s < 6; ++s)
					{
						if (v[s] & bit)
						{
							if (p < nw)
								desc[p] = uint16_t(base | x | (s << 5));
							++p;
						}
					}
				}
			}
			else
			{
				while (any && int(p) < int(nw))
				{
					const uint32_t x = uint32_t(__ffs(any) - 1);
					any &= any - 1u;
					const uint32_t bit = 1u << x;
					const bool tw = (d & bit) != 0u;
#pragma unroll
					for (uint32_t s = 0; s < 6; ++s)
					{
						if (v[s] & bit)
						{
							if (p < nw)
								desc[p] = uint16_t(base | x | (s << 5));
							++p;
							if (tw)
							{
								if (p < nw)
									desc[p] = uint16_t(base | x | (s << 5) | 0x100u);
								++p;
							}
						}
					}
				}
			}
		}
		__syncwarp();

		return kA;
	}

	// Build the faces of the window [w0, w0 + nw) from their descriptors, one per lane, and stream them out through the warp's 16-face
	// staging tile with coalesced 16-byte stores (vertices, collider positions, indices).
	// MODE_LEAN: the table-driven flat path without tangents, compiled without the generic corner / attribute / tangent code (the
	// instantiation a planet remesh runs); MODE_GENERIC adds the generic corner providers and per-face attribute math (deformed
	// chunks, other block sizes); MODE_TANGENTS adds what GenerateTangents computes on top of that
	template<int MODE, class B>
	__device__ __forceinline__ void store_window(const MeshArgs& a, const B& sm, const FaceTable& table, bool fastFlat, const FaceCtx& cx, uint32_t w0, uint32_t nw, uint32_t kA,
	                                             const uint16_t* desc, float4* stage, unsigned long long firstFace, const LaneConst& lc, bool wantRender, bool wantCollider, bool wantTangents,
	                                             int lane)
	{
		// ---- build one face per lane, stream out through the 16-face staging tile
		const uint32_t nTiles = (nw + 31u) >> 5;
		for (uint32_t tt = 0; tt < nTiles; ++tt)
		{
			const uint32_t f = tt * 32u + lane;
			const uint32_t nf = min(32u, nw - tt * 32u);
			const bool valid = f < nw;
			FaceOut fo;
			uint32_t row = 0u;
			int x = 0, s = 0;
			bool twin = false;
			if (valid)
			{
				const uint32_t de = desc[f];
				row = sm.ne[kA + (de >> 9)] >> NE_ROW_SHIFT;
				x = int(de & 31u);
				s = int((de >> 5) & 7u);
				twin = ((de >> 8) & 1u) != 0u;
			}
			// a chunk the deformation touches still has plenty of blocks it leaves alone: their faces need no corner evaluation
			bool flatFace = false;
			constexpr bool LEAN = MODE == MODE_LEAN;
			if (!LEAN && !fastFlat && cx.exactDeform && valid)
				flatFace = region_is_undeformed(cx, float(x) * cx.tile, float(row >> 5) * cx.tile, float(row & 31u) * cx.tile, cx.tile);
			if (LEAN || fastFlat)
			{
				if (valid)
					draw_face_flat_r(a, cx, table, LEAN ? float(CS) * 0.5f : cx.flatOrigin, x, int(row & 31u), int(row >> 5), s, twin, sm.ids[row * CS + x], fo);
			}
			else
			{
				// Generic corner providers (DeformedChunk: a normalisation per corner): the 8 corners and the centre of a block are the
				// same for all of its faces (up to 12), and the faces of a block are neighbours in the emission order.  The warp finds
				// the distinct blocks of the tile (those not already known to be undeformed), spreads their 8 * nBlocks corner evaluations over its 32 lanes and shares the
				// results through the staging tile, which is idle at this point: [0, 32) block keys, then 9 F3 per block (8 corners by
				// code dX | dY << 1 | dZ << 2, centre), 16 blocks per round, then 16 "block left undeformed" flags.
				constexpr uint32_t FULLM = 0xffffffffu;
				uint32_t* keys = reinterpret_cast<uint32_t*>(stage);
				float* cache = reinterpret_cast<float*>(stage) + 32;
				uint32_t* blockFlat = reinterpret_cast<uint32_t*>(stage) + 32 + 16 * 27; // 16 flags: all 8 corners of the block undeformed
				const bool blockTable = cx.deformed && a.faceTable != nullptr;
				const bool needGeo = valid && !flatFace;
				const uint32_t key = needGeo ? ((row << 5) | uint32_t(x)) : 0xFFFFFFFFu;
				const uint32_t prev = __shfl_up_sync(FULLM, key, 1);
				const bool leader = needGeo && (lane == 0 || key != prev);
				const uint32_t leaders = __ballot_sync(FULLM, leader);
				const uint32_t nBlocks = __popc(leaders);
				const uint32_t myBlock = __popc(leaders & (0xFFFFFFFFu >> (31 - lane))) - 1u; // ordinal of this lane's block in the tile
				if (leader)
					keys[myBlock] = key;
				__syncwarp();
				F3 pos[4], blockCenter = f3(0.f, 0.f, 0.f);
#pragma unroll
				for (int i = 0; i < 4; ++i)
					pos[i] = f3(0.f, 0.f, 0.f);
				bool flatBlock = false;
				for (uint32_t b0 = 0; b0 < nBlocks; b0 += 16u)
				{
					const uint32_t nb = min(16u, nBlocks - b0);
					for (uint32_t e0 = 0; e0 < nb * 8u; e0 += 32u)
					{
						const uint32_t e = e0 + uint32_t(lane);
						bool same = false;
						if (e < nb * 8u)
						{
							const uint32_t k = keys[b0 + (e >> 3)];
							const int bx = int(k & 31u), by = int((k >> 5) & 31u), bz = int(k >> 10);
							const F3 c = blockTable ? deformed_corner(cx, bx, by, bz, e & 7u, same) : voxel_corner(cx, bx, by, bz, e & 7u);
							float* dst = cache + ((e >> 3) * 9u + (e & 7u)) * 3u;
							dst[0] = c.x;
							dst[1] = c.y;
							dst[2] = c.z;
						}
						// a block is flat when all 8 of its corners (8 neighbouring lanes) stayed in place
						const uint32_t sameMask = __ballot_sync(FULLM, same);
						if (e < nb * 8u && (e & 7u) == 0u)
							blockFlat[e >> 3] = ((sameMask >> (lane & 24)) & 0xFFu) == 0xFFu ? 1u : 0u;
					}
					__syncwarp();
					if (uint32_t(lane) < nb)
					{
						// block centre: the 8 corners summed in the enum order of Nz::BoxCorner (codes 4, 5, 0, 1, 6, 7, 2, 3), / 8 (Chunk.cpp:186-188)
						const float* cb = cache + uint32_t(lane) * 27u;
						F3 sum = f3(0.f, 0.f, 0.f);
						constexpr int order[8] = { 4, 5, 0, 1, 6, 7, 2, 3 };
#pragma unroll
						for (int q = 0; q < 8; ++q)
							sum = sum + f3(cb[order[q] * 3 + 0], cb[order[q] * 3 + 1], cb[order[q] * 3 + 2]);
						sum = sum / 8.f;
						float* dst = cache + (uint32_t(lane) * 9u + 8u) * 3u;
						dst[0] = sum.x;
						dst[1] = sum.y;
						dst[2] = sum.z;
					}
					__syncwarp();
					if (needGeo && myBlock >= b0 && myBlock < b0 + nb)
					{
						const float* cb = cache + (myBlock - b0) * 27u;
						blockCenter = f3(cb[24], cb[25], cb[26]);
						flatBlock = blockFlat[myBlock - b0] != 0u;
#pragma unroll
						for (int i = 0; i < 4; ++i)
						{
							const unsigned code = c_faceCorners[s][twin ? (i ^ 1) : i];
							pos[i] = f3(cb[code * 3 + 0], cb[code * 3 + 1], cb[code * 3 + 2]);
						}
					}
					__syncwarp();
				}
				if (flatFace)
					draw_face_flat_r(a, cx, table, cx.flatOrigin, x, int(row & 31u), int(row >> 5), s, twin, sm.ids[row * CS + x], fo);
				else if (valid)
				{
					if (flatBlock)
						draw_face_cached_table(a, cx, table, pos, s, twin, sm.ids[row * CS + x], fo);
					else
						draw_face_cached(a, cx, blockCenter, pos, sm.ids[row * CS + x], fo);
				}
			}
			const unsigned long long face0 = firstFace + w0 + tt * 32u;
			if (wantRender)
			{
#pragma unroll
				for (uint32_t h = 0; h < 2; ++h)
				{
					if (nf <= h * HALF_FACES)
						break;
					const uint32_t nh = min(uint32_t(HALF_FACES), nf - h * HALF_FACES);
					if (valid && (uint32_t(lane) >> 4) == h)
					{
						if (MODE == MODE_TANGENTS)
							store_face(stage + (lane & 15) * STAGE_FACE_F4, fo, wantTangents);
						else
							store_face_compact(stage + (lane & 15) * STAGE_FACE_F4, fo);
					}
					__syncwarp();
					float4* dst = a.vertices + (face0 + h * HALF_FACES) * 12ull + lane;
					if (nh == HALF_FACES)
					{
#pragma unroll
						for (int k = 0; k < 6; ++k)
							ST_OUT(dst + 32 * k, stage[lc.stageOff[k]]);
					}
					else
					{
						const uint32_t n4 = nh * 12u;
#pragma unroll
						for (int k = 0; k < 6; ++k)
							if (uint32_t(lane) + 32u * k < n4)
								ST_OUT(dst + 32 * k, stage[lc.stageOff[k]]);
					}
					__syncwarp();
				}
			}
			if (wantCollider && valid)
			{
				// positions only: the 12 floats of this lane's face straight from its registers, 48 bytes at face * 48.  Three 16-byte stores
				// per lane at a 48-byte stride: the warp's three instructions cover 1,536 contiguous bytes between them (L2 merges the
				// sectors); gathering them from the staging tile into fully coalesced stores cost 12 % of the launch's instructions
				// (31^3 render + collider 3.03 -> 2.81 ms, collider only 2.35 -> 1.93 ms)
				float4* dst = a.collider + (face0 + lane) * 3ull;
				ST_OUT(dst + 0, make_float4(fo.pos[0].x, fo.pos[0].y, fo.pos[0].z, fo.pos[1].x));
				ST_OUT(dst + 1, make_float4(fo.pos[1].y, fo.pos[1].z, fo.pos[2].x, fo.pos[2].y));
				ST_OUT(dst + 2, make_float4(fo.pos[2].z, fo.pos[3].x, fo.pos[3].y, fo.pos[3].z));
			}
			{
				// indices are a pure function of the chunk-relative face ordinal (Chunk.cpp:95-101): k*4 + {0,2,1,1,2,3}
				uint2* dst = a.indices + face0 * 3ull + lane;
				const uint32_t base = (w0 + tt * 32u) * 4u;
				const uint32_t n2 = nf * 3u;
#pragma unroll
				for (int k = 0; k < 3; ++k)
					if (nf == 32u || uint32_t(lane) + 32u * k < n2)
						ST_OUT(dst + 32 * k, make_uint2(lc.idxPat[k].x + base, lc.idxPat[k].y + base));
			}
		}
	}

	// Decoupled look-back over the jobs before `job` (one warp): publish this job's face count, add up the counts of the
	// predecessors back to the nearest one whose inclusive prefix is known, publish the own inclusive prefix, return the
	// exclusive one.  Each step inspects 32 * LBK predecessors with independent loads (one L2 round trip); position
	// p = lane * LBK + k counts backwards from job - 1.  Measured on the 31^3 planet: LBK = 1 is best (2 the same, 8 is 25 % slower
	// in the count-only pass) — the cost is waiting for the nearest predecessors to publish, not the distance walked.
	template<int LBK>
	__device__ __forceinline__ unsigned long long chain_lookback(unsigned long long* status, uint32_t job, uint32_t F, int lane)
	{
		constexpr uint32_t FULL = 0xffffffffu;
		constexpr unsigned long long FLAG_AGG = 1ull << 62, FLAG_INC = 2ull << 62, VAL = (1ull << 62) - 1ull;
		volatile unsigned long long* st = status;
		unsigned long long excl = 0ull;
		if (job == 0u)
		{
			if (lane == 0)
				st[0] = FLAG_INC | F;
			return 0ull;
		}
		if (lane == 0)
			st[job] = FLAG_AGG | F;
		long long base = (long long)job - 1;
		for (;;)
		{
			unsigned long long s[LBK];
			uint32_t pInc; // position of the nearest inclusive prefix in this window, 32 * LBK if none
			bool retry;
			do
			{
#pragma unroll
				for (int k = 0; k < LBK; ++k)
				{
					const long long i = base - (long long)lane * LBK - k;
					s[k] = i >= 0 ? st[i] : FLAG_INC; // before job 0: inclusive prefix 0
				}
				pInc = 32u * LBK;
				uint32_t pWait = 32u * LBK; // position of the nearest entry that is not published yet
#pragma unroll
				for (int k = 0; k < LBK; ++k)
				{
					const unsigned fl = unsigned(s[k] >> 62);
					const uint32_t inc = __ballot_sync(FULL, fl == 2u);
					const uint32_t wait = __ballot_sync(FULL, fl == 0u);
					if (inc)
						pInc = min(pInc, uint32_t(__ffs(inc) - 1) * LBK + k);
					if (wait)
						pWait = min(pWait, uint32_t(__ffs(wait) - 1) * LBK + k);
				}
				retry = pWait < pInc || (pInc == 32u * LBK && pWait < 32u * LBK);
			} while (retry);
			unsigned long long val = 0ull;
#pragma unroll
			for (int k = 0; k < LBK; ++k)
				if (uint32_t(lane) * LBK + k <= pInc)
					val += s[k] & VAL;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1)
				val += __shfl_xor_sync(FULL, val, o);
			excl += val;
			if (pInc < 32u * LBK)
				break;
			base -= 32 * LBK;
		}
		if (lane == 0)
			st[job] = FLAG_INC | (excl + F);
		return excl;
	}

	// one lane: take the next job and look up what it names (two dependent global loads paid once, before the CTA barrier)
	// The next job of this CTA: ticket j from the global counter, then its entry jobSlot[j] = slot | has-content << 31 (one dependent
	// load behind the atomic)
	__device__ __forceinline__ void take_ticket(const MeshArgs& a, FusedSmem& sm)
	{
		uint32_t j = atomicAdd(a.jobCounter, 1u);
		if (a.order != nullptr && j < a.nJobs)
		{
			// ticket -> job through the bucket lists (heavy chunks first)
			uint32_t t = j, q = 0u;
#pragma unroll
			for (int b = 0; b < JOB_BUCKETS - 1; ++b)
			{
				const uint32_t cnt = a.bucketCount[b];
				if (q == uint32_t(b) && t >= cnt)
				{
					t -= cnt;
					q = uint32_t(b) + 1u;
				}
			}
			j = a.order[size_t(q) * a.nJobs + t];
		}
		const uint32_t entry = j < a.nJobs ? a.jobSlot[j] : 0u;
		sm.slot = entry & JOB_SLOT_MASK;
		// a chunk whose summary proves that it has no face (job_classify_kernel) is treated like one without content: 0 faces, no stage;
		// 2 = all opaque: the row masks are known up front
		sm.hasContent = (entry & (JOB_HAS_CONTENT | JOB_SKIP)) == JOB_HAS_CONTENT ? ((entry & JOB_OPAQUE) ? 2u : 1u) : 0u;
		sm.job = j;
	}

	template<int MODE>
	__global__ void __launch_bounds__(FUSED_THREADS, 2) mesh_fused_kernel(const MeshArgs a)
	{
		extern __shared__ __align__(128) unsigned char smemRaw[];
		FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smemRaw);
		const int tid = threadIdx.x;
		const int lane = tid & 31;
		const int warp = tid >> 5;
		constexpr uint32_t FULL = 0xffffffffu;

		if (tid == 0)
		{
			mbar_init(&sm.bar, 1);
			fence_mbar_init();
		}
		for (int i = tid; i < LUTN; i += FUSED_THREADS)
			sm.lut32[i] = pack_flags(a.blkFlags[uint32_t(i) < a.nTypes ? i : a.nTypes - 1], unsigned(i));
		// the attribute table serves every face of a FlatChunk and, on a DeformedChunk, the faces of blocks the deformation left in place
		const bool fastFlat = a.faceTable != nullptr && !(a.flags & MESH_F_DEFORMED);
		if (a.faceTable != nullptr)
		{
			const uint32_t* src = reinterpret_cast<const uint32_t*>(a.faceTable);
			uint32_t* dst = reinterpret_cast<uint32_t*>(&sm.table);
			for (int i = tid; i < int(sizeof(FaceTable) / 4); i += FUSED_THREADS)
				dst[i] = src[i];
		}
		if (a.texFits8)
			for (int i = tid; i < TEX8N * 6; i += FUSED_THREADS)
				sm.tex8[i] = uint8_t(a.blkTex[size_t(uint32_t(i / 6) < a.nTypes ? i / 6 : a.nTypes - 1) * 6 + i % 6]);
		const LaneConst lc = lane_constants(lane, MODE != MODE_TANGENTS);
		const bool wantRender = (a.flags & MESH_F_RENDER) != 0;
		const bool wantCollider = (a.flags & MESH_F_COLLIDER) != 0;
		const bool wantTangents = (a.flags & MESH_F_TANGENTS) != 0;
		const bool ordered = (a.flags & MESH_F_ORDERED) != 0; // arena slices in job order (look-back) instead of first come, first served
		__syncthreads();

		if (tid == 0)
		{
			sm.ffSeq = 0u;
			sm.warpsDone = 0u;
			take_ticket(a, sm);
		}
		uint32_t parity = 0;
		for (;;)
		{
			__syncthreads(); // the ticket is visible; every read of the previous chunk's shared state is done
			const uint32_t job = sm.job;
			if (job >= a.nJobs)
				break;
			const uint32_t slot = sm.slot;
			const bool hasContent = sm.hasContent != 0u;
			// every block non-Empty, opaque, single sided (the producer's summary): mT = 0, mE = all ones, mD = 0 without looking at the ids.
			// The ids are still staged (the emission reads the id of every face for its texture slice), but nothing waits for them before
			// the faces are counted: the copy runs behind the halo classification, the count and the scan.
			const bool opaque = sm.hasContent == 2u;

			uint32_t F = 0, K = 0;
			if (hasContent)
			{
				// ---- stage the ids (TMA) and classify the halo slabs meanwhile
				if (tid == 0)
				{
					fence_proxy_async(); // the generic reads of the previous chunk (ordered by the barrier above) precede the async writes
					mbar_arrive_expect_tx(&sm.bar, NB * 2);
					const uint16_t* src = a.blocks + size_t(slot) * NB;
#pragma unroll
					for (int p = 0; p < 4; ++p)
						tma_load_1d(sm.ids + p * (NB / 4), src + p * (NB / 4), NB / 2, &sm.bar);
				}
				build_halo_masks<FUSED_THREADS>(a, sm.lut32, slot, sm.hT, tid);
				if (!opaque)
				{
					mbar_wait(&sm.bar, parity);
					parity ^= 1;
					build_masks_fused<FUSED_THREADS>(a, sm.lut32, sm, tid);
				}
				else
				{
#pragma unroll
					for (int k = 0; k < ROWS / FUSED_THREADS; ++k)
					{
						sm.mT[k * FUSED_THREADS + tid] = 0u;
						sm.mE[k * FUSED_THREADS + tid] = 0xFFFFFFFFu;
						sm.mD[k * FUSED_THREADS + tid] = 0u;
					}
				}
				__syncthreads();

				// ---- faces per row (thread t owns rows 4t .. 4t+3) and one packed scan: faces | rows-with-faces << 20
				uint32_t c[4];
#pragma unroll
				for (int j = 0; j < 4; ++j)
				{
					uint32_t v[6];
					const uint32_t d = row_vis(a, sm, slot, 4 * tid + j, v);
					c[j] = row_faces(v, d);
				}
				const uint32_t mine = (c[0] + c[1] + c[2] + c[3]) | (uint32_t((c[0] != 0) + (c[1] != 0) + (c[2] != 0) + (c[3] != 0)) << 20);
				uint32_t incl = mine;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1)
				{
					const uint32_t n = __shfl_up_sync(FULL, incl, o);
					if (lane >= o)
						incl += n;
				}
				if (lane == 31)
					sm.warpSum[warp] = incl;
				__syncthreads();
				uint32_t wbase = 0, total = 0;
#pragma unroll
				for (int w = 0; w < FUSED_WARPS; ++w)
				{
					const uint32_t ws = sm.warpSum[w];
					if (w < warp)
						wbase += ws;
					total += ws;
				}
				F = total & 0xFFFFFu;
				K = total >> 20;
				const uint32_t excl = wbase + incl - mine;
				uint32_t off = excl & 0xFFFFFu, k = excl >> 20;
#pragma unroll
				for (int j = 0; j < 4; ++j)
				{
					if (c[j])
						sm.ne[k++] = off | (uint32_t(4 * tid + j) << NE_ROW_SHIFT);
					off += c[j];
				}
				if (tid == 0)
					sm.ne[K] = F;
			}

			__syncthreads(); // ne[] is complete; sm.job has been read by everyone
			if (hasContent && opaque)
			{
				mbar_wait(&sm.bar, parity); // the ids of an all-opaque chunk are first needed here
				parity ^= 1;
			}

			// Unordered mode: no chunk waits on another, so the next ticket (and the dependent load behind it) can be taken
			// now and its latency hides behind this chunk's emission.  (In ordered mode a ticket held during an emission would
			// delay that job's aggregate and stall the look-back of every later job: there the last finishing warp takes it.)
			if (!ordered && tid == 32)
				take_ticket(a, sm);

			// ---- place the chunk in the arena.  Ordered mode: decoupled look-back over the jobs before this one (warp 0; the other warps
			// already build their first descriptor window and pick firstFace up when they are about to store)
			unsigned long long firstFace = 0ull;
			bool haveFirst = false;
			if (warp == 0)
			{
				unsigned long long excl;
				if (ordered)
					excl = chain_lookback<1>(a.status, job, F, lane);
				else
				{
					// first come, first served: one atomic on the arena cursor, no dependency on any other chunk
					excl = 0ull;
					if (lane == 0 && F > 0)
						excl = atomicAdd(a.totalFacesOut, static_cast<unsigned long long>(F));
					excl = __shfl_sync(FULL, excl, 0);
				}
				if (lane == 0)
				{
					sm.firstFace = excl;
					__threadfence_block();
					*reinterpret_cast<volatile uint32_t*>(&sm.ffSeq) = job + 1u;
					a.faceCount[job] = F;
					a.firstFaceOut[job] = excl;
					a.lastFaces[slot] = F;
					if (ordered && job == a.nJobs - 1u)
						*a.totalFacesOut = excl + F;
					if (F > 0 && excl + F > a.arenaFaces)
						atomicExch(a.overflow, 1u);
				}
				firstFace = excl;
				haveFirst = true;
			}

			if (F > 0)
			{
				FaceCtx cx;
				cx.tile = a.tile;
				cx.radius = a.deformRadius;
				cx.deformed = (a.flags & MESH_F_DEFORMED) != 0;
				cx.wantAttr = wantRender;
				cx.rawUp = false;
				cx.tex8 = a.texFits8 ? sm.tex8 : nullptr;
				cx.flatOrigin = float(CS) * 0.5f;
				if (a.gravity)
					cx.gravity = f3(a.gravity[size_t(job) * 3 + 0], a.gravity[size_t(job) * 3 + 1], a.gravity[size_t(job) * 3 + 2]);
				else
				{
					// container centre - GetChunkOffset(indices) (ClientChunkEntities.cpp:160, ChunkContainer.inl:53-56)
					const int4 ci = a.chunkIdx[slot];
					cx.gravity = f3(a.center[0] - float(ci.x) * float(CS) * a.tile, a.center[1] - float(ci.y) * float(CS) * a.tile, a.center[2] - float(ci.z) * float(CS) * a.tile);
					cx.rawUp = fastFlat && a.center[0] == 0.f && a.center[1] == 0.f && a.center[2] == 0.f && max(max(abs(ci.x), abs(ci.y)), abs(ci.z)) < (1 << 14);
				}

				// A DeformedChunk the deformation leaves alone is a FlatChunk whose corners start at 0 instead of -size / 2
				bool tableJob = fastFlat;
				cx.exactDeform = false;
				if (MODE != MODE_LEAN && cx.deformed && a.faceTable != nullptr)
				{
					cx.flatOrigin = 0.f;
					cx.exactDeform = deform_is_exact(cx);
					tableJob = cx.exactDeform && region_is_undeformed(cx, 0.f, 0.f, 0.f, float(CS) * cx.tile);
					// faceUp of an undeformed face from the raw centre-to-face vector (see FaceCtx::rawUp): exact and well separated when
					// the centre sits on the half-block grid within 2^20 half-blocks, like the default centre (a multiple of 32 blocks)
					const float toHalf = 2.f / cx.tile;
					auto halfGrid = [toHalf](float v) { const float t = v * toHalf; return fabsf(t) < 1048576.f && t == truncf(t); };
					cx.rawUp = cx.exactDeform && halfGrid(cx.gravity.x) && halfGrid(cx.gravity.y) && halfGrid(cx.gravity.z);
				}

				float4* stage = sm.stage[warp];
				uint16_t* desc = sm.desc[warp];

				// every warp takes an equal, contiguous share of the chunk's faces (boundaries on multiples of 16 faces = whole staging
				// tiles) and walks it in windows of up to 128: the warps reach the end of the chunk together.  (Fixed 128-face windows
				// dealt round-robin left the CTA waiting for the warps that got one window more: 12.6 % of all stall samples on the
				// planet, where a chunk has ~14 windows for 8 warps.)
				const uint32_t shareLo = min(F, (((F * uint32_t(warp)) / FUSED_WARPS + 15u) >> 4) << 4);
				const uint32_t shareHi = warp == FUSED_WARPS - 1 ? F : min(F, (((F * uint32_t(warp + 1)) / FUSED_WARPS + 15u) >> 4) << 4);
				for (uint32_t w0 = shareLo; w0 < shareHi; w0 += WIN)
				{
					const uint32_t nw = min(uint32_t(WIN), shareHi - w0);
					const uint32_t kA = expand_window(a, sm, slot, K, w0, nw, desc, lane);
					__syncwarp();

					if (!haveFirst)
					{
						while (*reinterpret_cast<volatile uint32_t*>(&sm.ffSeq) != job + 1u) {}
						__threadfence_block();
						firstFace = *reinterpret_cast<volatile unsigned long long*>(&sm.firstFace);
						haveFirst = true;
					}
					if (firstFace + F > a.arenaFaces)
						break; // arena too small: this launch only counts (the host grows the arena and runs the kernel again)

					store_window<MODE>(a, sm, sm.table, tableJob, cx, w0, nw, kA, desc, stage, firstFace, lc, wantRender, wantCollider, wantTangents, lane);
					__syncwarp(); // the descriptors are consumed before the next window overwrites them
				}
			}
			// next ticket: taken by the LAST warp to finish this chunk, so no extra CTA barrier is needed — and not earlier: a
			// ticket held during an emission delays this job's aggregate and with it the look-back of every later job (measured).
			__syncwarp();
			if (ordered && lane == 0)
			{
				__threadfence_block();
				if (atomicAdd(&sm.warpsDone, 1u) == FUSED_WARPS - 1)
				{
					sm.warpsDone = 0u;
					take_ticket(a, sm);
				}
			}
		}
	}

	// =====================================================================================================================
	// mesh_aabb_kernel — Nz::StaticMesh::GenerateAABB (src/ClientLib/ClientChunkEntities.cpp:168): the bounding box of a chunk's vertex
	// positions, taken from the floats tsom_mesh wrote (so it is the box of exactly the vertices the caller uploads), one CTA per chunk.
	// Source: the render arena (position = first 12 bytes of each 48-byte vertex) or, without it, the collider arena (packed xyz).
	__global__ void __launch_bounds__(256) mesh_aabb_kernel(const float4* __restrict__ vertices, const float* __restrict__ collider, const unsigned long long* __restrict__ firstFace,
	                                                        const uint32_t* __restrict__ faceCount, uint32_t n, float* __restrict__ out)
	{
		const uint32_t job = blockIdx.x;
		if (job >= n)
			return;
		const unsigned long long first = firstFace[job];
		const uint32_t nVerts = faceCount[job] * 4u;
		float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
		for (uint32_t v = threadIdx.x; v < nVerts; v += blockDim.x)
		{
			float p[3];
			if (vertices)
			{
				const float4 q = vertices[(first * 4ull + v) * 3ull];
				p[0] = q.x, p[1] = q.y, p[2] = q.z;
			}
			else
			{
				const float* q = collider + (first * 4ull + v) * 3ull;
				p[0] = q[0], p[1] = q[1], p[2] = q[2];
			}
#pragma unroll
			for (int k = 0; k < 3; ++k)
			{
				lo[k] = fminf(lo[k], p[k]);
				hi[k] = fmaxf(hi[k], p[k]);
			}
		}
		__shared__ float red[8][6];
#pragma unroll
		for (int k = 0; k < 3; ++k)
		{
#pragma unroll
			for (int o = 16; o > 0; o >>= 1)
			{
				lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
				hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
			}
		}
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		if (lane == 0)
		{
#pragma unroll
			for (int k = 0; k < 3; ++k)
			{
				red[warp][k] = lo[k];
				red[warp][3 + k] = hi[k];
			}
		}
		__syncthreads();
		if (threadIdx.x < 6)
		{
			float r = red[0][threadIdx.x];
			for (int w = 1; w < 8; ++w)
				r = threadIdx.x < 3 ? fminf(r, red[w][threadIdx.x]) : fmaxf(r, red[w][threadIdx.x]);
			out[size_t(job) * 6 + threadIdx.x] = nVerts ? r : 0.f;
		}
	}

	// =====================================================================================================================
	// mesh_checksum_kernel — one CTA per view: two 64-bit position-weighted sums, over the 32-bit words of the chunk's vertices (or
	// collider positions) and over its indices.  Independent of where the chunk's faces sit in the arena, so meshes built by
	// different launches / different GPUs compare chunk by chunk on the device (sharded parity without moving the arenas to the host).
	__global__ void __launch_bounds__(256) mesh_checksum_kernel(const uint32_t* __restrict__ vertexWords, uint32_t wordsPerFace, const uint32_t* __restrict__ indices,
	                                                            const unsigned long long* __restrict__ firstFace, const uint32_t* __restrict__ faceCount, uint32_t n,
	                                                            unsigned long long* __restrict__ out)
	{
		const uint32_t job = blockIdx.x;
		if (job >= n)
			return;
		const unsigned long long first = firstFace[job];
		const unsigned long long nv = static_cast<unsigned long long>(faceCount[job]) * wordsPerFace, ni = static_cast<unsigned long long>(faceCount[job]) * 6ull;
		auto term = [](uint32_t w, unsigned long long i) { return (static_cast<unsigned long long>(w) + 0x9E3779B97F4A7C15ull) * (2ull * i + 1ull); };
		unsigned long long hv = 0ull, hi = 0ull;
		const uint32_t* v = vertexWords + first * wordsPerFace;
		for (unsigned long long i = threadIdx.x; i < nv; i += blockDim.x)
			hv += term(v[i], i);
		const uint32_t* ix = indices + first * 6ull;
		for (unsigned long long i = threadIdx.x; i < ni; i += blockDim.x)
			hi += term(ix[i], i);
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			hv += __shfl_xor_sync(0xffffffffu, hv, o);
			hi += __shfl_xor_sync(0xffffffffu, hi, o);
		}
		__shared__ unsigned long long red[8][2];
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		if (lane == 0)
		{
			red[warp][0] = hv;
			red[warp][1] = hi;
		}
		__syncthreads();
		if (threadIdx.x < 2)
		{
			unsigned long long r = 0ull;
			for (int w = 0; w < 8; ++w)
				r += red[w][threadIdx.x];
			out[size_t(job) * 2 + threadIdx.x] = r;
		}
	}

	void launch_mesh_checksum(const uint32_t* vertexWords, uint32_t wordsPerFace, const uint32_t* indices, const unsigned long long* firstFace, const uint32_t* faceCount, uint32_t n,
	                          unsigned long long* out, cudaStream_t st)
	{
		if (n)
			mesh_checksum_kernel<<<n, 256, 0, st>>>(vertexWords, wordsPerFace, indices, firstFace, faceCount, n, out);
	}

	void launch_mesh_aabb(const float4* vertices, const float* collider, const unsigned long long* firstFace, const uint32_t* faceCount, uint32_t n, float* out, cudaStream_t st)
	{
		mesh_aabb_kernel<<<n, 256, 0, st>>>(vertices, collider, firstFace, faceCount, n, out);
	}

	// =====================================================================================================================
	// voxel_corners_kernel — Chunk::ComputeVoxelCorners (Chunk.hpp:54) as a query: the pure helper the game uses for block picking
	// (src/Game/States/GameState.cpp:859, src/ServerLib/Session/PlayerSessionHandler.cpp:367,489), answered by the same device
	// functions the mesher uses.  One thread per (block, corner); out = n x 8 x xyz in Nz::BoxCorner enum order.
	__global__ void voxel_corners_kernel(const int4* __restrict__ blocks /* local x, y, z, unused */, const float* __restrict__ centers /* n x 3 */, uint32_t n, float tile, float radius,
	                                     int deformed, float* __restrict__ out)
	{
		const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
		if (t >= n * 8u)
			return;
		const uint32_t q = t >> 3, k = t & 7u;
		constexpr unsigned enumOrder[8] = { 4u, 5u, 0u, 1u, 6u, 7u, 2u, 3u }; // Nz::BoxCorner -> corner code dX | dY << 1 | dZ << 2
		FaceCtx cx;
		cx.tile = tile;
		cx.radius = radius;
		cx.deformed = deformed != 0;
		cx.wantAttr = false;
		cx.gravity = f3(centers[q * 3 + 0], centers[q * 3 + 1], centers[q * 3 + 2]);
		const int4 b = blocks[q];
		const F3 p = voxel_corner(cx, b.x, b.y, b.z, enumOrder[k]);
		out[t * 3 + 0] = p.x;
		out[t * 3 + 1] = p.y;
		out[t * 3 + 2] = p.z;
	}

	void launch_voxel_corners(const int4* blocks, const float* centers, uint32_t n, float tile, float radius, int deformed, float* out, cudaStream_t st)
	{
		voxel_corners_kernel<<<(n * 8u + 255u) / 256u, 256, 0, st>>>(blocks, centers, n, tile, radius, deformed, out);
	}

	// =====================================================================================================================
	// job_classify_kernel — one thread per job: a chunk that is all Empty, or all opaque with six all-opaque neighbours, has no face
	// (Chunk.cpp:190-201: a face needs a non-Empty block next to a transparent one or to a missing chunk).  The summaries come for free
	// from the producers (generate_kernel); two thirds of a large planet are such chunks and skip their 64 KiB stage this way.
	// It also deals the tickets of the unordered mode: jobs go into JOB_BUCKETS lists by the face count their chunk had the last time it
	// was meshed (heavy first, jobs without faces last), so that the persistent CTAs of mesh_fused_kernel end on small chunks instead of
	// on whatever the caller's list ends with (a planet's list ends with its +z surface).  One atomic per warp and bucket.
	__global__ void __launch_bounds__(256) job_classify_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t n, const uint8_t* __restrict__ summary,
	                                                           const int32_t* __restrict__ nbrSlot, const uint32_t* __restrict__ lastFaces, uint32_t* __restrict__ order,
	                                                           uint32_t* __restrict__ bucketCount)
	{
		const uint32_t j = blockIdx.x * 256u + threadIdx.x;
		const bool live = j < n;
		uint32_t entry = live ? src[j] & ~(JOB_SKIP | JOB_OPAQUE) : 0u;
		if (live && (entry & JOB_HAS_CONTENT))
		{
			const uint32_t slot = entry & JOB_SLOT_MASK;
			const uint8_t s = summary[slot];
			bool skip = s == SUM_EMPTY;
			if (s == SUM_OPAQUE)
			{
				skip = true;
#pragma unroll
				for (int d = 0; d < 6; ++d)
				{
					const int32_t ns = nbrSlot[size_t(slot) * 6 + d];
					skip = skip && ns >= 0 && summary[ns] == SUM_OPAQUE;
				}
			}
			if (skip)
				entry |= JOB_SKIP;
			else if (s == SUM_OPAQUE)
				entry |= JOB_OPAQUE;
		}
		if (live)
			dst[j] = entry;
		if (order != nullptr)
		{
			int b = JOB_BUCKETS - 1;
			if (live && (entry & (JOB_HAS_CONTENT | JOB_SKIP)) == JOB_HAS_CONTENT)
			{
				const uint32_t lf = lastFaces[entry & JOB_SLOT_MASK];
				b = lf >= 4096u ? 0 : (lf >= 512u ? 1 : 2);
			}
			const int lane = threadIdx.x & 31;
#pragma unroll
			for (int q = 0; q < JOB_BUCKETS; ++q)
			{
				const uint32_t members = __ballot_sync(0xffffffffu, live && b == q);
				if (members == 0u)
					continue;
				uint32_t base = 0u;
				if (lane == __ffs(members) - 1)
					base = atomicAdd(&bucketCount[q], uint32_t(__popc(members)));
				base = __shfl_sync(0xffffffffu, base, __ffs(members) - 1);
				if (live && b == q)
					order[size_t(q) * n + base + __popc(members & ((1u << lane) - 1u))] = j;
			}
		}
	}

	void launch_job_classify(const uint32_t* src, uint32_t* dst, uint32_t n, const uint8_t* summary, const int32_t* nbrSlot, const uint32_t* lastFaces, uint32_t* order,
	                         uint32_t* bucketCount, cudaStream_t st)
	{
		if (n)
			job_classify_kernel<<<(n + 255u) / 256u, 256, 0, st>>>(src, dst, n, summary, nbrSlot, lastFaces, order, bucketCount);
	}

	// ---- launchers ----
	size_t mesh_fused_smem_bytes() { return sizeof(FusedSmem); }

	template<int MODE>
	static cudaError_t init_fused()
	{
		cudaError_t e = cudaFuncSetAttribute(mesh_fused_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(FusedSmem)));
		return e != cudaSuccess ? e : cudaFuncSetAttribute(mesh_fused_kernel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
	}

	cudaError_t mesh_kernels_init()
	{
		cudaError_t e;
		if ((e = init_fused<MODE_LEAN>()) != cudaSuccess || (e = init_fused<MODE_GENERIC>()) != cudaSuccess)
			return e;
		return init_fused<MODE_TANGENTS>();
	}

	void launch_mesh_fused(const MeshArgs& a, int grid, cudaStream_t st)
	{
		if (a.flags & MESH_F_TANGENTS)
			mesh_fused_kernel<MODE_TANGENTS><<<grid, FUSED_THREADS, sizeof(FusedSmem), st>>>(a);
		else if (a.faceTable != nullptr && !(a.flags & MESH_F_DEFORMED))
			mesh_fused_kernel<MODE_LEAN><<<grid, FUSED_THREADS, sizeof(FusedSmem), st>>>(a);
		else
			mesh_fused_kernel<MODE_GENERIC><<<grid, FUSED_THREADS, sizeof(FusedSmem), st>>>(a);
	}
}
