// oracle/nazara_tangents.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Restatement of Nz::SubMesh::GenerateTangents, which ClientChunkEntities::BuildMesh runs on every chunk mesh right after
// Chunk::BuildMesh (reference src/ClientLib/ClientChunkEntities.cpp:169).  NazaraEngine (xmake package nazaraengine >= 2023.08.15,
// un-vendored) is absent from /root/reference and from this box, and the reference ships no test for it: PARITY UNPINNED.  This is
// the one CPU statement of the algorithm; oracle/capi.cpp and oracle/ref_shim/ref_capi.cpp both export it as orc_generate_tangents.
//
// Algorithm as published in the engine (src/Nazara/Core/SubMesh.cpp): zero every vertex tangent; for every triangle (i0, i1, i2) of
// the index buffer: dv0 = p[i1] - p[i0], dv1 = p[i2] - p[i0], duv0 = uv[i1] - uv[i0], duv1 = uv[i2] - uv[i0],
// coef = 1 / (duv0.x * duv1.y - duv1.x * duv0.y), tangent = coef * (dv0 * duv1.y + dv1 * (-duv0.y)), added to the three vertices;
// finally every vertex tangent is normalised (multiplied by 1 / length when the length is > 0).  uv = the first two components of
// the chunk vertex's TexCoord (u, v); the third one is the texture slice.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace nazara_tangents
{
	inline void Generate(const float* pos, const float* uvw, const std::uint32_t* indices, std::size_t vertexCount, std::size_t indexCount, float* tangents)
	{
		for (std::size_t i = 0; i < vertexCount * 3; ++i)
			tangents[i] = 0.f;
		for (std::size_t t = 0; t + 2 < indexCount; t += 3)
		{
			const std::uint32_t i0 = indices[t], i1 = indices[t + 1], i2 = indices[t + 2];
			const float* p0 = pos + std::size_t(i0) * 3;
			float dv0[3], dv1[3];
			for (int k = 0; k < 3; ++k)
			{
				dv0[k] = pos[std::size_t(i1) * 3 + k] - p0[k];
				dv1[k] = pos[std::size_t(i2) * 3 + k] - p0[k];
			}
			const float u0 = uvw[std::size_t(i0) * 3], v0 = uvw[std::size_t(i0) * 3 + 1];
			const float du0 = uvw[std::size_t(i1) * 3] - u0, dw0 = uvw[std::size_t(i1) * 3 + 1] - v0;
			const float du1 = uvw[std::size_t(i2) * 3] - u0, dw1 = uvw[std::size_t(i2) * 3 + 1] - v0;
			const float coef = 1.f / (du0 * dw1 - du1 * dw0);
			for (int k = 0; k < 3; ++k)
			{
				const float tg = coef * (dv0[k] * dw1 + dv1[k] * (-dw0));
				tangents[std::size_t(i0) * 3 + k] += tg;
				tangents[std::size_t(i1) * 3 + k] += tg;
				tangents[std::size_t(i2) * 3 + k] += tg;
			}
		}
		for (std::size_t i = 0; i < vertexCount; ++i)
		{
			float* t = tangents + i * 3;
			const float len = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
			if (len > 0.f)
			{
				const float inv = 1.f / len;
				t[0] *= inv;
				t[1] *= inv;
				t[2] *= inv;
			}
		}
	}
}
