// oracle/ref_shim — stand-in for <Nazara/Physics3D/Collider3D.hpp>: the collider classes keep what they were given so the
// harness can read the boxes / triangles back (Jolt itself is out of scope).
#pragma once
#include <Nazara/Math/Vector3.hpp>
#include <memory>
#include <vector>

namespace Nz
{
	class Collider3D
	{
		public:
			virtual ~Collider3D() = default;
	};

	class BoxCollider3D : public Collider3D
	{
		public:
			explicit BoxCollider3D(const Vector3f& lengths, float /*convexRadius*/ = 0.f) : m_lengths(lengths) {}
			const Vector3f& GetLengths() const { return m_lengths; }

		private:
			Vector3f m_lengths;
	};

	class CompoundCollider3D : public Collider3D
	{
		public:
			struct ChildCollider
			{
				std::shared_ptr<Collider3D> collider;
				Quaternionf rotation = Quaternionf::Identity();
				Vector3f offset = Vector3f::Zero();
			};
			explicit CompoundCollider3D(std::vector<ChildCollider> children) : m_children(std::move(children)) {}
			const std::vector<ChildCollider>& GetGeoms() const { return m_children; }

		private:
			std::vector<ChildCollider> m_children;
	};

	class MeshCollider3D : public Collider3D
	{
		public:
			MeshCollider3D(const Vector3f* vertices, std::size_t vertexCount, const UInt32* indices, std::size_t indexCount) :
			m_vertices(vertices, vertices + vertexCount), m_indices(indices, indices + indexCount)
			{
			}
			const std::vector<Vector3f>& GetVertices() const { return m_vertices; }
			const std::vector<UInt32>& GetIndices() const { return m_indices; }

		private:
			std::vector<Vector3f> m_vertices;
			std::vector<UInt32> m_indices;
	};
}
