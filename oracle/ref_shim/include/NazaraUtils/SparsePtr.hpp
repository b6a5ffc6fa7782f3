// oracle/ref_shim — stand-in for <NazaraUtils/SparsePtr.hpp>: pointer with a byte stride
#pragma once
#include <cstddef>
#include <cstdint>

namespace Nz
{
	template<typename T> class SparsePtr
	{
		public:
			SparsePtr() : m_ptr(nullptr), m_stride(0) {}
			SparsePtr(T* ptr, int stride = sizeof(T)) : m_ptr(reinterpret_cast<std::uint8_t*>(ptr)), m_stride(stride) {}
			explicit operator bool() const { return m_ptr != nullptr; }
			T& operator[](std::size_t i) const { return *reinterpret_cast<T*>(m_ptr + std::ptrdiff_t(i) * m_stride); }
			T* GetPtr() const { return reinterpret_cast<T*>(m_ptr); }

		private:
			std::uint8_t* m_ptr;
			int m_stride;
	};
}
