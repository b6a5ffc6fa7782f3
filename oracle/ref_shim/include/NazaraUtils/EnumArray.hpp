// oracle/ref_shim — stand-in for <NazaraUtils/EnumArray.hpp>: a std::array indexed by an enum whose last value is E::Max
#pragma once
#include <NazaraUtils/Prerequisites.hpp>
#include <array>
#include <utility>

namespace Nz
{
	template<typename E> constexpr std::size_t EnumValueCount_v = static_cast<std::size_t>(E::Max) + 1;

	template<typename E, typename V> class EnumArray : public std::array<V, EnumValueCount_v<E>>
	{
		using Base = std::array<V, EnumValueCount_v<E>>;

		public:
			using Base::operator[];
			constexpr V& operator[](E e) { return Base::operator[](static_cast<std::size_t>(e)); }
			constexpr const V& operator[](E e) const { return Base::operator[](static_cast<std::size_t>(e)); }

			template<typename Arr, typename Ref> struct KvIterator
			{
				Arr* arr;
				std::size_t i;
				constexpr std::pair<E, Ref> operator*() const { return { static_cast<E>(i), (*arr)[static_cast<E>(i)] }; }
				constexpr KvIterator& operator++() { ++i; return *this; }
				constexpr bool operator!=(const KvIterator& o) const { return i != o.i; }
			};
			template<typename Arr, typename Ref> struct KvRange
			{
				Arr* arr;
				constexpr KvIterator<Arr, Ref> begin() const { return { arr, 0 }; }
				constexpr KvIterator<Arr, Ref> end() const { return { arr, EnumValueCount_v<E> }; }
			};
			constexpr KvRange<EnumArray, V&> iter_kv() { return { this }; }
			constexpr KvRange<const EnumArray, const V&> iter_kv() const { return { this }; }
	};
}
