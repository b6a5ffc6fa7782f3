// oracle/ref_shim — stand-in for <tsl/hopscotch_map.h>: std::unordered_map with the heterogeneous find() the reference uses
// (BlockLibrary looks names up by std::string_view) and iterators that also answer tsl's .key() / .value() (Ship.cpp:86).
// Iteration order is never relied upon by the chunk pipeline.
#pragma once
#include <functional>
#include <unordered_map>

namespace tsl
{
	template<typename Key, typename T, typename Hash = std::hash<Key>, typename KeyEqual = std::equal_to<Key>>
	class hopscotch_map : public std::unordered_map<Key, T, std::hash<Key>>
	{
		using Base = std::unordered_map<Key, T, std::hash<Key>>;

		public:
			struct iterator : Base::iterator
			{
				iterator() = default;
				iterator(typename Base::iterator it) : Base::iterator(it) {}
				const Key& key() const { return (**this).first; }
				T& value() const { return (**this).second; }
			};
			struct const_iterator : Base::const_iterator
			{
				const_iterator() = default;
				const_iterator(typename Base::const_iterator it) : Base::const_iterator(it) {}
				const Key& key() const { return (**this).first; }
				const T& value() const { return (**this).second; }
			};
			iterator begin() { return Base::begin(); }
			iterator end() { return Base::end(); }
			const_iterator begin() const { return Base::begin(); }
			const_iterator end() const { return Base::end(); }
			iterator find(const Key& k) { return Base::find(k); }
			const_iterator find(const Key& k) const { return Base::find(k); }
			template<typename K, typename = std::enable_if_t<!std::is_same_v<std::decay_t<K>, Key>>> iterator find(const K& k) { return Base::find(Key(k)); }
			template<typename K, typename = std::enable_if_t<!std::is_same_v<std::decay_t<K>, Key>>> const_iterator find(const K& k) const { return Base::find(Key(k)); }
			template<typename K> bool contains(const K& k) const { return find(k) != end(); }
	};
}
