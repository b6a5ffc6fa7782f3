// oracle/ref_shim — stand-in for <tsl/hopscotch_map.h>: std::unordered_map with the heterogeneous find() the reference uses
// (BlockLibrary looks names up by std::string_view).  Iteration order is never relied upon by the chunk pipeline.
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
			using Base::find;
			template<typename K, typename = std::enable_if_t<!std::is_same_v<std::decay_t<K>, Key>>> auto find(const K& k) { return Base::find(Key(k)); }
			template<typename K, typename = std::enable_if_t<!std::is_same_v<std::decay_t<K>, Key>>> auto find(const K& k) const { return Base::find(Key(k)); }
			template<typename K> bool contains(const K& k) const { return find(k) != Base::end(); }
	};
}
