// TEST INFRASTRUCTURE (oracle build only) -- minimal stand-in for "kgd/utils/assertequal.hpp".
// The reference's assertEqual friends (node.hpp:222-235, speciesdata.hpp:49-57, ...) delegate
// to utils::assertEqual for leaves and standard containers; user types are reached through ADL.
#ifndef APT_ORACLE_SHIM_ASSERTEQUAL_HPP
#define APT_ORACLE_SHIM_ASSERTEQUAL_HPP

#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <type_traits>
#include <vector>

namespace utils {

template <typename T>
std::enable_if_t<std::is_arithmetic<T>::value || std::is_enum<T>::value, void>
assertEqual (const T &lhs, const T &rhs, bool /*deepcopy*/) {
  if (!(lhs == rhs)) throw std::logic_error("assertEqual: values differ");
}

/// Stateless helper types (e.g. NoUserData) are always equal
template <typename T>
std::enable_if_t<std::is_class<T>::value && std::is_empty<T>::value, void>
assertEqual (const T &, const T &, bool) {}

template <typename T> void assertEqual (const std::vector<T> &l, const std::vector<T> &r, bool d);
template <typename K, typename V> void assertEqual (const std::map<K,V> &l, const std::map<K,V> &r, bool d);
template <typename T> void assertEqual (const std::set<T> &l, const std::set<T> &r, bool d);
template <typename T> void assertEqual (const std::shared_ptr<T> &l, const std::shared_ptr<T> &r, bool d);
template <typename T> void assertEqual (const std::unique_ptr<T> &l, const std::unique_ptr<T> &r, bool d);

template <typename T>
void assertEqual (const std::vector<T> &l, const std::vector<T> &r, bool d) {
  if (l.size() != r.size()) throw std::logic_error("assertEqual: vector sizes differ");
  for (size_t i = 0; i < l.size(); i++) assertEqual(l[i], r[i], d);
}

template <typename K, typename V>
void assertEqual (const std::map<K,V> &l, const std::map<K,V> &r, bool d) {
  if (l.size() != r.size()) throw std::logic_error("assertEqual: map sizes differ");
  auto il = l.begin(), ir = r.begin();
  for (; il != l.end(); ++il, ++ir) {
    assertEqual(il->first, ir->first, d);
    assertEqual(il->second, ir->second, d);
  }
}

template <typename T>
void assertEqual (const std::set<T> &l, const std::set<T> &r, bool d) {
  if (l.size() != r.size()) throw std::logic_error("assertEqual: set sizes differ");
  auto il = l.begin(), ir = r.begin();
  for (; il != l.end(); ++il, ++ir) assertEqual(*il, *ir, d);
}

template <typename P>
void assertEqualPtr (const P &l, const P &r, bool d) {
  if (bool(l) != bool(r)) throw std::logic_error("assertEqual: null-ness differs");
  if (!l) return;
  if (d && l.get() == r.get()) throw std::logic_error("assertEqual: deep copy shares storage");
  assertEqual(*l, *r, d);
}

template <typename T>
void assertEqual (const std::shared_ptr<T> &l, const std::shared_ptr<T> &r, bool d) { assertEqualPtr(l, r, d); }
template <typename T>
void assertEqual (const std::unique_ptr<T> &l, const std::unique_ptr<T> &r, bool d) { assertEqualPtr(l, r, d); }

} // namespace utils

#endif
