// join_helpers.hpp -- joined-table types and the host-side result check of the Join dwarfs.
//
// Same public names as the reference's join/join_helpers/join_helpers.hpp (ColJoinedTableTy,
// RowJoinedTableTy, is_malformed, get_size, to_row_store, to_col_store, seq_join, eq, operator==, zip,
// operator<<) so code written against it compiles here.  The O(n log n) companion the dwarfs use to fill
// Result::valid lives in sort_join.hpp.
#pragma once

#include <algorithm>
#include <cstddef>
#include <ostream>
#include <stdexcept>
#include <utility>
#include <vector>

namespace join_helpers {

template <class Key, class Val1, class Val2>
using ColJoinedTableTy = std::pair<std::vector<Key>, std::pair<std::vector<Val1>, std::vector<Val2>>>;

template <class Key, class V1, class V2> using RowJoinedTableTy = std::vector<std::pair<Key, std::pair<V1, V2>>>;

template <class K, class V1, class V2> bool is_malformed(const ColJoinedTableTy<K, V1, V2> &t) {
  const size_t n = t.first.size();
  return t.second.first.size() != n || t.second.second.size() != n;
}

template <class K, class V1, class V2> size_t get_size(const ColJoinedTableTy<K, V1, V2> &t) {
  if (is_malformed(t)) throw std::invalid_argument("ColJoinedTableTy is malformed.");
  return t.first.size();
}

template <class K, class V1, class V2> size_t get_size(const RowJoinedTableTy<K, V1, V2> &t) { return t.size(); }

template <class Key, class V1, class V2> RowJoinedTableTy<Key, V1, V2> to_row_store(const ColJoinedTableTy<Key, V1, V2> &t) {
  const size_t n = get_size(t);
  RowJoinedTableTy<Key, V1, V2> rows;
  rows.reserve(n);
  for (size_t i = 0; i < n; ++i) rows.emplace_back(t.first[i], std::make_pair(t.second.first[i], t.second.second[i]));
  return rows;
}

template <class Key, class V1, class V2> ColJoinedTableTy<Key, V1, V2> to_col_store(const RowJoinedTableTy<Key, V1, V2> &t) {
  ColJoinedTableTy<Key, V1, V2> cols;
  cols.first.reserve(t.size());
  cols.second.first.reserve(t.size());
  cols.second.second.reserve(t.size());
  for (const auto &row : t) {
    cols.first.push_back(row.first);
    cols.second.first.push_back(row.second.first);
    cols.second.second.push_back(row.second.second);
  }
  return cols;
}

template <class Key, class Val1, class Val2> std::ostream &operator<<(std::ostream &os, const ColJoinedTableTy<Key, Val1, Val2> &t) {
  const size_t n = get_size(t);
  for (size_t i = 0; i < n; ++i) os << (i ? "\n" : "") << t.first[i] << " " << t.second.first[i] << " " << t.second.second[i];
  return os;
}

template <class Key, class Val1, class Val2> std::ostream &operator<<(std::ostream &os, const RowJoinedTableTy<Key, Val1, Val2> &t) {
  for (const auto &row : t) os << row.first << ' ' << row.second.first << ' ' << row.second.second << std::endl;
  return os;
}

// The definition of the join result: every (i, j) with a_keys[i] == b_keys[j], emitted i-major.
template <class K, class V1, class V2>
ColJoinedTableTy<K, V1, V2> seq_join(const std::vector<K> &a_keys, const std::vector<V1> &a_vals, const std::vector<K> &b_keys,
                                     const std::vector<V2> &b_vals) {
  ColJoinedTableTy<K, V1, V2> out;
  for (size_t i = 0; i < a_keys.size(); ++i)
    for (size_t j = 0; j < b_keys.size(); ++j)
      if (a_keys[i] == b_keys[j]) {
        out.first.push_back(a_keys[i]);
        out.second.first.push_back(a_vals[i]);
        out.second.second.push_back(b_vals[j]);
      }
  return out;
}

// Order-insensitive equality: sort both row lists, compare.
template <class K, class V1, class V2> bool eq(const RowJoinedTableTy<K, V1, V2> &t1, const RowJoinedTableTy<K, V1, V2> &t2) {
  if (t1.size() != t2.size()) return false;
  RowJoinedTableTy<K, V1, V2> s1 = t1, s2 = t2;
  std::sort(s1.begin(), s1.end());
  std::sort(s2.begin(), s2.end());
  return s1 == s2;
}

template <class K, class V1, class V2> bool operator==(const ColJoinedTableTy<K, V1, V2> &t1, const ColJoinedTableTy<K, V1, V2> &t2) {
  if (is_malformed(t1) || is_malformed(t2)) throw std::invalid_argument("ColJoinedTableTy is malformed.");
  return eq(to_row_store(t1), to_row_store(t2));
}

template <class K, class V1, class V2>
ColJoinedTableTy<K, V1, V2> zip(const std::vector<K> &keys, const std::vector<V1> &v1, const std::vector<V2> &v2) {
  return ColJoinedTableTy<K, V1, V2>{keys, {v1, v2}};
}

}  // namespace join_helpers
