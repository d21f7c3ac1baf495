// Minimal host-only stand-in for <CL/sycl.hpp>.
//
// TEST INFRASTRUCTURE ONLY.  It exists so that the reference's own headers
// (common/dpcpp/hashtable.hpp, common/dpcpp/hashfunctions.hpp,
// join/join_helpers/join_helpers.hpp) compile UNMODIFIED with plain g++ into
// oracle/_ref/libref_join.so -- no SYCL compiler exists in this image.  Only
// the handful of names those headers touch are provided: global_ptr (a raw
// pointer), sycl::atomic<T> over a global pointer (GCC __atomic builtins, the
// same seq_cst semantics the SYCL 1.2.1 class defaults to relaxed-or-stronger),
// sycl::ext::intel::ctz, and opaque device / device_selector types that
// dpcpp_common.hpp merely declares functions over.
#pragma once
#include <climits>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace cl {
namespace sycl {

template <class T> using global_ptr = T *;

template <class T> class atomic {
public:
  explicit atomic(T *p) : p_(p) {}
  T fetch_or(T v) { return __atomic_fetch_or(p_, v, __ATOMIC_SEQ_CST); }
  T fetch_and(T v) { return __atomic_fetch_and(p_, v, __ATOMIC_SEQ_CST); }
  T fetch_add(T v) { return __atomic_fetch_add(p_, v, __ATOMIC_SEQ_CST); }
  void store(T v) { __atomic_store_n(p_, v, __ATOMIC_SEQ_CST); }
  T load() const { return __atomic_load_n(p_, __ATOMIC_SEQ_CST); }
  bool compare_exchange_strong(T &expected, T desired) {
    return __atomic_compare_exchange_n(p_, &expected, desired, false, __ATOMIC_SEQ_CST,
                                       __ATOMIC_SEQ_CST);
  }

private:
  T *p_;
};

class device {};
class device_selector {
public:
  virtual ~device_selector() = default;
};

namespace ext {
namespace intel {
// Count trailing zeros; the bit width for a zero argument (SYCL/OpenCL ctz).
template <class T> inline T ctz(T x) {
  static_assert(std::is_unsigned<T>::value && sizeof(T) <= 8, "unsigned integral expected");
  if (x == 0) return static_cast<T>(sizeof(T) * CHAR_BIT);
  return static_cast<T>(__builtin_ctzll(static_cast<unsigned long long>(x)));
}
} // namespace intel
} // namespace ext

} // namespace sycl
} // namespace cl

namespace sycl = cl::sycl;
