// Host-side helpers of the estimator layer (no device code).
//
// b2svm_host_std: the population standard deviation of a C-contiguous float32 / float64 matrix, bit-identical to
// numpy's `X.std()` -- gamma='scale' is 1 / (n_features * X.std()) (reference svc.py:212-217), and a last-bit change
// of gamma changes every kernel value.  numpy reduces a contiguous array with one pairwise summation over all
// elements (8 interleaved accumulators on blocks of <= 128, recursive halving above that, numpy
// core/src/umath/loops_utils.h.src `pairwise_sum`), computes the mean, then the same summation over (x - mean)^2.
// The recursion is reproduced exactly; its independent sub-trees run on several threads, so the 48 ms numpy needs
// for the 200 000 x 128 matrix of BASELINE.json configs[2] become a few ms of the end-to-end fit.
#include <cmath>
#include <cstdint>
#include <future>
#include <thread>

#include "b2svm_internal.h"

namespace b2svm {
namespace {

template <typename T, typename F>
T pairwise_block(F elem, int64_t lo, int64_t n) {   // n <= 128
  if (n < 8) {
    T res = T(0);
    for (int64_t i = 0; i < n; ++i) res = res + elem(lo + i);
    return res;
  }
  T r[8];
  for (int k = 0; k < 8; ++k) r[k] = elem(lo + k);
  int64_t i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int k = 0; k < 8; ++k) r[k] = r[k] + elem(lo + i + k);
  T res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res = res + elem(lo + i);
  return res;
}

template <typename T, typename F>
T pairwise(F elem, int64_t lo, int64_t n, int par_depth) {
  if (n <= 128) return pairwise_block<T>(elem, lo, n);
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  if (par_depth > 0 && n > (1 << 16)) {
    auto left = std::async(std::launch::async, [=]() { return pairwise<T>(elem, lo, n2, par_depth - 1); });
    const T right = pairwise<T>(elem, lo + n2, n - n2, par_depth - 1);
    return left.get() + right;
  }
  return pairwise<T>(elem, lo, n2, 0) + pairwise<T>(elem, lo + n2, n - n2, 0);
}

template <typename T>
double std_T(const T* x, int64_t n, int par_depth) {
  const T sum = pairwise<T>([x](int64_t i) { return x[i]; }, 0, n, par_depth);
  const T mean = sum / (T)n;
  // (x - mean) and its square are rounded to T separately, like numpy's temporaries (no fused multiply-add)
  const T ss = pairwise<T>(
      [x, mean](int64_t i) {
        volatile T d = x[i] - mean;
        const T dd = d;
        return dd * dd;
      },
      0, n, par_depth);
  const T var = ss / (T)n;
  return (double)(T)std::sqrt(var);
}

}  // namespace
}  // namespace b2svm

extern "C" int b2svm_host_std(const void* x, int64_t n_elements, int32_t x_dtype, double* out) {
  using namespace b2svm;
  if (!x || !out || n_elements <= 0) return fail(B2SVM_E_INVALID, "b2svm_host_std: empty input");
  unsigned hw = std::thread::hardware_concurrency();
  int depth = 0;
  while ((1u << (depth + 1)) <= hw && depth < 5) ++depth;   // up to 32 sub-trees in flight
  if (x_dtype == B2SVM_F32) *out = std_T<float>(static_cast<const float*>(x), n_elements, depth);
  else if (x_dtype == B2SVM_F64) *out = std_T<double>(static_cast<const double*>(x), n_elements, depth);
  else return fail(B2SVM_E_INVALID, "b2svm_host_std: dtype");
  return B2SVM_OK;
}
