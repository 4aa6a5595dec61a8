// Device-side helpers shared by the b2svm kernels (sm_100a only).
//
// Compiled with -fmad=false: every fused multiply-add in this tree is an explicit fma()/fmaf()
// call (dot products), everything else rounds once per operation like the NumPy reference does
// (pysvm/solver.py:86-146 is plain float64 arithmetic).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace b2svm {

// ------------------------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA, 1-D) wrappers.  SASS: SYNCS.* / UBLKCP.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// global -> shared bulk copy completing on an mbarrier (bytes and both addresses 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// the same copy with an L2 eviction-priority hint: X is re-read by every sweep, and most of it fits the 126 MB L2
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// 16-byte shared-memory units that are polled by another warp of the CTA
__device__ __forceinline__ void st_volatile_shared_v4(void* p, uint4 v) {
  asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_shared_v4(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// 16-byte relaxed (L1-bypassing, single-transaction) global accesses: the working-set records carry
// their own epoch tag in each 16-byte half, so no fence is needed around them (b2svm_smo.cuh).
__device__ __forceinline__ uint4 ld_relaxed_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_v4(void* p, uint4 v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// system-scope variants for records that travel between GPUs over NVLink peer mappings
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_v4(void* p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------
// 16-byte chunks of a row in the element type T
// ------------------------------------------------------------------------------------------
template <typename T>
struct Chunk;
template <>
struct Chunk<float> {
  using V = float4;
  static constexpr int N = 4;
  static __device__ __forceinline__ V zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ float dot(const V& a, const V& b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
    return acc;
  }
  // component-wise accumulation: four independent FMA chains per vector
  static __device__ __forceinline__ V fma_each(const V& a, const V& b, V acc) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.y = fmaf(a.y, b.y, acc.y);
    acc.z = fmaf(a.z, b.z, acc.z);
    acc.w = fmaf(a.w, b.w, acc.w);
    return acc;
  }
  static __device__ __forceinline__ float hsum(const V& a) { return (a.x + a.y) + (a.z + a.w); }
};
template <>
struct Chunk<double> {
  using V = double2;
  static constexpr int N = 2;
  static __device__ __forceinline__ V zero() { return make_double2(0.0, 0.0); }
  static __device__ __forceinline__ double dot(const V& a, const V& b, double acc) {
    acc = fma(a.x, b.x, acc);
    acc = fma(a.y, b.y, acc);
    return acc;
  }
  static __device__ __forceinline__ V fma_each(const V& a, const V& b, V acc) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.y = fma(a.y, b.y, acc.y);
    return acc;
  }
  static __device__ __forceinline__ double hsum(const V& a) { return a.x + a.y; }
};

// ------------------------------------------------------------------------------------------
// kernel functions, evaluated in the input dtype with the reference's operation order
// (pysvm/svm/svc.py:230-241): gamma / coef0 / degree are Python scalars there, so under NumPy's
// promotion rules they are rounded to the array dtype first -- KernelParams<T> holds them in T.
// ------------------------------------------------------------------------------------------
enum { K_LINEAR = 0, K_POLY = 1, K_RBF = 2, K_SIGMOID = 3, K_PRECOMPUTED = 4 };

template <typename T>
struct KernelParams {
  int kind;
  T gamma, coef0, degree;
};

__device__ __forceinline__ float t_exp(float x) { return expf(x); }
__device__ __forceinline__ double t_exp(double x) { return exp(x); }
__device__ __forceinline__ float t_pow(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double t_pow(double x, double y) { return pow(x, y); }
__device__ __forceinline__ float t_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ double t_tanh(double x) { return tanh(x); }

template <typename T>
__device__ __forceinline__ T kernel_value(const KernelParams<T>& k, T dot, T norm_a, T norm_b) {
  switch (k.kind) {
    case K_RBF:  // exp(-gamma * ((|a|^2 + |b|^2) - 2 a.b))      svc.py:230-232
      return t_exp(-k.gamma * ((norm_a + norm_b) - T(2) * dot));
    case K_POLY:  // (gamma * a.b + coef0) ** degree              svc.py:238
      return t_pow(k.gamma * dot + k.coef0, k.degree);
    case K_SIGMOID:  // tanh(gamma * a.b + coef0)                 svc.py:240
      return t_tanh(k.gamma * dot + k.coef0);
    default:  // a.b                                              svc.py:237
      return dot;
  }
}

// same, with the kind resolved at compile time (dense epilogues: keeps pow / tanh out of the unrolled rbf loop)
template <int KIND, typename T>
__device__ __forceinline__ T kernel_value_k(const KernelParams<T>& k, T dot, T norm_a, T norm_b) {
  if constexpr (KIND == K_RBF) return t_exp(-k.gamma * ((norm_a + norm_b) - T(2) * dot));
  else if constexpr (KIND == K_POLY) return t_pow(k.gamma * dot + k.coef0, k.degree);
  else if constexpr (KIND == K_SIGMOID) return t_tanh(k.gamma * dot + k.coef0);
  else return dot;
}

// ------------------------------------------------------------------------------------------
// working-set candidates.  Order: value first, then the LOWER index (np.argmax / np.argmin over an
// index-sorted subset return the first extremum: solver.py:68-69).  idx < 0 = empty set.
// ------------------------------------------------------------------------------------------
struct Cand {        // lane / warp level: value, variable index, flags byte of that variable
  double g;
  int idx;
  int flags;
};

// Order-preserving map double -> uint64 (so arg-max over doubles becomes integer REDUX.MAX), and back.
__device__ __forceinline__ unsigned long long dkey(double g) {
  const long long b = __double_as_longlong(g);
  return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
  const long long b = (k >> 63) ? (long long)(k ^ 0x8000000000000000ull) : (long long)~k;
  return __longlong_as_double(b);
}

// Arg-max (is_max) / arg-min over the lanes in `mask` with ties going to the lower index: three
// REDUX instructions (hi word, lo word, packed index) instead of a 5-round shuffle butterfly.
// Every lane of `mask` must call it with the same mask; all of them receive the winner.
__device__ __forceinline__ void redux_select(unsigned mask, bool is_max, Cand& c) {
  unsigned long long key = c.idx >= 0 ? dkey(c.g) : 0ull;      // 0 is below every real value's key
  if (!is_max && c.idx >= 0) key = ~key;
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned ik = c.idx >= 0 ? (((unsigned)c.idx << 3) | (unsigned)c.flags) : 0xffffffffu;
  const unsigned mh = __reduce_max_sync(mask, hi);
  const unsigned ml = __reduce_max_sync(mask, hi == mh ? lo : 0u);
  const unsigned mi = __reduce_min_sync(mask, (hi == mh && lo == ml) ? ik : 0xffffffffu);
  if (mi == 0xffffffffu) {
    c.g = 0.0; c.idx = -1; c.flags = 0;
  } else {
    unsigned long long k = ((unsigned long long)mh << 32) | ml;
    if (!is_max) k = ~k;
    c.g = dkey_inv(k);
    c.idx = (int)(mi >> 3);
    c.flags = (int)(mi & 7u);
  }
}
__device__ __forceinline__ bool better_max(double g, int idx, double bg, int bidx) {
  return idx >= 0 && (bidx < 0 || g > bg || (g == bg && idx < bidx));
}
__device__ __forceinline__ bool better_min(double g, int idx, double bg, int bidx) {
  return idx >= 0 && (bidx < 0 || g < bg || (g == bg && idx < bidx));
}

// flags byte per variable: bit0 y>0, bit1 alpha>0, bit2 alpha<C   (solver.py:55-64)
enum { F_YPOS = 1, F_AGT0 = 2, F_ALTC = 4 };
__device__ __forceinline__ uint8_t make_flags(bool ypos, double alpha, double C) {
  return (ypos ? F_YPOS : 0) | (alpha > 0 ? F_AGT0 : 0) | (alpha < C ? F_ALTC : 0);
}
__device__ __forceinline__ bool in_up(int f) { return (f & F_YPOS) ? (f & F_ALTC) : (f & F_AGT0); }
__device__ __forceinline__ bool in_low(int f) { return (f & F_YPOS) ? (f & F_AGT0) : (f & F_ALTC); }

// ------------------------------------------------------------------------------------------
// The analytic two-variable step with LIBSVM box clipping, literal order of the reference
// (Solver.update solver.py:86-145; NuSolver.update solver.py:343-373).
// ------------------------------------------------------------------------------------------
struct StepOut {
  double ai, aj, di, dj;
};
__device__ __forceinline__ StepOut smo_step(bool nu, double Kii, double Kjj, double Kij, double yi, double yj,
                                            double gi, double gj, double ai0, double aj0, double C) {
  double quad = (Kii + Kjj) - 2.0 * Kij;
  if (quad <= 0) quad = 1e-12;
  double ai = ai0, aj = aj0;
  if (!nu && yi * yj == -1.0) {
    const double delta = (gi * yi + gj * yj) / quad;
    const double diff = ai - aj;
    ai += delta;
    aj += delta;
    if (diff > 0) {
      if (aj < 0) { aj = 0; ai = diff; }
    } else {
      if (ai < 0) { ai = 0; aj = -diff; }
    }
    if (diff > 0) {
      if (ai > C) { ai = C; aj = C - diff; }
    } else {
      if (aj > C) { aj = C; ai = C + diff; }
    }
  } else {
    const double delta = nu ? (yi * (gj - gi)) / quad : (gj * yj - gi * yi) / quad;
    const double sum = ai + aj;
    ai -= delta;
    aj += delta;
    if (sum > C) {
      if (ai > C) { ai = C; aj = sum - C; }
    } else {
      if (aj < 0) { aj = 0; ai = sum; }
    }
    if (sum > C) {
      if (aj > C) { aj = C; ai = sum - C; }
    } else {
      if (ai < 0) { ai = 0; aj = sum; }
    }
  }
  StepOut o;
  o.ai = ai;
  o.aj = aj;
  o.di = ai - ai0;
  o.dj = aj - aj0;
  return o;
}

// ------------------------------------------------------------------------------------------
// Transposed warp reduction.  Every lane of a group of LPR lanes holds CNT partial sums; after
// log2(LPR) exchange steps each lane holds CNT/LPR complete sums: lane `r` of the group ends with
// the values whose index has r as its top bits (index = r * (CNT/LPR) + low).  One shuffle per
// value per halving instead of log2(LPR) shuffles per value.
// ------------------------------------------------------------------------------------------
template <typename T, int CNT, int M>
struct TransposeReduce {
  template <int NV>
  static __device__ __forceinline__ void run(T (&v)[NV], int lane_in_group) {
    constexpr int HALF = CNT / 2;
    const bool upper = (lane_in_group & M) != 0;
#pragma unroll
    for (int k = 0; k < HALF; ++k) {
      const T send = upper ? v[k] : v[k + HALF];
      const T keep = upper ? v[k + HALF] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, M);
    }
    TransposeReduce<T, HALF, M / 2>::run(v, lane_in_group);
  }
};
template <typename T, int CNT>
struct TransposeReduce<T, CNT, 0> {
  template <int NV>
  static __device__ __forceinline__ void run(T (&)[NV], int) {}
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

}  // namespace b2svm
