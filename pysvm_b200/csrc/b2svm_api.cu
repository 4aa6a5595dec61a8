// libb2svm.so: C-ABI entry points of the solver (see include/b2svm.h) and the small support
// kernels around the persistent SMO kernel.  Build: pysvm_b200/build.py (nvcc, sm_100a).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "b2svm_internal.h"
#include "b2svm_smo.cuh"

namespace b2svm {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int chunks_for(int x_dtype, int d) {
  const int per = x_dtype == B2SVM_F64 ? 2 : 4;
  const int need = (d + per - 1) / per;
  for (int ch = 4; ch <= 128; ch *= 2)
    if (need <= ch) return ch;
  return 0;
}

// ------------------------------------------------------------------------------------------
// support kernels
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void pad_rows_kernel(const T* __restrict__ src, int64_t ld, T* __restrict__ dst, int64_t l, int d, int dp) {
  const int64_t total = l * dp;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / dp;
    const int c = (int)(e - r * dp);
    dst[e] = c < d ? src[r * ld + c] : T(0);
  }
}

// |x_r|^2 in T (reference: (x**2).sum(1), svc.py:231) and K(x_r, x_r) in T, one warp per row.  The diagonal is
// constant over the solve, so the select -> step chain reads it instead of re-evaluating it for every winner.
template <typename T>
__global__ void row_norms_kernel(const T* __restrict__ X, int64_t l, int dp, T* __restrict__ norms, T* __restrict__ kdiag,
                                 KernelParams<T> kp) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int EPC = 16 / (int)sizeof(T);   // elements per 16-byte chunk
  for (int64_t r = warp; r < l; r += nwarps) {
    T s = 0, dd = 0;
    for (int c = lane; c < dp; c += 32) {
      const T v = X[r * dp + c];
      s += v * v;
    }
    // x.x for the diagonal in the summation order the persistent kernel uses for x_i.x_j (lane owns the 16-byte
    // chunks lane, lane + 32, ...; fused multiply-add along each chunk; xor-butterfly over the lanes)
    for (int cc = lane; cc * EPC < dp; cc += 32)
      for (int e = 0; e < EPC; ++e) {
        const T v = X[r * dp + cc * EPC + e];
        dd = fma(v, v, dd);
      }
    s = warp_sum(s);
    dd = warp_sum(dd);
    if (lane == 0) {
      norms[r] = s;
      kdiag[r] = kernel_value(kp, dd, s, s);
    }
  }
}

template <typename T>
__global__ void diag_precomputed_kernel(const T* __restrict__ K, int64_t l, T* __restrict__ kdiag) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < l; r += (int64_t)gridDim.x * blockDim.x) kdiag[r] = K[r * l + r];
}

__global__ void refresh_flags_kernel(const double* __restrict__ y, const double* __restrict__ alpha, double C,
                                     uint8_t* __restrict__ flags, int64_t nv) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x)
    flags[v] = make_flags(y[v] > 0, alpha[v], C);
}

// Solver.__init__: alpha = 0, neg_y_grad = -y * p   (solver.py:42-45)
__global__ void init_state_kernel(const double* __restrict__ y, const double* __restrict__ p,
                                  double* __restrict__ alpha, double* __restrict__ G, int64_t nv) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    alpha[v] = 0.0;
    G[v] = -y[v] * p[v];
  }
}

// func(i): Q column as float64, out[v] = y_v * K(x_row(v), x_row(i)) * y_i   (svc.py:257-258)
template <typename T>
__global__ void column_kernel(const T* __restrict__ X, const T* __restrict__ norms, const double* __restrict__ y,
                              int64_t l, int64_t nv, int dp, KernelParams<T> kp, int64_t i,
                              double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t ri = i >= l ? i - l : i;
  const double yi = y[i];
  for (int64_t r = warp; r < l; r += nwarps) {
    T s = 0;
    for (int c = lane; c < dp; c += 32) s = fma(X[r * dp + c], X[ri * dp + c], s);
    s = warp_sum(s);
    if (lane == 0) {
      const double k = (double)kernel_value(kp, s, norms[r], norms[ri]);
      out[r] = y[r] * k * yi;
      if (nv > l) out[r + l] = y[r + l] * k * yi;
    }
  }
}

template <typename T>
__global__ void column_precomputed_kernel(const T* __restrict__ K, const double* __restrict__ y, int64_t l, int64_t nv,
                                          int64_t i, double* __restrict__ out) {
  const int64_t ri = i >= l ? i - l : i;
  const double yi = y[i];
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < l; r += (int64_t)gridDim.x * blockDim.x) {
    const double k = (double)K[r * l + ri];
    out[r] = y[r] * k * yi;
    if (nv > l) out[r + l] = y[r + l] * k * yi;
  }
}

// m = K w for a precomputed kernel matrix (initial gradients of the nu / one-class starts), one warp per row
template <typename T>
__global__ void dense_matvec_kernel(const T* __restrict__ K, const double* __restrict__ w, int64_t l, double* __restrict__ m) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < l; r += nwarps) {
    double s = 0.0;
    for (int64_t c = lane; c < l; c += 32) s += w[c] * (double)K[r * l + c];
    s = warp_sum(s);
    if (lane == 0) m[r] = s;
  }
}

// ------------------------------------------------------------------------------------------
// Opt-in second-order working-set selection (LIBSVM's WSS 3; Fan, Chen, Lin, JMLR 2005, Algorithm 2).  NOT what the
// reference's Solver does (it takes the maximal violating pair, solver.py:47-75); BASELINE.json's north_star names it,
// so it is offered as an opt-in with its own iteration count:
//   i = argmax_{t in I_up} g_t;  j = argmin_{t in I_low, g_t < g_i} -(g_i - g_t)^2 / a_it,
//   a_it = K_ii + K_tt - 2 K_it (<= 0 -> 1e-12); stop when g_i - min_{I_low} g < tol.  Ties go to the lower index.
// Two block-parallel passes per iteration (the second evaluates K(x_i, .), i.e. one extra sweep of X).
// ------------------------------------------------------------------------------------------
struct WssRec {
  double v;
  long long idx;
};
__device__ __forceinline__ bool wss_less(const WssRec& a, const WssRec& b) {   // a strictly better for a MIN search
  return a.idx >= 0 && (b.idx < 0 || a.v < b.v || (a.v == b.v && a.idx < b.idx));
}
__device__ __forceinline__ WssRec wss_block_min(WssRec r, WssRec* sh) {
  for (int m = 16; m >= 1; m >>= 1) {
    WssRec o;
    o.v = __shfl_down_sync(0xffffffffu, r.v, m);
    o.idx = __shfl_down_sync(0xffffffffu, r.idx, m);
    if (wss_less(o, r)) r = o;
  }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = r;
  __syncthreads();
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : WssRec{0.0, -1};
    for (int m = 16; m >= 1; m >>= 1) {
      WssRec o;
      o.v = __shfl_down_sync(0xffffffffu, r.v, m);
      o.idx = __shfl_down_sync(0xffffffffu, r.idx, m);
      if (wss_less(o, r)) r = o;
    }
  }
  return r;   // valid in thread 0
}

// pass 1: per block, arg-max of g over I_up (as a min of -g) and min of g over I_low
__global__ void __launch_bounds__(256) wss3_first_kernel(const double* __restrict__ G, const uint8_t* __restrict__ flags, int64_t nv,
                                                         WssRec* __restrict__ pmax, WssRec* __restrict__ pmin) {
  __shared__ WssRec sh[8];
  WssRec up{0.0, -1}, low{0.0, -1};
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    const int f = flags[v];
    const double g = G[v];
    if (in_up(f)) { WssRec c{-g, v}; if (wss_less(c, up)) up = c; }
    if (in_low(f)) { WssRec c{g, v}; if (wss_less(c, low)) low = c; }
  }
  up = wss_block_min(up, sh);
  if (threadIdx.x == 0) pmax[blockIdx.x] = up;
  __syncthreads();
  low = wss_block_min(low, sh);
  if (threadIdx.x == 0) pmin[blockIdx.x] = low;
}

// out[0] = i (or -1), out[1] = g_i, out[2] = min over I_low of g (as doubles), from the per-block partials
__global__ void wss3_reduce_first_kernel(const WssRec* __restrict__ pmax, const WssRec* __restrict__ pmin, int nb, double* __restrict__ out) {
  __shared__ WssRec sh[8];
  WssRec a{0.0, -1}, b{0.0, -1};
  for (int k = threadIdx.x; k < nb; k += blockDim.x) {
    if (wss_less(pmax[k], a)) a = pmax[k];
    if (wss_less(pmin[k], b)) b = pmin[k];
  }
  a = wss_block_min(a, sh);
  __syncthreads();
  WssRec a0 = a;
  b = wss_block_min(b, sh);
  if (threadIdx.x == 0) {
    out[0] = (double)a0.idx;
    out[1] = -a0.v;
    out[2] = b.idx >= 0 ? b.v : 0.0;
    out[3] = (double)b.idx;
  }
}

// pass 2: one warp per row r: K(x_r, x_i), then the second-order gain of the row's variables; per-block arg-min
template <typename T>
__global__ void __launch_bounds__(256) wss3_second_kernel(const T* __restrict__ X, const T* __restrict__ norms, const T* __restrict__ kdiag,
                                                          const T* __restrict__ Kpre, const double* __restrict__ G,
                                                          const uint8_t* __restrict__ flags, int64_t l, int64_t nv, int dp,
                                                          KernelParams<T> kp, const double* __restrict__ first, WssRec* __restrict__ pgain) {
  __shared__ WssRec sh[8];
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t i = (int64_t)first[0];
  const double gi = first[1];
  WssRec best{0.0, -1};
  if (i >= 0) {
    const int64_t ri = i >= l ? i - l : i;
    const double kii = (double)kdiag[ri];
    for (int64_t r = warp; r < l; r += nwarps) {
      T k;
      if (Kpre) {
        k = Kpre[r * l + ri];
      } else {
        T s = 0;
        for (int c = lane; c < dp; c += 32) s = fma(X[r * dp + c], X[ri * dp + c], s);
        s = warp_sum(s);
        k = kernel_value(kp, s, norms[r], norms[ri]);
      }
      if (lane == 0) {
        double a = (kii + (double)kdiag[r]) - 2.0 * (double)k;
        if (a <= 0) a = 1e-12;
        for (int64_t v = r; v < nv; v += l) {
          const double g = G[v];
          if (in_low((int)flags[v]) && g < gi) {
            const double b = gi - g;
            WssRec c{-(b * b) / a, v};
            if (wss_less(c, best)) best = c;
          }
        }
      }
    }
  }
  best = wss_block_min(best, sh);
  if (threadIdx.x == 0) pgain[blockIdx.x] = best;
}

__global__ void wss3_reduce_second_kernel(const WssRec* __restrict__ pgain, int nb, double* __restrict__ out) {
  __shared__ WssRec sh[8];
  WssRec a{0.0, -1};
  for (int k = threadIdx.x; k < nb; k += blockDim.x)
    if (wss_less(pgain[k], a)) a = pgain[k];
  a = wss_block_min(a, sh);
  if (threadIdx.x == 0) out[4] = (double)a.idx;
}

__global__ void iota_kernel(int* __restrict__ a, int64_t n) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) a[v] = (int)v;
}

// Everything calculate_rho (solver.py:160-177) and calculate_rho_b (solver.py:378-419) need,
// reduced in a fixed order by one block.
struct BiasStats {
  // C variant
  double sum_free, min_lb, max_ub;
  long long cnt_free, cnt_lb, cnt_ub;
  // nu variant, per class (0: y=+1, 1: y=-1); grad = -y * g, bound tests against the literal 1
  double nsum[2], nmax_hi[2], nmin_lo[2];
  long long ncnt[2], nhi[2], nlo[2];
};

__device__ __forceinline__ void merge_stats(BiasStats& a, const BiasStats& b) {
  a.sum_free += b.sum_free;
  a.cnt_free += b.cnt_free;
  if (b.cnt_lb) { a.min_lb = a.cnt_lb ? fmin(a.min_lb, b.min_lb) : b.min_lb; a.cnt_lb += b.cnt_lb; }
  if (b.cnt_ub) { a.max_ub = a.cnt_ub ? fmax(a.max_ub, b.max_ub) : b.max_ub; a.cnt_ub += b.cnt_ub; }
  for (int s = 0; s < 2; ++s) {
    a.nsum[s] += b.nsum[s];
    a.ncnt[s] += b.ncnt[s];
    if (b.nhi[s]) { a.nmax_hi[s] = a.nhi[s] ? fmax(a.nmax_hi[s], b.nmax_hi[s]) : b.nmax_hi[s]; a.nhi[s] += b.nhi[s]; }
    if (b.nlo[s]) { a.nmin_lo[s] = a.nlo[s] ? fmin(a.nmin_lo[s], b.nmin_lo[s]) : b.nmin_lo[s]; a.nlo[s] += b.nlo[s]; }
  }
}

// block-level pieces of the statistics: every block reduces a contiguous chunk in a fixed order (lanes, then warps) ...
__global__ void __launch_bounds__(256) bias_stats_kernel(const double* __restrict__ y, const double* __restrict__ alpha,
                                                         const double* __restrict__ G, double C, int64_t nv, int64_t per_block,
                                                         BiasStats* __restrict__ partial) {
  __shared__ BiasStats sh[8];
  BiasStats st;
  memset(&st, 0, sizeof(st));
  const int64_t v0 = (int64_t)blockIdx.x * per_block, v1 = min(nv, v0 + per_block);
  for (int64_t v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    const double a = alpha[v], g = G[v], yy = y[v];
    BiasStats one;
    memset(&one, 0, sizeof(one));
    if (a > 0 && a < C) { one.sum_free = g; one.cnt_free = 1; }
    const bool ub = (a == 0 && yy < 0) || (a == C && yy > 0);
    const bool lb = (a == 0 && yy > 0) || (a == C && yy < 0);
    if (ub) { one.max_ub = g; one.cnt_ub = 1; }
    if (lb) { one.min_lb = g; one.cnt_lb = 1; }
    const int s = yy == 1.0 ? 0 : (yy == -1.0 ? 1 : -1);
    if (s >= 0) {
      const double grad = -yy * g;
      if (a > 0 && a < 1) { one.nsum[s] = grad; one.ncnt[s] = 1; }
      if (a == 1) { one.nmax_hi[s] = grad; one.nhi[s] = 1; }
      if (a == 0) { one.nmin_lo[s] = grad; one.nlo[s] = 1; }
    }
    merge_stats(st, one);
  }
  for (int m = 16; m >= 1; m >>= 1) {
    BiasStats o;
    const int* src = reinterpret_cast<const int*>(&st);
    int* dst = reinterpret_cast<int*>(&o);
    for (int w = 0; w < (int)(sizeof(BiasStats) / 4); ++w) dst[w] = __shfl_down_sync(0xffffffffu, src[w], m);
    if ((threadIdx.x & 31) + m < 32) merge_stats(st, o);
  }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = st;
  __syncthreads();
  if (threadIdx.x == 0) {
    BiasStats tot = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) merge_stats(tot, sh[w]);
    partial[blockIdx.x] = tot;
  }
}
// ... and one thread merges the block results in block order (a fixed summation tree: the result does not depend on timing)
__global__ void bias_stats_merge_kernel(const BiasStats* __restrict__ partial, int nb, BiasStats* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    BiasStats tot = partial[0];
    for (int k = 1; k < nb; ++k) merge_stats(tot, partial[k]);
    *out = tot;
  }
}

// w[r] = sum over the variables that share row r of y_v * alpha_v  (non-mirror: y_r alpha_r;
// mirror: alpha_r - alpha_{r+l}).  (Q alpha)[t] = y_t * (K w)[row(t)].
__global__ void row_weights_kernel(const double* __restrict__ y, const double* __restrict__ alpha, int64_t l,
                                   int64_t nv, double* __restrict__ w) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < l; r += (int64_t)gridDim.x * blockDim.x) {
    double s = y[r] * alpha[r];
    if (nv > l) s += y[r + l] * alpha[r + l];
    w[r] = s;
  }
}

// G[t] = -y_t * (y_t * m[row(t)] + p_t)     (solver.py:280 / 470-472, oc_svm.py:95)
__global__ void finish_gradient_kernel(const double* __restrict__ y, const double* __restrict__ p,
                                       const double* __restrict__ m, int64_t l, int64_t nv, double* __restrict__ G) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = v >= l ? v - l : v;
    const double q = y[v] * m[r];
    G[v] = -y[v] * (q + (p ? p[v] : 0.0));
  }
}

}  // namespace b2svm

using namespace b2svm;

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
struct b2svm_handle {
  b2svm_problem_desc d;
  cudaStream_t stream = nullptr;
  int ch = 0;           // 16-byte chunks per padded row
  int dp = 0;           // padded feature count
  bool f64 = false;
  void* Xown = nullptr;         // padded copy (nullptr when the caller's X is used in place)
  const void* Xuse = nullptr;
  void* norms = nullptr;
  void* kdiag = nullptr;
  uint8_t* flags = nullptr;
  int* abort_flag = nullptr;
  SolveResult* result = nullptr;
  SolveResult* result_host = nullptr;
  ProblemDev* prob_dev = nullptr;
  BiasStats* bias_dev = nullptr;
  double* scratch = nullptr;    // [2 * n_rows] row weights + mat-vec result
  double* p_dev = nullptr;      // staging for host-side p
  unsigned char* wss_buf = nullptr;   // scratch of b2svm_select_wss3
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int num_sms = 0;
  int ncta = 1, tiles_per_cta = 1, stages = 2, sub = 1, stages1 = 2, tile_bytes = 0;
  size_t smem_bytes1 = 0;
  long long ntiles = 0;
  size_t smem_bytes = 0;
  int launches = 0;
  // ---- exchange buffers (b2svm_smo.cuh): `l1` = private level-1 mailboxes of the CTAs (local); `block` = private
  // level-2 mailboxes + bulk slots of the box (world > 1 only: a plain cudaMalloc, shared over CUDA IPC)
  unsigned char* l1 = nullptr;
  size_t l1_bytes = 0;
  unsigned char* block = nullptr;
  size_t block_bytes = 0;
  long long rows_per_rank = 0;
  std::vector<void*> peer_blocks;          // opened IPC mappings (nullptr for this rank)
  void** peer_tables = nullptr;            // device: [world] exported block of every rank (own entry = block)
  bool attached = false, prepared = false;
  bool precomputed = false;       // kernel.kind == 4: X is the kernel matrix itself and doubles as a full cache
  void* cache = nullptr;          // Q-column cache [cache_slots][n_rows] in the input dtype
  int* slot_of = nullptr;
  int* row_of_slot = nullptr;
  bool cache_pooled = true;
  long long cache_slots = 0, cache_used = 0, cache_hand = 0;
  long long* timeline = nullptr;   // debug: set by B2SVM_TIMELINE=1
  double* dbg_steps = nullptr;     // debug: set by B2SVM_DEBUG_STEPS=1
};

namespace {

// Handle buffers come from the stream-ordered allocator, whose pool keeps freed memory for the next handle: a fit is
// create -> solve -> destroy, and ~15 cudaMalloc / cudaFree pairs (each a device-wide synchronisation, and slow on
// some hosts) used to cost more than the data transfer.  Only the level-2 box of a multi-GPU handle stays a plain
// cudaMalloc: CUDA IPC exports whole cudaMalloc allocations.
template <typename P>
inline cudaError_t pool_alloc(P** p, size_t bytes, cudaStream_t stream) {
  if (getenv("B2SVM_NO_POOL")) return cudaMalloc(reinterpret_cast<void**>(p), bytes ? bytes : 1);
  keep_pool_memory();
  return cudaMallocAsync(reinterpret_cast<void**>(p), bytes ? bytes : 1, stream);
}
inline void pool_free(void* p, cudaStream_t stream) {
  if (!p) return;
  if (getenv("B2SVM_NO_POOL")) cudaFree(p);
  else cudaFreeAsync(p, stream);
}
inline int sm_count(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device] > 0) return cached[device];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) n = 0;
  if (device >= 0 && device < 64) cached[device] = n;
  return n;
}
// pinned landing zone for the small device -> host result copies, one per host thread (handles of different threads
// run concurrently; a handle is not thread-safe)
inline SolveResult* pinned_result() {
  static thread_local SolveResult* buf = nullptr;
  if (!buf && cudaMallocHost(&buf, sizeof(SolveResult) * 16) != cudaSuccess) buf = nullptr;
  return buf;
}

inline int grid_for(int64_t n, int block = 256, int cap = 148 * 8) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

template <typename T, int CH>
struct Geometry {
  using C = Cfg<T, CH>;
  static size_t smem_for(int stages, int sub) {
    return PROBLEM_SMEM_BYTES + (size_t)stages * sub * C::TILE_BYTES + 2 * MAX_STAGES * sizeof(uint64_t) + MAX_STAGES * sizeof(unsigned) +
           NCAND * (size_t)C::ROWB +
           sizeof(Shared<T>) + 128;
  }
};

constexpr int kBiasBlocks = 128;              // partial results of the bias statistics
constexpr size_t kSmemBudget = 227 * 1024;   // opt-in maximum per CTA on sm_100
constexpr size_t kRingBudget = 216 * 1024;   // 13 stages of 16 KB (+ ~6 KB of barriers / state) under the 227 KB cap: 18.30 -> 18.05 us at C3 vs 12 stages

template <typename T, int CH>
int plan_T(b2svm_handle* h) {
  using C = Cfg<T, CH>;
  h->ntiles = (h->d.n_rows + C::BR - 1) / C::BR;
  auto geometry = [](int sub, int& stages, size_t& smem) {
    size_t ring_budget = kRingBudget;
    if (const char* e = getenv("B2SVM_RING_KB")) ring_budget = (size_t)atoi(e) * 1024;   // experiment knob
    stages = (int)std::min<size_t>(MAX_STAGES, ring_budget / ((size_t)sub * C::TILE_BYTES));
    if (stages < 2) stages = 2;
    while (stages > 2 && Geometry<T, CH>::smem_for(stages, sub) > kSmemBudget) --stages;
    smem = Geometry<T, CH>::smem_for(stages, sub);
  };
  geometry(1, h->stages1, h->smem_bytes1);   // batched mode (one CTA per problem) always runs with single tiles
  // every rank of a sharded solve must launch the same number of CTAs (the mailbox is indexed rank * ncta + cta):
  // plan with the common slice size, not with this rank's (possibly shorter) last slice
  const long long plan_tiles = h->d.world > 1 ? (h->rows_per_rank + C::BR - 1) / C::BR : h->ntiles;
  // CTAs per problem: as many as leave every CTA at least five tiles.  (Measured on slices of 4k-25k rows, d = 128:
  // spreading the sweep over more SMs wins 5-13 % over one tile per warp although every CTA then polls more records;
  // `scripts/gpu_min_tiles.py`.)
  int min_tiles = 5;
  if (const char* e = getenv("B2SVM_MIN_TILES_PER_CTA")) min_tiles = std::max(1, atoi(e));
  int max_ctas = h->d.max_ctas > 0 ? h->d.max_ctas : h->num_sms;
  if (const char* e = getenv("B2SVM_MAX_CTAS")) max_ctas = std::max(1, atoi(e));
  max_ctas = std::min(max_ctas, h->num_sms);
  // narrow rows: a 32-row tile is only 2-8 KB, and per-tile costs (ticket, mbarrier round trip, one bulk copy
  // issued by the single producer lane) dominate a long sweep.  Move several tiles per ring stage ("macro tile")
  // once every warp still gets at least two of them per sweep.
  int sub = 1;
  const int max_sub = std::max<int>(1, 16384 / C::TILE_BYTES);
  while (sub * 2 <= max_sub && plan_tiles >= (long long)max_ctas * NW * (sub * 2) * 2) sub *= 2;
  if (const char* e = getenv("B2SVM_SUB")) sub = std::max(1, std::min(max_sub, atoi(e)));
  h->sub = sub;
  h->tile_bytes = C::TILE_BYTES;
  int stages = 2;
  geometry(sub, stages, h->smem_bytes);
  h->stages = stages;
  long long ncta = std::max<long long>(1, std::min<long long>(max_ctas, plan_tiles / min_tiles));
  // prefer a split whose slices fit the ring, so X stays resident in shared memory for the whole solve
  const long long fit = ((plan_tiles + sub - 1) / sub + stages - 1) / stages;
  if (fit <= max_ctas) ncta = std::max(ncta, fit);
  // ... and, when the GPU has room, one tile per consumer warp (no intra-CTA imbalance)
  const long long one_each = (plan_tiles + NW - 1) / NW;
  if (one_each <= max_ctas) ncta = std::max(ncta, one_each);
  if (ncta > MAX_CTAS) ncta = MAX_CTAS;   // the level-1 poll reads at most MAX_CTAS records per type
  h->ncta = (int)ncta;
  h->tiles_per_cta = (int)((plan_tiles + ncta - 1) / ncta);   // informational: the kernel deals tiles out evenly
  return B2SVM_OK;
}

int plan(b2svm_handle* h) {
#define B2_PLAN_CASE(CHV) case CHV: return h->f64 ? plan_T<double, CHV>(h) : plan_T<float, CHV>(h);
  switch (h->ch) {
    B2_PLAN_CASE(4) B2_PLAN_CASE(8) B2_PLAN_CASE(16) B2_PLAN_CASE(32) B2_PLAN_CASE(64) B2_PLAN_CASE(128)
    default: return fail(B2SVM_E_UNSUPPORTED, "unsupported row width");
  }
#undef B2_PLAN_CASE
}

// The 48 instantiations of the persistent kernel are compiled in four translation units
// (b2svm_launch_{f32,f64}_{plain,generic}.cu) so the build parallelises.
int launch(bool f64, int ch, bool cache, bool generic, const ProblemDev* probs_dev, int nprob, int ncta_pp, int stages,
           size_t smem, cudaStream_t stream, bool coop = false) {
  if (f64) return generic ? launch_f64_generic(ch, cache, probs_dev, nprob, ncta_pp, stages, smem, stream, coop)
                          : launch_f64_plain(ch, cache, probs_dev, nprob, ncta_pp, stages, smem, stream, coop);
  return generic ? launch_f32_generic(ch, cache, probs_dev, nprob, ncta_pp, stages, smem, stream, coop)
                 : launch_f32_plain(ch, cache, probs_dev, nprob, ncta_pp, stages, smem, stream, coop);
}

inline bool needs_generic(const b2svm_handle* h) {
  return h->d.variant == B2SVM_SOLVER_NU || h->d.n_vars > h->d.n_rows;
}

void fill_problem(const b2svm_handle* h, ProblemDev& P, long long max_iter, long long fi, long long fj) {
  P.X = h->Xuse;
  P.norms = h->norms;
  P.kdiag = h->kdiag;
  P.alpha = h->d.alpha;
  P.G = h->d.neg_y_grad;
  P.flags = h->flags;
  P.l = h->d.n_rows;
  P.nv = h->d.n_vars;
  P.mirror = h->d.n_vars > h->d.n_rows ? 1 : 0;
  P.nu = h->d.variant == B2SVM_SOLVER_NU ? 1 : 0;
  P.C = h->d.C;
  P.tol = h->d.tol;
  P.kind = h->d.kernel.kind;
  P.gamma = h->d.kernel.gamma;
  P.coef0 = h->d.kernel.coef0;
  P.degree = h->d.kernel.degree;
  P.ncta = h->ncta;
  P.tiles_per_cta = h->tiles_per_cta;
  P.sub = h->sub;
  P.ntiles = h->ntiles;
  // L2 residency plan.  Every sweep re-reads the same X; the first 48 MB of it (the same leading macro tiles of every
  // CTA's slice) are fetched with an evict_last hint and stay in the 126 MB L2 across sweeps, the rest streams through
  // with evict_first.  Measured at C3: DRAM reads 83 -> 40 MB / iteration, 19.9 -> 18.7 us (flat between 40 and 72 MB).
  // B2SVM_L2_PIN_MB overrides the size (0 = no hints), B2SVM_L2_PIN_SKIP the first pinned macro tile.
  {
    double mb = 48.0;
    if (const char* e = getenv("B2SVM_L2_PIN_MB")) mb = atof(e);
    const long long per_cta = (long long)(mb * 1048576.0 / ((double)h->ncta * h->tile_bytes * h->sub));
    const int skip = getenv("B2SVM_L2_PIN_SKIP") ? atoi(getenv("B2SVM_L2_PIN_SKIP")) : 0;
    P.pin_lo = skip;
    P.pin_hi = skip + (int)std::min<long long>(per_cta, 1 << 30);
    // a slice that fits L2 comfortably stays resident on its own: hints would only push its tail out
    const double slice_mb = (double)h->ntiles * h->tile_bytes / 1048576.0;
    if (mb <= 0 || (slice_mb <= 64.0 && !getenv("B2SVM_L2_PIN_MB"))) P.pin_lo = P.pin_hi = 0;
  }
  P.abort_flag = h->abort_flag;
  P.result = h->result;
  P.max_iter = max_iter;
  P.forced_i = fi;
  P.forced_j = fj;
  P.cache = h->cache;
  P.slot_of = h->slot_of;
  P.cache_slots = h->cache_slots;
  P.cache_used = h->cache_used;
  P.cache_hand = h->cache_hand;
  P.row_of_slot = h->row_of_slot;
  P.world = h->d.world > 1 ? h->d.world : 1;
  P.rank = h->d.world > 1 ? h->d.rank : 0;
  P.row_offset = h->d.world > 1 ? h->d.row_offset : 0;
  P.l_global = h->d.world > 1 ? h->d.n_rows_global : h->d.n_rows;
  P.rows_per_rank = h->d.world > 1 ? h->rows_per_rank : h->d.n_rows;
  P.mbox = reinterpret_cast<uint4*>(h->l1);
  P.box = reinterpret_cast<uint4*>(h->block);
  P.peer_box = reinterpret_cast<uint4* const*>(h->peer_tables);
  P.timeline = h->timeline;
  P.dbg_steps = h->dbg_steps;
  P.debug_flags = getenv("B2SVM_DEBUG_FLAGS") ? atoi(getenv("B2SVM_DEBUG_FLAGS")) : 0;
  auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
  P.poll_delay = env_int("B2SVM_POLL_DELAY", 500);
  P.poll_gap = env_int("B2SVM_POLL_GAP", 0);
  P.poll_delay2 = env_int("B2SVM_POLL_DELAY2", 0);
  P.poll_gap2 = env_int("B2SVM_POLL_GAP2", 0);
}

// one persistent launch for one handle; fills *res from the device result
// everything that must be in place before the persistent kernel starts; with several GPUs the host
// puts a barrier over all ranks between this and the launch (a peer must not write into a mailbox
// that is still being cleared, nor read a published alpha that is still being copied)
int prepare_launch(b2svm_handle* h) {
  B2_CUDA(cudaSetDevice(h->d.device));
  const int64_t nv = h->d.n_vars;
  refresh_flags_kernel<<<grid_for(nv), 256, 0, h->stream>>>(h->d.y, h->d.alpha, h->d.C, h->flags, nv);
  B2_CUDA(cudaGetLastError());
  B2_CUDA(cudaMemsetAsync(h->l1, 0, h->l1_bytes, h->stream));
  if (h->block) B2_CUDA(cudaMemsetAsync(h->block, 0, h->block_bytes, h->stream));
  B2_CUDA(cudaMemsetAsync(h->abort_flag, 0, sizeof(int), h->stream));
  if (h->d.world > 1) B2_CUDA(cudaStreamSynchronize(h->stream));
  h->prepared = true;
  return B2SVM_OK;
}

int run_solver(b2svm_handle* h, long long max_iter, long long fi, long long fj, SolveResult* res, float* ms) {
  B2_CUDA(cudaSetDevice(h->d.device));
  if (h->d.world > 1) {
    if (!h->attached) return fail(B2SVM_E_INVALID, "multi-GPU handle: call b2svm_mailbox_attach first");
    if (!h->prepared) return fail(B2SVM_E_INVALID, "multi-GPU handle: call b2svm_prepare on every rank, then a barrier, before each solve");
    if (fi >= 0) return fail(B2SVM_E_UNSUPPORTED, "step-wise update is single-GPU only");
  } else {
    int rc0 = prepare_launch(h);
    if (rc0) return rc0;
  }
  h->prepared = false;
  ProblemDev P;
  memset(&P, 0, sizeof(P));
  fill_problem(h, P, max_iter, fi, fj);
  B2_CUDA(cudaMemcpyAsync(h->prob_dev, &P, sizeof(P), cudaMemcpyHostToDevice, h->stream));
  B2_CUDA(cudaEventRecord(h->ev0, h->stream));
  int rc = launch(h->f64, h->ch, h->cache_slots > 0, needs_generic(h), h->prob_dev, 1, h->ncta, h->stages, h->smem_bytes,
                  h->stream);
  if (rc != B2SVM_OK) return rc;
  B2_CUDA(cudaEventRecord(h->ev1, h->stream));
  B2_CUDA(cudaMemcpyAsync(h->result_host, h->result, sizeof(SolveResult), cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  h->launches = 2;
  *res = *h->result_host;
  h->cache_used = res->cache_used;
  h->cache_hand = res->cache_hand;
  if (ms) B2_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  if (res->status != 0) return fail(B2SVM_E_TIMEOUT, "device watchdog: a grid barrier did not complete");
  return B2SVM_OK;
}

}  // namespace

static void fill_stats(const b2svm_handle* h, const SolveResult& r, int n_ctas, float ms, int launches, b2svm_solve_stats* stats) {
  memset(stats, 0, sizeof(*stats));
  stats->n_iter = r.n_iter;
  stats->converged = r.converged;
  stats->n_ctas = n_ctas;
  stats->last_i = r.last_i;
  stats->last_j = r.last_j;
  stats->last_delta_i = r.di;
  stats->last_delta_j = r.dj;
  stats->cold_sweeps = r.cold_sweeps;
  stats->cache_hits = 2 * r.warm_sweeps;                 // both columns read from the cache
  stats->cache_misses = 2 * r.cold_sweeps;               // cold sweep: both columns recomputed in one pass over X
  const double esz = h->f64 ? 8.0 : 4.0;
  // DESIGN.md section 4 / SURVEY 8(d): B_cold = l d sizeof(T) + n_vars 18 (+ sizeof(T) l per stored column);
  // B_warm = n_vars 18 + 2 sizeof(T) l
  stats->algorithmic_bytes =
      (double)r.cold_sweeps * ((double)h->d.n_rows * h->d.n_features * esz + (double)h->d.n_vars * 18.0) +
      (double)r.cols_stored * (double)h->d.n_rows * esz +
      (double)r.warm_sweeps * ((double)h->d.n_vars * 18.0 + 2.0 * (double)h->d.n_rows * esz);
  stats->device_ms = ms;
  stats->launches = launches;
  stats->cache_stores = r.cols_stored;
  stats->cache_evictions = r.cols_evicted;
}

// ------------------------------------------------------------------------------------------
// exported functions
// ------------------------------------------------------------------------------------------
extern "C" {

int b2svm_abi_version(void) { return B2SVM_ABI_VERSION; }
const char* b2svm_last_error(void) { return g_last_error.c_str(); }

int b2svm_create(const b2svm_problem_desc* desc, b2svm_handle** out) {
  if (!desc || !out) return fail(B2SVM_E_INVALID, "null argument");
  if (desc->struct_bytes != sizeof(b2svm_problem_desc))
    return fail(B2SVM_E_INVALID, "b2svm_problem_desc size mismatch (ABI version?)");
  if (desc->n_rows <= 0 || desc->n_features <= 0) return fail(B2SVM_E_INVALID, "empty problem");
  if (desc->n_vars != desc->n_rows && desc->n_vars != 2 * desc->n_rows)
    return fail(B2SVM_E_INVALID, "n_vars must be n_rows or 2*n_rows");
  if (desc->n_vars >= (1ll << 29)) return fail(B2SVM_E_UNSUPPORTED, "n_vars must be below 2^29");
  if (!desc->X || !desc->y || !desc->alpha || !desc->neg_y_grad) return fail(B2SVM_E_INVALID, "null device pointer");
  if (desc->x_dtype != B2SVM_F32 && desc->x_dtype != B2SVM_F64) return fail(B2SVM_E_INVALID, "bad x_dtype");
  if (desc->kernel.kind < 0 || desc->kernel.kind > 4) return fail(B2SVM_E_INVALID, "bad kernel kind");
  const bool precomputed = desc->kernel.kind == 4;
  if (precomputed && (desc->world > 1 || desc->n_features != desc->n_rows || desc->ldx != desc->n_rows))
    return fail(B2SVM_E_INVALID, "precomputed kernel: X must be the dense n_rows x n_rows kernel matrix (single GPU)");
  if (desc->world > 1) {
    if (desc->world > 8) return fail(B2SVM_E_UNSUPPORTED, "at most 8 ranks (one NVSwitch box)");
    if (desc->rank < 0 || desc->rank >= desc->world) return fail(B2SVM_E_INVALID, "bad rank");
  }
  const int ch = precomputed ? 4 : chunks_for(desc->x_dtype, desc->n_features);
  if (ch == 0) return fail(B2SVM_E_UNSUPPORTED, "n_features above the compiled maximum (512 f32 / 256 f64)");

  B2_CUDA(cudaSetDevice(desc->device));
  b2svm_handle* h = new b2svm_handle();
  h->d = *desc;
  h->stream = (cudaStream_t)desc->stream;
  h->ch = ch;
  h->f64 = desc->x_dtype == B2SVM_F64;
  h->precomputed = precomputed;
  const int esz = h->f64 ? 8 : 4;
  h->dp = ch * 16 / esz;
  h->num_sms = sm_count(desc->device);
  if (h->num_sms <= 0) { delete h; return fail(B2SVM_E_CUDA, "cannot query the SM count"); }
  const int64_t l = desc->n_rows, nv = desc->n_vars;

  auto cleanup_fail = [&](int code, const std::string& msg) {
    b2svm_destroy(h);
    return fail(code, msg);
  };
#define B2_TRY(expr)                                                                              \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) return cleanup_fail(B2SVM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

  if (desc->world > 1) {
    // equal slices of whole tiles: rank r owns global rows [r * rpr, min(n, (r + 1) * rpr))
    const long long rpr = (((desc->n_rows_global + desc->world - 1) / desc->world) + 31) / 32 * 32;
    h->rows_per_rank = rpr;
    const long long want = std::min<long long>(rpr, desc->n_rows_global - (long long)desc->rank * rpr);
    if (desc->row_offset != (long long)desc->rank * rpr || desc->n_rows != want)
      return cleanup_fail(B2SVM_E_INVALID, "shard must be rows [rank*rpr, min(n,(rank+1)*rpr)), rpr = roundup32(ceil(n/world))");
  }
  // rows of X held by this GPU: its own slice only, also in a row-sharded solve (the winning rows of an iteration
  // travel with their candidate records)
  const int64_t lx = l;
  const bool in_place = precomputed || (desc->ldx == h->dp && desc->n_features == h->dp && ((uintptr_t)desc->X % 16 == 0));
  if (in_place) {
    h->Xuse = desc->X;
  } else {
    B2_TRY(pool_alloc(&h->Xown, (size_t)lx * h->dp * esz, h->stream));
    void* xdst = h->Xown;
    if (h->f64)
      pad_rows_kernel<double><<<grid_for(lx * h->dp), 256, 0, h->stream>>>((const double*)desc->X, desc->ldx,
                                                                           (double*)xdst, lx, desc->n_features, h->dp);
    else
      pad_rows_kernel<float><<<grid_for(lx * h->dp), 256, 0, h->stream>>>((const float*)desc->X, desc->ldx,
                                                                          (float*)xdst, lx, desc->n_features, h->dp);
    B2_TRY(cudaGetLastError());
    h->Xuse = xdst;
  }
  B2_TRY(pool_alloc(&h->norms, (size_t)lx * esz, h->stream));
  B2_TRY(pool_alloc(&h->kdiag, (size_t)lx * esz, h->stream));
  if (precomputed) {
    B2_TRY(cudaMemsetAsync(h->norms, 0, (size_t)lx * esz, h->stream));
    if (h->f64) diag_precomputed_kernel<double><<<grid_for(lx), 256, 0, h->stream>>>((const double*)h->Xuse, lx, (double*)h->kdiag);
    else diag_precomputed_kernel<float><<<grid_for(lx), 256, 0, h->stream>>>((const float*)h->Xuse, lx, (float*)h->kdiag);
  } else if (h->f64) {
    KernelParams<double> kp{desc->kernel.kind, desc->kernel.gamma, desc->kernel.coef0, desc->kernel.degree};
    row_norms_kernel<double><<<grid_for(lx * 32), 256, 0, h->stream>>>((const double*)h->Xuse, lx, h->dp, (double*)h->norms, (double*)h->kdiag, kp);
  } else {
    KernelParams<float> kp{desc->kernel.kind, (float)desc->kernel.gamma, (float)desc->kernel.coef0, (float)desc->kernel.degree};
    row_norms_kernel<float><<<grid_for(lx * 32), 256, 0, h->stream>>>((const float*)h->Xuse, lx, h->dp, (float*)h->norms, (float*)h->kdiag, kp);
  }
  B2_TRY(cudaGetLastError());
  B2_TRY(pool_alloc(&h->flags, (size_t)nv, h->stream));
  if (precomputed) {
    // the kernel matrix IS a completely filled column cache: every iteration takes the warm path
    h->cache = const_cast<void*>(desc->X);
    B2_TRY(pool_alloc(&h->slot_of, sizeof(int) * (size_t)l, h->stream));
    iota_kernel<<<grid_for(l), 256, 0, h->stream>>>(h->slot_of, l);
    B2_TRY(cudaGetLastError());
    h->cache_slots = l;
    h->cache_used = l;
  } else if (desc->cache_bytes > 0) {
    // device-resident Q-column cache: as many whole columns (this GPU's rows) as the budget holds
    const size_t col_bytes = (size_t)l * esz;
    // the same number of slots on every rank of a sharded solve (slices may differ by a few rows): the replicated
    // slot tables then take the same store / evict decisions everywhere
    const size_t plan_col_bytes = (size_t)(desc->world > 1 ? h->rows_per_rank : l) * esz;
    long long slots = (long long)(desc->cache_bytes / plan_col_bytes);
    const long long lglob = desc->world > 1 ? desc->n_rows_global : l;
    slots = std::min<long long>(slots, lglob);
    if (slots > 0) {
      // (a multi-GB cache stays out of the pool: freed into it, the pool would trim back to its release threshold at the
      // next synchronisation -- hundreds of ms of unmapping inside somebody else's timed region)
      h->cache_pooled = (size_t)slots * col_bytes <= ((size_t)1 << 30);
      if (h->cache_pooled) B2_TRY(pool_alloc(&h->cache, (size_t)slots * col_bytes, h->stream));
      else B2_TRY(cudaMalloc(&h->cache, (size_t)slots * col_bytes));
      B2_TRY(pool_alloc(&h->slot_of, sizeof(int) * (size_t)lglob, h->stream));
      B2_TRY(cudaMemsetAsync(h->slot_of, 0xff, sizeof(int) * (size_t)lglob, h->stream));
      B2_TRY(pool_alloc(&h->row_of_slot, sizeof(int) * (size_t)slots, h->stream));
      B2_TRY(cudaMemsetAsync(h->row_of_slot, 0xff, sizeof(int) * (size_t)slots, h->stream));
      h->cache_slots = slots;
    }
  }
  B2_TRY(pool_alloc(&h->abort_flag, sizeof(int), h->stream));
  B2_TRY(pool_alloc(&h->result, sizeof(SolveResult), h->stream));
  h->result_host = pinned_result();
  if (!h->result_host) return cleanup_fail(B2SVM_E_CUDA, "cudaMallocHost failed");
  B2_TRY(pool_alloc(&h->prob_dev, sizeof(ProblemDev), h->stream));
  B2_TRY(pool_alloc(&h->bias_dev, sizeof(BiasStats) * (kBiasBlocks + 1), h->stream));
  B2_TRY(cudaEventCreate(&h->ev0));
  B2_TRY(cudaEventCreate(&h->ev1));
  if (getenv("B2SVM_DEBUG_STEPS")) {
    B2_TRY(cudaMalloc(&h->dbg_steps, sizeof(double) * DBG_EPOCHS * (h->num_sms * DBG_W + 2)));
    B2_TRY(cudaMemset(h->dbg_steps, 0, sizeof(double) * DBG_EPOCHS * (h->num_sms * DBG_W + 2)));
  }
  if (getenv("B2SVM_TIMELINE")) {
    B2_TRY(cudaMalloc(&h->timeline, sizeof(long long) * TL_TOTAL));
    B2_TRY(cudaMemset(h->timeline, 0, sizeof(long long) * TL_TOTAL));
  }
  int rc = plan(h);
  if (rc != B2SVM_OK) {
    b2svm_destroy(h);
    return rc;
  }
  // exchange buffers, sized by the number of CTAs the plan chose (the same on every rank of a sharded solve)
  h->l1_bytes = sizeof(uint4) * 2 * l1_units(h->ncta);
  B2_TRY(pool_alloc(&h->l1, h->l1_bytes, h->stream));
  B2_TRY(cudaMemsetAsync(h->l1, 0, h->l1_bytes, h->stream));
  if (desc->world > 1) {
    h->block_bytes = sizeof(uint4) * (l2_record_units(h->ncta) + BOX_BULK_UNITS);
    B2_TRY(cudaMalloc(&h->block, h->block_bytes));
    B2_TRY(cudaMemsetAsync(h->block, 0, h->block_bytes, h->stream));
  }
#undef B2_TRY
  *out = h;
  return B2SVM_OK;
}

int b2svm_destroy(b2svm_handle* h) {
  if (!h) return B2SVM_OK;
  cudaSetDevice(h->d.device);
  if (h->stream) cudaStreamSynchronize(h->stream); else cudaDeviceSynchronize();
  pool_free(h->Xown, h->stream);
  pool_free(h->norms, h->stream);
  pool_free(h->kdiag, h->stream);
  for (void* pb : h->peer_blocks)
    if (pb) cudaIpcCloseMemHandle(pb);
  cudaFree(h->peer_tables);
  cudaFree(h->block);
  pool_free(h->l1, h->stream);
  pool_free(h->flags, h->stream);
  if (!h->precomputed) { if (h->cache_pooled) pool_free(h->cache, h->stream); else cudaFree(h->cache); }
  pool_free(h->slot_of, h->stream);
  pool_free(h->row_of_slot, h->stream);
  pool_free(h->abort_flag, h->stream);
  pool_free(h->result, h->stream);
  pool_free(h->prob_dev, h->stream);
  pool_free(h->bias_dev, h->stream);
  pool_free(h->scratch, h->stream);
  pool_free(h->p_dev, h->stream);
  pool_free(h->wss_buf, h->stream);
  cudaFree(h->timeline);
  cudaFree(h->dbg_steps);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
  return B2SVM_OK;
}

static int stage_p(b2svm_handle* h, const double* p, const double** p_use) {
  // p may live on the host or the device: stage through a device buffer (UVA-aware copy)
  if (!p) { *p_use = nullptr; return B2SVM_OK; }
  if (!h->p_dev) B2_CUDA(pool_alloc(&h->p_dev, sizeof(double) * h->d.n_vars, h->stream));
  B2_CUDA(cudaMemcpyAsync(h->p_dev, p, sizeof(double) * h->d.n_vars, cudaMemcpyDefault, h->stream));
  *p_use = h->p_dev;
  return B2SVM_OK;
}

int b2svm_init_state(b2svm_handle* h, const double* p) {
  if (!h || !p) return fail(B2SVM_E_INVALID, "null argument");
  B2_CUDA(cudaSetDevice(h->d.device));
  const double* pu;
  int rc = stage_p(h, p, &pu);
  if (rc) return rc;
  init_state_kernel<<<grid_for(h->d.n_vars), 256, 0, h->stream>>>(h->d.y, pu, h->d.alpha, h->d.neg_y_grad, h->d.n_vars);
  B2_CUDA(cudaGetLastError());
  return B2SVM_OK;
}

int b2svm_init_gradient(b2svm_handle* h, const double* p) {
  if (!h) return fail(B2SVM_E_INVALID, "null handle");
  if (h->d.world > 1) return fail(B2SVM_E_UNSUPPORTED, "init_gradient is single-GPU only (a sharded handle holds its own rows of X; see sharded.rows_kernel_sum)");
  B2_CUDA(cudaSetDevice(h->d.device));
  const int64_t l = h->d.n_rows, nv = h->d.n_vars;
  const double* pu;
  int rc = stage_p(h, p, &pu);
  if (rc) return rc;
  if (!h->scratch) B2_CUDA(pool_alloc(&h->scratch, sizeof(double) * 2 * l, h->stream));
  double* w = h->scratch;
  double* m = h->scratch + l;
  row_weights_kernel<<<grid_for(l), 256, 0, h->stream>>>(h->d.y, h->d.alpha, l, nv, w);
  B2_CUDA(cudaGetLastError());
  if (h->precomputed) {
    if (h->f64) dense_matvec_kernel<double><<<grid_for(l * 32), 256, 0, h->stream>>>((const double*)h->Xuse, w, l, m);
    else dense_matvec_kernel<float><<<grid_for(l * 32), 256, 0, h->stream>>>((const float*)h->Xuse, w, l, m);
    B2_CUDA(cudaGetLastError());
  } else {
    rc = kernel_matvec_impl(h->stream, h->d.kernel, h->d.x_dtype, h->dp, h->Xuse, l, h->dp, h->Xuse, l, h->dp, w, m);
    if (rc) return rc;
  }
  finish_gradient_kernel<<<grid_for(nv), 256, 0, h->stream>>>(h->d.y, pu, m, l, nv, h->d.neg_y_grad);
  B2_CUDA(cudaGetLastError());
  return B2SVM_OK;
}

int b2svm_select(b2svm_handle* h, int64_t* i, int64_t* j) {
  if (!h || !i || !j) return fail(B2SVM_E_INVALID, "null argument");
  SolveResult r;
  int rc = run_solver(h, 0, -1, -1, &r, nullptr);
  if (rc) return rc;
  *i = r.last_i;
  *j = r.last_j;
  return B2SVM_OK;
}

int b2svm_update(b2svm_handle* h, int64_t i, int64_t j, double* delta_i, double* delta_j) {
  if (!h) return fail(B2SVM_E_INVALID, "null handle");
  if (i < 0 || j < 0 || i >= h->d.n_vars || j >= h->d.n_vars) return fail(B2SVM_E_INVALID, "index out of range");
  SolveResult r;
  int rc = run_solver(h, 1, i, j, &r, nullptr);
  if (rc) return rc;
  if (delta_i) *delta_i = r.di;
  if (delta_j) *delta_j = r.dj;
  return B2SVM_OK;
}

int b2svm_select_wss3(b2svm_handle* h, int64_t* i, int64_t* j) {
  if (!h || !i || !j) return fail(B2SVM_E_INVALID, "null argument");
  if (h->d.world > 1) return fail(B2SVM_E_UNSUPPORTED, "second-order selection is single-GPU only");
  if (h->d.variant != B2SVM_SOLVER_C) return fail(B2SVM_E_UNSUPPORTED, "second-order selection is offered for the C variant (Solver) only");
  B2_CUDA(cudaSetDevice(h->d.device));
  const int64_t l = h->d.n_rows, nv = h->d.n_vars;
  const int nb = (int)std::max<int64_t>(1, std::min<int64_t>(148 * 4, (l * 32 + 255) / 256));
  // scratch: [nb] max partials, [nb] min partials, [nb] gain partials, 8 doubles of results
  const size_t need = sizeof(WssRec) * 3 * (size_t)nb + 64;
  if (!h->wss_buf) B2_CUDA(pool_alloc(&h->wss_buf, sizeof(WssRec) * 3 * 148 * 4 + 64, h->stream));
  (void)need;
  WssRec* pmax = reinterpret_cast<WssRec*>(h->wss_buf);
  WssRec* pmin = pmax + nb;
  WssRec* pgain = pmin + nb;
  double* res = reinterpret_cast<double*>(pgain + nb);
  refresh_flags_kernel<<<grid_for(nv), 256, 0, h->stream>>>(h->d.y, h->d.alpha, h->d.C, h->flags, nv);
  wss3_first_kernel<<<nb, 256, 0, h->stream>>>(h->d.neg_y_grad, h->flags, nv, pmax, pmin);
  wss3_reduce_first_kernel<<<1, 256, 0, h->stream>>>(pmax, pmin, nb, res);
  if (h->f64) {
    KernelParams<double> kp{h->d.kernel.kind, h->d.kernel.gamma, h->d.kernel.coef0, h->d.kernel.degree};
    wss3_second_kernel<double><<<nb, 256, 0, h->stream>>>((const double*)h->Xuse, (const double*)h->norms, (const double*)h->kdiag,
                                                         h->precomputed ? (const double*)h->Xuse : nullptr, h->d.neg_y_grad, h->flags, l, nv,
                                                         h->dp, kp, res, pgain);
  } else {
    KernelParams<float> kp{h->d.kernel.kind, (float)h->d.kernel.gamma, (float)h->d.kernel.coef0, (float)h->d.kernel.degree};
    wss3_second_kernel<float><<<nb, 256, 0, h->stream>>>((const float*)h->Xuse, (const float*)h->norms, (const float*)h->kdiag,
                                                        h->precomputed ? (const float*)h->Xuse : nullptr, h->d.neg_y_grad, h->flags, l, nv,
                                                        h->dp, kp, res, pgain);
  }
  wss3_reduce_second_kernel<<<1, 256, 0, h->stream>>>(pgain, nb, res);
  B2_CUDA(cudaGetLastError());
  double host[5];
  B2_CUDA(cudaMemcpyAsync(host, res, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  const int64_t si = (int64_t)host[0], low_any = (int64_t)host[3], sj = (int64_t)host[4];
  // stopping rule of the reference (solver.py:72-75) with the same two extremes: m(alpha) - M(alpha) < tol
  if (si < 0 || low_any < 0 || host[1] - host[2] < h->d.tol || sj < 0) { *i = -1; *j = -1; }
  else { *i = si; *j = sj; }
  return B2SVM_OK;
}

int b2svm_kernel_column(b2svm_handle* h, int64_t i, double* out) {
  if (!h || !out) return fail(B2SVM_E_INVALID, "null argument");
  if (i < 0 || i >= h->d.n_vars) return fail(B2SVM_E_INVALID, "index out of range");
  if (h->d.world > 1) return fail(B2SVM_E_UNSUPPORTED, "kernel_column is single-GPU only (a sharded handle holds its own rows of X)");
  B2_CUDA(cudaSetDevice(h->d.device));
  const int64_t l = h->d.n_rows;
  if (h->precomputed) {
    if (h->f64) column_precomputed_kernel<double><<<grid_for(l), 256, 0, h->stream>>>((const double*)h->Xuse, h->d.y, l, h->d.n_vars, i, out);
    else column_precomputed_kernel<float><<<grid_for(l), 256, 0, h->stream>>>((const float*)h->Xuse, h->d.y, l, h->d.n_vars, i, out);
    B2_CUDA(cudaGetLastError());
    return B2SVM_OK;
  }
  if (h->f64) {
    KernelParams<double> kp{h->d.kernel.kind, h->d.kernel.gamma, h->d.kernel.coef0, h->d.kernel.degree};
    column_kernel<double><<<grid_for(l * 32), 256, 0, h->stream>>>((const double*)h->Xuse, (const double*)h->norms,
                                                                   h->d.y, l, h->d.n_vars, h->dp, kp, i, out);
  } else {
    KernelParams<float> kp{h->d.kernel.kind, (float)h->d.kernel.gamma, (float)h->d.kernel.coef0,
                           (float)h->d.kernel.degree};
    column_kernel<float><<<grid_for(l * 32), 256, 0, h->stream>>>((const float*)h->Xuse, (const float*)h->norms,
                                                                  h->d.y, l, h->d.n_vars, h->dp, kp, i, out);
  }
  B2_CUDA(cudaGetLastError());
  return B2SVM_OK;
}

int b2svm_solve(b2svm_handle* h, int64_t max_iter, b2svm_solve_stats* stats) {
  if (!h) return fail(B2SVM_E_INVALID, "null handle");
  if (max_iter < 0) return fail(B2SVM_E_INVALID, "max_iter < 0");
  if (max_iter >= B2SVM_MAX_ITER) return fail(B2SVM_E_UNSUPPORTED, "max_iter must be below 2^29 per call (mailbox records carry a 29-bit epoch tag)");
  SolveResult r;
  float ms = 0.f;
  int rc = run_solver(h, max_iter, -1, -1, &r, &ms);
  if (rc) return rc;
  if (stats) fill_stats(h, r, h->ncta, ms, h->launches, stats);
  return B2SVM_OK;
}

/* The ranks of a row-sharded solve emulated on ONE GPU: `world` handles (created with rank = 0..world-1 of `world`,
 * all on this device, max_ctas <= num_sms / world) run as the CTA groups of one cooperative launch, their level-2
 * boxes attached to each other in local memory instead of over CUDA IPC.  Kernels that wait on one another must never
 * be separate launches on one GPU; this entry point is how the two-level exchange protocol is tested on a 1-GPU box. */
int b2svm_solve_ranks_on_one_gpu(b2svm_handle* const* handles, int32_t world, int64_t max_iter, b2svm_solve_stats* stats) {
  if (!handles || world < 2 || world > BOX_RANKS) return fail(B2SVM_E_INVALID, "need 2..8 rank handles");
  if (max_iter < 0 || max_iter >= B2SVM_MAX_ITER) return fail(B2SVM_E_UNSUPPORTED, "max_iter must be in [0, 2^29) per call");
  b2svm_handle* h0 = handles[0];
  if (!h0) return fail(B2SVM_E_INVALID, "null handle");
  B2_CUDA(cudaSetDevice(h0->d.device));
  for (int r = 0; r < world; ++r) {
    b2svm_handle* h = handles[r];
    if (!h) return fail(B2SVM_E_INVALID, "null handle");
    if (h->d.world != world || h->d.rank != r) return fail(B2SVM_E_INVALID, "handle r must be rank r of `world`");
    if (h->f64 != h0->f64 || h->ch != h0->ch || h->d.device != h0->d.device || h->stream != h0->stream ||
        h->ncta != h0->ncta || h->stages != h0->stages || h->sub != h0->sub || (h->cache_slots > 0) != (h0->cache_slots > 0) ||
        needs_generic(h) != needs_generic(h0))
      return fail(B2SVM_E_INVALID, "emulated ranks must share dtype, row width, device, stream and launch geometry");
  }
  if ((long long)world * h0->ncta > h0->num_sms)
    return fail(B2SVM_E_INVALID, "world * n_ctas exceeds the SM count: create the handles with max_ctas <= num_sms / world");
  std::vector<void*> tab((size_t)world);
  for (int r = 0; r < world; ++r) tab[r] = handles[r]->block;
  std::vector<ProblemDev> P((size_t)world);
  ProblemDev* pd = nullptr;
  B2_CUDA(pool_alloc(&pd, sizeof(ProblemDev) * world, h0->stream));
  int rc = B2SVM_OK;
  for (int r = 0; r < world && rc == B2SVM_OK; ++r) {
    b2svm_handle* h = handles[r];
    if (!h->peer_tables) {
      cudaError_t e = cudaMalloc(&h->peer_tables, sizeof(void*) * world);
      if (e != cudaSuccess) { rc = fail(B2SVM_E_CUDA, cudaGetErrorString(e)); break; }
    }
    cudaMemcpyAsync(h->peer_tables, tab.data(), sizeof(void*) * world, cudaMemcpyHostToDevice, h0->stream);
    h->attached = true;
    rc = prepare_launch(h);
    h->prepared = false;
    memset(&P[r], 0, sizeof(ProblemDev));
    fill_problem(h, P[r], max_iter, -1, -1);
  }
  if (rc == B2SVM_OK) {
    cudaMemcpyAsync(pd, P.data(), sizeof(ProblemDev) * world, cudaMemcpyHostToDevice, h0->stream);
    cudaEventRecord(h0->ev0, h0->stream);
    rc = launch(h0->f64, h0->ch, h0->cache_slots > 0, needs_generic(h0), pd, world, h0->ncta, h0->stages, h0->smem_bytes,
                h0->stream, true);
  }
  if (rc == B2SVM_OK) {
    cudaEventRecord(h0->ev1, h0->stream);
    for (int r = 0; r < world; ++r)
      cudaMemcpyAsync(pinned_result() + r, handles[r]->result, sizeof(SolveResult), cudaMemcpyDeviceToHost, h0->stream);
    cudaError_t e = cudaStreamSynchronize(h0->stream);
    if (e != cudaSuccess) rc = fail(B2SVM_E_CUDA, cudaGetErrorString(e));
  }
  if (rc == B2SVM_OK) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h0->ev0, h0->ev1);
    for (int r = 0; r < world; ++r) {
      const SolveResult& R = pinned_result()[r];
      handles[r]->cache_used = R.cache_used;
      handles[r]->cache_hand = R.cache_hand;
      if (R.status != 0) rc = fail(B2SVM_E_TIMEOUT, "device watchdog tripped in an emulated sharded solve");
      if (stats) fill_stats(handles[r], R, h0->ncta, ms, 1, &stats[r]);
    }
  }
  pool_free(pd, h0->stream);
  return rc;
}

static int bias_stats(b2svm_handle* h, BiasStats* host) {
  B2_CUDA(cudaSetDevice(h->d.device));
  // multi-block (one block took 0.69 ms at 200k variables): chunks of >= 4096 variables, at most 128 blocks
  const int64_t nvb = h->d.n_vars;
  const int nb = (int)std::max<int64_t>(1, std::min<int64_t>(kBiasBlocks, (nvb + 4095) / 4096));
  const int64_t per_block = (nvb + nb - 1) / nb;
  bias_stats_kernel<<<nb, 256, 0, h->stream>>>(h->d.y, h->d.alpha, h->d.neg_y_grad, h->d.C, nvb, per_block, h->bias_dev + 1);
  bias_stats_merge_kernel<<<1, 32, 0, h->stream>>>(h->bias_dev + 1, nb, h->bias_dev);
  B2_CUDA(cudaGetLastError());
  B2_CUDA(cudaMemcpyAsync(host, h->bias_dev, sizeof(BiasStats), cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  return B2SVM_OK;
}

/* Raw ingredients of calculate_rho / calculate_rho_b over THIS handle's variables, so that a
 * row-sharded solve can combine them across ranks (sums and counts add, extremes min / max):
 * out[0..5]  = sum_free, cnt_free, min_lb, cnt_lb, max_ub, cnt_ub            (solver.py:160-177)
 * out[6+6c..] = nsum, ncnt, nmax_hi, nhi, nmin_lo, nlo for class c = 0 (+1), 1 (-1)   (solver.py:378-419)
 * Empty extremes are reported as +inf (min) / -inf (max). */
int b2svm_release_cache(b2svm_handle* h) {
  // After a fit the solver object lives on (rho, state export, the reference keeps it inside its decision_function
  // closure), but its Q-column cache -- up to 70 % of the device memory -- is dead weight: hand it back.  Later
  // solves on the handle run without a cache.
  if (!h) return fail(B2SVM_E_INVALID, "null handle");
  if (h->precomputed || h->cache_slots == 0) return B2SVM_OK;
  B2_CUDA(cudaSetDevice(h->d.device));
  if (h->stream) B2_CUDA(cudaStreamSynchronize(h->stream)); else B2_CUDA(cudaDeviceSynchronize());
  if (h->cache_pooled) pool_free(h->cache, h->stream); else cudaFree(h->cache);
  pool_free(h->slot_of, h->stream);
  pool_free(h->row_of_slot, h->stream);
  h->cache = nullptr;
  h->slot_of = nullptr;
  h->row_of_slot = nullptr;
  h->cache_slots = 0;
  h->cache_used = 0;
  h->cache_hand = 0;
  return B2SVM_OK;
}

int b2svm_bias_partials(b2svm_handle* h, double* out /* host [18] */) {
  if (!h || !out) return fail(B2SVM_E_INVALID, "null argument");
  BiasStats s;
  int rc = bias_stats(h, &s);
  if (rc) return rc;
  const double inf = INFINITY;
  out[0] = s.sum_free; out[1] = (double)s.cnt_free;
  out[2] = s.cnt_lb ? s.min_lb : inf; out[3] = (double)s.cnt_lb;
  out[4] = s.cnt_ub ? s.max_ub : -inf; out[5] = (double)s.cnt_ub;
  for (int c = 0; c < 2; ++c) {
    double* o = out + 6 + 6 * c;
    o[0] = s.nsum[c]; o[1] = (double)s.ncnt[c];
    o[2] = s.nhi[c] ? s.nmax_hi[c] : -inf; o[3] = (double)s.nhi[c];
    o[4] = s.nlo[c] ? s.nmin_lo[c] : inf; o[5] = (double)s.nlo[c];
  }
  return B2SVM_OK;
}

int b2svm_rho(b2svm_handle* h, double* rho) {
  if (!h || !rho) return fail(B2SVM_E_INVALID, "null argument");
  BiasStats s;
  int rc = bias_stats(h, &s);
  if (rc) return rc;
  if (s.cnt_free > 0) {
    *rho = -(s.sum_free / (double)s.cnt_free);
  } else {
    if (s.cnt_lb == 0 || s.cnt_ub == 0)
      return fail(B2SVM_E_INVALID, "calculate_rho: empty bound set (the reference raises ValueError here)");
    *rho = -(s.min_lb + s.max_ub) / 2;
  }
  return B2SVM_OK;
}

int b2svm_rho_b(b2svm_handle* h, double* rho, double* b) {
  if (!h || !rho || !b) return fail(B2SVM_E_INVALID, "null argument");
  BiasStats s;
  int rc = bias_stats(h, &s);
  if (rc) return rc;
  double r[2];
  for (int c = 0; c < 2; ++c) {
    if (s.ncnt[c] > 0) r[c] = s.nsum[c] / (double)s.ncnt[c];
    else if (s.nhi[c] > 0 && s.nlo[c] > 0) r[c] = (s.nmax_hi[c] + s.nmin_lo[c]) / 2;
    else r[c] = 0;   // the reference's bare except (solver.py:393-394, 411-412)
  }
  *rho = (r[0] + r[1]) / 2;
  *b = (r[1] - r[0]) / 2;
  return B2SVM_OK;
}

int b2svm_solve_batched(b2svm_handle* const* handles, int32_t count, int64_t max_iter, b2svm_solve_stats* stats) {
  if (!handles || count <= 0) return fail(B2SVM_E_INVALID, "empty batch");
  if (max_iter < 0 || max_iter >= B2SVM_MAX_ITER) return fail(B2SVM_E_UNSUPPORTED, "max_iter must be in [0, 2^29) per call");
  b2svm_handle* h0 = handles[0];
  B2_CUDA(cudaSetDevice(h0->d.device));
  for (int k = 0; k < count; ++k) {
    b2svm_handle* h = handles[k];
    if (!h) return fail(B2SVM_E_INVALID, "null handle in batch");
    if (h->f64 != h0->f64 || h->ch != h0->ch || h->d.device != h0->d.device || h->stream != h0->stream)
      return fail(B2SVM_E_INVALID, "batched problems must share dtype, row width, device and stream");
  }
  // one CTA per problem, all problems in one launch (no inter-CTA waits -> ordinary launch)
  std::vector<ProblemDev> P(count);
  ProblemDev* pd = nullptr;
  SolveResult* rd = nullptr;
  B2_CUDA(pool_alloc(&pd, sizeof(ProblemDev) * count, h0->stream));
  B2_CUDA(pool_alloc(&rd, sizeof(SolveResult) * count, h0->stream));
  for (int k = 0; k < count; ++k) {
    b2svm_handle* h = handles[k];
    refresh_flags_kernel<<<grid_for(h->d.n_vars), 256, 0, h0->stream>>>(h->d.y, h->d.alpha, h->d.C, h->flags, h->d.n_vars);
    cudaMemsetAsync(h->l1, 0, h->l1_bytes, h0->stream);
    cudaMemsetAsync(h->abort_flag, 0, sizeof(int), h0->stream);
    memset(&P[k], 0, sizeof(ProblemDev));
    fill_problem(h, P[k], max_iter, -1, -1);
    P[k].ncta = 1;
    P[k].sub = 1;
    P[k].tiles_per_cta = (int)h->ntiles;
    P[k].result = rd + k;
  }
  B2_CUDA(cudaGetLastError());
  B2_CUDA(cudaMemcpyAsync(pd, P.data(), sizeof(ProblemDev) * count, cudaMemcpyHostToDevice, h0->stream));
  B2_CUDA(cudaEventRecord(h0->ev0, h0->stream));
  bool any_cache = false;
  for (int k = 0; k < count; ++k) any_cache = any_cache || handles[k]->cache_slots > 0;
  bool any_generic = false;
  for (int k = 0; k < count; ++k) any_generic = any_generic || needs_generic(handles[k]);
  int rc = launch(h0->f64, h0->ch, any_cache, any_generic, pd, count, 1, h0->stages1, h0->smem_bytes1, h0->stream);
  if (rc == B2SVM_OK) {
    cudaEventRecord(h0->ev1, h0->stream);
    std::vector<SolveResult> R(count);
    cudaError_t e = cudaMemcpyAsync(R.data(), rd, sizeof(SolveResult) * count, cudaMemcpyDeviceToHost, h0->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h0->stream);
    if (e != cudaSuccess) rc = fail(B2SVM_E_CUDA, cudaGetErrorString(e));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h0->ev0, h0->ev1);
    for (int k = 0; rc == B2SVM_OK && k < count; ++k) {
      handles[k]->cache_used = R[k].cache_used;   // the device slot table persists across launches (as in run_solver)
      handles[k]->cache_hand = R[k].cache_hand;
      if (R[k].status != 0) rc = fail(B2SVM_E_TIMEOUT, "device watchdog tripped in a batched solve");
      if (stats) {
        memset(&stats[k], 0, sizeof(stats[k]));
        stats[k].n_iter = R[k].n_iter;
        stats[k].converged = R[k].converged;
        stats[k].n_ctas = 1;
        stats[k].last_i = R[k].last_i;
        stats[k].last_j = R[k].last_j;
        stats[k].last_delta_i = R[k].di;
        stats[k].last_delta_j = R[k].dj;
        stats[k].cold_sweeps = R[k].cold_sweeps;
        stats[k].cache_misses = 2 * R[k].cold_sweeps;
        const double esz = handles[k]->f64 ? 8.0 : 4.0;
        stats[k].algorithmic_bytes = (double)R[k].cold_sweeps * ((double)handles[k]->d.n_rows * handles[k]->d.n_features * esz +
                                                                 (double)handles[k]->d.n_vars * 18.0);
        stats[k].device_ms = ms;
        stats[k].launches = count + 1;
      }
    }
  }
  pool_free(pd, h0->stream);
  pool_free(rd, h0->stream);
  return rc;
}

/* debug only (not part of the ABI header): copy the leader timeline of the last solve to the host */
int b2svm_debug_timeline(b2svm_handle* h, long long* out, int n) {
  if (!h || !h->timeline) return fail(B2SVM_E_INVALID, "timeline not enabled (B2SVM_TIMELINE=1)");
  B2_CUDA(cudaMemcpy(out, h->timeline, sizeof(long long) * std::min(n, TL_TOTAL), cudaMemcpyDeviceToHost));
  return B2SVM_OK;
}

int b2svm_debug_steps(b2svm_handle* h, double* out, int epochs) {
  if (!h || !h->dbg_steps) return fail(B2SVM_E_INVALID, "debug steps not enabled (B2SVM_DEBUG_STEPS=1)");
  B2_CUDA(cudaMemcpy(out, h->dbg_steps, sizeof(double) * (size_t)std::min(epochs, DBG_EPOCHS) * h->ncta * DBG_W, cudaMemcpyDeviceToHost));
  B2_CUDA(cudaMemcpy(out + (size_t)std::min(epochs, DBG_EPOCHS) * h->ncta * DBG_W, h->dbg_steps + (size_t)DBG_EPOCHS * h->ncta * DBG_W,
                     sizeof(double) * (size_t)std::min(epochs, DBG_EPOCHS) * 2, cudaMemcpyDeviceToHost));
  return B2SVM_OK;
}

int b2svm_mailbox_export(b2svm_handle* h, void* ipc_handle_out) {
  if (!h || !ipc_handle_out) return fail(B2SVM_E_INVALID, "null argument");
  if (h->d.world <= 1 || !h->block) return fail(B2SVM_E_INVALID, "not a multi-GPU handle");
  B2_CUDA(cudaSetDevice(h->d.device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t mh;
  B2_CUDA(cudaIpcGetMemHandle(&mh, h->block));
  memcpy(ipc_handle_out, &mh, sizeof(mh));
  return B2SVM_OK;
}

int b2svm_mailbox_attach(b2svm_handle* h, const void* ipc_handles) {
  if (!h || !ipc_handles) return fail(B2SVM_E_INVALID, "null argument");
  if (h->d.world <= 1 || !h->block) return fail(B2SVM_E_INVALID, "not a multi-GPU handle");
  if (h->attached) return B2SVM_OK;
  B2_CUDA(cudaSetDevice(h->d.device));
  const int world = h->d.world;
  h->peer_blocks.assign(world, nullptr);
  std::vector<void*> tab((size_t)world);
  for (int r = 0; r < world; ++r) {
    unsigned char* base = h->block;
    if (r != h->d.rank) {
      cudaIpcMemHandle_t mh;
      memcpy(&mh, reinterpret_cast<const unsigned char*>(ipc_handles) + 64 * (size_t)r, sizeof(mh));
      void* p = nullptr;
      B2_CUDA(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
      h->peer_blocks[r] = p;
      base = reinterpret_cast<unsigned char*>(p);
    }
    tab[r] = base;
  }
  B2_CUDA(cudaMalloc(&h->peer_tables, sizeof(void*) * tab.size()));
  B2_CUDA(cudaMemcpy(h->peer_tables, tab.data(), sizeof(void*) * tab.size(), cudaMemcpyHostToDevice));
  h->attached = true;
  return B2SVM_OK;
}

int b2svm_prepare(b2svm_handle* h) {
  if (!h) return fail(B2SVM_E_INVALID, "null handle");
  return prepare_launch(h);
}

}  // extern "C"
