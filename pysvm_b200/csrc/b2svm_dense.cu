// Dense kernel evaluation away from the SMO loop: decision_function / initial gradients (K7) and
// random Fourier features (K8).  First-generation CUDA-core kernels (FFMA / DFMA with shared-memory
// tiling); they never materialise the n_a x n_b kernel matrix the reference builds on the host
// (pysvm/svm/svc.py:269-272).
#include <algorithm>
#include <cstring>
#include <string>

#include "b2svm_device.cuh"
#include "b2svm_internal.h"

namespace b2svm {

// ------------------------------------------------------------------------------------------
// ordered compaction of the rows with a non-zero coefficient (support vectors)
// ------------------------------------------------------------------------------------------
// Ordered, multi-block: (1) every block counts the non-zeros of its contiguous chunk, (2) one block turns the counts
// into exclusive offsets (and the total), (3) every block writes its indices at its offset, in index order.
constexpr int CMP_THREADS = 256, CMP_ITEMS = 16;          // 4096 coefficients per block
__global__ void __launch_bounds__(CMP_THREADS) compact_count_kernel(const double* __restrict__ coef, int64_t n, int* __restrict__ counts) {
  __shared__ int warp_tot[CMP_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * CMP_THREADS * CMP_ITEMS;
  int c = 0;
  for (int k = 0; k < CMP_ITEMS; ++k) {
    const int64_t e = base + (int64_t)k * CMP_THREADS + threadIdx.x;
    c += (e < n && coef[e] != 0.0) ? 1 : 0;
  }
  for (int m = 16; m >= 1; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
  if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < CMP_THREADS / 32; ++w) t += warp_tot[w];
    counts[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(1024) compact_scan_kernel(int* __restrict__ counts, int nblocks, int* __restrict__ total) {
  __shared__ int sh[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int start = 0; start < nblocks; start += 1024) {
    const int k = start + threadIdx.x;
    const int v = k < nblocks ? counts[k] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {          // inclusive Hillis-Steele scan
      const int add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += add;
      __syncthreads();
    }
    if (k < nblocks) counts[k] = carry + sh[threadIdx.x] - v;   // exclusive offset of block k
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(CMP_THREADS) compact_write_kernel(const double* __restrict__ coef, int64_t n, const int* __restrict__ offsets,
                                                                    int* __restrict__ idx) {
  __shared__ int warp_tot[CMP_THREADS / 32];
  __shared__ int base_sh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * CMP_THREADS * CMP_ITEMS;
  if (threadIdx.x == 0) base_sh = offsets[blockIdx.x];
  __syncthreads();
  for (int k = 0; k < CMP_ITEMS; ++k) {                 // chunk of CMP_THREADS consecutive coefficients: order preserved
    const int64_t e = base + (int64_t)k * CMP_THREADS + threadIdx.x;
    const bool nz = e < n && coef[e] != 0.0;
    const unsigned bal = __ballot_sync(0xffffffffu, nz);
    const int before = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base_sh;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (nz) idx[off + before] = (int)e;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < CMP_THREADS / 32; ++w) t += warp_tot[w];
      base_sh += t;
    }
    __syncthreads();
  }
}

// gather support rows into a dense [nsv][dp] block (zero padded), with norms and coefficients
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ Xa, int64_t lda, int d, int dp, const int* __restrict__ idx,
                                   int nsv, const double* __restrict__ coef, T* __restrict__ sv, T* __restrict__ svn,
                                   double* __restrict__ svc) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp; k < nsv; k += nwarps) {
    const int64_t r = idx[k];
    T s = 0;
    for (int c = lane; c < dp; c += 32) {
      const T v = c < d ? Xa[r * lda + c] : T(0);
      sv[k * dp + c] = v;
      s += v * v;
    }
    s = warp_sum(s);
    if (lane == 0) {
      svn[k] = s;
      svc[k] = coef[r];
    }
  }
}

// out_partial[y][q] = sum_{s in slice y} coef_s K(sv_s, x_q); one query per thread, TS support rows
// per shared-memory tile, query tile kept in shared memory chunk-major (conflict-free 16-B reads).
constexpr int MV_TS = 32;

template <typename T, int TQ, int KIND>
__global__ void __launch_bounds__(TQ) matvec_kernel(const T* __restrict__ sv, const T* __restrict__ svn,
                                                    const double* __restrict__ svc, int nsv, int dp,
                                                    const T* __restrict__ Xb, int64_t ldb, int d, int64_t nb,
                                                    KernelParams<T> kp, int slice, double* __restrict__ partial) {
  using CK = Chunk<T>;
  using V = typename CK::V;
  extern __shared__ __align__(16) unsigned char smem_mv[];
  const int nch = dp / CK::N;
  V* qs = reinterpret_cast<V*>(smem_mv);            // [nch][TQ]
  V* ts = qs + (size_t)nch * TQ;                     // [MV_TS][nch]
  T* tn = reinterpret_cast<T*>(ts + (size_t)MV_TS * nch);
  double* tc = reinterpret_cast<double*>(tn + MV_TS + (MV_TS & 1));
  const int tid = threadIdx.x;
  const int64_t q0 = (int64_t)blockIdx.x * TQ;
  // stage the query tile (arbitrary d / ldb: scalar, bounds-checked, zero padded)
  T* qs_s = reinterpret_cast<T*>(qs);
  for (int e = tid; e < TQ * dp; e += TQ) {
    const int qq = e / dp, c = e - qq * dp;
    const int64_t q = q0 + qq;
    const T v = (q < nb && c < d) ? Xb[q * ldb + c] : T(0);
    qs_s[((size_t)(c / CK::N) * TQ + qq) * CK::N + (c % CK::N)] = v;
  }
  __syncthreads();
  T nq = 0;
  for (int c = 0; c < nch; ++c) nq = CK::dot(qs[(size_t)c * TQ + tid], qs[(size_t)c * TQ + tid], nq);

  const int s0 = blockIdx.y * slice;
  const int s1 = min(nsv, s0 + slice);
  double sum = 0.0;
  for (int t0 = s0; t0 < s1; t0 += MV_TS) {
    __syncthreads();
    for (int e = tid; e < MV_TS * nch; e += TQ) {
      const int s = e / nch;
      ts[e] = (t0 + s < s1) ? reinterpret_cast<const V*>(sv + (size_t)(t0 + s) * dp)[e - s * nch] : CK::zero();
    }
    for (int s = tid; s < MV_TS; s += TQ) {
      tn[s] = (t0 + s < s1) ? svn[t0 + s] : T(0);
      tc[s] = (t0 + s < s1) ? svc[t0 + s] : 0.0;
    }
    __syncthreads();
    T acc[MV_TS];
#pragma unroll
    for (int s = 0; s < MV_TS; ++s) acc[s] = 0;
    for (int c = 0; c < nch; ++c) {
      const V vq = qs[(size_t)c * TQ + tid];
#pragma unroll
      for (int s = 0; s < MV_TS; ++s) acc[s] = CK::dot(vq, ts[s * nch + c], acc[s]);
    }
#pragma unroll
    for (int s = 0; s < MV_TS; ++s) {
      if (t0 + s < s1) {
        const T kv = kernel_value_k<KIND>(kp, acc[s], tn[s], nq);
        sum = fma(tc[s], (double)kv, sum);
      }
    }
  }
  const int64_t q = q0 + tid;
  if (q < nb) partial[(size_t)blockIdx.y * nb + q] = sum;
}

__global__ void reduce_partials_kernel(const double* __restrict__ partial, int ny, int64_t nb, double* __restrict__ out) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nb; q += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int y = 0; y < ny; ++y) s += partial[(size_t)y * nb + q];
    out[q] = s;
  }
}

template <typename T>
static int matvec_T(cudaStream_t stream, const b2svm_kernel_desc& k, int d, const T* Xa, int64_t na, int64_t lda,
                    const T* Xb, int64_t nb, int64_t ldb, const double* coef, double* out) {
  using CK = Chunk<T>;
  if (na >= (1ll << 31)) return fail(B2SVM_E_UNSUPPORTED, "too many rows");
  const int dp = (d + CK::N - 1) / CK::N * CK::N;
  int* idx = nullptr;
  int* count_dev = nullptr;
  keep_pool_memory();
  const int cmp_blocks = (int)((na + CMP_THREADS * CMP_ITEMS - 1) / (CMP_THREADS * CMP_ITEMS));
  B2_CUDA(cudaMallocAsync(&idx, sizeof(int) * (size_t)na, stream));
  B2_CUDA(cudaMallocAsync(&count_dev, sizeof(int) * (size_t)(cmp_blocks + 1), stream));
  compact_count_kernel<<<cmp_blocks, CMP_THREADS, 0, stream>>>(coef, na, count_dev + 1);
  compact_scan_kernel<<<1, 1024, 0, stream>>>(count_dev + 1, cmp_blocks, count_dev);
  compact_write_kernel<<<cmp_blocks, CMP_THREADS, 0, stream>>>(coef, na, count_dev + 1, idx);
  // (the count comes back to the host: the work buffers and the launch geometry of the mat-vec depend on it)
  int nsv = 0;
  B2_CUDA(cudaMemcpyAsync(&nsv, count_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
  B2_CUDA(cudaStreamSynchronize(stream));
  int rc = B2SVM_OK;
  T *sv = nullptr, *svn = nullptr;
  double *svc = nullptr, *partial = nullptr;
  auto done = [&](int code) {
    for (void* ptr : {(void*)idx, (void*)count_dev, (void*)sv, (void*)svn, (void*)svc, (void*)partial})
      if (ptr) cudaFreeAsync(ptr, stream);
    return code;
  };
  if (nsv == 0) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double) * (size_t)nb, stream);
    return done(e == cudaSuccess ? B2SVM_OK : fail(B2SVM_E_CUDA, cudaGetErrorString(e)));
  }
  if constexpr (sizeof(T) == 4) {
    // float32, d <= 128, enough work to fill 128 x 128 tiles: tcgen05 path (b2svm_matvec_tc.cu)
    const int mode = matvec_tc_mode(B2SVM_F32, d);
    if (mode == 2 || (mode == 1 && nb >= 1024 && nsv >= 1024))
      return done(matvec_tc(stream, k, d, (const float*)Xa, lda, idx, nsv, coef, (const float*)Xb, nb, ldb, out));
  }
#define MV_TRY(expr)                                                                     \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) return done(fail(B2SVM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e))); \
  } while (0)
  MV_TRY(cudaMallocAsync(&sv, sizeof(T) * (size_t)nsv * dp, stream));
  MV_TRY(cudaMallocAsync(&svn, sizeof(T) * (size_t)nsv, stream));
  MV_TRY(cudaMallocAsync(&svc, sizeof(double) * (size_t)nsv, stream));
  {
    int g = (int)std::min<int64_t>(148 * 8, ((int64_t)nsv * 32 + 255) / 256);
    gather_rows_kernel<T><<<std::max(g, 1), 256, 0, stream>>>(Xa, lda, d, dp, idx, nsv, coef, sv, svn, svc);
    MV_TRY(cudaGetLastError());
  }
  KernelParams<T> kp{k.kind, (T)k.gamma, (T)k.coef0, (T)k.degree};
  const int nch = dp / CK::N;
  int TQ = 128;
  auto smem_for = [&](int tq) {
    return (size_t)nch * tq * 16 + (size_t)MV_TS * nch * 16 + sizeof(T) * (MV_TS + 1) + sizeof(double) * MV_TS + 32;
  };
  while (TQ > 32 && smem_for(TQ) > 160 * 1024) TQ /= 2;
  if (smem_for(TQ) > 200 * 1024) return done(fail(B2SVM_E_UNSUPPORTED, "n_features too large for kernel_matvec"));
  const int gx = (int)((nb + TQ - 1) / TQ);
  const int tiles = (nsv + MV_TS - 1) / MV_TS;
  int ny = std::max(1, std::min(tiles, (2 * 148 + gx - 1) / gx));
  ny = std::min(ny, 65535);
  const int slice = ((tiles + ny - 1) / ny) * MV_TS;
  ny = (nsv + slice - 1) / slice;
  MV_TRY(cudaMallocAsync(&partial, sizeof(double) * (size_t)ny * nb, stream));
  const size_t smem = smem_for(TQ);
  dim3 grid(gx, ny);
#define MV_LAUNCH2(TQV, KIND)                                                                                      \
  do {                                                                                                             \
    MV_TRY(cudaFuncSetAttribute(matvec_kernel<T, TQV, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    matvec_kernel<T, TQV, KIND><<<grid, TQV, smem, stream>>>(sv, svn, svc, nsv, dp, Xb, ldb, d, nb, kp, slice, partial); \
  } while (0)
#define MV_LAUNCH(TQV)                                 \
  switch (k.kind) {                                    \
    case K_RBF: MV_LAUNCH2(TQV, K_RBF); break;         \
    case K_POLY: MV_LAUNCH2(TQV, K_POLY); break;       \
    case K_SIGMOID: MV_LAUNCH2(TQV, K_SIGMOID); break; \
    default: MV_LAUNCH2(TQV, K_LINEAR); break;         \
  }
  if (TQ == 128) { MV_LAUNCH(128) }
  else if (TQ == 64) { MV_LAUNCH(64) }
  else { MV_LAUNCH(32) }
  MV_TRY(cudaGetLastError());
  {
    int g = (int)std::min<int64_t>(148 * 8, (nb + 255) / 256);
    reduce_partials_kernel<<<std::max(g, 1), 256, 0, stream>>>(partial, ny, nb, out);
    MV_TRY(cudaGetLastError());
  }
  MV_TRY(cudaStreamSynchronize(stream));
#undef MV_LAUNCH
#undef MV_LAUNCH2
#undef MV_TRY
  (void)rc;
  return done(B2SVM_OK);
}

int kernel_matvec_impl(cudaStream_t stream, const b2svm_kernel_desc& k, int x_dtype, int d, const void* Xa,
                       int64_t na, int64_t lda, const void* Xb, int64_t nb, int64_t ldb, const double* coef,
                       double* out) {
  if (nb <= 0 || na <= 0) return B2SVM_OK;
  if (x_dtype == B2SVM_F64)
    return matvec_T<double>(stream, k, d, (const double*)Xa, na, lda, (const double*)Xb, nb, ldb, coef, out);
  return matvec_T<float>(stream, k, d, (const float*)Xa, na, lda, (const float*)Xb, nb, ldb, coef, out);
}

// ------------------------------------------------------------------------------------------
// random Fourier features (rff.py:19-20): float64 like the reference (W, b are float64 there, so
// X is promoted before the product).  Block: 16 rows x 64 features.
// ------------------------------------------------------------------------------------------
constexpr int RF_ROWS = 16, RF_COLS = 64;

template <typename T, bool DECISION>
__global__ void __launch_bounds__(RF_COLS * 4) rff_kernel(const T* __restrict__ X, int64_t n, int64_t ldx, int d,
                                                         const double* __restrict__ W, const double* __restrict__ b,
                                                         int D, const double* __restrict__ wz,
                                                         double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_rf[];
  double* xs = reinterpret_cast<double*>(smem_rf);   // [RF_ROWS][d]
  double* ws = xs + (size_t)RF_ROWS * d;             // [RF_COLS][d + 1]
  double* red = ws + (size_t)RF_COLS * (d + 1);      // [RF_ROWS][RF_COLS] (decision only)
  const int tx = threadIdx.x % RF_COLS;              // feature within tile
  const int ty = threadIdx.x / RF_COLS;              // 0..3 -> rows ty*4 .. ty*4+3
  const int64_t r0 = (int64_t)blockIdx.x * RF_ROWS;
  for (int e = threadIdx.x; e < RF_ROWS * d; e += blockDim.x) {
    const int rr = e / d, c = e - rr * d;
    xs[e] = (r0 + rr < n) ? (double)X[(r0 + rr) * ldx + c] : 0.0;
  }
  const double scale = sqrt(2.0 / (double)D);
  double dsum[4] = {0, 0, 0, 0};
  const int k_begin = DECISION ? 0 : blockIdx.y * RF_COLS;
  const int k_end = DECISION ? D : min(D, k_begin + RF_COLS);
  for (int k0 = k_begin; k0 < k_end; k0 += RF_COLS) {
    __syncthreads();
    for (int e = threadIdx.x; e < RF_COLS * d; e += blockDim.x) {
      const int kk = e / d, c = e - kk * d;
      ws[kk * (d + 1) + c] = (k0 + kk < D) ? W[(size_t)(k0 + kk) * d + c] : 0.0;
    }
    __syncthreads();
    const int k = k0 + tx;
    double acc[4] = {0, 0, 0, 0};
    for (int c = 0; c < d; ++c) {
      const double w = ws[tx * (d + 1) + c];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fma(xs[(ty * 4 + r) * d + c], w, acc[r]);
    }
    if (k < D) {
      const double bk = b[k];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const double z = scale * cos(acc[r] + bk);
        if (DECISION) dsum[r] += z * wz[k];
        else if (r0 + ty * 4 + r < n) out[(size_t)(r0 + ty * 4 + r) * D + k] = z;
      }
    }
  }
  if (DECISION) {
#pragma unroll
    for (int r = 0; r < 4; ++r) red[(ty * 4 + r) * RF_COLS + tx] = dsum[r];
    __syncthreads();
    if (threadIdx.x < RF_ROWS) {
      double s = 0.0;
      for (int c = 0; c < RF_COLS; ++c) s += red[threadIdx.x * RF_COLS + c];
      if (r0 + threadIdx.x < n) out[r0 + threadIdx.x] = s;
    }
  }
}

template <bool DECISION>
static int rff_launch(int device, void* stream_, int x_dtype, const void* X, int64_t n, int64_t ldx, int d,
                      const double* W, const double* b, int D, const double* wz, double* out) {
  if (n <= 0 || D <= 0 || d <= 0) return fail(B2SVM_E_INVALID, "empty RFF problem");
  B2_CUDA(cudaSetDevice(device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t smem = sizeof(double) * ((size_t)RF_ROWS * d + (size_t)RF_COLS * (d + 1) + (size_t)RF_ROWS * RF_COLS);
  if (smem > 200 * 1024) return fail(B2SVM_E_UNSUPPORTED, "n_features too large for the RFF kernel");
  dim3 grid((unsigned)((n + RF_ROWS - 1) / RF_ROWS), DECISION ? 1u : (unsigned)((D + RF_COLS - 1) / RF_COLS));
  if (x_dtype == B2SVM_F64) {
    B2_CUDA(cudaFuncSetAttribute(rff_kernel<double, DECISION>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rff_kernel<double, DECISION><<<grid, RF_COLS * 4, smem, stream>>>((const double*)X, n, ldx, d, W, b, D, wz, out);
  } else {
    B2_CUDA(cudaFuncSetAttribute(rff_kernel<float, DECISION>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rff_kernel<float, DECISION><<<grid, RF_COLS * 4, smem, stream>>>((const float*)X, n, ldx, d, W, b, D, wz, out);
  }
  B2_CUDA(cudaGetLastError());
  return B2SVM_OK;
}

// ------------------------------------------------------------------------------------------
// dense kernel matrix out[i][j] = K(a_i, b_j) in the input dtype (the reference's dense-Q mode builds
// y y^T * kernel_func(X, X) on the host: svc.py:251-253, oc_svm.py:85-89; with rff=True the kernel is z(a).z(b),
// svc.py:225-228).  CUDA-core version: float64 (exact like the reference's BLAS product up to summation order) and
// every shape the tensor-core path (tile_tc, float32, d <= 96) does not take.  64 x 64 tile per block, 4 x 4 per
// thread, K in chunks of 16 through shared memory; products are accumulated in increasing feature order.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void sq_norms_kernel(const T* __restrict__ X, int64_t ld, int64_t n, int d, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n; r += nwarps) {
    T s = 0;
    for (int c = lane; c < d; c += 32) {
      const T v = X[r * ld + c];
      s += v * v;
    }
    s = warp_sum(s);
    if (lane == 0) out[r] = s;
  }
}

template <typename T, int KIND>
__global__ void __launch_bounds__(256) kernel_matrix_kernel(const T* __restrict__ A, int64_t lda, int64_t na, const T* __restrict__ B,
                                                            int64_t ldb, int64_t nb, int d, const T* __restrict__ an,
                                                            const T* __restrict__ bn, KernelParams<T> kp, T* __restrict__ out,
                                                            int64_t ldo) {
  __shared__ T As[16][65], Bs[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * 64, j0 = (int64_t)blockIdx.x * 64;
  T acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0;
  for (int k0 = 0; k0 < d; k0 += 16) {
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      const int r = e >> 4, k = e & 15;
      As[k][r] = (i0 + r < na && k0 + k < d) ? A[(i0 + r) * lda + k0 + k] : T(0);
      Bs[k][r] = (j0 + r < nb && k0 + k < d) ? B[(j0 + r) * ldb + k0 + k] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      T av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[k][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[k][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t i = i0 + ty * 4 + a;
    if (i >= na) continue;
    const T ni = KIND == K_RBF ? an[i] : T(0);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t j = j0 + tx * 4 + b;
      if (j < nb) out[i * ldo + j] = kernel_value_k<KIND>(kp, acc[a][b], ni, KIND == K_RBF ? bn[j] : T(0));
    }
  }
}

template <typename T>
static int kernel_matrix_T(cudaStream_t stream, const b2svm_kernel_desc& k, int d, const T* A, int64_t na, int64_t lda, const T* B,
                           int64_t nb, int64_t ldb, T* out, int64_t ldo) {
  keep_pool_memory();
  T *an = nullptr, *bn = nullptr;
  if (k.kind == K_RBF) {
    B2_CUDA(cudaMallocAsync(&an, sizeof(T) * (size_t)na, stream));
    B2_CUDA(cudaMallocAsync(&bn, sizeof(T) * (size_t)nb, stream));
    sq_norms_kernel<T><<<(int)std::max<int64_t>(1, std::min<int64_t>(1184, (na * 32 + 255) / 256)), 256, 0, stream>>>(A, lda, na, d, an);
    sq_norms_kernel<T><<<(int)std::max<int64_t>(1, std::min<int64_t>(1184, (nb * 32 + 255) / 256)), 256, 0, stream>>>(B, ldb, nb, d, bn);
  }
  KernelParams<T> kp{k.kind, (T)k.gamma, (T)k.coef0, (T)k.degree};
  dim3 grid((unsigned)((nb + 63) / 64), (unsigned)((na + 63) / 64));
  switch (k.kind) {
    case K_RBF: kernel_matrix_kernel<T, K_RBF><<<grid, 256, 0, stream>>>(A, lda, na, B, ldb, nb, d, an, bn, kp, out, ldo); break;
    case K_POLY: kernel_matrix_kernel<T, K_POLY><<<grid, 256, 0, stream>>>(A, lda, na, B, ldb, nb, d, an, bn, kp, out, ldo); break;
    case K_SIGMOID: kernel_matrix_kernel<T, K_SIGMOID><<<grid, 256, 0, stream>>>(A, lda, na, B, ldb, nb, d, an, bn, kp, out, ldo); break;
    default: kernel_matrix_kernel<T, K_LINEAR><<<grid, 256, 0, stream>>>(A, lda, na, B, ldb, nb, d, an, bn, kp, out, ldo); break;
  }
  cudaError_t e = cudaGetLastError();
  if (an) cudaFreeAsync(an, stream);
  if (bn) cudaFreeAsync(bn, stream);
  if (e != cudaSuccess) return fail(B2SVM_E_CUDA, cudaGetErrorString(e));
  return B2SVM_OK;
}

// out[c] = sum_r w[r] X[r][c] in float64 (primal weights of the linear estimators, svc.py:76 / svr.py:89-90 telescoped;
// z(x).(Z^T coef) of the RFF decision).  Deterministic: fixed row slices, partial sums reduced in slice order.
template <typename T>
__global__ void __launch_bounds__(256) weighted_rows_kernel(const T* __restrict__ X, int64_t ld, int64_t n, int d,
                                                            const double* __restrict__ w, int64_t rows_per_slice,
                                                            double* __restrict__ partial) {
  __shared__ double red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slice, r1 = min(n, r0 + rows_per_slice);
  double acc = 0.0;
  if (c < d)
    for (int64_t r = r0 + ty; r < r1; r += 8) acc = fma(w[r], (double)X[r * ld + c], acc);
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < d) {
    double s = 0.0;
    for (int y = 0; y < 8; ++y) s += red[y][tx];
    partial[(size_t)blockIdx.y * d + c] = s;
  }
}

}  // namespace b2svm

using namespace b2svm;

extern "C" {

int b2svm_kernel_matvec(int32_t device, void* stream, const b2svm_kernel_desc* k, int32_t x_dtype, int32_t n_features,
                        const void* Xa, int64_t na, int64_t lda, const void* Xb, int64_t nb, int64_t ldb,
                        const double* coef, double* out) {
  if (!k || !Xa || !Xb || !coef || !out) return fail(B2SVM_E_INVALID, "null argument");
  if (x_dtype != B2SVM_F32 && x_dtype != B2SVM_F64) return fail(B2SVM_E_INVALID, "bad x_dtype");
  if (n_features <= 0 || lda < n_features || ldb < n_features) return fail(B2SVM_E_INVALID, "bad leading dimension");
  B2_CUDA(cudaSetDevice(device));
  return kernel_matvec_impl((cudaStream_t)stream, *k, x_dtype, n_features, Xa, na, lda, Xb, nb, ldb, coef, out);
}

// does the tensor-core tile kernel take this RFF transform?  B2SVM_RFF_TC=0 keeps the float64 kernel, =2 forces it
static bool rff_transform_on_tc(int32_t x_dtype, int64_t n, int32_t d, int32_t D, int64_t ldo, bool out_f64, const void* out) {
  int mode = x_dtype == B2SVM_F32 ? 1 : 0;
  if (const char* e = getenv("B2SVM_RFF_TC")) mode = std::min(mode ? 2 : 0, atoi(e));
  if (!tile_tc_ok(d, ldo, out_f64, out)) return false;
  return mode == 2 || (mode == 1 && n >= 1024 && D >= 128);
}

int b2svm_rff_transform(int32_t device, void* stream, int32_t x_dtype, const void* X, int64_t n, int64_t ldx, int32_t d,
                        const double* W, const double* b, int32_t D, double* out) {
  if (!X || !W || !b || !out) return fail(B2SVM_E_INVALID, "null argument");
  if (n <= 0 || D <= 0 || d <= 0) return fail(B2SVM_E_INVALID, "empty RFF problem");
  if (rff_transform_on_tc(x_dtype, n, d, D, D, true, out)) {
    B2_CUDA(cudaSetDevice(device));
    b2svm_kernel_desc k{};
    return tile_tc((cudaStream_t)stream, k, d, (const float*)X, ldx, n, nullptr, 0, W, b, D, std::sqrt(2.0 / (double)D), out, D, true);
  }
  return rff_launch<false>(device, stream, x_dtype, X, n, ldx, d, W, b, D, nullptr, out);
}

int b2svm_rff_transform_f32(int32_t device, void* stream, const float* X, int64_t n, int64_t ldx, int32_t d, const double* W,
                            const double* b, int32_t D, float* out, int64_t ldo) {
  if (!X || !W || !b || !out) return fail(B2SVM_E_INVALID, "null argument");
  if (n <= 0 || D <= 0 || d <= 0 || ldo < D) return fail(B2SVM_E_INVALID, "empty RFF problem / bad output pitch");
  if (!tile_tc_ok(d, ldo, false, out))
    return fail(B2SVM_E_UNSUPPORTED, "rff_transform_f32 needs n_features <= 96 and a 16-byte aligned output pitch (tensor-core path only)");
  B2_CUDA(cudaSetDevice(device));
  b2svm_kernel_desc k{};
  return tile_tc((cudaStream_t)stream, k, d, X, ldx, n, nullptr, 0, W, b, D, std::sqrt(2.0 / (double)D), out, ldo, false);
}

int b2svm_kernel_matrix(int32_t device, void* stream, const b2svm_kernel_desc* k, int32_t x_dtype, int32_t n_features,
                        const void* Xa, int64_t na, int64_t lda, const void* Xb, int64_t nb, int64_t ldb, void* out, int64_t ldo) {
  if (!k || !Xa || !Xb || !out) return fail(B2SVM_E_INVALID, "null argument");
  if (x_dtype != B2SVM_F32 && x_dtype != B2SVM_F64) return fail(B2SVM_E_INVALID, "bad x_dtype");
  if (k->kind < 0 || k->kind > 3) return fail(B2SVM_E_INVALID, "bad kernel kind");
  if (n_features <= 0 || na <= 0 || nb <= 0 || lda < n_features || ldb < n_features || ldo < nb) return fail(B2SVM_E_INVALID, "bad shape");
  B2_CUDA(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  if (x_dtype == B2SVM_F32) {
    int mode = 1;
    if (const char* e = getenv("B2SVM_GRAM_TC")) mode = atoi(e);
    if (mode > 0 && tile_tc_ok(n_features, ldo, false, out) && (mode == 2 || (na >= 1024 && nb >= 1024)))
      return tile_tc(s, *k, n_features, (const float*)Xa, lda, na, (const float*)Xb, ldb, nullptr, nullptr, nb, 1.0, out, ldo, false);
    int rc = kernel_matrix_T<float>(s, *k, n_features, (const float*)Xa, na, lda, (const float*)Xb, nb, ldb, (float*)out, ldo);
    if (rc) return rc;
  } else {
    int rc = kernel_matrix_T<double>(s, *k, n_features, (const double*)Xa, na, lda, (const double*)Xb, nb, ldb, (double*)out, ldo);
    if (rc) return rc;
  }
  B2_CUDA(cudaStreamSynchronize(s));
  return B2SVM_OK;
}

int b2svm_weighted_rows(int32_t device, void* stream, int32_t x_dtype, const void* X, int64_t n, int64_t ldx, int32_t d,
                        const double* w, double* out) {
  if (!X || !w || !out) return fail(B2SVM_E_INVALID, "null argument");
  if (x_dtype != B2SVM_F32 && x_dtype != B2SVM_F64) return fail(B2SVM_E_INVALID, "bad x_dtype");
  if (n <= 0 || d <= 0 || ldx < d) return fail(B2SVM_E_INVALID, "bad shape");
  B2_CUDA(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  keep_pool_memory();
  const int gx = (d + 31) / 32;
  int ny = (int)std::max<int64_t>(1, std::min<int64_t>(n / 256 + 1, (4 * 148 + gx - 1) / gx));
  const int64_t rps = (n + ny - 1) / ny;
  ny = (int)((n + rps - 1) / rps);
  double* partial = nullptr;
  B2_CUDA(cudaMallocAsync(&partial, sizeof(double) * (size_t)ny * d, s));
  if (x_dtype == B2SVM_F64) weighted_rows_kernel<double><<<dim3(gx, ny), 256, 0, s>>>((const double*)X, ldx, n, d, w, rps, partial);
  else weighted_rows_kernel<float><<<dim3(gx, ny), 256, 0, s>>>((const float*)X, ldx, n, d, w, rps, partial);
  reduce_partials_kernel<<<std::max(1, (d + 255) / 256), 256, 0, s>>>(partial, ny, d, out);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(partial, s);
  if (e != cudaSuccess) return fail(B2SVM_E_CUDA, cudaGetErrorString(e));
  B2_CUDA(cudaStreamSynchronize(s));
  return B2SVM_OK;
}

int b2svm_rff_decision(int32_t device, void* stream, int32_t x_dtype, const void* Xq, int64_t nq, int64_t ldx, int32_t d,
                       const double* W, const double* b, int32_t D, const double* wz, double* out) {
  if (!Xq || !W || !b || !wz || !out) return fail(B2SVM_E_INVALID, "null argument");
  {
    // float32 queries, enough work for 128 x 128 tiles: the GEMM + cos epilogue runs on the tensor cores (3xTF32,
    // argument error ~1e-6); B2SVM_RFF_TC=0 keeps the float64 kernel, =2 forces the tensor path
    int mode = matvec_tc_mode(x_dtype, d);
    if (const char* e = getenv("B2SVM_RFF_TC")) mode = std::min(mode ? 2 : 0, atoi(e));
    if (mode == 2 || (mode == 1 && nq >= 1024 && D >= 256)) {
      B2_CUDA(cudaSetDevice(device));
      return rff_decision_tc((cudaStream_t)stream, d, W, b, D, wz, (const float*)Xq, nq, ldx, out);
    }
  }
  return rff_launch<true>(device, stream, x_dtype, Xq, nq, ldx, d, W, b, D, wz, out);
}

}  // extern "C"
