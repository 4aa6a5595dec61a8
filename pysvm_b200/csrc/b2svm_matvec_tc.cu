// K7 on the 5th-generation tensor cores: out[q] = sum_s coef_s K(sv_s, x_q) for float32 inputs.
//
//   D[q][s] = x_q . sv_s  is a GEMM (M = 128 queries per CTA, N = 128 support rows per tile, K = padded d)
//   issued as tcgen05.mma.kind::tf32 by one elected thread, operands staged in shared memory by TMA
//   (cp.async.bulk.tensor.2d, 128-byte swizzle), accumulators in TMEM (two buffers of 128 columns);
//   four epilogue warps read the accumulator back with tcgen05.ld -- TMEM lane = query row, so every thread
//   owns one query and walks the 128 support columns of the tile: kernel function, times coef, summed in
//   float64 (the reference multiplies a float64 coefficient vector into the kernel block, svc.py:269-272).
//   No cross-thread reduction, no n_a x n_b matrix.
//
// Precision: TF32 keeps 10 mantissa bits, far too few for exp(-gamma |a-b|^2) at 1e-4.  Every operand is split
// x = hi + lo (hi = x with the low 13 bits cleared, lo = x - hi exactly) and the product is accumulated as
// lo.hi + hi.lo + hi.hi ("3xTF32"): relative error ~2^-21, the same order as the FFMA path.
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include <algorithm>
#include <string>

#include "b2svm_device.cuh"
#include "b2svm_internal.h"

namespace b2svm {

constexpr int K_RFFCOS = 16;   // internal epilogue: cos(dot + phase_s), the random-Fourier-feature form (rff.py:19-20)
constexpr int TC_M = 128, TC_N = 128, TC_KBLK = 32;           // tile: 128 x 128, K blocks of 32 floats = 128 B
constexpr int TC_TILE_BYTES = TC_M * TC_KBLK * 4;              // one [128 rows][128 B] swizzled operand block
constexpr int TC_EPI_WARPS = 16;                                // four per TMEM lane quarter, 32 columns each
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;             // warp 0: TMA, warp 1: MMA + TMEM, warps 2-17: epilogue

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c_inner, int c_outer,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c_inner), "r"(c_outer), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded spin: a broken pipeline traps (reported as a CUDA error) instead of hanging the device
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand block [rows][128 B], 128-byte swizzle: 8-row groups are 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem_tile) {
  const uint32_t addr = smem_u32(smem_tile);
  const uint32_t lo = ((addr & 0x3FFFFu) >> 4) | (1u << 16);          // start address, LBO = 1 (unused)
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);          // SBO, descriptor version, SWIZZLE_128B
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
// kind::tf32, fp32 accumulate, both operands K-major, M = 128, N = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((TC_N >> 3) << 17) | ((TC_M >> 4) << 24);

#define B2_TMEM_LD32(taddr, v)                                                                                   \
  asm volatile(                                                                                                  \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                  \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                 \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                  \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),         \
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),   \
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])  \
      : "r"(taddr)                                                                                               \
      : "memory")

// ------------------------------------------------------------------------------------------ operand preparation
// rows -> (hi, lo) split, zero padded to [rows_pad][dp32]; optional row gather (support vectors), norms of the
// ORIGINAL float32 values (they feed the rbf epilogue exactly like the FFMA path)
__global__ void split_rows_kernel(const float* __restrict__ X, int64_t ld, int d, int dp, const int* __restrict__ idx,
                                  int64_t rows, int64_t rows_pad, float* __restrict__ hi, float* __restrict__ lo,
                                  float* __restrict__ norms, float norm_scale) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp; k < rows_pad; k += nwarps) {
    const int64_t r = k < rows ? (idx ? (int64_t)idx[k] : k) : -1;
    float s = 0.f;
    for (int c = lane; c < dp; c += 32) {
      const float v = (r >= 0 && c < d) ? X[r * ld + c] : 0.f;
      const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
      hi[k * dp + c] = h;
      lo[k * dp + c] = v - h;
      s += v * v;
    }
    s = warp_sum(s);
    if (lane == 0 && norms) norms[k] = s * norm_scale;
  }
}

// float64 rows (the RFF projection matrix W): hi = float(w) with the low 13 bits cleared, lo = float(w - hi)
__global__ void split_rows_f64_kernel(const double* __restrict__ W, int d, int dp, int64_t rows, int64_t rows_pad,
                                      float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t total = rows_pad * dp;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = e / dp;
    const int c = (int)(e - k * dp);
    const double w = (k < rows && c < d) ? W[k * d + c] : 0.0;
    const float h = __uint_as_float(__float_as_uint((float)w) & 0xffffe000u);
    hi[e] = h;
    lo[e] = (float)(w - (double)h);
  }
}

__global__ void rff_meta_kernel(const double* __restrict__ b, const double* __restrict__ wz, int D, int D_pad, double scale,
                                float* __restrict__ phase, double* __restrict__ coef) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < D_pad; k += gridDim.x * blockDim.x) {
    phase[k] = k < D ? (float)b[k] : 0.f;
    coef[k] = k < D ? wz[k] * scale : 0.0;
  }
}

__global__ void gather_coef_kernel(const double* __restrict__ coef, const int* __restrict__ idx, int n, int n_pad,
                                   double* __restrict__ out) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_pad; k += gridDim.x * blockDim.x)
    out[k] = k < n ? coef[idx[k]] : 0.0;
}

// ------------------------------------------------------------------------------------------ the kernel
struct TcParams {
  const float* svn;      // [nsv_pad] squared norms of the support rows
  const double* svc;     // [nsv_pad] coefficients (0 in the padding)
  const float* qn;       // [nq_pad]
  double* partial;       // [slices][nq]
  int64_t nq;
  int nsv;               // true number of support rows
  int tiles_total;       // support tiles of 128 rows
  int tiles_per_slice;   // blockIdx.y * tiles_per_slice .. + tiles_per_slice
  int kb;                // K blocks of 32 floats (1 .. 4)
  int stages;            // B pipeline depth
  float c2;              // rbf: 2 gamma log2(e)
  int debug;             // B2SVM_TC_DEBUG: 1 = epilogue adds raw dots, 2 = no MMAs, 4 = no TMA of B
  KernelParams<float> kp;
};

// Epilogue value of one (query, support row) pair from the tensor-core dot product.
// rbf: exp(-gamma (|a|^2 + |b|^2 - 2 a.b)) = 2^(c2 * a.b + na' + nb') with the norms pre-scaled by -gamma log2(e)
// (split_rows_kernel) and c2 = 2 gamma log2(e): one add, one fma, one MUFU.EX2.  The rounding of the exponent
// is |x| 2^-24 relative -- the same order as the float32 rounding of -gamma * dist in the reference (svc.py:230-232).
// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f with the 1.5 * 2^23 trick, 2^f by the Cephes
// exp2f polynomial on [-0.5, 0.5] (relative error ~1e-7, the same order as MUFU.EX2's 2^-22), 2^n by an integer add
// into the exponent field.  ~12 issue slots against 1 for ex2.approx -- used for ONE IN FOUR pairs of the rbf epilogue,
// which looked bound by the MUFU pipe (XU 62 % busy, FMA 23 %: profiles/r2_matvec_tc_ncu_full.md).  MEASURED SLOWER
// (1M x 500k, d = 32: 202.5 -> 226.3 ms): the epilogue is issue / latency bound (16 warps, 28 % warps active), so the
// extra issue slots cost more than the MUFU slots they free.  Kept for the record, not used (POLY = false everywhere).
__device__ __forceinline__ float exp2_fma(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = 1.535336188319500e-4f;
  p = fmaf(p, f, 1.339887440266574e-3f);
  p = fmaf(p, f, 9.618437357674640e-3f);
  p = fmaf(p, f, 5.550332471162809e-2f);
  p = fmaf(p, f, 2.402264791363012e-1f);
  p = fmaf(p, f, 6.931472028550421e-1f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int KIND, bool POLY = false>
__device__ __forceinline__ float epi_value(float c2, const KernelParams<float>& kp, uint32_t dot_bits, float norm_s, float norm_q) {
  const float dot = __uint_as_float(dot_bits);
  if constexpr (KIND == K_RBF) {
    const float x = fmaf(c2, dot, norm_s + norm_q);
    if constexpr (POLY) return exp2_fma(x);
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
  } else if constexpr (KIND == K_RFFCOS) {
    // cos(w_s . x_q + b_s): reduce to [-pi, pi] with a two-term 2 pi, then MUFU.COS (abs error ~5e-7)
    const float t = dot + norm_s;
    const float k = rintf(t * 0.15915494309189535f);
    float u = fmaf(k, -6.2831854820251465f, t);
    u = fmaf(k, 1.7484555e-7f, u);
    float r;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(u));
    return r;
  } else {
    return kernel_value_k<KIND>(kp, dot, norm_s, norm_q);
  }
}

// 32 accumulator columns (registers v) against the 32 support rows whose metadata (norm / phase, coefficient) the
// warp staged in shared memory (n4, c2): broadcast vector reads, two independent float64 accumulation chains
// 32 accumulator columns (registers v) against the 32 support rows whose metadata (norm / phase, coefficient) the
// warp staged in shared memory: broadcast vector reads.  The 32 products coef * k are summed in float32 (four
// independent chains, coefficients rounded to float32) and the chunk sum is added to the float64 accumulator: the
// rounding this adds (~1e-7 of the chunk) is below the 3xTF32 error of the dot products, and the per-pair work drops
// from {FADD, FFMA, MUFU, 3 integer ops, DFMA} to {FADD, FFMA, MUFU, FFMA} -- the epilogue is issue- and MUFU-bound.
template <int KIND>
__device__ __forceinline__ double accum32(const TcParams& p, const uint32_t (&v)[32], const float4* n4, const float4* c4,
                                          float nq) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 n = n4[j];                                      // shared-memory broadcast reads
    const float4 c = c4[j];
    s0 = fmaf(c.x, epi_value<KIND>(p.c2, p.kp, v[4 * j + 0], n.x, nq), s0);
    s1 = fmaf(c.y, epi_value<KIND>(p.c2, p.kp, v[4 * j + 1], n.y, nq), s1);
    s2 = fmaf(c.z, epi_value<KIND>(p.c2, p.kp, v[4 * j + 2], n.z, nq), s2);
    s3 = fmaf(c.w, epi_value<KIND>(p.c2, p.kp, v[4 * j + 3], n.w, nq), s3);
  }
  return (double)((s0 + s1) + (s2 + s3));
}

// MT = query tiles (of 128 rows) per CTA.  Every support tile that TMA brings in is multiplied against MT query
// tiles, so the L2 -> shared-memory traffic per kernel evaluation drops by MT: at d = 32 one CTA with MT = 1 needs
// 32 KB of support data per 16 384 evaluations (~1000 MUFU cycles), i.e. ~9 TB/s over 148 SMs -- more than the L2
// delivers; MT = 2 halves that.  TMEM: MT x 2 accumulators of 128 columns (512 columns = all of it at MT = 2).
template <int KIND, int MT>
__global__ void __launch_bounds__(TC_THREADS, 1)
matvec_tc_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                 const __grid_constant__ CUtensorMap tm_s_hi, const __grid_constant__ CUtensorMap tm_s_lo, TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_tc[];
  const int KB = p.kb, ST = p.stages;
  // [A_hi MT*KB blocks][A_lo MT*KB blocks][stage: one K block of B_hi, B_lo] ...
  unsigned char* a_hi = smem_tc;
  unsigned char* a_lo = a_hi + (size_t)MT * KB * TC_TILE_BYTES;
  unsigned char* b_base = a_lo + (size_t)MT * KB * TC_TILE_BYTES;
  const size_t b_stage_bytes = (size_t)2 * TC_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)ST * b_stage_bytes);
  uint64_t* a_full = bars;            // 1
  uint64_t* b_full = bars + 1;        // ST
  uint64_t* b_empty = b_full + ST;    // ST
  uint64_t* t_full = b_empty + ST;    // 2
  uint64_t* t_empty = t_full + 2;     // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  unsigned char* meta = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);   // [epilogue warp][32 norms | 32 coefficients] float32, 16-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TC_M * MT;
  const int tile_begin = blockIdx.y * p.tiles_per_slice;
  const int tile_end = min(p.tiles_total, tile_begin + p.tiles_per_slice);
  const int ntile = max(0, tile_end - tile_begin);
  constexpr uint32_t TMEM_COLS = MT == 1 ? 256 : 512;

  if (threadIdx.x == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < ST; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&t_full[a], 1); mbar_init(&t_empty[a], TC_EPI_WARPS); }
    mbar_fence_init();
  }
  if (warp == 1) {   // TMEM: MT x two accumulators of 128 fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(a_full, (uint32_t)(2 * MT * KB * TC_TILE_BYTES));
      for (int mt = 0; mt < MT; ++mt)
        for (int kb = 0; kb < KB; ++kb) {
          tma_load_2d(a_hi + (size_t)(mt * KB + kb) * TC_TILE_BYTES, &tm_q_hi, kb * TC_KBLK, q0 + mt * TC_M, a_full);
          tma_load_2d(a_lo + (size_t)(mt * KB + kb) * TC_TILE_BYTES, &tm_q_lo, kb * TC_KBLK, q0 + mt * TC_M, a_full);
        }
      // one pipeline stage = one K block (32 floats) of a support tile, hi and lo copies
      for (int step = 0; step < ntile * KB; ++step) {
        const int s = step % ST, it = step / KB, kb = step - it * KB;
        const uint32_t ph = (uint32_t)((step / ST) & 1);
        mbar_wait(&b_empty[s], ph ^ 1u);
        unsigned char* bh = b_base + (size_t)s * b_stage_bytes;
        if (p.debug & 4) { mbar_arrive(&b_full[s]); continue; }
        mbar_arrive_expect_tx(&b_full[s], (uint32_t)b_stage_bytes);
        const int row = (tile_begin + it) * TC_N;
        tma_load_2d(bh, &tm_s_hi, kb * TC_KBLK, row, &b_full[s]);
        tma_load_2d(bh + TC_TILE_BYTES, &tm_s_lo, kb * TC_KBLK, row, &b_full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected thread) =====
    if (lane == 0) {
      mbar_wait(a_full, 0);
      int step = 0;
      for (int it = 0; it < ntile; ++it) {
        const int acc = it & 1;
        mbar_wait(&t_empty[acc], (uint32_t)(((it >> 1) & 1) ^ 1));   // epilogue drained this accumulator set
        uint32_t accumulate = 0;
        for (int kb = 0; kb < KB; ++kb, ++step) {
          const int s = step % ST;
          mbar_wait(&b_full[s], (uint32_t)((step / ST) & 1));
          tc_fence_after();
          const unsigned char* bh = b_base + (size_t)s * b_stage_bytes;
          const unsigned char* bl = bh + TC_TILE_BYTES;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((mt * 2 + acc) * TC_N);
            uint32_t accm = accumulate;
            // lo.hi + hi.lo + hi.hi per K block, small terms first
            for (int pass = (p.debug & 2) ? 3 : 0; pass < 3; ++pass) {
              const uint64_t da = umma_desc_sw128((pass == 0 ? a_lo : a_hi) + (size_t)(mt * KB + kb) * TC_TILE_BYTES);
              const uint64_t db = umma_desc_sw128(pass == 1 ? bl : bh);
#pragma unroll
              for (int k = 0; k < TC_KBLK / 8; ++k) {   // UMMA_K = 8 tf32 = 32 bytes: advance the start address
                tc_mma_tf32(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), TC_IDESC, accm);
                accm = 1;
              }
            }
          }
          accumulate = 1;
          tc_commit(&b_empty[s]);   // the smem stage may be refilled once these MMAs have read it
        }
        tc_commit(&t_full[acc]);    // the accumulators are complete
      }
    }
  } else {
    // ===== epilogue warps: TMEM lane = query row; the four warps of a lane quarter split the tile's columns =====
    const int quarter = warp & 3;                       // a warp may only touch TMEM lanes 32*(warp%4) ..
    const int part = (warp - 2) >> 2;                   // columns part*32 .. +32 of the tile
    const int qrow = quarter * 32 + lane;
    float nq[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) nq[mt] = p.qn[q0 + mt * TC_M + qrow];
    // the metadata of this warp's 32 support rows is fetched one tile ahead (coalesced, one value per lane) and
    // staged in shared memory: the per-pair reads are then short-latency broadcasts instead of L1/L2 round trips
    float* wn = reinterpret_cast<float*>(meta + (size_t)(warp - 2) * 256);
    float* wc = wn + 32;
    float mn = 0.f, mc = 0.f;
    if (ntile > 0) {
      mn = __ldg(p.svn + tile_begin * TC_N + part * 32 + lane);
      mc = (float)__ldg(p.svc + tile_begin * TC_N + part * 32 + lane);
    }
    double sum[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) sum[mt] = 0.0;
    for (int it = 0; it < ntile; ++it) {
      const int acc = it & 1;
      mbar_wait(&t_full[acc], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const int s0 = (tile_begin + it) * TC_N + part * 32;
      wn[lane] = mn;
      wc[lane] = mc;
      __syncwarp();
      if (it + 1 < ntile) {
        mn = __ldg(p.svn + s0 + TC_N + lane);
        mc = (float)__ldg(p.svc + s0 + TC_N + lane);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((mt * 2 + acc) * TC_N + part * 32);
        B2_TMEM_LD32(taddr, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (mt == MT - 1) {
          // the last accumulator slice is in registers: hand the TMEM buffers back before the math
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[acc]);
        }
        if (p.debug & 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[mt] += (double)__uint_as_float(v[j]);
          continue;
        }
        if (s0 + 32 <= p.nsv) {
          sum[mt] += accum32<KIND>(p, v, reinterpret_cast<const float4*>(wn), reinterpret_cast<const float4*>(wc), nq[mt]);
        } else {   // last, ragged tile: padding rows carry coefficient 0, but their kernel value need not be finite
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (s0 + j < p.nsv) sum[mt] = fma(__ldg(p.svc + s0 + j), (double)epi_value<KIND>(p.c2, p.kp, v[j], __ldg(p.svn + s0 + j), nq[mt]), sum[mt]);
        }
      }
      __syncwarp();   // every lane is done with the staged metadata before the next tile overwrites it
    }
    // the column parts of a query row meet in shared memory (the A tiles are dead by now: every MMA that read
    // them completed before the last t_full flipped)
    double* red = reinterpret_cast<double*>(a_hi);      // [MT][3][128]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
      if (part > 0) red[(mt * 3 + part - 1) * TC_M + qrow] = sum[mt];
    asm volatile("bar.sync 1, %0;" ::"r"(32 * TC_EPI_WARPS) : "memory");
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
      if (part == 0 && q0 + mt * TC_M + qrow < p.nq)
        p.partial[(size_t)blockIdx.y * p.nq + q0 + mt * TC_M + qrow] =
            ((sum[mt] + red[(mt * 3) * TC_M + qrow]) + red[(mt * 3 + 1) * TC_M + qrow]) + red[(mt * 3 + 2) * TC_M + qrow];
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

__global__ void reduce_partials_tc_kernel(const double* __restrict__ partial, int ny, int64_t nb, double* __restrict__ out) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nb; q += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int y = 0; y < ny; ++y) s += partial[(size_t)y * nb + q];
    out[q] = s;
  }
}

// ------------------------------------------------------------------------------------------ tile-output variant
// The same GEMM pipeline (TMA -> tcgen05.mma 3xTF32 -> TMEM), but the epilogue WRITES the tile: out[m][n] =
// f(a_m . b_n, row scalar, column scalar) -- random Fourier features sqrt(2/D) cos(x w^T + b) (rff.py:19-20) and the
// dense kernel matrix K(X, X) of the reference's dense-Q mode (svc.py:251-253).  Each epilogue warp owns 32 rows x 32
// columns of the accumulator; every thread converts its row's values, writes them into a 128-byte-swizzled staging
// tile in shared memory (conflict-free 16-byte stores) and one lane hands the tile to the TMA store engine
// (cp.async.bulk.tensor.2d.global.shared::cta, SASS UTMASTG): full-line writes, no LSU traffic, clipped to the
// tensor bounds by the hardware (ragged last tiles need no masks).  The kernel is bound by the HBM write stream.
struct TileParams {
  const float* coln;     // [n_pad] per-column scalar: phase b (rff) / squared norm pre-scaled (rbf)
  const float* rown;     // [m_pad] per-row scalar
  int tiles_total;       // column tiles of 128
  int tiles_per_slice;
  int kb, stages;
  float c2, scale;       // rbf exponent factor; output scale (rff: sqrt(2/D))
  int nbuf;              // staging buffers per epilogue warp (1 or 2)
  KernelParams<float> kp;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* smem_src, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(c_inner), "r"(c_outer), "r"(smem_u32(smem_src))
               : "memory");
}

template <int KIND, typename OUT>
__global__ void __launch_bounds__(TC_THREADS, 1)
tile_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
               const __grid_constant__ CUtensorMap tm_out, TileParams p) {
  extern __shared__ __align__(1024) unsigned char smem_tc[];
  const int KB = p.kb, ST = p.stages;
  constexpr int BOXC = 128 / (int)sizeof(OUT);        // columns per 128-byte staging row: 16 doubles / 32 floats
  constexpr int NBOX = 32 / BOXC;                      // TMA stores per 32-column part: 2 / 1
  unsigned char* a_hi = smem_tc;
  unsigned char* a_lo = a_hi + (size_t)KB * TC_TILE_BYTES;
  unsigned char* b_base = a_lo + (size_t)KB * TC_TILE_BYTES;
  const size_t b_stage_bytes = (size_t)2 * TC_TILE_BYTES;
  unsigned char* stage_base = b_base + (size_t)ST * b_stage_bytes;                 // [epilogue warp][nbuf][32 rows][128 B], 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + (size_t)TC_EPI_WARPS * p.nbuf * 4096);
  uint64_t* a_full = bars;
  uint64_t* b_full = bars + 1;
  uint64_t* b_empty = b_full + ST;
  uint64_t* t_full = b_empty + ST;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  unsigned char* meta = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_M;
  const int tile_begin = blockIdx.y * p.tiles_per_slice;
  const int tile_end = min(p.tiles_total, tile_begin + p.tiles_per_slice);
  const int ntile = max(0, tile_end - tile_begin);

  if (threadIdx.x == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < ST; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&t_full[a], 1); mbar_init(&t_empty[a], TC_EPI_WARPS); }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {   // ===== TMA producer (as in matvec_tc_kernel)
      mbar_arrive_expect_tx(a_full, (uint32_t)(2 * KB * TC_TILE_BYTES));
      for (int kb = 0; kb < KB; ++kb) {
        tma_load_2d(a_hi + (size_t)kb * TC_TILE_BYTES, &tm_a_hi, kb * TC_KBLK, m0, a_full);
        tma_load_2d(a_lo + (size_t)kb * TC_TILE_BYTES, &tm_a_lo, kb * TC_KBLK, m0, a_full);
      }
      for (int step = 0; step < ntile * KB; ++step) {
        const int s = step % ST, it = step / KB, kb = step - it * KB;
        mbar_wait(&b_empty[s], (uint32_t)(((step / ST) & 1) ^ 1));
        unsigned char* bh = b_base + (size_t)s * b_stage_bytes;
        mbar_arrive_expect_tx(&b_full[s], (uint32_t)b_stage_bytes);
        const int row = (tile_begin + it) * TC_N;
        tma_load_2d(bh, &tm_b_hi, kb * TC_KBLK, row, &b_full[s]);
        tma_load_2d(bh + TC_TILE_BYTES, &tm_b_lo, kb * TC_KBLK, row, &b_full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ===== MMA issuer
      mbar_wait(a_full, 0);
      int step = 0;
      for (int it = 0; it < ntile; ++it) {
        const int acc = it & 1;
        mbar_wait(&t_empty[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_N);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < KB; ++kb, ++step) {
          const int s = step % ST;
          mbar_wait(&b_full[s], (uint32_t)((step / ST) & 1));
          tc_fence_after();
          const unsigned char* bh = b_base + (size_t)s * b_stage_bytes;
          const unsigned char* bl = bh + TC_TILE_BYTES;
          for (int pass = 0; pass < 3; ++pass) {   // lo.hi + hi.lo + hi.hi, small terms first
            const uint64_t da = umma_desc_sw128((pass == 0 ? a_lo : a_hi) + (size_t)kb * TC_TILE_BYTES);
            const uint64_t db = umma_desc_sw128(pass == 1 ? bl : bh);
#pragma unroll
            for (int k = 0; k < TC_KBLK / 8; ++k) {
              tc_mma_tf32(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), TC_IDESC, accumulate);
              accumulate = 1;
            }
          }
          tc_commit(&b_empty[s]);
        }
        tc_commit(&t_full[acc]);
      }
    }
  } else {
    // ===== epilogue warps: TMEM lane = output row; warp (quarter, part) owns rows quarter*32.. and columns part*32..
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int part = ew >> 2;
    const int mrow = quarter * 32 + lane;
    const float rs = p.rown ? p.rown[m0 + mrow] : 0.f;
    float* wn = reinterpret_cast<float*>(meta + (size_t)ew * 128);
    unsigned char* stg = stage_base + (size_t)ew * p.nbuf * 4096;
    float mn = 0.f;
    if (ntile > 0) mn = __ldg(p.coln + tile_begin * TC_N + part * 32 + lane);
    int buf = 0;
    for (int it = 0; it < ntile; ++it) {
      const int acc = it & 1;
      mbar_wait(&t_full[acc], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const int n0 = (tile_begin + it) * TC_N + part * 32;
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TC_N + part * 32);
      B2_TMEM_LD32(taddr, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);     // accumulator slice is in registers: hand the TMEM buffer back
      wn[lane] = mn;
      __syncwarp();
      if (it + 1 < ntile) mn = __ldg(p.coln + n0 + TC_N + lane);
#pragma unroll
      for (int b = 0; b < NBOX; ++b) {
        unsigned char* tile = stg + (size_t)buf * 4096;
        // the staging tile may still be read by the TMA store issued nbuf boxes ago
        if (lane == 0) {
          if (p.nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
        // my row: BOXC values = 8 chunks of 16 bytes, chunk position XORed with (row & 7) (the 128-byte TMA swizzle)
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          constexpr int EPC = 16 / (int)sizeof(OUT);     // values per 16-byte chunk: 2 doubles / 4 floats
          OUT o[EPC];
#pragma unroll
          for (int e = 0; e < EPC; ++e) {
            const int j = b * BOXC + ch * EPC + e;
            o[e] = (OUT)(p.scale * epi_value<KIND>(p.c2, p.kp, v[j], wn[j], rs));
          }
          *reinterpret_cast<uint4*>(tile + lane * 128 + ((ch ^ (lane & 7)) << 4)) = *reinterpret_cast<const uint4*>(o);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tm_out, tile, n0 + b * BOXC, m0 + quarter * 32);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        buf = (buf + 1 == p.nbuf) ? 0 : buf + 1;
      }
      __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// Work buffers come from the stream-ordered allocator and its pool keeps them between calls: a prediction over
// millions of queries is many calls, and cudaMalloc / cudaFree (a device-wide synchronisation each) cost more than
// the kernel.
void keep_pool_memory() {
  static bool done = false;
  if (done) return;
  done = true;
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = 4ull << 30;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
}

static bool make_map(CUtensorMap* m, const float* base, int64_t rows, int dp) {
  const cuuint64_t dims[2] = {(cuuint64_t)dp, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)dp * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)TC_KBLK, (cuuint32_t)TC_M};
  const cuuint32_t estr[2] = {1, 1};
  return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int matvec_tc_mode(int x_dtype, int d) {
  if (x_dtype != B2SVM_F32 || d > 128) return 0;
  int mode = 1;
  if (const char* e = getenv("B2SVM_MATVEC_TC")) mode = atoi(e);
  if (mode <= 0) return 0;
  return encode_fn() != nullptr ? std::min(mode, 2) : 0;
}

// idx / nsv: ordered list of rows of Xa with a non-zero coefficient (device), from the caller's compaction.
// rff != nullptr: the "support rows" are the D rows of the float64 projection matrix W, the per-row scalar is the
// phase b, the coefficient wz * sqrt(2 / D), and the epilogue is cos (b2svm_rff_decision on the tensor cores).
struct RffOperands {
  const double* W;
  const double* b;
  const double* wz;
};

static int matvec_tc_core(cudaStream_t stream, const b2svm_kernel_desc& k, int d, const float* Xa, int64_t lda, const int* idx,
                          int nsv, const double* coef, const RffOperands* rff, const float* Xb, int64_t nb, int64_t ldb,
                          double* out) {
  keep_pool_memory();
  const int dp = (d + TC_KBLK - 1) / TC_KBLK * TC_KBLK;
  const int KB = dp / TC_KBLK;
  const int64_t nsv_pad = ((int64_t)nsv + TC_N - 1) / TC_N * TC_N;
  // two query tiles per CTA when their operands fit beside a useful B pipeline (K blocks <= 2, i.e. d <= 64) and there
  // are enough queries to fill the GPU twice over: halves the support-tile traffic per kernel evaluation
  int MT = (KB <= 2 && nb >= 2 * 148 * 2 * TC_M) ? 2 : 1;
  if (const char* e = getenv("B2SVM_TC_MT")) MT = std::max(1, std::min(KB <= 2 ? 2 : 1, atoi(e)));
  const int64_t nq_pad = (nb + TC_M * MT - 1) / (TC_M * MT) * (TC_M * MT);
  float *s_hi = nullptr, *s_lo = nullptr, *q_hi = nullptr, *q_lo = nullptr, *svn = nullptr, *qn = nullptr;
  double *svc = nullptr, *partial = nullptr;
  auto done = [&](int code) {
    for (void* ptr : {(void*)s_hi, (void*)s_lo, (void*)q_hi, (void*)q_lo, (void*)svn, (void*)qn, (void*)svc, (void*)partial})
      if (ptr) cudaFreeAsync(ptr, stream);
    return code;
  };
#define TC_TRY(expr)                                                                                             \
  do {                                                                                                           \
    cudaError_t _e = (expr);                                                                                     \
    if (_e != cudaSuccess) return done(fail(B2SVM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)));  \
  } while (0)
  TC_TRY(cudaMallocAsync(&s_hi, sizeof(float) * nsv_pad * dp, stream));
  TC_TRY(cudaMallocAsync(&s_lo, sizeof(float) * nsv_pad * dp, stream));
  TC_TRY(cudaMallocAsync(&q_hi, sizeof(float) * nq_pad * dp, stream));
  TC_TRY(cudaMallocAsync(&q_lo, sizeof(float) * nq_pad * dp, stream));
  TC_TRY(cudaMallocAsync(&svn, sizeof(float) * nsv_pad, stream));
  TC_TRY(cudaMallocAsync(&qn, sizeof(float) * nq_pad, stream));
  TC_TRY(cudaMallocAsync(&svc, sizeof(double) * nsv_pad, stream));
  const float norm_scale = k.kind == K_RBF ? (float)(-(double)(float)k.gamma * 1.4426950408889634) : 1.0f;
  auto grid_rows = [](int64_t rows) { return (int)std::max<int64_t>(1, std::min<int64_t>(148 * 8, (rows * 32 + 255) / 256)); };
  if (rff) {
    split_rows_f64_kernel<<<grid_rows(nsv_pad), 256, 0, stream>>>(rff->W, d, dp, nsv, nsv_pad, s_hi, s_lo);
    rff_meta_kernel<<<std::max(1, (int)((nsv_pad + 255) / 256)), 256, 0, stream>>>(rff->b, rff->wz, nsv, (int)nsv_pad,
                                                                               std::sqrt(2.0 / (double)nsv), svn, svc);
  } else {
    split_rows_kernel<<<grid_rows(nsv_pad), 256, 0, stream>>>(Xa, lda, d, dp, idx, nsv, nsv_pad, s_hi, s_lo, svn, norm_scale);
    gather_coef_kernel<<<std::max(1, (int)std::min<int64_t>(1184, (nsv_pad + 255) / 256)), 256, 0, stream>>>(coef, idx, nsv, (int)nsv_pad, svc);
  }
  split_rows_kernel<<<grid_rows(nq_pad), 256, 0, stream>>>(Xb, ldb, d, dp, nullptr, nb, nq_pad, q_hi, q_lo, qn, rff ? 0.f : norm_scale);
  TC_TRY(cudaGetLastError());

  CUtensorMap m_qh, m_ql, m_sh, m_sl;
  if (!make_map(&m_qh, q_hi, nq_pad, dp) || !make_map(&m_ql, q_lo, nq_pad, dp) || !make_map(&m_sh, s_hi, nsv_pad, dp) ||
      !make_map(&m_sl, s_lo, nsv_pad, dp))
    return done(fail(B2SVM_E_CUDA, "cuTensorMapEncodeTiled failed"));

  TcParams p;
  p.svn = svn; p.svc = svc; p.qn = qn; p.nq = nb; p.nsv = nsv;
  p.tiles_total = (int)(nsv_pad / TC_N);
  const int gx = (int)(nq_pad / (TC_M * MT));
  int ny = std::max(1, std::min(p.tiles_total, (2 * 148 + gx - 1) / gx));
  // one slice of support rows (hi + lo copies) should sit in L2 while the query tiles sweep over it
  const int l2_tiles = std::max(1, (int)((32u << 20) / (2u * KB * TC_TILE_BYTES)));
  ny = std::max(ny, (p.tiles_total + l2_tiles - 1) / l2_tiles);
  ny = std::min(ny, 65535);
  p.tiles_per_slice = (p.tiles_total + ny - 1) / ny;
  ny = (p.tiles_total + p.tiles_per_slice - 1) / p.tiles_per_slice;
  p.kb = KB;
  p.c2 = (float)(2.0 * (double)(float)k.gamma * 1.4426950408889634);
  p.debug = getenv("B2SVM_TC_DEBUG") ? atoi(getenv("B2SVM_TC_DEBUG")) : 0;
  const size_t a_bytes = (size_t)2 * MT * KB * TC_TILE_BYTES, b_stage = (size_t)2 * TC_TILE_BYTES;
  p.stages = (int)std::min<size_t>(6, (200 * 1024 - a_bytes) / b_stage);
  if (p.stages < 2) return done(fail(B2SVM_E_UNSUPPORTED, "kernel_matvec (tensor path): n_features too large"));
  p.kp = KernelParams<float>{k.kind, (float)k.gamma, (float)k.coef0, (float)k.degree};
  TC_TRY(cudaMallocAsync(&partial, sizeof(double) * (size_t)ny * nb, stream));
  p.partial = partial;
  const size_t smem = a_bytes + p.stages * b_stage + (size_t)(2 * p.stages + 8) * sizeof(uint64_t) + 32 + TC_EPI_WARPS * 256 + 1024;
#define TC_LAUNCH2(KIND, MTV)                                                                                             \
  do {                                                                                                                    \
    TC_TRY(cudaFuncSetAttribute(matvec_tc_kernel<KIND, MTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    matvec_tc_kernel<KIND, MTV><<<dim3(gx, ny), TC_THREADS, smem, stream>>>(m_qh, m_ql, m_sh, m_sl, p);                   \
  } while (0)
#define TC_LAUNCH(KIND)              \
  do {                               \
    if (MT == 2) TC_LAUNCH2(KIND, 2); \
    else TC_LAUNCH2(KIND, 1);         \
  } while (0)
  switch (rff ? K_RFFCOS : k.kind) {
    case K_RFFCOS: TC_LAUNCH(K_RFFCOS); break;
    case K_RBF: TC_LAUNCH(K_RBF); break;
    case K_POLY: TC_LAUNCH(K_POLY); break;
    case K_SIGMOID: TC_LAUNCH(K_SIGMOID); break;
    default: TC_LAUNCH(K_LINEAR); break;
  }
#undef TC_LAUNCH
#undef TC_LAUNCH2
  TC_TRY(cudaGetLastError());
  reduce_partials_tc_kernel<<<std::max(1, (int)std::min<int64_t>(1184, (nb + 255) / 256)), 256, 0, stream>>>(partial, ny, nb, out);
  TC_TRY(cudaGetLastError());
  TC_TRY(cudaStreamSynchronize(stream));
#undef TC_TRY
  return done(B2SVM_OK);
}

__global__ void phase_kernel(const double* __restrict__ b, int D, int D_pad, float* __restrict__ phase) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < D_pad; k += gridDim.x * blockDim.x) phase[k] = k < D ? (float)b[k] : 0.f;
}

static bool make_out_map(CUtensorMap* m, void* base, int64_t rows, int64_t cols, int64_t ld, bool f64) {
  const size_t esz = f64 ? 8 : 4;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / esz), 32u};
  const cuuint32_t estr[2] = {1, 1};
  return encode_fn()(m, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// can the tile-output tensor path take this shape?  (float32 operands, K blocks that leave room for the staging tiles,
// a 16-byte friendly output pitch)
bool tile_tc_ok(int d, int64_t ldo, bool out_f64, const void* out) {
  if (encode_fn() == nullptr || d > 96) return false;
  const size_t esz = out_f64 ? 8 : 4;
  return (ldo * esz) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
}

// out[m][n] = scale * f(a_m . b_n): f = cos(. + phase_n) (W / b given as float64: NormalRFF.transform) or the kernel
// function of `k` (Bf rows float32: the dense kernel matrix).  out is float64 or float32 with row pitch ldo (elements).
int tile_tc(cudaStream_t stream, const b2svm_kernel_desc& k, int d, const float* A, int64_t lda, int64_t m, const float* Bf,
            int64_t ldb, const double* Wd, const double* phase_d, int64_t n, double scale, void* out, int64_t ldo, bool out_f64) {
  keep_pool_memory();
  const bool rff = Wd != nullptr;
  const int dp = (d + TC_KBLK - 1) / TC_KBLK * TC_KBLK;
  const int KB = dp / TC_KBLK;
  const int64_t m_pad = (m + TC_M - 1) / TC_M * TC_M, n_pad = (n + TC_N - 1) / TC_N * TC_N;
  float *a_hi = nullptr, *a_lo = nullptr, *b_hi = nullptr, *b_lo = nullptr, *rown = nullptr, *coln = nullptr;
  auto done = [&](int code) {
    for (void* ptr : {(void*)a_hi, (void*)a_lo, (void*)b_hi, (void*)b_lo, (void*)rown, (void*)coln})
      if (ptr) cudaFreeAsync(ptr, stream);
    return code;
  };
#define TT_TRY(expr)                                                                                             \
  do {                                                                                                           \
    cudaError_t _e = (expr);                                                                                     \
    if (_e != cudaSuccess) return done(fail(B2SVM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)));  \
  } while (0)
  TT_TRY(cudaMallocAsync(&a_hi, sizeof(float) * m_pad * dp, stream));
  TT_TRY(cudaMallocAsync(&a_lo, sizeof(float) * m_pad * dp, stream));
  TT_TRY(cudaMallocAsync(&b_hi, sizeof(float) * n_pad * dp, stream));
  TT_TRY(cudaMallocAsync(&b_lo, sizeof(float) * n_pad * dp, stream));
  TT_TRY(cudaMallocAsync(&rown, sizeof(float) * m_pad, stream));
  TT_TRY(cudaMallocAsync(&coln, sizeof(float) * n_pad, stream));
  const float norm_scale = (!rff && k.kind == K_RBF) ? (float)(-(double)(float)k.gamma * 1.4426950408889634) : 1.0f;
  auto grid_rows = [](int64_t rows) { return (int)std::max<int64_t>(1, std::min<int64_t>(148 * 8, (rows * 32 + 255) / 256)); };
  split_rows_kernel<<<grid_rows(m_pad), 256, 0, stream>>>(A, lda, d, dp, nullptr, m, m_pad, a_hi, a_lo, rown, norm_scale);
  if (rff) {
    split_rows_f64_kernel<<<grid_rows(n_pad), 256, 0, stream>>>(Wd, d, dp, n, n_pad, b_hi, b_lo);
    phase_kernel<<<std::max(1, (int)((n_pad + 255) / 256)), 256, 0, stream>>>(phase_d, (int)n, (int)n_pad, coln);
  } else {
    split_rows_kernel<<<grid_rows(n_pad), 256, 0, stream>>>(Bf, ldb, d, dp, nullptr, n, n_pad, b_hi, b_lo, coln, norm_scale);
  }
  TT_TRY(cudaGetLastError());
  CUtensorMap m_ah, m_al, m_bh, m_bl, m_out;
  if (!make_map(&m_ah, a_hi, m_pad, dp) || !make_map(&m_al, a_lo, m_pad, dp) || !make_map(&m_bh, b_hi, n_pad, dp) ||
      !make_map(&m_bl, b_lo, n_pad, dp) || !make_out_map(&m_out, out, m, n, ldo, out_f64))
    return done(fail(B2SVM_E_CUDA, "cuTensorMapEncodeTiled failed"));
  TileParams p;
  p.coln = coln;
  p.rown = rown;
  p.tiles_total = (int)(n_pad / TC_N);
  const int gx = (int)(m_pad / TC_M);
  int ny = std::max(1, std::min(p.tiles_total, (2 * 148 + gx - 1) / gx));
  ny = std::min(ny, 65535);
  p.tiles_per_slice = (p.tiles_total + ny - 1) / ny;
  ny = (p.tiles_total + p.tiles_per_slice - 1) / p.tiles_per_slice;
  p.kb = KB;
  p.c2 = (float)(2.0 * (double)(float)k.gamma * 1.4426950408889634);
  p.scale = (float)scale;
  p.kp = KernelParams<float>{k.kind, (float)k.gamma, (float)k.coef0, (float)k.degree};
  const size_t a_bytes = (size_t)2 * KB * TC_TILE_BYTES, b_stage = (size_t)2 * TC_TILE_BYTES;
  const size_t fixed = (size_t)(2 * 6 + 8) * sizeof(uint64_t) + 32 + TC_EPI_WARPS * 128 + 1024;
  const size_t budget = 227 * 1024;
  p.nbuf = (a_bytes + 2 * b_stage + (size_t)TC_EPI_WARPS * 2 * 4096 + fixed <= budget) ? 2 : 1;
  const size_t stage_bytes = (size_t)TC_EPI_WARPS * p.nbuf * 4096;
  if (a_bytes + 2 * b_stage + stage_bytes + fixed > budget) return done(fail(B2SVM_E_UNSUPPORTED, "tile_tc: n_features too large"));
  p.stages = (int)std::min<size_t>(6, (budget - a_bytes - stage_bytes - fixed) / b_stage);
  const size_t smem = a_bytes + p.stages * b_stage + stage_bytes + fixed;
#define TT_LAUNCH(KIND, OUT)                                                                                         \
  do {                                                                                                               \
    TT_TRY(cudaFuncSetAttribute(tile_tc_kernel<KIND, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    tile_tc_kernel<KIND, OUT><<<dim3(gx, ny), TC_THREADS, smem, stream>>>(m_ah, m_al, m_bh, m_bl, m_out, p);          \
  } while (0)
#define TT_KIND(KIND)                                \
  do {                                               \
    if (out_f64) TT_LAUNCH(KIND, double);            \
    else TT_LAUNCH(KIND, float);                     \
  } while (0)
  switch (rff ? K_RFFCOS : k.kind) {
    case K_RFFCOS: TT_KIND(K_RFFCOS); break;
    case K_RBF: TT_KIND(K_RBF); break;
    case K_POLY: TT_KIND(K_POLY); break;
    case K_SIGMOID: TT_KIND(K_SIGMOID); break;
    default: TT_KIND(K_LINEAR); break;
  }
#undef TT_KIND
#undef TT_LAUNCH
  TT_TRY(cudaGetLastError());
  TT_TRY(cudaStreamSynchronize(stream));
#undef TT_TRY
  return done(B2SVM_OK);
}

int matvec_tc(cudaStream_t stream, const b2svm_kernel_desc& k, int d, const float* Xa, int64_t lda, const int* idx, int nsv,
              const double* coef, const float* Xb, int64_t nb, int64_t ldb, double* out) {
  return matvec_tc_core(stream, k, d, Xa, lda, idx, nsv, coef, nullptr, Xb, nb, ldb, out);
}

int rff_decision_tc(cudaStream_t stream, int d, const double* W, const double* b, int D, const double* wz, const float* Xq,
                    int64_t nq, int64_t ldq, double* out) {
  b2svm_kernel_desc k{};
  k.kind = K_LINEAR;
  const RffOperands r{W, b, wz};
  return matvec_tc_core(stream, k, d, nullptr, 0, nullptr, D, nullptr, &r, Xq, nq, ldq, out);
}

}  // namespace b2svm
