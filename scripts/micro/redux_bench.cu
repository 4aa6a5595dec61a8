// Micro-benchmark: cycles per warp-wide arg-max (value double, lowest index wins ties) for three formulations.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I pysvm_b200/csrc scripts/micro/redux_bench.cu -o gpurun_out/redux_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "b2svm_device.cuh"
using namespace b2svm;

__device__ __forceinline__ void select_v2(bool is_max, Cand& c) {   // hi word first, vote, fall back on ties
  unsigned long long key = c.idx >= 0 ? dkey(c.g) : 0ull;
  if (!is_max && c.idx >= 0) key = ~key;
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  unsigned cand = __ballot_sync(0xffffffffu, hi == mh && c.idx >= 0);
  if (cand == 0) { c.g = 0.0; c.idx = -1; c.flags = 0; return; }
  if (__popc(cand) > 1) {
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    cand = __ballot_sync(0xffffffffu, hi == mh && lo == ml && c.idx >= 0);
    if (__popc(cand) > 1) {
      const unsigned mi = __reduce_min_sync(0xffffffffu, ((cand >> (threadIdx.x & 31)) & 1u) ? (unsigned)c.idx : 0xffffffffu);
      cand = __ballot_sync(0xffffffffu, ((cand >> (threadIdx.x & 31)) & 1u) && (unsigned)c.idx == mi);
    }
  }
  const int src = __ffs(cand) - 1;
  c.g = __shfl_sync(0xffffffffu, c.g, src);
  c.idx = __shfl_sync(0xffffffffu, c.idx, src);
  c.flags = __shfl_sync(0xffffffffu, c.flags, src);
}
__device__ __forceinline__ void select_shfl(bool is_max, Cand& c) {   // 5-round butterfly
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double g = __shfl_xor_sync(0xffffffffu, c.g, o);
    const int idx = __shfl_xor_sync(0xffffffffu, c.idx, o);
    const int f = __shfl_xor_sync(0xffffffffu, c.flags, o);
    const bool take = is_max ? better_max(g, idx, c.g, c.idx) : better_min(g, idx, c.g, c.idx);
    if (take) { c.g = g; c.idx = idx; c.flags = f; }
  }
}
template <int MODE>
__global__ void bench(double* out, long long* cyc, int iters, double seed) {
  Cand c;
  c.g = seed * ((threadIdx.x * 2654435761u) % 1000) + 0.001 * threadIdx.x;
  c.idx = threadIdx.x * 7 + 3;
  c.flags = threadIdx.x & 7;
  double acc = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    Cand m = c;
    m.g += acc * 1e-30;      // dependence chain across iterations
    if (MODE == 0) redux_select(0xffffffffu, true, m);
    else if (MODE == 1) select_v2(true, m);
    else select_shfl(true, m);
    acc += m.g + m.idx;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { *cyc = (t1 - t0) / iters; }
  out[threadIdx.x] = acc;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
  const char* names[3] = {"3 x CREDUX (current)", "hi-word CREDUX + vote + shfl", "shuffle butterfly"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) bench<0><<<1, 32>>>(out, cyc, 20000, 1.0);
      if (mode == 1) bench<1><<<1, 32>>>(out, cyc, 20000, 1.0);
      if (mode == 2) bench<2><<<1, 32>>>(out, cyc, 20000, 1.0);
      cudaDeviceSynchronize();
    }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-32s %lld cycles per select (one warp, dependent chain)\n", names[mode], h);
  }
  return 0;
}
