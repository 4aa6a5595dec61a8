// Latency of the ways one SM can read a 16-byte record another SM has just written (what the in-kernel exchange
// polls), on an L2-resident line.  One warp of CTA 0 measures; CTA 1 rewrites the lines between samples so that they
// are never L1-resident in CTA 0's SM.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a poll_bench.cu -o poll_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define LD4(NAME, PTX)                                                                                              \
  __device__ __forceinline__ uint4 NAME(const void* p) {                                                            \
    uint4 v;                                                                                                        \
    asm volatile(PTX " {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); \
    return v;                                                                                                       \
  }
LD4(ld_relaxed_gpu, "ld.relaxed.gpu.global.v4.u32")
LD4(ld_relaxed_sys, "ld.relaxed.sys.global.v4.u32")
LD4(ld_volatile, "ld.volatile.global.v4.u32")
LD4(ld_cg, "ld.global.cg.v4.u32")
LD4(ld_cv, "ld.global.cv.v4.u32")
LD4(ld_weak, "ld.global.v4.u32")
LD4(ld_acquire_gpu, "ld.acquire.gpu.global.v4.u32")
__device__ __forceinline__ void st_relaxed_gpu(void* p, uint4 v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

constexpr int NREC = 160, REPS = 64, NVAR = 20;

__global__ void __launch_bounds__(64, 1) bench(uint4* rec, uint4* scratch, double* gbuf, long long* out, volatile int* turn) {
  __shared__ __align__(128) uint4 sbuf[NREC * 2];
  __shared__ uint64_t bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (blockIdx.x == 1) {   // writer: refresh the records whenever asked, so they sit in L2 (not in CTA 0's L1)
    if (threadIdx.x == 0) {
      int seen = 0;
      while (seen < REPS * NVAR) {
        while (*turn <= seen) {}
        seen = *turn;
        for (int k = 0; k < NREC * 2; ++k) st_relaxed_gpu(rec + k, make_uint4(seen, k, seen, ~seen));
        __threadfence();
        *(turn + 32) = seen;
      }
    }
    return;
  }
  if (warp != 0) return;
  if (lane == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); }
  __syncwarp();
  uint32_t phase = 0;
  int tick = 0;
  for (int var = 0; var < NVAR; ++var) {
    long long acc = 0, mx = 0;
    for (int rep = 0; rep < REPS; ++rep) {
      ++tick;
      if (lane == 0) { *turn = tick; while (*(turn + 32) < tick) {} }
      __syncwarp();
      // optional preamble: what the kernel does right before polling
      if (var == 1 || var == 3) st_relaxed_gpu(scratch + lane, make_uint4(tick, 1, 2, 3));            // own strong store
      if (var == 2 || var == 3) for (int k = 0; k < 8; ++k) gbuf[(k * 32 + lane) * 8] = (double)tick;  // sweep-like weak stores
      __syncwarp();
      const long long t0 = clock64();
      uint4 v = make_uint4(0, 0, 0, 0), v2 = v, v3 = v, v4 = v, v5 = v;
      switch (var) {
        case 0: case 1: case 2: case 3: v = ld_relaxed_gpu(rec + 2 * lane); break;
        case 4: v = ld_relaxed_gpu(rec + 2 * lane); v2 = ld_relaxed_gpu(rec + 2 * (lane + 32)); v3 = ld_relaxed_gpu(rec + 2 * (lane + 64));
                v4 = ld_relaxed_gpu(rec + 2 * (lane + 96)); v5 = ld_relaxed_gpu(rec + 2 * (lane + 128)); break;
        case 5: v = ld_relaxed_sys(rec + 2 * lane); break;
        case 6: v = ld_volatile(rec + 2 * lane); break;
        case 7: v = ld_cg(rec + 2 * lane); break;
        case 8: v = ld_cv(rec + 2 * lane); break;
        case 9: v = ld_weak(rec + 2 * lane); break;
        case 10: v = ld_acquire_gpu(rec + 2 * lane); break;
        case 11: v = ld_cg(rec + 2 * lane); v2 = ld_cg(rec + 2 * (lane + 32)); v3 = ld_cg(rec + 2 * (lane + 64)); v4 = ld_cg(rec + 2 * (lane + 96)); v5 = ld_cg(rec + 2 * (lane + 128)); break;
        case 12: {   // one bulk copy of all records into shared memory, then one LDS
          if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(NREC * 32) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sbuf)), "l"(rec), "r"(NREC * 32), "r"(smem_u32(&bar)) : "memory");
          }
          uint32_t ok = 0;
          while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
          phase ^= 1;
          v = sbuf[2 * lane];
          break;
        }
        case 13: {   // LDGSTS 16 B per lane (cp.async.cg) + wait
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sbuf + lane)), "l"(rec + 2 * lane) : "memory");
          asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
          v = sbuf[lane];
          break;
        }
        case 14: {   // 5 x LDGSTS per lane
          for (int k = 0; k < 5; ++k) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sbuf + lane + 32 * k)), "l"(rec + 2 * (lane + 32 * k)) : "memory");
          asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
          v = sbuf[lane]; v2 = sbuf[lane + 32]; v3 = sbuf[lane + 64]; v4 = sbuf[lane + 96]; v5 = sbuf[lane + 128];
          break;
        }
        case 15: v = ld_relaxed_gpu(rec + 2 * lane); v2 = ld_relaxed_gpu(rec + 2 * lane + 1); break;   // record + its neighbour unit
        case 16: if (lane == 0) st_relaxed_gpu(rec, make_uint4(tick, 0, tick, ~tick)); v = ld_relaxed_gpu(rec + 2 * lane); break;   // lane 0: own store, same address
        case 17: if (lane == 0) st_relaxed_gpu(rec + 1, make_uint4(tick, 0, tick, ~tick)); v = ld_relaxed_gpu(rec + 2 * lane); break;   // own store, same sector
        case 18: v = ld_relaxed_gpu(rec + 8 * (lane % 20)); break;   // 128-byte stride (one line per lane), 20 records
        case 19: if (lane == 0) st_relaxed_gpu(rec + 2 * 40, make_uint4(tick, 0, tick, ~tick)); v = ld_relaxed_gpu(rec + 2 * lane); break;   // own store, other line
      }
      const uint32_t s = v.x + v.z + v2.x + v3.x + v4.x + v5.x + v.w + v2.w;
      const uint32_t any = __reduce_add_sync(0xffffffffu, s);
      const long long t1 = clock64();
      if (any == 0x12345) out[63] = any;
      if (rep >= 8) { acc += t1 - t0; if (t1 - t0 > mx) mx = t1 - t0; }
      if (var <= 15 && var != 9 && rep == REPS - 1 && lane == 1 && (v.x != (uint32_t)tick)) out[62] = 1000 + var;   // stale data seen
    }
    if (lane == 0) { out[2 * var] = acc / (REPS - 8); out[2 * var + 1] = mx; }
  }
}

int main() {
  uint4 *rec, *scratch; double* gbuf; long long* out; int* turn;
  cudaMalloc(&rec, NREC * 2 * sizeof(uint4)); cudaMalloc(&scratch, 4096); cudaMalloc(&gbuf, 32 * 8 * 8 * 8 + 4096);
  cudaMalloc(&out, 64 * 8); cudaMalloc(&turn, 256 * 4);
  cudaMemset(rec, 0, NREC * 2 * sizeof(uint4)); cudaMemset(out, 0, 64 * 8); cudaMemset(turn, 0, 256 * 4);
  void* args[] = {&rec, &scratch, &gbuf, &out, &turn};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)bench, dim3(2), dim3(64), args, 0, 0);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  long long h[64]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[NVAR] = {"ld.relaxed.gpu.v4 x1", "  ... after own st.relaxed.gpu", "  ... after 8 weak stores/lane", "  ... after both",
                             "ld.relaxed.gpu.v4 x5", "ld.relaxed.sys.v4 x1", "ld.volatile.v4 x1", "ld.global.cg.v4 x1", "ld.global.cv.v4 x1",
                             "ld.global.v4 (weak) x1", "ld.acquire.gpu.v4 x1", "ld.global.cg.v4 x5", "cp.async.bulk 5 KB -> smem + LDS",
                             "cp.async.cg 16 B x1 + wait", "cp.async.cg 16 B x5 + wait", "ld.relaxed.gpu.v4 x2 (same sector)", "own st same address, then ld", "own st same sector, then ld", "ld 128 B stride", "own st other line, then ld"};
  for (int v = 0; v < NVAR; ++v) printf("%-36s mean %6lld  max %6lld cycles\n", names[v], h[2 * v], h[2 * v + 1]);
  printf("stale flag: %lld\n", h[62]);
  return 0;
}
