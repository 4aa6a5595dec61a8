// One-way "store on SM A -> seen by a polling load on SM B" latency on B200, measured as half a ping-pong round
// trip, for the store / load flavours the in-kernel exchange could use.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a pingpong.cu -o pingpong
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { ST_RELAXED_GPU, ST_VOLATILE, ST_WEAK_FENCE, ATOM_EXCH, ST_RELEASE_GPU, RED_MAX, ST_RELAXED_SYS, ST_CG_V4, NMODE };

__device__ __forceinline__ void put(int mode, unsigned* p, unsigned v) {
  switch (mode) {
    case ST_RELAXED_GPU: asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); break;
    case ST_VOLATILE: asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); break;
    case ST_WEAK_FENCE: *p = v; __threadfence(); break;
    case ATOM_EXCH: atomicExch(p, v); break;
    case ST_RELEASE_GPU: asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); break;
    case RED_MAX: asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); break;
    case ST_RELAXED_SYS: asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); break;
    case ST_CG_V4: asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(v) : "memory"); break;
  }
}
__device__ __forceinline__ unsigned get(int mode, unsigned* p) {
  unsigned v;
  if (mode == ST_RELAXED_SYS) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  else if (mode == ST_VOLATILE) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  else if (mode == ST_CG_V4) { unsigned a, b, c; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v), "=r"(a), "=r"(b), "=r"(c) : "l"(p) : "memory"); v = v + 0 * (a + b + c); }
  else asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void pingpong(unsigned* flags, long long* out, int peer_cta, int busy) {
  const int REPS = 2000;
  if (threadIdx.x != 0) {
    // optional background traffic from the other warps of the two CTAs: strided global stores like a sweep's G updates
    if (busy && (blockIdx.x == 0 || blockIdx.x == peer_cta)) {
      double* g = reinterpret_cast<double*>(flags + 4096) + (size_t)blockIdx.x * 65536;
      for (int it = 0; it < 200000 && *((volatile unsigned*)flags + 1024) == 0; ++it) g[(threadIdx.x * 37 + it * 8191) & 65535] = it;
    }
    return;
  }
  for (int mode = 0; mode < NMODE; ++mode) {
    unsigned* a = flags + 64 * mode;        // written by CTA 0
    unsigned* b = flags + 64 * mode + 32;   // written by the peer (another 128-byte line)
    if (blockIdx.x == 0) {
      long long t0 = 0;
      for (int r = 1; r <= REPS; ++r) {
        if (r == 100) t0 = clock64();
        put(mode, a, (unsigned)r);
        while (get(mode, b) < (unsigned)r) {}
      }
      out[mode] = (clock64() - t0) / (2 * (REPS - 99));
    } else if (blockIdx.x == peer_cta) {
      for (int r = 1; r <= REPS; ++r) {
        while (get(mode, a) < (unsigned)r) {}
        put(mode, b, (unsigned)r);
      }
    }
  }
  if (blockIdx.x == 0) *((volatile unsigned*)flags + 1024) = 1;
}

int main() {
  unsigned* flags; long long* out;
  cudaMalloc(&flags, 4096 * 4 + 160 * 65536 * 8); cudaMalloc(&out, 64 * 8);
  const char* names[NMODE] = {"st.relaxed.gpu / ld.relaxed.gpu", "st.volatile / ld.volatile", "st + __threadfence / ld.relaxed.gpu", "atomicExch / ld.relaxed.gpu",
                              "st.release.gpu / ld.relaxed.gpu", "red.max / ld.relaxed.gpu", "st.relaxed.sys / ld.relaxed.sys", "st.relaxed.gpu.v4 / ld.relaxed.gpu.v4"};
  for (int busy = 0; busy < 2; ++busy)
    for (int peer : {1, 75, 147}) {
      cudaMemset(flags, 0, 4096 * 4); cudaMemset(out, 0, 64 * 8);
      void* args[] = {&flags, &out, &peer, &busy};
      cudaLaunchCooperativeKernel((void*)pingpong, dim3(148), dim3(busy ? 256 : 32), args, 0, 0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[NMODE]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      printf("peer CTA %3d, %s: one-way latency (cycles):", peer, busy ? "with store traffic from 7 more warps" : "quiet");
      for (int m = 0; m < NMODE; ++m) printf("  %lld", h[m]);
      printf("\n");
    }
  for (int m = 0; m < NMODE; ++m) printf("  mode %d = %s\n", m, names[m]);
  return 0;
}
