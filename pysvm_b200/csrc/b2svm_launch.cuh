// Launch helper shared by the four translation units that instantiate the persistent kernel.
#pragma once
#include "b2svm_internal.h"
#include "b2svm_smo.cuh"

namespace b2svm {

template <typename T, int CH, bool CACHE, bool GENERIC>
int launch_T(const ProblemDev* probs_dev, int nprob, int ncta_pp, int stages, size_t smem, cudaStream_t stream, bool coop) {
  auto kern = smo_persistent<T, CH, CACHE, GENERIC>;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = nprob * ncta_pp;
  if (ncta_pp > 1 || coop) {
    // CTAs of one problem (or the emulated ranks of one sharded problem) wait on each other: co-residency must be guaranteed
    void* args[] = {(void*)&probs_dev, (void*)&ncta_pp, (void*)&stages};
    B2_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(NTHREADS), args, smem, stream));
  } else {
    kern<<<grid, NTHREADS, smem, stream>>>(probs_dev, ncta_pp, stages);
    B2_CUDA(cudaGetLastError());
  }
  return B2SVM_OK;
}

template <typename T, bool GENERIC>
int launch_variant(int ch, bool cache, const ProblemDev* p, int nprob, int ncta_pp, int stages, size_t smem,
                   cudaStream_t stream, bool coop) {
#define B2_CASE(CHV)                                                                             \
  case CHV:                                                                                      \
    return cache ? launch_T<T, CHV, true, GENERIC>(p, nprob, ncta_pp, stages, smem, stream, coop)      \
                 : launch_T<T, CHV, false, GENERIC>(p, nprob, ncta_pp, stages, smem, stream, coop);
  switch (ch) {
    B2_CASE(4) B2_CASE(8) B2_CASE(16) B2_CASE(32) B2_CASE(64) B2_CASE(128)
    default: return fail(B2SVM_E_UNSUPPORTED, "unsupported row width");
  }
#undef B2_CASE
}

}  // namespace b2svm
