#include "b2svm_launch.cuh"
namespace b2svm {
int launch_f32_generic(int ch, bool cache, const ProblemDev* p, int nprob, int ncta_pp, int stages, size_t smem,
                      cudaStream_t stream, bool coop) {
  return launch_variant<float, true>(ch, cache, p, nprob, ncta_pp, stages, smem, stream, coop);
}
}  // namespace b2svm
