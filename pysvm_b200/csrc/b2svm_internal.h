// Host-side internals shared by the translation units of libb2svm.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/b2svm.h"

namespace b2svm {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define B2_CUDA(expr)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ::b2svm::fail(B2SVM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// number of 16-byte chunks per padded row for (dtype, d); 0 when d exceeds the compiled maximum
int chunks_for(int x_dtype, int d);

// dense kernel mat-vec (b2svm_dense.cu): out[q] = sum_s coef[s] K(Xa[s], Xb[q])
int kernel_matvec_impl(cudaStream_t stream, const b2svm_kernel_desc& k, int x_dtype, int d, const void* Xa,
                       int64_t na, int64_t lda, const void* Xb, int64_t nb, int64_t ldb, const double* coef,
                       double* out);

// tensor-core (tcgen05, 3xTF32) variant for float32 rows with d <= 128 (b2svm_matvec_tc.cu).
// mode: 0 = not available / disabled, 1 = use for large problems, 2 = forced (B2SVM_MATVEC_TC=2)
int matvec_tc_mode(int x_dtype, int d);
int matvec_tc(cudaStream_t stream, const b2svm_kernel_desc& k, int d, const float* Xa, int64_t lda, const int* idx, int nsv,
              const double* coef, const float* Xb, int64_t nb, int64_t ldb, double* out);

// z(x_q) . wz = sqrt(2/D) sum_j wz_j cos(w_j . x_q + b_j) on the tensor cores (float32 queries, d <= 128)
int rff_decision_tc(cudaStream_t stream, int d, const double* W, const double* b, int D, const double* wz, const float* Xq,
                    int64_t nq, int64_t ldq, double* out);

// tile-output tensor-core kernel: out[m][n] = scale * f(a_m . b_n) written with TMA stores (b2svm_matvec_tc.cu)
bool tile_tc_ok(int d, int64_t ldo, bool out_f64, const void* out);
int tile_tc(cudaStream_t stream, const b2svm_kernel_desc& k, int d, const float* A, int64_t lda, int64_t m, const float* Bf,
            int64_t ldb, const double* Wd, const double* phase_d, int64_t n, double scale, void* out, int64_t ldo, bool out_f64);

void keep_pool_memory();   // stream-ordered allocator: keep up to 4 GB of work buffers between calls

struct ProblemDev;
#define B2_DECL_LAUNCH(name)                                                                                  \
  int name(int ch, bool cache, const ProblemDev* p, int nprob, int ncta_pp, int stages, size_t smem, cudaStream_t stream, bool coop)
B2_DECL_LAUNCH(launch_f32_plain);
B2_DECL_LAUNCH(launch_f32_generic);
B2_DECL_LAUNCH(launch_f64_plain);
B2_DECL_LAUNCH(launch_f64_generic);
#undef B2_DECL_LAUNCH

}  // namespace b2svm
