/*
 * b2svm.h -- C ABI of the B200-native SMO solver / kernel-evaluation engine.
 *
 * This is the drop-in boundary for the solve path of Kaslanarian/PySVM.  The reference has no FFI
 * of its own (it is pure Python/NumPy); each entry point below names the reference interface
 * (file:line under the reference root) whose work it takes over.  INTEGRATION.md shows the
 * ctypes binding a PySVM maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative B2SVM_E_* code otherwise;
 *     b2svm_last_error() returns a thread-local human readable message for the last failure;
 *   - no exceptions, no torch / C++ types in any signature: plain pointers and sizes;
 *   - "dev" pointers are device pointers (e.g. torch.Tensor.data_ptr()) owned by the caller and
 *     must stay alive while the handle uses them; "host" pointers are ordinary host memory;
 *     "any" pointers may be either (copied with cudaMemcpyDefault);
 *   - one handle per (device, stream); a handle is not thread-safe, independent handles are;
 *   - all work is queued on the stream given at creation; calls that return values to the host
 *     synchronise that stream before returning.
 *
 * Problem solved (reference pysvm/solver.py:5-24 and :232-254)
 *     min_a 0.5 a'Qa + p'a   s.t.  y'a = 0, 0 <= a <= C              (B2SVM_SOLVER_C)
 *     ... and additionally e'a = t                                    (B2SVM_SOLVER_NU)
 * with Q[s][t] = y_s y_t K(x_{s mod l}, x_{t mod l}).  n_vars == n_rows for classification /
 * one-class (y == 1), n_vars == 2*n_rows ("mirror") for the regression estimators
 * (pysvm/svm/svr.py:158-179).
 */
#ifndef B2SVM_H_
#define B2SVM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2SVM_ABI_VERSION 1
/* b2svm_solve / b2svm_solve_batched accept max_iter in [0, B2SVM_MAX_ITER): the candidate records of the in-kernel
 * exchange carry a 29-bit iteration tag.  Longer solves are chunked by the caller (state lives in alpha / neg_y_grad). */
#define B2SVM_MAX_ITER (1ll << 29)

enum {
  B2SVM_OK = 0,
  B2SVM_E_INVALID = -1,      /* bad argument                                   */
  B2SVM_E_CUDA = -2,         /* a CUDA runtime call failed                     */
  B2SVM_E_UNSUPPORTED = -3,  /* e.g. n_features above the compiled maximum     */
  B2SVM_E_TIMEOUT = -4,      /* device-side watchdog tripped (grid / peer wait) */
  B2SVM_E_NOMEM = -5
};

enum { B2SVM_F32 = 0, B2SVM_F64 = 1 };
enum { B2SVM_KERNEL_LINEAR = 0, B2SVM_KERNEL_POLY = 1, B2SVM_KERNEL_RBF = 2, B2SVM_KERNEL_SIGMOID = 3,
       /* X is the dense n_rows x n_rows kernel matrix itself (the reference's dense-Q mode, solver.py:25-45
        * with cache_size == 0, and the RFF kernel z(a).z(b) of svc.py:225-228 for any D) */
       B2SVM_KERNEL_PRECOMPUTED = 4 };
enum { B2SVM_SOLVER_C = 0, B2SVM_SOLVER_NU = 1 };

/* Kernel function description: pysvm/svm/svc.py:210-241 (register_kernel).
 * gamma is already resolved ('scale' -> 1/(d*X.std()), 'auto' -> 1/d) by the host. */
typedef struct b2svm_kernel_desc {
  int32_t kind;      /* B2SVM_KERNEL_*                                        */
  int32_t _pad;
  double gamma;      /* rbf / poly / sigmoid                                  */
  double coef0;      /* poly / sigmoid                                        */
  double degree;     /* poly                                                  */
} b2svm_kernel_desc;

/* Everything Solver.__init__ / NuSolver.__init__ and the estimator's fit() hand to the solver
 * (pysvm/solver.py:25-45, :255-281; pysvm/svm/svc.py:243-258). */
typedef struct b2svm_problem_desc {
  uint32_t struct_bytes;   /* = sizeof(b2svm_problem_desc), checked                         */
  int32_t device;          /* CUDA device ordinal                                           */
  void* stream;            /* cudaStream_t (0 = legacy default stream)                      */
  int64_t n_rows;          /* l: rows of X                                                  */
  int32_t n_features;      /* d                                                             */
  int32_t x_dtype;         /* B2SVM_F32 / B2SVM_F64: kernel values are evaluated in this type */
  const void* X;           /* dev, row-major [n_rows][ldx] (with world > 1: this shard's rows only)  */
  int64_t ldx;             /* leading dimension of X in elements                            */
  int64_t n_vars;          /* n_rows, or 2*n_rows (mirror: vars t and t+l share row t)      */
  int32_t variant;         /* B2SVM_SOLVER_C / B2SVM_SOLVER_NU                              */
  int32_t _pad0;
  const double* y;         /* dev [n_vars], entries +1 / -1                                 */
  double* alpha;           /* dev [n_vars]  caller-owned state, updated in place            */
  double* neg_y_grad;      /* dev [n_vars]  caller-owned state  (-y * grad f), in place     */
  double C;
  double tol;
  b2svm_kernel_desc kernel;
  uint64_t cache_bytes;    /* HBM budget for the device-resident Q-column cache; 0 = off    */
  int32_t max_ctas;        /* 0 = auto (one CTA per SM at most)                             */
  int32_t rank;            /* row-sharded multi-GPU: this handle's rank ...                 */
  int32_t world;           /* ... out of `world` (1 = single GPU)                           */
  int32_t _pad1;
  int64_t row_offset;      /* global index of this shard's first row (multi-GPU)            */
  int64_t n_rows_global;   /* total rows over all ranks (== n_rows when world == 1)         */
} b2svm_problem_desc;

/* Counters returned by b2svm_solve (iterations, cache hits, bytes). */
typedef struct b2svm_solve_stats {
  int64_t n_iter;            /* updates performed by this call                              */
  int32_t converged;         /* 1 when working-set selection returned (-1,-1)               */
  int32_t n_ctas;            /* CTAs of the persistent launch                               */
  int64_t last_i, last_j;    /* pair chosen for the NEXT update (-1,-1 when converged)      */
  double last_delta_i, last_delta_j;
  int64_t cold_sweeps;       /* iterations that streamed X (at least one column missed)     */
  int64_t cache_hits, cache_misses;
  double algorithmic_bytes;  /* sum over iterations of B_iter          */
  float device_ms;           /* CUDA-event time of the persistent kernel(s) of this call    */
  int32_t launches;          /* kernels launched by this call                               */
  int64_t cache_stores;      /* Q columns written into the HBM column cache                 */
  int64_t cache_evictions;   /* ... of which replaced an older column (FIFO, cache full)    */
} b2svm_solve_stats;

typedef struct b2svm_handle b2svm_handle;

int b2svm_abi_version(void);
const char* b2svm_last_error(void);

/* Solver.__init__ / SolverWithCache.__init__ / NuSolver.__init__ (solver.py:25-45,209-216,255-281).
 * Pads / converts nothing numerically: X is used as given (copied only when the row pitch is not
 * 16-byte friendly).  Does not touch alpha / neg_y_grad. */
int b2svm_create(const b2svm_problem_desc* desc, b2svm_handle** out);
int b2svm_destroy(b2svm_handle* h);

/* alpha = 0, neg_y_grad = -y*p  (solver.py:42-45).  p: any [n_vars]. */
int b2svm_init_state(b2svm_handle* h, const double* p);

/* neg_y_grad = -y * (Q alpha + p) for the current alpha, using only the non-zero alphas
 * (NuSolverWithCache.__init__ solver.py:468-473; OneClassSVM.fit oc_svm.py:94-95).  p: any or NULL. */
int b2svm_init_gradient(b2svm_handle* h, const double* p);

/* Solver.working_set_select (solver.py:47-75) or NuSolver.working_set_select (solver.py:283-339)
 * according to the handle's variant.  (-1,-1) = stop. */
int b2svm_select(b2svm_handle* h, int64_t* i, int64_t* j);

/* OPT-IN, not the reference's rule: second-order working-set selection (LIBSVM WSS 3, Fan / Chen / Lin 2005):
 * i = argmax_{I_up} g, j = argmin_{t in I_low, g_t < g_i} -(g_i - g_t)^2 / (K_ii + K_tt - 2 K_it); the stopping test is
 * the reference's (g_i - min_{I_low} g < tol -> (-1,-1)).  C variant, single GPU; one extra sweep of X per call.
 * Pair it with b2svm_update; iteration counts differ from the reference's (typically 2-3x fewer). */
int b2svm_select_wss3(b2svm_handle* h, int64_t* i, int64_t* j);

/* Solver.update (solver.py:77-147) / NuSolver.update (solver.py:341-376): analytic two-variable
 * step with box clipping + rank-2 gradient update.  Returns (delta_i, delta_j). */
int b2svm_update(b2svm_handle* h, int64_t i, int64_t j, double* delta_i, double* delta_j);

/* func(i) / get_Q(i, func): the Q column y * K(X, x_i) * y_i as float64 [n_vars]
 * (svc.py:257-258, svr.py:174-179, oc_svm.py:72-73).  out: dev. */
int b2svm_kernel_column(b2svm_handle* h, int64_t i, double* out);

/* The estimator's `for n_iter in range(max_iter)` loop (svc.py:260-267, :426-433; svr.py:181-189,
 * :286-294; oc_svm.py:97-105) fused into one persistent kernel. */
int b2svm_solve(b2svm_handle* h, int64_t max_iter, b2svm_solve_stats* stats);

/* Solver.calculate_rho (solver.py:149-177) and NuSolver.calculate_rho_b (solver.py:378-419). */
int b2svm_rho(b2svm_handle* h, double* rho);
int b2svm_rho_b(b2svm_handle* h, double* rho, double* b);
/* The same statistics un-combined (host double[18]: sums, counts, extremes), for a row-sharded solve
 * whose ranks merge them before forming rho / (rho, b). */
int b2svm_bias_partials(b2svm_handle* h, double* out);

/* Frees the handle's Q-column cache (the reference's LRU of solver.py:207, here up to 70 % of the device memory) while
 * keeping the handle usable: rho / rho_b, state access and further solves (then without a cache) still work.  The
 * estimators call it at the end of fit(), where the reference keeps the solver alive inside the decision_function
 * closure (svc.py:269-272).  No-op for precomputed kernels and cache-less handles. */
int b2svm_release_cache(b2svm_handle* h);

/* decision_function core (svc.py:269-272, :436-439; svr.py:191-194, :297-300; oc_svm.py:108-111):
 *   out[q] = sum_s coef[s] * K(Xa[s], Xb[q])      (no n_a x n_b matrix is materialised)
 * Xa: dev [na][lda], Xb: dev [nb][ldb] (dtype x_dtype), coef: dev float64 [na], out: dev float64 [nb].
 * Rows with coef == 0 are skipped.  float32 rows with n_features <= 128 and at least 1024 queries and non-zero
 * coefficients run on the tensor cores (tcgen05, 3xTF32: ~1e-6 of the output scale against float64); everything else
 * on the FP32 / FP64 pipes.  The call synchronises the stream before it returns. */
int b2svm_kernel_matvec(int32_t device, void* stream, const b2svm_kernel_desc* k, int32_t x_dtype,
                        int32_t n_features, const void* Xa, int64_t na, int64_t lda,
                        const void* Xb, int64_t nb, int64_t ldb, const double* coef, double* out);

/* NormalRFF.transform (rff.py:19-20): out[n][D] = sqrt(2/D) * cos(X W^T + b), float64 like the
 * reference (W, b float64 [D][d], [D]; X in x_dtype).  float64 X (the sklearn datasets): float64 arithmetic, exact to
 * 1e-13.  float32 X with n_features <= 96, n >= 1024, D >= 128: tcgen05 GEMM (3xTF32) + cos epilogue + TMA stores; the
 * cos ARGUMENT is then accurate to ~1e-6 (float32 X against float64 W: the reference promotes X), the values to ~2e-6
 * of sqrt(2/D) -- enough for the kernel approximation the features serve, not bit parity. */
int b2svm_rff_transform(int32_t device, void* stream, int32_t x_dtype, const void* X, int64_t n,
                        int64_t ldx, int32_t d, const double* W, const double* b, int32_t D, double* out);

/* The same features as float32 (prediction-only pipelines, and half the bytes of the HBM-write-bound transform).
 * Tensor-core path only: float32 X with n_features <= 96; out: dev [n][ldo] float32, ldo*4 a multiple of 16. */
int b2svm_rff_transform_f32(int32_t device, void* stream, const float* X, int64_t n, int64_t ldx, int32_t d,
                            const double* W, const double* b, int32_t D, float* out, int64_t ldo);

/* Dense kernel matrix out[i][j] = K(Xa[i], Xb[j]) in the dtype of X (dev [na][ldo]): what the reference's dense-Q
 * mode (cache_size == 0: svc.py:251-253, :419-421; oc_svm.py:85-89) and its rff=True kernel z(a).z(b) (svc.py:225-228)
 * build on the host with kernel_func(X, X).  float32 with n_features <= 96 and >= 1024 rows on both sides: tcgen05 GEMM
 * (3xTF32) + kernel epilogue + TMA stores; float64 and everything else: a tiled FMA kernel (products accumulated in
 * feature order).  kind: linear / poly / rbf / sigmoid.  Synchronises the stream. */
int b2svm_kernel_matrix(int32_t device, void* stream, const b2svm_kernel_desc* k, int32_t x_dtype, int32_t n_features,
                        const void* Xa, int64_t na, int64_t lda, const void* Xb, int64_t nb, int64_t ldb,
                        void* out, int64_t ldo);

/* out[c] = sum_r w[r] * X[r][c] as float64 (dev [d]): the primal weights w = X^T (alpha*y) of the Linear* estimators
 * (svc.py:76, svr.py:89-90, telescoped) and Z^T coef of the RFF decision.  Deterministic summation order. */
int b2svm_weighted_rows(int32_t device, void* stream, int32_t x_dtype, const void* X, int64_t n, int64_t ldx, int32_t d,
                        const double* w, double* out);

/* z(q) . wz  without materialising z(q):  out[q] = sum_k sqrt(2/D) cos(x_q . W_k + b_k) * wz[k].
 * float32 queries (n_features <= 128, nq >= 1024, D >= 256) run as a tcgen05 GEMM with a cos epilogue, argument error
 * ~1e-6; float64 queries and small problems in float64 like b2svm_rff_transform. */
int b2svm_rff_decision(int32_t device, void* stream, int32_t x_dtype, const void* Xq, int64_t nq,
                       int64_t ldx, int32_t d, const double* W, const double* b, int32_t D,
                       const double* wz, double* out);

/* OvO / OvR orchestration (svc.py:332-342 + sklearn.multiclass): `count` independent problems of
 * the same dtype / n_features / kernel kind solved concurrently, one CTA group each. */
int b2svm_solve_batched(b2svm_handle* const* handles, int32_t count, int64_t max_iter,
                        b2svm_solve_stats* stats /* [count] */);

/* Row-sharded multi-GPU (new).  Each rank exports the IPC handle of its
 * exchange mailbox, the host side all-gathers them (torch.distributed), and every rank attaches
 * its peers' mailboxes before the first solve. */
int b2svm_mailbox_export(b2svm_handle* h, void* ipc_handle_out /* 64 bytes */);
int b2svm_mailbox_attach(b2svm_handle* h, const void* ipc_handles /* world * 64 bytes */);
/* Multi-GPU only, before EVERY b2svm_solve: clears this rank's mailbox.  The host then runs a barrier
 * over all ranks (torch.distributed) and calls b2svm_solve on each.
 * Shards are equal slices of whole 32-row tiles: rank r sweeps global rows
 * [r*rpr, min(n, (r+1)*rpr)), rpr = roundup32(ceil(n_rows_global / world)), and owns the matching slices of
 * y, alpha, neg_y_grad (a mirror problem keeps [first half | second half] of its LOCAL rows) and of every
 * cached column, AND ONLY THOSE ROWS of X (desc.X = the shard's rows, n_rows of them): the CTA that owns a GPU's
 * winning candidate pushes record + alpha + norm + the row itself into every GPU's box, so the two selected rows
 * reach all ranks with the exchange and nothing else crosses NVLink. */
int b2svm_prepare(b2svm_handle* h);

/* Test / bring-up entry: the `world` ranks of a row-sharded solve emulated on ONE GPU.  handles[r] must have been
 * created with rank = r, world = `world`, all on the same device and stream, with max_ctas <= num_sms / world; they run
 * as the CTA groups of one cooperative launch and exchange candidates through each other's level-2 boxes in local
 * memory (no CUDA IPC, no b2svm_mailbox_attach / b2svm_prepare needed).  Same device code as the multi-GPU solve.
 * stats: [world]. */
int b2svm_solve_ranks_on_one_gpu(b2svm_handle* const* handles, int32_t world, int64_t max_iter,
                                 b2svm_solve_stats* stats);

/* Host helper of register_kernel (svc.py:212-217): gamma='scale' is 1 / (n_features * X.std()).  Population
 * standard deviation of n_elements contiguous host values, bit-identical to numpy's X.std() on a C-contiguous
 * array (same pairwise summation tree, evaluated on several host threads).  *out holds the float32 result
 * exactly when x_dtype is B2SVM_F32. */
int b2svm_host_std(const void* x_host, int64_t n_elements, int32_t x_dtype, double* out);

#ifdef __cplusplus
}
#endif
#endif /* B2SVM_H_ */
