// The persistent SMO kernel (K6 = K1 + K3 + K4 + K5 of SURVEY section 2.1) for sm_100a.
//
// One launch runs the whole `for n_iter in range(max_iter)` loop of the reference estimators
// (pysvm/svm/svc.py:260-267): per iteration every CTA
//   1. (leader warp) reduces the per-CTA working-set candidates of the previous sweep into the
//      global pair (i, j)                                   -> Solver.working_set_select, solver.py:47-75
//   2. (leader warp) loads x_i, x_j, evaluates K_ii, K_jj, K_ij and takes the two-variable step
//      redundantly (deterministic, so all CTAs agree)       -> Solver.update, solver.py:82-145
//   3. sweeps its own slice of X once: both Q columns (dot products + kernel epilogue), the
//      rank-2 gradient update, and the next iteration's argmax/argmin candidates in the same pass
//                                                           -> func(i) svc.py:257-258, solver.py:146
//   4. publishes its candidates; a flag-based grid barrier closes the iteration.
//
// X rows reach shared memory through a ring of 1-D TMA bulk copies (cp.async.bulk + mbarrier,
// SASS UBLKCP) issued by a dedicated producer thread that runs ahead of the consumers across
// iteration boundaries, so HBM keeps streaming while the grid barrier and the serial
// select -> step chain are in flight.  When a CTA's slice fits in the ring it is loaded once and
// stays resident in shared memory for the whole solve (small problems, multi-GPU shards).
#pragma once
#include "b2svm_device.cuh"

namespace b2svm {

constexpr int NW = 8;                      // consumer warps per CTA
constexpr int NTHREADS = (NW + 1) * 32;    // + one producer warp
constexpr int MAX_STAGES = 64;
constexpr int NCAND = 4;                   // up/low (C variant) or up+/low+/up-/low- (nu variant)

struct __align__(16) CandRec {
  double g;
  double alpha;
  int idx;
  int flags;
  int pad[2];
};

struct SolveResult {
  long long n_iter;
  long long last_i, last_j;
  long long cold_sweeps;
  double di, dj;
  int converged;
  int status;   // 0 ok, 1 watchdog
  int pad[2];
};

struct ProblemDev {
  const void* X;        // [l][CH*16 bytes]
  const void* norms;    // T[l] squared row norms
  double* alpha;        // [nv]
  double* G;            // [nv]  neg_y_grad
  uint8_t* flags;       // [nv]
  long long l, nv;
  int mirror, nu;
  double C, tol;
  int kind;
  double gamma, coef0, degree;
  int ncta;             // CTAs cooperating on this problem
  int tiles_per_cta;
  long long ntiles;
  CandRec* cand;        // [2][ncta][NCAND]
  uint32_t* flag;       // [ncta] epoch counters (zeroed before every launch)
  int* abort_flag;
  SolveResult* result;
  long long max_iter;
  long long forced_i, forced_j;   // >= 0: skip the initial selection and take this pair first
};

template <typename T, int CH_>
struct Cfg {
  using CK = Chunk<T>;
  using V = typename CK::V;
  static constexpr int CH = CH_;                         // 16-byte chunks per row
  static constexpr int LPR = CH < 32 ? CH : 32;          // lanes per row
  static constexpr int CPL = CH / LPR;                   // chunks per lane
  static constexpr int GRP = 32 / LPR;                   // rows fetched by one warp-wide load
  static constexpr int FINAL = sizeof(T) == 4 ? 2 : 1;   // sums per lane after the transposed reduce
  static constexpr int RPL = FINAL == 2 ? LPR : LPR / 2; // rows per lane group per tile
  static constexpr int BR = GRP * RPL;                   // rows per tile (32 for f32, 16 for f64)
  static constexpr int NV = 2 * RPL;                     // partial sums per lane
  static constexpr int ROWB = CH * 16;
  static constexpr int TILE_BYTES = BR * ROWB;
};

template <typename T>
struct Shared {
  double ci, cj;        // y_i * delta_i, y_j * delta_j
  double ai, aj;        // new alpha_i, alpha_j
  T ni, nj;             // |x_i|^2, |x_j|^2
  int i, j, fi, fj;
  int stop;
  volatile int producer_stop;
  volatile long long consumed;
  Cand wc[NW][NCAND];
};

__device__ __forceinline__ void consider(bool nu, int f, double g, int v, Cand (&b)[NCAND]) {
  const bool second = nu && !(f & F_YPOS);
  if (in_up(f)) {
    if (!second) {
      if (better_max(g, v, b[0].g, b[0].idx)) { b[0].g = g; b[0].idx = v; }
    } else {
      if (better_max(g, v, b[2].g, b[2].idx)) { b[2].g = g; b[2].idx = v; }
    }
  }
  if (in_low(f)) {
    if (!second) {
      if (better_min(g, v, b[1].g, b[1].idx)) { b[1].g = g; b[1].idx = v; }
    } else {
      if (better_min(g, v, b[3].g, b[3].idx)) { b[3].g = g; b[3].idx = v; }
    }
  }
}

// butterfly reduce of one candidate over the warp; k even = max type, k odd = min type
__device__ __forceinline__ void warp_reduce_cand(Cand& c, bool is_max) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const double og = __shfl_xor_sync(0xffffffffu, c.g, m);
    const int oi = __shfl_xor_sync(0xffffffffu, c.idx, m);
    const bool take = is_max ? better_max(og, oi, c.g, c.idx) : better_min(og, oi, c.g, c.idx);
    if (take) { c.g = og; c.idx = oi; }
  }
}

struct Winner {
  double g, alpha;
  int idx, flags;
};
__device__ __forceinline__ void warp_reduce_winner(Winner& w, bool is_max) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const double og = __shfl_xor_sync(0xffffffffu, w.g, m);
    const double oa = __shfl_xor_sync(0xffffffffu, w.alpha, m);
    const int oi = __shfl_xor_sync(0xffffffffu, w.idx, m);
    const int of = __shfl_xor_sync(0xffffffffu, w.flags, m);
    const bool take = is_max ? better_max(og, oi, w.g, w.idx) : better_min(og, oi, w.g, w.idx);
    if (take) { w.g = og; w.alpha = oa; w.idx = oi; w.flags = of; }
  }
}

// Leader warp: load rows ra, rb (16-byte chunks spread over the lanes) and evaluate
// K(a,a), K(b,b), K(a,b) in T.  Chunks stay in xa / xb for the caller.
template <typename T, int CH>
__device__ __forceinline__ void pair_kernels(const ProblemDev& P, const KernelParams<T>& kp, long long ra,
                                             long long rb, int lane, typename Chunk<T>::V (&xa)[(CH + 31) / 32],
                                             typename Chunk<T>::V (&xb)[(CH + 31) / 32], T& na, T& nb, T& Kaa,
                                             T& Kbb, T& Kab) {
  using CK = Chunk<T>;
  using V = typename CK::V;
  constexpr int Q = (CH + 31) / 32;
  const V* Xa = reinterpret_cast<const V*>(P.X) + ra * CH;
  const V* Xb = reinterpret_cast<const V*>(P.X) + rb * CH;
  T daa = 0, dbb = 0, dab = 0;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int c = lane + 32 * q;
    xa[q] = (c < CH) ? __ldg(Xa + c) : CK::zero();
    xb[q] = (c < CH) ? __ldg(Xb + c) : CK::zero();
    daa = CK::dot(xa[q], xa[q], daa);
    dbb = CK::dot(xb[q], xb[q], dbb);
    dab = CK::dot(xa[q], xb[q], dab);
  }
  daa = warp_sum(daa);
  dbb = warp_sum(dbb);
  dab = warp_sum(dab);
  const T* norms = reinterpret_cast<const T*>(P.norms);
  na = __ldg(norms + ra);
  nb = __ldg(norms + rb);
  Kaa = kernel_value(kp, daa, na, na);
  Kbb = kernel_value(kp, dbb, nb, nb);
  Kab = kernel_value(kp, dab, na, nb);
}

template <typename T, int CH>
__global__ void __launch_bounds__(NTHREADS, 1) smo_persistent(const ProblemDev* __restrict__ probs, int ncta_pp,
                                                              int ring_stages) {
  using C = Cfg<T, CH>;
  using CK = typename C::CK;
  using V = typename C::V;
  constexpr int QL = (CH + 31) / 32;

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)ring_stages * C::TILE_BYTES);
  uint64_t* empty = full + MAX_STAGES;
  V* xs = reinterpret_cast<V*>(empty + MAX_STAGES);      // [2][CH]
  Shared<T>* sh = reinterpret_cast<Shared<T>*>(xs + 2 * CH);

  const ProblemDev& P = probs[blockIdx.x / ncta_pp];
  const int cta = blockIdx.x % ncta_pp;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = ring_stages;

  // this CTA's slice of rows, in whole tiles
  const long long tile0 = (long long)cta * P.tiles_per_cta;
  long long tile1 = tile0 + P.tiles_per_cta;
  if (tile1 > P.ntiles) tile1 = P.ntiles;
  const int NT = tile1 > tile0 ? (int)(tile1 - tile0) : 0;
  const long long row0 = tile0 * C::BR;
  long long row1 = tile1 * C::BR;
  if (row1 > P.l) row1 = P.l;
  if (row1 < row0) row1 = row0;
  const bool resident = NT <= S;
  const bool nu = P.nu != 0;
  const bool mirror = P.mirror != 0;
  const long long l = P.l;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_fence_init();
    sh->stop = 0;
    sh->producer_stop = 0;
    sh->consumed = 0;
  }
  __syncthreads();

  // ============================== producer warp ==============================================
  if (warp == NW) {
    if (lane == 0 && NT > 0) {
      const unsigned char* Xb = reinterpret_cast<const unsigned char*>(P.X);
      unsigned long long Tp = 0;
      auto issue = [&](unsigned long long Tg, int t, int stage) {
        const long long r0 = row0 + (long long)t * C::BR;
        long long nr = row1 - r0;
        if (nr > C::BR) nr = C::BR;
        const uint32_t bytes = (uint32_t)(nr * C::ROWB);
        mbar_arrive_expect_tx(&full[stage], bytes);
        bulk_g2s(ring + (size_t)stage * C::TILE_BYTES, Xb + r0 * C::ROWB, bytes, &full[stage]);
      };
      if (resident) {
        for (int t = 0; t < NT; ++t) issue(t, t, t);
        Tp = NT;
        while (!sh->producer_stop) __nanosleep(200);
      } else {
        bool stop = false;
        while (!stop) {
          const int stage = (int)(Tp % S);
          const uint32_t par = (uint32_t)((Tp / S) & 1);
          while (!mbar_try_wait(&empty[stage], par ^ 1u)) {
            if (sh->producer_stop) { stop = true; break; }
          }
          if (stop) break;
          if (sh->producer_stop) break;
          issue(Tp, (int)(Tp % NT), stage);
          ++Tp;
        }
      }
      // drain: never leave the CTA with bulk copies still landing in its shared memory
      const unsigned long long Tc = resident ? 0ull : (unsigned long long)sh->consumed;
      for (unsigned long long Tg = Tc; Tg < Tp; ++Tg) {
        const int stage = resident ? (int)Tg : (int)(Tg % S);
        const uint32_t par = resident ? 0u : (uint32_t)((Tg / S) & 1);
        while (!mbar_try_wait(&full[stage], par)) {}
      }
    }
    return;
  }

  // ============================== consumer warps =============================================
  KernelParams<T> kp;
  kp.kind = P.kind;
  kp.gamma = (T)P.gamma;
  kp.coef0 = (T)P.coef0;
  kp.degree = (T)P.degree;
  const T* __restrict__ norms = reinterpret_cast<const T*>(P.norms);
  double* __restrict__ G = P.G;
  double* __restrict__ alpha = P.alpha;
  uint8_t* __restrict__ flags = P.flags;
  const int ntypes = nu ? 4 : 2;
  const double Cbox = P.C;

  const int lig = lane % C::LPR;
  const int grp = lane / C::LPR;
  const int rib = C::FINAL == 2 ? lane : grp * C::RPL + (lig >> 1);   // my row within a tile
  const bool lane_active = C::FINAL == 2 ? true : !(lig & 1);

  long long n_iter = 0;
  long long cold = 0;
  unsigned long long Tbase = 0;
  uint32_t epoch = 0;
  int status = 0;
  int converged = 0;
  // leader-only state (warp 0, replicated in all its lanes)
  double l_ai = 0, l_aj = 0, l_di = 0, l_dj = 0;
  int l_i = -1, l_j = -1;

  Cand best[NCAND];
  auto reset_best = [&]() {
#pragma unroll
    for (int k = 0; k < NCAND; ++k) { best[k].g = 0.0; best[k].idx = -1; }
  };

  // ---- initial selection pass over G / flags only (no X) -------------------------------------
  reset_best();
  const bool forced = P.forced_i >= 0;
  if (!forced) {
    for (long long r = row0 + warp * 32 + lane; r < row1; r += NW * 32) {
      consider(nu, flags[r], G[r], (int)r, best);
      if (mirror) consider(nu, flags[r + l], G[r + l], (int)(r + l), best);
    }
  }

  for (;;) {
    // ---- (B) block-level candidate reduce, publish, grid barrier, global reduce (leader) ------
    ++epoch;
#pragma unroll
    for (int k = 0; k < NCAND; ++k)
      if (k < ntypes) warp_reduce_cand(best[k], (k & 1) == 0);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NCAND; ++k) sh->wc[warp][k] = best[k];
    }
    named_bar_sync(1, NW * 32);

    if (warp == 0) {
      // CTA-level winner per type -> record in global memory
      Cand mine[NCAND];
#pragma unroll
      for (int k = 0; k < NCAND; ++k) {
        if (lane < NW) mine[k] = sh->wc[lane][k];
        else { mine[k].g = 0.0; mine[k].idx = -1; }
        if (k < ntypes) warp_reduce_cand(mine[k], (k & 1) == 0);
      }
      if (lane < ntypes) {
        Cand m = mine[0];
#pragma unroll
        for (int k = 1; k < NCAND; ++k)
          if (lane == k) m = mine[k];
        CandRec rec;
        rec.g = m.g;
        rec.idx = m.idx;
        rec.alpha = 0.0;
        rec.flags = 0;
        rec.pad[0] = rec.pad[1] = 0;
        if (m.idx >= 0) {   // the winner lives in this CTA's slice: same-SM coherent reads
          rec.alpha = alpha[m.idx];
          rec.flags = flags[m.idx];
        }
        P.cand[((size_t)(epoch & 1) * P.ncta + cta) * NCAND + lane] = rec;
        __threadfence();
      }
      __syncwarp();
      if (lane == 0) st_release_gpu(P.flag + cta, epoch);

      // wait for every CTA of this problem, fold their records
      Winner w[NCAND];
#pragma unroll
      for (int k = 0; k < NCAND; ++k) { w[k].g = 0.0; w[k].alpha = 0.0; w[k].idx = -1; w[k].flags = 0; }
      const uint64_t t_start = globaltimer_ns();
      if (*(volatile int*)P.abort_flag) status = 1;
      for (int c = lane; c < P.ncta && !status; c += 32) {
        uint32_t polls = 0;
        while (ld_acquire_gpu(P.flag + c) < epoch) {
          if ((++polls & 255u) == 0) {
            if (*(volatile int*)P.abort_flag) { status = 1; break; }
            if (globaltimer_ns() - t_start > 4000000000ull) {
              *(volatile int*)P.abort_flag = 1;
              status = 1;
              break;
            }
          }
        }
        if (status) break;
        const CandRec* rp = P.cand + ((size_t)(epoch & 1) * P.ncta + c) * NCAND;
#pragma unroll
        for (int k = 0; k < NCAND; ++k) {
          if (k < ntypes) {
            const int4 lo = __ldcg(reinterpret_cast<const int4*>(rp + k));
            const int4 hi = __ldcg(reinterpret_cast<const int4*>(rp + k) + 1);
            const double g = __hiloint2double(lo.y, lo.x);
            const double a = __hiloint2double(lo.w, lo.z);
            const int idx = hi.x, f = hi.y;
            const bool take = (k & 1) == 0 ? better_max(g, idx, w[k].g, w[k].idx) : better_min(g, idx, w[k].g, w[k].idx);
            if (take) { w[k].g = g; w[k].alpha = a; w[k].idx = idx; w[k].flags = f; }
          }
        }
      }
      status = __any_sync(0xffffffffu, status) ? 1 : 0;
#pragma unroll
      for (int k = 0; k < NCAND; ++k)
        if (k < ntypes) warp_reduce_winner(w[k], (k & 1) == 0);

      // ---- decide the next pair (working_set_select) ------------------------------------------
      int si = -1, sj = -1;
      double gi = 0, gj = 0, ai = 0, aj = 0;
      int fi = 0, fj = 0;
      bool stop_sel = false;
      V xi[QL], xj[QL];
      T ni = 0, nj = 0, Kii = 0, Kjj = 0, Kij = 0;
      bool have_k = false;
      if (forced && epoch == 1) {
        si = (int)P.forced_i;
        sj = (int)P.forced_j;
        gi = __ldcg(G + si); gj = __ldcg(G + sj);
        ai = __ldcg(alpha + si); aj = __ldcg(alpha + sj);
        fi = __ldcg(flags + si); fj = __ldcg(flags + sj);
      } else if (!nu) {
        // solver.py:66-75
        si = w[0].idx; sj = w[1].idx;
        gi = w[0].g; gj = w[1].g; ai = w[0].alpha; aj = w[1].alpha; fi = w[0].flags; fj = w[1].flags;
        stop_sel = si < 0 || sj < 0 || (gi - gj < P.tol);
      } else {
        // solver.py:300-339: no tol test, gain comparison between the class-wise pairs
        const bool pos_ok = w[0].idx >= 0 && w[1].idx >= 0;
        const bool neg_ok = w[2].idx >= 0 && w[3].idx >= 0;
        int pick = -1;   // 0 = '+' pair, 1 = '-' pair
        if (!pos_ok && !neg_ok) stop_sel = true;
        else if (!pos_ok) pick = 1;
        else if (!neg_ok) pick = 0;
        else {
          V xa[QL], xb[QL];
          T na, nb, Kaa, Kbb, Kab;
          const long long rpa = w[0].idx >= l && mirror ? w[0].idx - l : w[0].idx;
          const long long rpb = w[1].idx >= l && mirror ? w[1].idx - l : w[1].idx;
          pair_kernels<T, CH>(P, kp, rpa, rpb, lane, xa, xb, na, nb, Kaa, Kbb, Kab);
          double quad_p = ((double)Kaa + (double)Kbb) - 2.0 * (double)Kab;
          if (quad_p <= 0) quad_p = 1e-12;
          const double dp = w[0].g - w[1].g;
          const double gain_p = -(dp * dp) / quad_p;
          const long long rna = w[2].idx >= l && mirror ? w[2].idx - l : w[2].idx;
          const long long rnb = w[3].idx >= l && mirror ? w[3].idx - l : w[3].idx;
          pair_kernels<T, CH>(P, kp, rna, rnb, lane, xi, xj, ni, nj, Kii, Kjj, Kij);
          double quad_n = ((double)Kii + (double)Kjj) - 2.0 * (double)Kij;
          if (quad_n <= 0) quad_n = 1e-12;
          const double dn = w[2].g - w[3].g;
          const double gain_n = -(dn * dn) / quad_n;
          if (gain_p < gain_n) {
            pick = 0;
#pragma unroll
            for (int q = 0; q < QL; ++q) { xi[q] = xa[q]; xj[q] = xb[q]; }
            ni = na; nj = nb; Kii = Kaa; Kjj = Kbb; Kij = Kab;
          } else {
            pick = 1;
          }
          have_k = true;
        }
        if (pick >= 0) {
          const Winner& a = pick == 0 ? w[0] : w[2];
          const Winner& b = pick == 0 ? w[1] : w[3];
          si = a.idx; sj = b.idx; gi = a.g; gj = b.g; ai = a.alpha; aj = b.alpha; fi = a.flags; fj = b.flags;
        }
      }

      if (stop_sel) converged = 1;
      const bool exit_now = status || stop_sel || n_iter >= P.max_iter;
      if (exit_now) {
        if (lane == 0) {
          sh->stop = 1;
          if (cta == 0) {
            SolveResult r;
            r.n_iter = n_iter;
            r.last_i = stop_sel ? -1 : si;
            r.last_j = stop_sel ? -1 : sj;
            r.cold_sweeps = cold;
            r.di = l_di;
            r.dj = l_dj;
            r.converged = converged;
            r.status = status;
            r.pad[0] = r.pad[1] = 0;
            *P.result = r;
          }
        }
      } else {
        // ---- Solver.update / NuSolver.update: the analytic step (solver.py:82-145, 342-373) ----
        const long long ri = (mirror && si >= l) ? si - l : si;
        const long long rj = (mirror && sj >= l) ? sj - l : sj;
        if (!have_k) pair_kernels<T, CH>(P, kp, ri, rj, lane, xi, xj, ni, nj, Kii, Kjj, Kij);
        const double yi = (fi & F_YPOS) ? 1.0 : -1.0;
        const double yj = (fj & F_YPOS) ? 1.0 : -1.0;
        const StepOut so = smo_step(nu, (double)Kii, (double)Kjj, (double)Kij, yi, yj, gi, gj, ai, aj, Cbox);
        l_ai = so.ai; l_aj = so.aj; l_di = so.di; l_dj = so.dj; l_i = si; l_j = sj;
#pragma unroll
        for (int q = 0; q < QL; ++q) {
          const int c = lane + 32 * q;
          if (c < CH) { xs[c] = xi[q]; xs[CH + c] = xj[q]; }
        }
        if (lane == 0) {
          sh->ci = yi * so.di;
          sh->cj = yj * so.dj;
          sh->ai = so.ai;
          sh->aj = so.aj;
          sh->ni = ni;
          sh->nj = nj;
          sh->i = si;
          sh->j = sj;
          sh->fi = make_flags(yi > 0, so.ai, Cbox);
          sh->fj = make_flags(yj > 0, so.aj, Cbox);
        }
      }
    }
    named_bar_sync(2, NW * 32);   // (A) selection + step visible to all consumer warps
    if (sh->stop) break;

    // ---- sweep: both Q columns, gradient update, next candidates --------------------------------
    const double ci = sh->ci, cj = sh->cj;
    const int wi = sh->i, wj = sh->j, nfi = sh->fi, nfj = sh->fj;
    const double nai = sh->ai, naj = sh->aj;
    const T ni = sh->ni, nj = sh->nj;
    V xi[C::CPL], xj[C::CPL];
#pragma unroll
    for (int q = 0; q < C::CPL; ++q) {
      xi[q] = xs[lig + q * C::LPR];
      xj[q] = xs[CH + lig + q * C::LPR];
    }
    reset_best();

    for (int t = warp; t < NT; t += NW) {
      const unsigned long long Tg = Tbase + (unsigned long long)t;
      const int stage = resident ? t : (int)(Tg % S);
      const uint32_t par = resident ? 0u : (uint32_t)((Tg / S) & 1);
      const long long row = row0 + (long long)t * C::BR + rib;
      const bool valid = lane_active && row < row1;
      double g0 = 0, g1 = 0;
      int f0 = 0, f1 = 0;
      T nt = 0;
      if (valid) {
        g0 = G[row];
        f0 = flags[row];
        nt = norms[row];
        if (mirror) { g1 = G[row + l]; f1 = flags[row + l]; }
      }
      {
        uint32_t spins = 0;
        while (!mbar_try_wait(&full[stage], par)) {
          if ((++spins & 1023u) == 0 && *(volatile int*)P.abort_flag) break;
        }
      }
      const unsigned char* base = ring + (size_t)stage * C::TILE_BYTES + (size_t)grp * C::RPL * C::ROWB + lig * 16;
      T acc[C::NV];
#pragma unroll
      for (int k = 0; k < C::NV; ++k) acc[k] = 0;
#pragma unroll
      for (int k = 0; k < C::RPL; ++k) {
#pragma unroll
        for (int q = 0; q < C::CPL; ++q) {
          const V v = *reinterpret_cast<const V*>(base + k * C::ROWB + q * C::LPR * 16);
          acc[2 * k] = CK::dot(v, xi[q], acc[2 * k]);
          acc[2 * k + 1] = CK::dot(v, xj[q], acc[2 * k + 1]);
        }
      }
      TransposeReduce<T, C::NV, C::LPR / 2>::run(acc, lig);
      __syncwarp();
      if (!resident && lane == 0) mbar_arrive(&empty[stage]);
      T dti, dtj;
      if (C::FINAL == 2) {
        dti = acc[0];
        dtj = acc[1];
      } else {
        const T o = __shfl_xor_sync(0xffffffffu, acc[0], 1);
        dti = (lig & 1) ? o : acc[0];
        dtj = (lig & 1) ? acc[0] : o;
      }
      if (valid) {
        const T kti = kernel_value(kp, dti, nt, ni);
        const T ktj = kernel_value(kp, dtj, nt, nj);
        // y_t * (delta_i Q_i[t] + delta_j Q_j[t]) = (y_i delta_i) K_ti + (y_j delta_j) K_tj   (solver.py:146)
        const double c = ci * (double)kti + cj * (double)ktj;
        {
          const int v = (int)row;
          const double g = g0 - c;
          G[v] = g;
          int f = f0;
          if (v == wi) { f = nfi; alpha[v] = nai; flags[v] = (uint8_t)nfi; }
          else if (v == wj) { f = nfj; alpha[v] = naj; flags[v] = (uint8_t)nfj; }
          consider(nu, f, g, v, best);
        }
        if (mirror) {
          const int v = (int)(row + l);
          const double g = g1 - c;
          G[v] = g;
          int f = f1;
          if (v == wi) { f = nfi; alpha[v] = nai; flags[v] = (uint8_t)nfi; }
          else if (v == wj) { f = nfj; alpha[v] = naj; flags[v] = (uint8_t)nfj; }
          consider(nu, f, g, v, best);
        }
      }
    }
    Tbase += (unsigned long long)NT;
    ++n_iter;
    ++cold;
  }

  // ---- exit: tell the producer how far the consumers got, let it drain --------------------------
  if (threadIdx.x == 0) {
    sh->consumed = (long long)Tbase;
    __threadfence_block();
    sh->producer_stop = 1;
  }
}

}  // namespace b2svm
