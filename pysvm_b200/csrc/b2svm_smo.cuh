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
//   4. publishes its candidates as self-validating records; reading every CTA's record IS the
//      grid barrier (no atomics, no fences).
//
// X rows reach shared memory through a ring of 1-D TMA bulk copies (cp.async.bulk + mbarrier,
// SASS UBLKCP) issued by a dedicated producer thread that runs ahead of the consumers across
// iteration boundaries, so HBM keeps streaming while the exchange and the serial
// select -> step chain are in flight.  When a CTA's slice fits in the ring it is loaded once and
// stays resident in shared memory for the whole solve (small problems, multi-GPU shards).  Narrow rows move several
// 32-row tiles per ring stage (one bulk copy, one ticket); the copies of the leading 48 MB of X carry an L2
// evict_last hint so that part of every sweep is served from the 126 MB L2.
//
// Candidate exchange (one level, same code on one GPU and on a row-sharded box).  Every CTA writes, per
// candidate type, one 16-byte record {g, idx, epoch<<3 | flags} and one 16-byte unit {alpha, epoch} into the
// mailbox of EVERY GPU (its own included; peers over NVLink through CUDA-IPC mappings), each with a single
// relaxed 16-byte store, double-buffered by epoch parity.  A reader accepts a word only when it carries
// the epoch it waits for, so data and "flag" travel together: no atomics and no release/acquire fences
// (MEMBAR.SC + CCTL.IVALL cost 2-4 k cycles each on B200) on the per-iteration critical path.  All consumer
// warps of a CTA share the polling of the world * nCTA records (one L2 round trip), reduce through shared
// memory, and the leader warp finishes.  X (read-only) is replicated on every GPU of a sharded solve, so
// the winning rows are local reads; only the sweep, the gradient, alpha and the column cache are sharded.
// G / flags / alpha of a row are only ever touched by the CTA (= SM) that owns the row.
#pragma once
#include "b2svm_device.cuh"

namespace b2svm {

constexpr int NW = 11;                     // consumer warps per CTA (+1 producer warp = 384 threads, 168 regs)
constexpr int NTHREADS = (NW + 1) * 32;    // + one producer warp
constexpr int MAX_STAGES = 64;
constexpr int NCAND = 4;                   // up/low (C variant) or up+/low+/up-/low- (nu variant)
constexpr int TL_ITERS = 64, TL_SLOTS = 16; // debug timeline
constexpr int DBG_EPOCHS = 2048;
constexpr int TL_WARP_SLOTS = 4;           // + per-warp {start, end, tiles, wait} of one sweep
constexpr int TL_CTA_BASE = TL_ITERS * TL_SLOTS + 16 * TL_WARP_SLOTS;   // + per-CTA {sweep cycles, tiles} of one sweep
constexpr int TL_TOTAL = TL_CTA_BASE + 2 * 160;

struct SolveResult {
  long long n_iter;
  long long last_i, last_j;
  long long cold_sweeps;
  long long warm_sweeps, cols_stored, cache_used;
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
  int pin_lo, pin_hi;   // macro tiles [pin_lo, pin_hi) of every CTA's slice are fetched with L2 evict_last, the rest evict_first
  int sub;              // 32-row tiles per ring stage (narrow rows: several tiles travel in one bulk copy)
  long long ntiles;
  uint4* mbox;          // local mailbox [2][world * ncta][NCAND][2] {record, alpha unit} (zeroed before every launch)
  uint4* const* peer_mbox;   // [world] every rank's mailbox (peer mapping; own entry = mbox)
  int* abort_flag;
  SolveResult* result;
  long long max_iter;
  long long forced_i, forced_j;   // >= 0: skip the initial selection and take this pair first
  long long* timeline;            // optional [TL_ITERS][TL_SLOTS] clock64 stamps of CTA 0's leader lane
  // ---- device-resident Q-column cache (replaces the LRU of solver.py:207,227-229) -------------------
  void* cache;                    // T [cache_slots][l]: K(x_t, x_r) for this GPU's rows t, one slot per cached row r
  int* slot_of;                   // [l_global] slot of global row r, -1 = not cached (replicated per GPU)
  long long cache_slots;          // 0 = cache off
  long long cache_used;           // slots already filled by earlier solves on this handle
  // ---- row-sharded multi-GPU (world > 1): this GPU sweeps global rows [row_offset, row_offset + l); X and
  // norms hold ALL l_global rows on every GPU
  int world, rank;
  long long row_offset, l_global, rows_per_rank;   // every rank's slice starts at rank * rows_per_rank
  double* dbg_steps;              // optional [DBG_EPOCHS][ncta][6] = si, sj, gi, gj, ai, aj seen by each leader
  int debug_flags;                // B2SVM_DEBUG_FLAGS: 1 = static tile assignment, 2 = L2 loads for G/flags
};
constexpr size_t PROBLEM_SMEM_BYTES = (sizeof(ProblemDev) + 127) / 128 * 128;   // descriptor copy in front of the ring

// Tile geometry.  A tile is one contiguous bulk copy of BR rows of CH 16-byte chunks.  LPR lanes share
// a row (lane p of the group owns chunks p, p+LPR, ...: the 8 lanes of a quarter-warp read 128
// contiguous bytes, conflict-free), a warp-wide load therefore fetches GRP = 32/LPR rows, and a lane
// group walks RPL rows per tile.  x_i / x_j live in registers (CPL chunks each per lane), so the tile
// is the only shared-memory traffic.  The RPL*2 partial sums per lane are combined with a transposed
// reduction (log2(LPR) exchange rounds, NV - FINAL shuffles in total) that leaves lane r of the warp
// holding both dot products of row r of the tile.
template <typename T, int CH_>
struct Cfg {
  using CK = Chunk<T>;
  using V = typename CK::V;
  static constexpr int CH = CH_;                                   // 16-byte chunks per row
  static constexpr int LPR = CH <= 8 ? CH : (CH <= 32 ? 8 : CH / 4);   // lanes per row: 4, 8, 8, 8, 16, 32
  static constexpr int CPL = CH / LPR;                             // chunks per lane: 1, 1, 2, 4, 4, 4
  static constexpr int GRP = 32 / LPR;                             // rows fetched by one warp-wide load
  static constexpr int FINAL = (sizeof(T) == 8 && LPR == 32) ? 1 : 2;   // sums per lane after the reduce
  static constexpr int RPL = FINAL == 2 ? LPR : LPR / 2;           // rows per lane group per tile
  static constexpr int BR = GRP * RPL;                             // rows per tile (32, or 16)
  static constexpr int NV = 2 * RPL;                               // partial sums per lane
  static constexpr int ROWB = CH * 16;
  static constexpr int TILE_BYTES = BR * ROWB;
};

struct WCand {
  double g, alpha;
  int idx, flags;
};

// State only the leader warp needs between iterations.  It lives in shared memory and is loaded into
// registers for the duration of the leader block only, so it costs the sweep loop no registers.
struct LeaderState {
  double di, dj;          // last step
  int ri, rj, si, sj;     // ... its rows / cache slots (same deferral)
  int n_cached;           // cache slots handed out so far (identical in every leader)
  int converged;
  long long warm_cnt, stored_cnt;
};

template <typename T>
struct Shared {
  double ci, cj;        // y_i * delta_i, y_j * delta_j
  double ai, aj;        // new alpha_i, alpha_j
  T ni, nj;             // |x_i|^2, |x_j|^2
  int i, j, fi, fj;
  int slot_i, slot_j;   // cache slots of rows i, j (-1 = none); store_* = fill the slot during this sweep
  int store_i, store_j, warm;
  int p_ri, p_rj, p_si, p_sj;   // previous pair's rows / slots: slot_of[] is updated one sweep late, like alpha
  int stop;
  volatile int producer_stop;
  volatile long long consumed;
  LeaderState leader;
  int ticket;           // next dynamically assigned tile of the current sweep
  uint4* peer[8];       // mailbox of every rank (copied once from global memory)
  volatile int status;  // a warp's watchdog tripped
  WCand wc[NW][NCAND];  // per-warp candidates of the sweep
  uint4 wa[NW][NCAND];  // ... their alpha, {lo, hi, epoch, 0}: written when the load lands, after the CTA join
  WCand cw[NCAND];      // this CTA's winners
  WCand ws[NW][NCAND];  // per-warp winners of the polled records (alpha field unused, flags = source slot)
  int wsrc[NW][NCAND];  // ... and the mailbox source (rank * ncta + cta) they came from
};

__device__ __forceinline__ void consider(bool nu, int f, double g, int v, Cand (&b)[NCAND]) {
  const bool second = nu && !(f & F_YPOS);
  if (in_up(f)) {
    if (!second) {
      if (better_max(g, v, b[0].g, b[0].idx)) { b[0].g = g; b[0].idx = v; b[0].flags = f; }
    } else {
      if (better_max(g, v, b[2].g, b[2].idx)) { b[2].g = g; b[2].idx = v; b[2].flags = f; }
    }
  }
  if (in_low(f)) {
    if (!second) {
      if (better_min(g, v, b[1].g, b[1].idx)) { b[1].g = g; b[1].idx = v; b[1].flags = f; }
    } else {
      if (better_min(g, v, b[3].g, b[3].idx)) { b[3].g = g; b[3].idx = v; b[3].flags = f; }
    }
  }
}

// Where the leader finds variable `idx`: row and norm in the (replicated) X / norms arrays.
template <typename T>
struct VarRef {
  const typename Chunk<T>::V* row;
  const T* norm;
};
template <typename T, int CH>
__device__ __forceinline__ VarRef<T> locate(const ProblemDev& P, int idx) {
  using V = typename Chunk<T>::V;
  VarRef<T> r;
  const long long grow = (P.mirror && idx >= P.l_global) ? idx - P.l_global : idx;
  r.row = reinterpret_cast<const V*>(P.X) + grow * CH;
  r.norm = reinterpret_cast<const T*>(P.norms) + grow;
  return r;
}

// Leader warp: load rows a, b (16-byte chunks spread over the lanes) and evaluate
// K(a,a), K(b,b), K(a,b) in T.  Chunks stay in xa / xb for the caller.
template <typename T, int CH>
__device__ __forceinline__ void pair_kernels(const KernelParams<T>& kp, const VarRef<T>& A, const VarRef<T>& B,
                                             int lane, typename Chunk<T>::V (&xa)[(CH + 31) / 32],
                                             typename Chunk<T>::V (&xb)[(CH + 31) / 32], T& na, T& nb, T& Kaa,
                                             T& Kbb, T& Kab) {
  using CK = Chunk<T>;
  constexpr int Q = (CH + 31) / 32;
  na = *A.norm;
  nb = *B.norm;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int c = lane + 32 * q;
    xa[q] = (c < CH) ? A.row[c] : CK::zero();
    xb[q] = (c < CH) ? B.row[c] : CK::zero();
  }
  T daa = 0, dbb = 0, dab = 0;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    daa = CK::dot(xa[q], xa[q], daa);
    dbb = CK::dot(xb[q], xb[q], dbb);
    dab = CK::dot(xa[q], xb[q], dab);
  }
  daa = warp_sum(daa);
  dbb = warp_sum(dbb);
  dab = warp_sum(dab);
  Kaa = kernel_value(kp, daa, na, na);
  Kbb = kernel_value(kp, dbb, nb, nb);
  Kab = kernel_value(kp, dab, na, nb);
}

template <typename T>
struct RowState {   // gradient-side state of one row (and its mirror variable), prefetched a tile ahead
  double g0, g1;
  int f0, f1;
  T nt;
  T ki, kj;   // cached kernel values of this row against x_i, x_j (warm sweeps only)
};

// CACHE: the Q-column cache code is compiled in.  GENERIC: nu variant and mirror (2l-variable) problems
// are handled at run time; GENERIC = false is the plain C-SVC / one-class shape (two candidate types,
// one variable per row), which frees enough registers to keep the sweep loop spill-free.
template <typename T, int CH, bool CACHE, bool GENERIC>
__global__ void __launch_bounds__(NTHREADS, 1) smo_persistent(const ProblemDev* __restrict__ probs, int ncta_pp,
                                                              int ring_stages) {
  using C = Cfg<T, CH>;
  using CK = typename C::CK;
  using V = typename C::V;
  constexpr int QL = (CH + 31) / 32;

  // The problem descriptor is read all along the latency-critical exchange / step chain; in global memory its
  // lines are evicted from L1 by every sweep and each first touch costs an L2 round trip.  Keep a copy in
  // shared memory, in front of the ring.
  extern __shared__ __align__(128) unsigned char smem_raw[];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(probs + blockIdx.x / ncta_pp);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem_raw);
    for (int k = threadIdx.x; k < (int)(sizeof(ProblemDev) / 4); k += NTHREADS) dst[k] = src[k];
  }
  __syncthreads();
  const ProblemDev& P = *reinterpret_cast<const ProblemDev*>(smem_raw);
  unsigned char* smem = smem_raw + PROBLEM_SMEM_BYTES;
  unsigned char* ring = smem;
  // tiles per ring stage: a run-time choice of the host for narrow rows, the constant 1 (folded away) for 16 KB tiles
  const int SUB = C::TILE_BYTES >= 16384 ? 1 : P.sub;
  const size_t STAGE_BYTES = (size_t)SUB * C::TILE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)ring_stages * STAGE_BYTES);
  uint64_t* empty = full + MAX_STAGES;
  // gen[s] = number of tiles consumed out of stage s so far.  A consumer waits for use u of a stage on
  // full[s] only once gen[s] == u: mbarrier parity can tell neighbouring phases apart, not phases two
  // apart, and with dynamically drawn tiles a warp may otherwise run a whole ring ahead of a slow one.
  volatile unsigned* gen = reinterpret_cast<volatile unsigned*>(empty + MAX_STAGES);
  V* xs = reinterpret_cast<V*>(const_cast<unsigned*>(gen) + MAX_STAGES);      // [2][CH]
  Shared<T>* sh = reinterpret_cast<Shared<T>*>(xs + 2 * CH);

  const int cta = blockIdx.x % ncta_pp;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const unsigned S = (unsigned)ring_stages;

  // this CTA's slice of rows, in whole tiles; tiles are dealt out as evenly as possible
  const int ncta = P.ncta;
  const long long tbase = P.ntiles / ncta, trem = P.ntiles % ncta;
  const long long tile0 = (long long)cta * tbase + (cta < trem ? cta : trem);
  const int NT = (int)(tbase + (cta < trem ? 1 : 0));
  const long long row0 = tile0 * C::BR;
  long long row1 = (tile0 + NT) * C::BR;
  if (row1 > P.l) row1 = P.l;
  const int NM = (NT + SUB - 1) / SUB;            // ring stages' worth of tiles ("macro tiles") in this slice
  const bool resident = (unsigned)NM <= S;
  const bool nu = GENERIC && P.nu != 0;
  const bool mirror = GENERIC && P.mirror != 0;
  const long long l = P.l;
  const long long goff = P.row_offset, lg = P.l_global;   // global variable index = goff + row (+ lg)
  const int world = P.world;

  if (threadIdx.x == 0) {
    for (unsigned s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      gen[s] = 0u;
    }
    mbar_fence_init();
    sh->stop = 0;
    sh->producer_stop = 0;
    sh->consumed = 0;
    sh->ticket = NW;
    sh->status = 0;
    for (int r = 0; r < P.world && r < 8; ++r) sh->peer[r] = const_cast<uint4*>(P.peer_mbox[r]);
    LeaderState L0;
    L0.di = L0.dj = 0.0;
    L0.ri = L0.rj = L0.si = L0.sj = -1;
    L0.n_cached = (int)P.cache_used;
    L0.converged = 0;
    L0.warm_cnt = L0.stored_cnt = 0;
    sh->leader = L0;
  }
  __syncthreads();

  // ============================== producer warp ==============================================
  if (warp == NW) {
    if (NM > 0 && P.kind != K_PRECOMPUTED) {   // a precomputed kernel matrix is only ever read through the cache path
      const unsigned char* Xb = reinterpret_cast<const unsigned char*>(P.X) + (size_t)P.row_offset * C::ROWB;
      unsigned long long Tp = 0;
      unsigned stage = 0, par = 0;   // stage = Tp % S, par = (Tp / S) & 1
      int t = 0;                     // Tp % NT
      // L2 residency plan: the same part of every slice is kept (evict_last) across sweeps, the rest streams through
      const int pin_lo = P.pin_lo, pin_hi = P.pin_hi;
      const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
      auto issue = [&](int tt, unsigned st) {   // macro tile tt -> stage st, one bulk copy
        const long long r0 = row0 + (long long)tt * SUB * C::BR;
        long long nr = row1 - r0;
        if (nr > (long long)SUB * C::BR) nr = (long long)SUB * C::BR;
        unsigned char* dst = ring + (size_t)st * STAGE_BYTES;
        if (lane == 0) {
          const uint32_t bytes = (uint32_t)(nr * C::ROWB);
          mbar_arrive_expect_tx(&full[st], bytes);
          if (pin_hi > pin_lo) bulk_g2s_hint(dst, Xb + r0 * C::ROWB, bytes, &full[st], (tt >= pin_lo && tt < pin_hi) ? pol_keep : pol_stream);
          else bulk_g2s(dst, Xb + r0 * C::ROWB, bytes, &full[st]);
        }
      };
      if (resident) {
        for (int tt = 0; tt < NM; ++tt) issue(tt, (unsigned)tt);
        Tp = NM;
        if (lane == 0)
          while (!sh->producer_stop) __nanosleep(200);
      } else {
        for (;;) {
          int go = 1;
          if (lane == 0) {
            while (!mbar_try_wait(&empty[stage], par ^ 1u)) {
              if (sh->producer_stop) { go = 0; break; }
            }
            if (sh->producer_stop) go = 0;
          }
          go = __shfl_sync(0xffffffffu, go, 0);
          if (!go) break;
          issue(t, stage);
          ++Tp;
          if (++t == NM) t = 0;
          if (++stage == S) { stage = 0; par ^= 1u; }
        }
      }
      // drain: never leave the CTA with bulk copies still landing in its shared memory
      if (lane == 0) {
        const unsigned long long Tc = resident ? 0ull : (unsigned long long)sh->consumed;
        for (unsigned long long Tg = Tc; Tg < Tp; ++Tg) {
          const unsigned st = resident ? (unsigned)Tg : (unsigned)(Tg % S);
          const uint32_t pp = resident ? 0u : (uint32_t)((Tg / S) & 1);
          while (!mbar_try_wait(&full[st], pp)) {}
        }
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
  const T* __restrict__ norms = reinterpret_cast<const T*>(P.norms) + P.row_offset;   // this GPU's rows
  double* __restrict__ G = P.G;
  double* __restrict__ alpha = P.alpha;
  uint8_t* __restrict__ flags = P.flags;
  const int ntypes = (GENERIC && nu) ? 4 : 2;
  const double Cbox = P.C;

  const int lig = lane % C::LPR;      // lane in group = which chunks of the row I own
  const int grp = lane / C::LPR;
  const int rib = C::FINAL == 2 ? lane : grp * C::RPL + (lig >> 1);   // my row within a tile (epilogue)
  const bool lane_active = C::FINAL == 2 ? true : !(lig & 1);

  unsigned tb_mod = 0, tb_div = 0;   // (tiles consumed so far) % S and / S, kept incrementally
#ifdef B2SVM_TIMELINE
  long long cta_sweep_start_prev = 0;
#endif
  long long cold = 0;                // cold sweeps so far (every warp counts; the producer drains cold * NT tiles)
  uint32_t epoch = 0;                // exchanges so far; sweeps completed = epoch - 1

  Cand best[NCAND];
  auto reset_best = [&]() {
#pragma unroll
    for (int k = 0; k < NCAND; ++k) { best[k].g = 0.0; best[k].idx = -1; best[k].flags = 0; }
  };
#ifdef B2SVM_TIMELINE
  auto stamp = [&](int slot) {
    if (P.timeline && cta == 0 && threadIdx.x == 0 && epoch >= 1 && epoch <= (uint32_t)TL_ITERS)
      P.timeline[(epoch - 1) * TL_SLOTS + slot] = clock64();
  };
#else
  auto stamp = [&](int) {};
#endif
  // local index of a variable this GPU owns
  auto local_var = [&](int idx) -> long long {
    const bool second = mirror && idx >= lg;
    return (second ? l : 0) + ((second ? idx - lg : idx) - goff);
  };
  // cache columns in use by the current / next sweep (set after bar (A))
  bool sweep_warm = false;
  const T* cache_i = nullptr;
  const T* cache_j = nullptr;
  auto load_row_state = [&](int t) {
    RowState<T> r;
    r.g0 = r.g1 = 0.0;
    r.f0 = r.f1 = 0;
    r.nt = 0;
    r.ki = r.kj = 0;
    const long long row = row0 + (long long)t * C::BR + rib;
    if (t < NT && lane_active && row < row1) {
      if (P.debug_flags & 2) {
        r.g0 = __ldcg(G + row);
        r.f0 = __ldcg(flags + row);
        if (mirror) { r.g1 = __ldcg(G + row + l); r.f1 = __ldcg(flags + row + l); }
      } else {
        r.g0 = G[row];
        r.f0 = flags[row];
        if (mirror) { r.g1 = G[row + l]; r.f1 = flags[row + l]; }
      }
      r.nt = norms[row];
      if (CACHE && sweep_warm) {
        r.ki = cache_i[row];
        r.kj = cache_j[row];
      }
    }
    return r;
  };

  // K(a,a), K(b,b), K(a,b) for the leader: from the rows, or straight from a precomputed kernel matrix
  auto pair_k = [&](int ia, int ib, V (&xa)[QL], V (&xb)[QL], T& na, T& nb, T& Kaa, T& Kbb, T& Kab) {
    if (CACHE && P.kind == K_PRECOMPUTED) {
      const long long ra = (mirror && ia >= lg) ? ia - lg : ia, rb = (mirror && ib >= lg) ? ib - lg : ib;
      const T* Kc = reinterpret_cast<const T*>(P.cache);
      na = nb = 0;
      Kaa = __ldg(Kc + ra * l + ra);
      Kbb = __ldg(Kc + rb * l + rb);
      Kab = __ldg(Kc + ra * l + rb);
    } else {
      pair_kernels<T, CH>(kp, locate<T, CH>(P, ia), locate<T, CH>(P, ib), lane, xa, xb, na, nb, Kaa, Kbb, Kab);
    }
  };

  // ---- initial selection pass over G / flags only (no X) -------------------------------------
  reset_best();
  const bool forced = P.forced_i >= 0;
  if (!forced) {
    for (long long r = row0 + warp * 32 + lane; r < row1; r += NW * 32) {
      consider(nu, flags[r], G[r], (int)(r + goff), best);
      if (mirror) consider(nu, flags[r + l], G[r + l], (int)(r + goff + lg), best);
    }
  }
  RowState<T> pf = load_row_state(warp * SUB);   // first tile of the first sweep (macro tiles 0..NW-1 are static)

  for (;;) {
    // ---- (B) warp -> CTA candidate reduce, publish to every GPU, shared polling, global reduce ----------
    ++epoch;
    stamp(0);   // this warp's sweep done
    const uint32_t par = epoch & 1u;
#pragma unroll
    for (int k = 0; k < NCAND; ++k)
      if (k < ntypes) redux_select(0xffffffffu, (k & 1) == 0, best[k]);
    __syncwarp();
    // one lane per type also fetches the candidate's alpha (a row of this warp: same-SM coherent).  The load is
    // only issued here: its value is parked in shared memory after the CTA join and travels in a separate 16-byte
    // unit, so neither the join nor the candidate record waits for it (the leaders need it at step time only).
    double cand_alpha = 0.0;
    if (lane < ntypes) {
      Cand m = best[0];
#pragma unroll
      for (int k = 1; k < NCAND; ++k)
        if (lane == k) m = best[k];
      if (m.idx >= 0) cand_alpha = alpha[local_var(m.idx)];
      WCand wv;
      wv.g = m.g;
      wv.idx = m.idx;
      wv.flags = m.flags;
      wv.alpha = 0.0;
      sh->wc[warp][lane] = wv;
    }
    named_bar_sync(1, NW * 32);
    stamp(1);   // all warps of this CTA done
    auto park_alpha = [&]() {
      if (lane < ntypes)
        st_volatile_shared_v4(&sh->wa[warp][lane],
                              make_uint4((uint32_t)__double2loint(cand_alpha), (uint32_t)__double2hiint(cand_alpha), epoch, 0u));
    };
    if (warp != 0) park_alpha();

    if (warp == 0) {
      // CTA-level winner per type.  Full-warp collectives only: with two half-warp masks in one instruction the
      // compiler falls back to a MATCH.ANY / WARPSYNC.EXCLUSIVE emulation that costs ~1000 cycles on this path.
      for (int k = 0; k < ntypes; ++k) {
        WCand mw;
        mw.g = 0.0; mw.alpha = 0.0; mw.idx = -1; mw.flags = 0;
        if (lane < NW) mw = sh->wc[lane][k];
        Cand m;
        m.g = mw.g; m.idx = mw.idx; m.flags = mw.flags;
        redux_select(0xffffffffu, (k & 1) == 0, m);
        const unsigned holders = __ballot_sync(0xffffffffu, m.idx >= 0 && mw.idx == m.idx);
        const int from = holders ? __ffs(holders) - 1 : 0;
        if (lane == 0) {
          WCand c;
          c.g = m.g; c.alpha = 0.0; c.idx = m.idx; c.flags = (m.flags & 7) | (from << 8);   // + the warp it came from
          sh->cw[k] = c;
        }
      }
      __syncwarp();
      stamp(7);   // CTA winners known
      // ... pushed into the mailbox of every GPU: one lane per (rank, type), the record first ...
      const int pr = lane / ntypes, pk = lane - pr * ntypes;
      uint4* rp = nullptr;
      int from_warp = 0;
      if (lane < world * ntypes) {
        const WCand c = sh->cw[pk];
        from_warp = c.flags >> 8;
        rp = sh->peer[pr] + ((((size_t)par * world + P.rank) * ncta + cta) * NCAND + pk) * 2;
        const uint4 rec = make_uint4((uint32_t)__double2loint(c.g), (uint32_t)__double2hiint(c.g), (uint32_t)c.idx,
                                     (epoch << 3) | (uint32_t)(c.flags & 7));
        if (world == 1) st_relaxed_v4(rp, rec);
        else st_relaxed_sys_v4(rp, rec);
      }
      if (lane == 0) sh->ticket = NW;   // next sweep: tiles 0..NW-1 are static, the rest are drawn here
      stamp(2);   // published
      // ... then the winner's alpha, as soon as the warp that owns it has parked it
      park_alpha();
      __syncwarp();
      if (lane < world * ntypes) {
        uint4 u = ld_volatile_shared_v4(&sh->wa[from_warp][pk]);
        while (u.z != epoch) u = ld_volatile_shared_v4(&sh->wa[from_warp][pk]);
        const uint4 au = make_uint4(u.x, u.y, 0u, epoch);
        if (world == 1) st_relaxed_v4(rp + 1, au);
        else st_relaxed_sys_v4(rp + 1, au);
      }
#ifdef B2SVM_TIMELINE
      if (P.timeline && threadIdx.x == 0 && (epoch == 11u || epoch == 41u) && cta < 160)   // two sweeps: is a slow CTA always slow?
        P.timeline[TL_CTA_BASE + 2 * cta + (epoch == 41u)] = clock64() - cta_sweep_start_prev;
#endif
    }

    // ---- every consumer warp polls its share of the world * ncta records (this is the grid / box barrier)
    {
      const int NS = world * ncta;
      const int per = (NS + NW - 1) / NW;
      const int s_begin = warp * per;
      const int s_end = min(NS, s_begin + per);
      constexpr int MAXM = 4;   // per <= ceil(8 * 148 / 11) = 108 <= 4 * 32
      const uint4* rbase = P.mbox + (size_t)par * NS * NCAND * 2;
      Cand w[NCAND];
      int wsrc[NCAND];
      const uint64_t t_start = globaltimer_ns();
      uint32_t rounds = 0;
      for (;;) {
        bool all = true;
#pragma unroll
        for (int k = 0; k < NCAND; ++k) { w[k].g = 0.0; w[k].idx = -1; w[k].flags = 0; wsrc[k] = 0; }
        // all loads of a pass are issued before any is looked at: one round trip per pass, not one per record
        for (int k0 = 0; k0 < ntypes; k0 += 2) {
          uint4 h[MAXM][2];
#pragma unroll
          for (int m = 0; m < MAXM; ++m) {
            const int sidx = s_begin + lane + 32 * m;
            if (sidx < s_end) {
              if (world == 1) {   // gpu scope is enough (and cheaper) when no peer writes into this mailbox
                h[m][0] = ld_relaxed_v4(rbase + ((size_t)sidx * NCAND + k0) * 2);
                h[m][1] = ld_relaxed_v4(rbase + ((size_t)sidx * NCAND + k0 + 1) * 2);
              } else {
                h[m][0] = ld_relaxed_sys_v4(rbase + ((size_t)sidx * NCAND + k0) * 2);
                h[m][1] = ld_relaxed_sys_v4(rbase + ((size_t)sidx * NCAND + k0 + 1) * 2);
              }
            }
          }
#pragma unroll
          for (int m = 0; m < MAXM; ++m) {
            const int sidx = s_begin + lane + 32 * m;
            if (sidx < s_end) {
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint4 hh = h[m][kk];
                if ((hh.w >> 3) != epoch) {
                  all = false;
                } else {
                  const double g = __hiloint2double((int)hh.y, (int)hh.x);
                  const int idx = (int)hh.z;
                  Cand& wk = (k0 == 0) ? w[kk] : w[2 + kk];
                  int& ws = (k0 == 0) ? wsrc[kk] : wsrc[2 + kk];
                  const bool take = kk == 0 ? better_max(g, idx, wk.g, wk.idx) : better_min(g, idx, wk.g, wk.idx);
                  if (take) { wk.g = g; wk.idx = idx; wk.flags = (int)(hh.w & 7u); ws = sidx; }
                }
              }
            }
          }
        }
        if (__all_sync(0xffffffffu, all)) break;
        if ((++rounds & 63u) == 0) {
          int bad = *(volatile int*)P.abort_flag | sh->status;
          if (!bad && globaltimer_ns() - t_start > 6000000000ull) {
            *(volatile int*)P.abort_flag = 1;
            bad = 1;
          }
          if (__any_sync(0xffffffffu, bad)) { sh->status = 1; break; }
        }
      }
      // warp winner per type, with the mailbox source it came from
#pragma unroll
      for (int k = 0; k < NCAND; ++k) {
        if (k < ntypes) {
          const int mine = w[k].idx;
          redux_select(0xffffffffu, (k & 1) == 0, w[k]);
          const unsigned holders = __ballot_sync(0xffffffffu, w[k].idx >= 0 && mine == w[k].idx);
          const int from = holders ? __ffs(holders) - 1 : 0;
          wsrc[k] = __shfl_sync(0xffffffffu, wsrc[k], from);
        }
      }
      // CTA 0 pulls the rows of its warp winners towards L2 now (a row of another GPU's slice is never swept here,
      // a local one may have been streamed out): the leaders' fetch of the pair's rows after the exchange then hits
      if (cta == 0 && P.kind != K_PRECOMPUTED) {
#pragma unroll
        for (int k = 0; k < NCAND; ++k) {
          if (k < ntypes && w[k].idx >= 0) {
            const long long grow = (mirror && w[k].idx >= lg) ? w[k].idx - lg : w[k].idx;
            if (lane * 128 < C::ROWB)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(P.X) + (size_t)grow * C::ROWB + lane * 128));
          }
        }
      }
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NCAND; ++k) {
          WCand c;
          c.g = w[k].g; c.alpha = 0.0; c.idx = w[k].idx; c.flags = w[k].flags;
          sh->ws[warp][k] = c;
          sh->wsrc[warp][k] = wsrc[k];
        }
      }
    }
    named_bar_sync(3, NW * 32);
    stamp(3);   // every record of the box read

    if (warp == 0) {
      LeaderState L = sh->leader;
      int status = sh->status;
      const long long n_iter = (long long)epoch - 1;
      // global winners: reduce the NW warp winners (16 lanes per type), keep the source slot of each
      Cand w[NCAND];
      int wsrc[NCAND];
#pragma unroll
      for (int k = 0; k < NCAND; ++k) { w[k].g = 0.0; w[k].idx = -1; w[k].flags = 0; wsrc[k] = 0; }
      for (int k = 0; k < ntypes; ++k) {   // full-warp collectives, one type at a time (see the CTA-level reduce)
        Cand m;
        m.g = 0.0; m.idx = -1; m.flags = 0;
        int msrc = 0;
        if (lane < NW) {
          const WCand c = sh->ws[lane][k];
          m.g = c.g; m.idx = c.idx; m.flags = c.flags;
          msrc = sh->wsrc[lane][k];
        }
        const int mine = m.idx;
        redux_select(0xffffffffu, (k & 1) == 0, m);
        const unsigned holders = __ballot_sync(0xffffffffu, m.idx >= 0 && mine == m.idx);
        const int from = holders ? __ffs(holders) - 1 : 0;
        msrc = __shfl_sync(0xffffffffu, msrc, from);
#pragma unroll
        for (int q = 0; q < NCAND; ++q)
          if (q == k) { w[q] = m; wsrc[q] = msrc; }
      }
      // alpha of a winner: the 16-byte unit its CTA pushed next to the record (polled until it carries the epoch)
      const uint4* abase = P.mbox + (size_t)par * world * ncta * NCAND * 2;
      auto alpha_unit = [&](int type) -> const uint4* { return abase + ((size_t)wsrc[type] * NCAND + type) * 2 + 1; };
      auto load_unit = [&](int type) -> uint4 {
        return world == 1 ? ld_relaxed_v4(alpha_unit(type)) : ld_relaxed_sys_v4(alpha_unit(type));
      };
      auto settle_alpha = [&](uint4 u, int type) -> double {   // re-poll until the unit carries this epoch
        uint32_t spins = 0;
        while (u.w != epoch) {
          u = load_unit(type);
          if ((++spins & 4095u) == 0 && *(volatile int*)P.abort_flag) break;
        }
        return __hiloint2double((int)u.y, (int)u.x);
      };
      stamp(4);   // global winners known

      // ---- decide the next pair (working_set_select) ------------------------------------------
      int si = -1, sj = -1;
      double gi = 0, gj = 0, ai = 0, aj = 0;
      int fi = 0, fj = 0;
      bool stop_sel = false;
      V xi[QL], xj[QL];
#pragma unroll
      for (int q = 0; q < QL; ++q) { xi[q] = CK::zero(); xj[q] = CK::zero(); }
      T ni = 0, nj = 0, Kii = 0, Kjj = 0, Kij = 0;
      bool have_k = false;
      int ti = 0, tj = 1;   // candidate types the pair came from
      if (forced && epoch == 1) {
        si = (int)P.forced_i;     // single GPU only (checked on the host)
        sj = (int)P.forced_j;
        gi = __ldcg(G + si); gj = __ldcg(G + sj);
        ai = __ldcg(alpha + si); aj = __ldcg(alpha + sj);
        fi = __ldcg(flags + si); fj = __ldcg(flags + sj);
      } else if (!nu) {
        // solver.py:66-75
        si = w[0].idx; sj = w[1].idx;
        gi = w[0].g; gj = w[1].g; fi = w[0].flags; fj = w[1].flags;
        stop_sel = si < 0 || sj < 0 || (gi - gj < P.tol);
        // (alpha_i, alpha_j are fetched together with the rows in the step below)
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
          pair_k(w[0].idx, w[1].idx, xa, xb, na, nb, Kaa, Kbb, Kab);
          double quad_p = ((double)Kaa + (double)Kbb) - 2.0 * (double)Kab;
          if (quad_p <= 0) quad_p = 1e-12;
          const double dp = w[0].g - w[1].g;
          const double gain_p = -(dp * dp) / quad_p;
          pair_k(w[2].idx, w[3].idx, xi, xj, ni, nj, Kii, Kjj, Kij);
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
          const Cand& a = pick == 0 ? w[0] : w[2];
          const Cand& b = pick == 0 ? w[1] : w[3];
          si = a.idx; sj = b.idx; gi = a.g; gj = b.g; fi = a.flags; fj = b.flags;
          ti = pick == 0 ? 0 : 2;
          tj = ti + 1;
        }
      }

#ifdef B2SVM_TIMELINE
      if (P.dbg_steps && lane == 0 && epoch <= (uint32_t)DBG_EPOCHS) {
        double* o = P.dbg_steps + ((size_t)(epoch - 1) * ncta + cta) * 6;
        o[0] = si; o[1] = sj; o[2] = gi; o[3] = gj; o[4] = ai; o[5] = aj;
      }
#endif
      if (stop_sel) L.converged = 1;
      const bool exit_now = status || stop_sel || n_iter >= P.max_iter;
      if (exit_now) {
        if (lane == 0) {
          sh->stop = 1;
          if (cta == 0) {
            SolveResult r;
            r.n_iter = n_iter;
            r.last_i = stop_sel ? -1 : si;
            r.last_j = stop_sel ? -1 : sj;
            r.cold_sweeps = n_iter - L.warm_cnt;
            r.warm_sweeps = L.warm_cnt;
            r.cols_stored = L.stored_cnt;
            r.cache_used = L.n_cached;
            if (CACHE && P.cache_slots > 0 && L.ri >= 0) {   // pending slot-table entries of the last step
              if (L.si >= 0) P.slot_of[L.ri] = L.si;
              if (L.sj >= 0) P.slot_of[L.rj] = L.sj;
            }
            r.di = L.di;
            r.dj = L.dj;
            r.converged = L.converged;
            r.status = status;
            r.pad[0] = r.pad[1] = 0;
            *P.result = r;
          }
        }
      } else {
        // ---- Solver.update / NuSolver.update: the analytic step (solver.py:82-145, 342-373) ----
        // alpha units and the two rows are requested together: one memory round trip, not three
        uint4 ua = make_uint4(0u, 0u, 0u, 0u), ub = ua;
        const bool from_records = !(forced && epoch == 1);
        if (from_records) { ua = load_unit(ti); ub = load_unit(tj); }
        // (the slot-table entries of the two rows ride along in the same round trip)
        const int gri = (int)((mirror && si >= lg) ? si - lg : si), grj = (int)((mirror && sj >= lg) ? sj - lg : sj);
        int raw_si = -1, raw_sj = -1;
        if (CACHE && P.cache_slots > 0) { raw_si = __ldcg(P.slot_of + gri); raw_sj = __ldcg(P.slot_of + grj); }
        stamp(8);
        if (!have_k) pair_k(si, sj, xi, xj, ni, nj, Kii, Kjj, Kij);
        stamp(9);
        if (from_records) { ai = settle_alpha(ua, ti); aj = settle_alpha(ub, tj); }
        stamp(10);
        const double yi = (fi & F_YPOS) ? 1.0 : -1.0;
        const double yj = (fj & F_YPOS) ? 1.0 : -1.0;
        const StepOut so = smo_step(nu, (double)Kii, (double)Kjj, (double)Kij, yi, yj, gi, gj, ai, aj, Cbox);
        stamp(11);
        // ---- Q-column cache: which of the two columns are already resident, which get stored now --------
        int slot_i = -1, slot_j = -1, store_i = 0, store_j = 0;
        if (CACHE && P.cache_slots > 0) {
          slot_i = gri == L.ri ? L.si : (gri == L.rj ? L.sj : raw_si);
          slot_j = grj == L.ri ? L.si : (grj == L.rj ? L.sj : raw_sj);
          if (slot_i < 0 && L.n_cached < P.cache_slots) { slot_i = L.n_cached++; store_i = 1; }
          if (grj == gri) { slot_j = slot_i; }
          else if (slot_j < 0 && L.n_cached < P.cache_slots) { slot_j = L.n_cached++; store_j = 1; }
        }
        const int warm = (CACHE && P.cache_slots > 0 && slot_i >= 0 && slot_j >= 0 && !store_i && !store_j) ? 1 : 0;
        if (lane == 0) {
          sh->slot_i = slot_i; sh->slot_j = slot_j; sh->store_i = store_i; sh->store_j = store_j; sh->warm = warm;
          sh->p_ri = L.ri; sh->p_rj = L.rj; sh->p_si = L.si; sh->p_sj = L.sj;
        }
        L.ri = gri; L.rj = grj; L.si = slot_i; L.sj = slot_j;
        if (warm) ++L.warm_cnt;
        L.stored_cnt += store_i + store_j;
        L.di = so.di; L.dj = so.dj;
#ifdef B2SVM_TIMELINE
        if (P.dbg_steps && lane == 0 && cta == 0 && epoch <= (uint32_t)DBG_EPOCHS) {
          double* o = P.dbg_steps + (size_t)DBG_EPOCHS * ncta * 6 + (size_t)(epoch - 1) * 2;
          o[0] = so.ai; o[1] = so.aj;
        }
#endif
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
      __syncwarp();
      if (lane == 0) sh->leader = L;
      stamp(5);   // leader done: step taken, x_i / x_j staged
    }
    named_bar_sync(2, NW * 32);   // (A) selection + step visible to all consumer warps
    if (sh->stop) break;
    stamp(6);   // sweep starts

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
    // ---- Q-column cache plan of this sweep ---------------------------------------------------------
    const bool warm = CACHE && sh->warm != 0;
    const int store_i = CACHE ? sh->store_i : 0, store_j = CACHE ? sh->store_j : 0;
    T* const cache_base = reinterpret_cast<T*>(P.cache);
    T* const col_i = sh->slot_i >= 0 ? cache_base + (size_t)sh->slot_i * (size_t)l : nullptr;
    T* const col_j = sh->slot_j >= 0 ? cache_base + (size_t)sh->slot_j * (size_t)l : nullptr;
    sweep_warm = warm;
    cache_i = col_i;
    cache_j = col_j;
    if (CACHE && P.cache_slots > 0 && cta == 0 && warp == NW - 1 && lane == 0 && sh->p_ri >= 0) {
      // the previous step's slot-table entries, one sweep late (a slow leader may still have been reading them)
      if (sh->p_si >= 0) P.slot_of[sh->p_ri] = sh->p_si;
      if (sh->p_sj >= 0) P.slot_of[sh->p_rj] = sh->p_sj;
      __threadfence();
    }
    if (warm) {   // the first tile's state was prefetched before this sweep's plan was known
      const long long r0w = row0 + (long long)warp * SUB * C::BR + rib;
      if (warp < NM && lane_active && r0w < row1) { pf.ki = col_i[r0w]; pf.kj = col_j[r0w]; }
    }
#ifdef B2SVM_TIMELINE
    long long cta_sweep_start = 0;
    if (P.timeline && threadIdx.x == 0 && (epoch == 10u || epoch == 40u)) cta_sweep_start = clock64() + (long long)(ci == 123.456);
    const bool dbg = P.timeline && cta == 0 && epoch == 10u;
    const long long dbg_start = dbg ? clock64() : 0;
    long long dbg_wait = 0;
    int dbg_tiles = 0;
#endif
    int m = warp, sub = 0;              // macro tile (first one static: its gradient state is already in pf), tile in it
    while (m < NM) {
      const int t = m * SUB + sub;
      // next tile: the following one of this macro tile, else draw a new macro tile -- now, so that its
      // gradient state can be prefetched a full tile of work ahead
      int mn = m, subn = sub + 1;
      if (subn == SUB || t + 1 >= NT) {
        subn = 0;
        if (P.debug_flags & 1) {
          mn = m + NW;
        } else {
          if (lane == 0) mn = atomicAdd(&sh->ticket, 1);
          mn = __shfl_sync(0xffffffffu, mn, 0);
        }
      }
      const RowState<T> cur = pf;
      pf = load_row_state(mn * SUB + subn);

      const long long row = row0 + (long long)t * C::BR + rib;
      const bool valid = lane_active && row < row1;
      T dti = 0, dtj = 0;
      if (!warm) {
        unsigned stage, use = 0u;
        if (resident) {
          stage = (unsigned)m;
        } else {
          const unsigned q = tb_mod + (unsigned)m;
          stage = q % S;
          use = tb_div + q / S;
        }
#ifdef B2SVM_TIMELINE
        const long long dbg_w0 = dbg ? clock64() : 0;
#endif
        {
          uint32_t spins = 0;
          uint64_t t_wait = 0;
          bool ok = false;
          while (!ok && sub == 0) {   // the later tiles of a macro tile arrived with its first
            ok = (resident || gen[stage] == use) && mbar_try_wait(&full[stage], use & 1u);
            if (!ok && (++spins & 1023u) == 0) {
              if (*(volatile int*)P.abort_flag) break;
              const uint64_t now = globaltimer_ns();
              if (t_wait == 0) t_wait = now;
              else if (now - t_wait > 4000000000ull) { *(volatile int*)P.abort_flag = 1; break; }
            }
          }
        }
#ifdef B2SVM_TIMELINE
        if (dbg) { dbg_wait += clock64() - dbg_w0; ++dbg_tiles; }
#endif
        const unsigned char* base = ring + (size_t)stage * STAGE_BYTES + (size_t)sub * C::TILE_BYTES + (size_t)grp * C::RPL * C::ROWB + lig * 16;
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
        if (C::FINAL == 2) {
          dti = acc[0];
          dtj = acc[1];
        } else {
          const T o = __shfl_xor_sync(0xffffffffu, acc[0], 1);
          dti = (lig & 1) ? o : acc[0];
          dtj = (lig & 1) ? acc[0] : o;
        }
        __syncwarp();
        if (!resident && lane == 0 && subn == 0) {   // leaving this macro tile: hand the stage back
          gen[stage] = use + 1u;
          mbar_arrive(&empty[stage]);
        }
      }
      if (valid) {
        // a cached value is bit-identical to the recomputed one (same code, same summation order)
        const T kti = warm ? cur.ki : kernel_value(kp, dti, cur.nt, ni);
        const T ktj = warm ? cur.kj : kernel_value(kp, dtj, cur.nt, nj);
        if (store_i) col_i[row] = kti;
        if (store_j) col_j[row] = ktj;
        // y_t * (delta_i Q_i[t] + delta_j Q_j[t]) = (y_i delta_i) K_ti + (y_j delta_j) K_tj   (solver.py:146)
        const double c = ci * (double)kti + cj * (double)ktj;
        {
          const int v = (int)(row + goff);
          const double g = cur.g0 - c;
          G[row] = g;
          int f = cur.f0;
          if (v == wi) { f = nfi; flags[row] = (uint8_t)nfi; alpha[row] = nai; }
          else if (v == wj) { f = nfj; flags[row] = (uint8_t)nfj; alpha[row] = naj; }
          consider(nu, f, g, v, best);
        }
        if (mirror) {
          const int v = (int)(row + goff + lg);
          const double g = cur.g1 - c;
          G[row + l] = g;
          int f = cur.f1;
          if (v == wi) { f = nfi; flags[row + l] = (uint8_t)nfi; alpha[row + l] = nai; }
          else if (v == wj) { f = nfj; flags[row + l] = (uint8_t)nfj; alpha[row + l] = naj; }
          consider(nu, f, g, v, best);
        }
      }
      m = mn;
      sub = subn;
    }
#ifdef B2SVM_TIMELINE
    if (cta_sweep_start) cta_sweep_start_prev = cta_sweep_start;
    if (dbg && lane == 0) {
      long long* o = P.timeline + TL_ITERS * TL_SLOTS + warp * TL_WARP_SLOTS;
      o[0] = dbg_start; o[1] = clock64(); o[2] = dbg_tiles; o[3] = dbg_wait;
    }
#endif
    sweep_warm = false;
    pf = load_row_state(warp * SUB);   // first tile of the next sweep (after this thread's own stores to it)
    if (!warm) {
      if (!resident) {
        const unsigned q = tb_mod + (unsigned)NM;
        tb_mod = q % S;
        tb_div = tb_div + q / S;
      }
      ++cold;
    }
  }

  // ---- exit: tell the producer how far the consumers got, let it drain --------------------------
  if (threadIdx.x == 0) {
    sh->consumed = cold * (long long)NM;
    __threadfence_block();
    sh->producer_stop = 1;
  }
}

}  // namespace b2svm
