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
// Candidate exchange, two levels, no atomics and no fences on the per-iteration critical path (MEMBAR.SC + CCTL.IVALL
// cost 2-4 k cycles each on B200): every word that is polled carries the iteration ("epoch") it belongs to, so data
// and flag travel together in single 16-byte relaxed stores, double-buffered by epoch parity.
//   level 1 (inside a GPU): every CTA pushes, per candidate type, one record {key, key, ~(idx<<3|flags), epoch}
//     into the PRIVATE mailbox of every CTA of the GPU (all-to-all: ncta stores per record,
//     five per lane); ONE warp per candidate type ("type warp": warp k handles type k, so the arg-max and the
//     arg-min chains run side by side) polls the ncta records of its own mailbox -- that read IS the grid barrier --
//     and reduces them with REDUX.  No line is ever polled by two SMs: with one shared mailbox the 2 x 148 polling
//     warps kept the L2 slices of those few lines so busy that the last record became visible only 3.6-4.9 us after
//     its store (0.7-0.9 us with 12 CTAs), measured with the global timer.
//     Polling is paced: the first poll goes out ~500 cycles after the alpha hand-over (~1000 after the publish), later
//     polls re-read only the records still missing.  Polling from the publish on, or with two polls in flight, gets
//     in the way of the records still travelling (measured: -4 % each).
//   level 2 (row-sharded box, world > 1): every CTA forwards its GPU's winner to its counterpart -- the CTA with the
//     same index -- on every GPU over NVLink (CUDA-IPC peer mappings): one 16-byte store per destination, one store
//     instruction (lane = rank), into that CTA's private level-2 mailbox; the CTA that owns the GPU-level winner
//     pushes, TOGETHER WITH IT, alpha, norm, K(x,x) and the row x itself into the box slot [rank][type] of every GPU
//     (candidates travel with their rows); the type warps poll the `world` records of their private
//     mailbox and then read the winning row from the local box (re-reading only the units that have not arrived).
//     Row units are {8 payload bytes, epoch, ~epoch}: self-validating like the records.  Each GPU holds ONLY ITS OWN
//     rows of X; nothing but these pushes crosses the link.
// A polled 16-byte unit is {payload, payload, epoch, ~epoch} (records: {g, g, idx, epoch<<3 | flags}) and the reader
// consumes ALL FOUR words: ptxas narrows a vector load whose components are partly dead into separate scalar loads
// (observed: ld.relaxed.sys.v4 with .y/.w unused became two LDG.32), and then payload and tag are no longer read by
// one memory transaction -- a stale payload can pass the tag check (found by comparing every leader's view per epoch).
// The same tags guard the alpha hand-over through SHARED memory, which is not cleared between launches: the slots are
// zeroed at kernel start (a unit parked by an earlier solve at the same epoch number once passed for the current one).
// The owner's row loads (world > 1) / L2 prefetch (world == 1) are issued before the level-1 poll, so the row is on
// hand (or L2-resident) by the time the winners are known.  G / flags / alpha of a row are only ever touched by the
// CTA (= SM) that owns the row; when every consumer warp owns at most one tile they stay in registers across sweeps.
//
// With the Q-column cache (CACHE) a sweep whose two columns are both resident ("warm") reads nothing but those columns
// and the gradient: it is pure memory latency, so it issues the loads of WARM_U tiles back to back before using the
// first, the candidate update is branch-free so that those tiles interleave, and the type warps prefetch the winners'
// cached columns into L2 as soon as the winners are known.
#pragma once
#include "b2svm_device.cuh"

namespace b2svm {

constexpr int NW = 11;                     // consumer warps per CTA (+1 producer warp = 384 threads, 168 regs)
constexpr int NTHREADS = (NW + 1) * 32;    // + one producer warp
constexpr int MAX_STAGES = 64;
constexpr int NCAND = 4;                   // up/low (C variant) or up+/low+/up-/low- (nu variant)
constexpr int TL_ITERS = 64, TL_SLOTS = 16; // debug timeline
constexpr int DBG_EPOCHS = 2048;
constexpr int DBG_W = 16;                  // debug record per (epoch, CTA): leader inputs, kernel values, consumer-side view
constexpr int TL_WARP_SLOTS = 4;           // + per-warp {start, end, tiles, wait} of one sweep
constexpr int TL_CTA_BASE = TL_ITERS * TL_SLOTS + 16 * TL_WARP_SLOTS;   // + per-CTA {sweep cycles, tiles} of one sweep
constexpr int TL_TOTAL = TL_CTA_BASE + 2 * 160;

struct SolveResult {
  long long n_iter;
  long long last_i, last_j;
  long long cold_sweeps;
  long long warm_sweeps, cols_stored, cache_used, cache_hand, cols_evicted;
  double di, dj;
  int converged;
  int status;   // 0 ok, 1 watchdog
  int pad[2];
};

struct ProblemDev {
  const void* X;        // [l][CH*16 bytes]
  const void* norms;    // T[l] squared row norms
  const void* kdiag;    // T[l] K(x_r, x_r): constant over the solve, so it is not recomputed on the select -> step chain
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
  uint4* mbox;          // level-1 private mailboxes [ncta targets][2][NCAND][nc8] records, then the alpha units in the same
                        // layout (nc8 = ncta rounded up to 8; zeroed before every launch)
  uint4* box;           // exported block of this GPU (world > 1; peers write into it over NVLink):
                        // [ncta targets][2][NCAND][8 ranks] private level-2 records, then [2][8 ranks][NCAND][BOX_UNITS] bulk slots
  uint4* const* peer_box;    // [world] every rank's exported block (peer mapping; own entry = box)
  int* abort_flag;
  SolveResult* result;
  long long max_iter;
  long long forced_i, forced_j;   // >= 0: skip the initial selection and take this pair first
  long long* timeline;            // optional [TL_ITERS][TL_SLOTS] clock64 stamps of CTA 0's leader lane
  // ---- device-resident Q-column cache (replaces the LRU of solver.py:207,227-229) -------------------
  void* cache;                    // T [cache_slots][l]: K(x_t, x_r) for this GPU's rows t, one slot per cached row r
  int* slot_of;                   // [l_global] slot of global row r, -1 = not cached (replicated per GPU)
  int* row_of_slot;               // [cache_slots] the row a slot holds (victim bookkeeping; replicated per GPU)
  long long cache_slots;          // 0 = cache off
  long long cache_used;           // slots already filled by earlier solves on this handle
  long long cache_hand;           // FIFO victim pointer carried over from earlier solves
  // ---- row-sharded multi-GPU (world > 1): this GPU holds and sweeps global rows [row_offset, row_offset + l) ONLY
  // (X, norms, alpha, G, flags and cached columns are all indexed by local row)
  int world, rank;
  long long row_offset, l_global, rows_per_rank;   // every rank's slice starts at rank * rows_per_rank
  double* dbg_steps;              // optional [DBG_EPOCHS][ncta][DBG_W]: si, sj, gi, gj, ai, aj, Kii, Kjj, Kij, ci, cj, row checksum | consumer view
  int debug_flags;                // B2SVM_DEBUG_FLAGS: 1 = static tile assignment
  // pacing of the exchange polls, in SM cycles (B2SVM_POLL_DELAY / _GAP / _DELAY2 / _GAP2): wait before the first poll of
  // a level and between two polls.  Polling competes with the very records it waits for.
  int poll_delay, poll_gap, poll_delay2, poll_gap2;
};
// level-2 box slot: {record, alpha unit, norm unit, K(x,x) unit, 2 half-chunk units per 16-byte chunk of the widest row (CH = 128)}
constexpr int BOX_UNITS = 4 + 2 * 128;
constexpr int BOX_RANKS = 8;
constexpr int MAX_CTAS = 160;              // level-1 poll: 5 records per lane
// sizes of the exchange buffers for a launch with `ncta` CTAs per problem
__host__ __device__ inline size_t l1_units(int ncta) { return (size_t)ncta * 2 * NCAND * ((ncta + 7) & ~7); }   // records; behind them one 128-byte outbox line per CTA for its alpha units
__host__ __device__ inline size_t l2_record_units(int ncta) { return (size_t)ncta * 2 * NCAND * BOX_RANKS; }
constexpr size_t BOX_BULK_UNITS = (size_t)2 * BOX_RANKS * NCAND * BOX_UNITS;
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

// a 16-byte exchange unit is valid for `epoch` when both tags match; using z AND w keeps every component of the
// vector load alive, so it stays one 128-bit memory transaction
__device__ __forceinline__ bool unit_ok(const uint4& u, uint32_t epoch) { return u.z == epoch && u.w == ~epoch; }
__device__ __forceinline__ uint4 make_unit(uint32_t lo, uint32_t hi, uint32_t epoch) { return make_uint4(lo, hi, epoch, ~epoch); }

// A candidate as it travels through the exchange: an order-preserving 64-bit key of the gradient (bit-flipped for
// the arg-min types, so every type is an arg-max over unsigned integers) and nik = ~((idx << 3) | flags), so that
// the lexicographic maximum of (key, nik) is the reference's choice: best value, then the LOWER index
// (np.argmax / np.argmin return the first extremum, solver.py:68-69).  nik == 0 marks an empty set.  Integer
// compares keep the polling loops branch-free (double compares cost a divergent branch per record).
struct KCand {
  uint32_t kh, kl, nik;
};
__device__ __forceinline__ KCand kc_empty() { return KCand{0u, 0u, 0u}; }
__device__ __forceinline__ KCand kc_make(bool is_max, const Cand& c) {
  if (c.idx < 0) return kc_empty();
  unsigned long long key = dkey(c.g);
  if (!is_max) key = ~key;
  return KCand{(uint32_t)(key >> 32), (uint32_t)key, ~(((uint32_t)c.idx << 3) | (uint32_t)(c.flags & 7))};
}
__device__ __forceinline__ bool kc_better(const KCand& a, const KCand& b) {   // a > b, branch-free
  return (a.kh > b.kh) | ((a.kh == b.kh) & ((a.kl > b.kl) | ((a.kl == b.kl) & (a.nik > b.nik))));
}
__device__ __forceinline__ void kc_take(bool take, const KCand& a, KCand& b) {
  b.kh = take ? a.kh : b.kh;
  b.kl = take ? a.kl : b.kl;
  b.nik = take ? a.nik : b.nik;
}
// warp-wide maximum (three REDUX.MAX); every lane receives the winner.  Full-warp collectives only.
__device__ __forceinline__ KCand kc_redux(const KCand& c) {
  KCand r;
  r.kh = __reduce_max_sync(0xffffffffu, c.kh);
  r.kl = __reduce_max_sync(0xffffffffu, c.kh == r.kh ? c.kl : 0u);
  r.nik = __reduce_max_sync(0xffffffffu, (c.kh == r.kh && c.kl == r.kl) ? c.nik : 0u);
  return r;
}
__device__ __forceinline__ int kc_idx(const KCand& c) { return c.nik ? (int)((~c.nik) >> 3) : -1; }
__device__ __forceinline__ int kc_flags(const KCand& c) { return (int)((~c.nik) & 7u); }
__device__ __forceinline__ double kc_g(bool is_max, const KCand& c) {
  unsigned long long key = ((unsigned long long)c.kh << 32) | c.kl;
  if (!is_max) key = ~key;
  return c.nik ? dkey_inv(key) : 0.0;
}

// State only the leader warp needs between iterations.  It lives in shared memory and is loaded into
// registers for the duration of the leader block only, so it costs the sweep loop no registers.
struct LeaderState {
  double di, dj;          // last step
  int ri, rj, si, sj;     // ... its rows / cache slots (same deferral)
  int n_cached;           // cache slots handed out so far (identical in every leader)
  int hand;               // next FIFO victim once every slot is taken
  int ei, ej;             // rows evicted by the last step (their slot-table entries are cleared one sweep late)
  int converged;
  long long warm_cnt, stored_cnt, evict_cnt;
};

// What a type warp found for its candidate type after the exchange (read by the leader = warp 0).
template <typename T>
struct Selected {
  double g, alpha;
  int idx, flags;
  int raw_slot, pad;    // slot_of[] entry of the row as read from global memory (-1 = not cached)
  T norm, kself;        // |x|^2 and K(x, x)
};

template <typename T>
struct Shared {
  double ci, cj;        // y_i * delta_i, y_j * delta_j
  double ai, aj;        // new alpha_i, alpha_j
  T ni, nj;             // |x_i|^2, |x_j|^2
  int i, j, fi, fj;
  int ti, tj;           // candidate types the pair came from: x_i = xs[ti], x_j = xs[tj]
  int slot_i, slot_j;   // cache slots of rows i, j (-1 = none); store_* = fill the slot during this sweep
  int store_i, store_j, warm;
  int p_ri, p_rj, p_si, p_sj;   // previous pair's rows / slots: slot_of[] is updated one sweep late, like alpha
  int p_ei, p_ej;               // ... and the rows that step evicted
  int stop;
  volatile int producer_stop;
  volatile long long consumed;
  LeaderState leader;
  int ticket;           // next dynamically assigned tile of the current sweep
  uint4* peer[BOX_RANKS];   // level-2 box of every rank (copied once from global memory)
  volatile int status;  // a warp's watchdog tripped
  KCand wc[NW][NCAND];  // per-warp candidates of the sweep
  uint4 wa[NW][NCAND];  // ... their alpha, {lo, hi, epoch, 0}: written when the load lands, after the CTA join
  Selected<T> sel[NCAND];   // the box-wide winner of every type
};

// `pa` = address of the variable's alpha.  Whenever a lane's running best changes, the line holding that alpha is
// pulled towards L2 (alpha is touched by two rows per iteration only, so the candidate's alpha is otherwise a DRAM
// miss sitting on the critical path between the sweep and the exchange: ~1600 cycles; a lane's best changes ~ln(rows
// per lane) times per sweep, so this costs a handful of prefetches).
__device__ __forceinline__ void touch_alpha(const double* pa) {
  if (pa) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(pa));   // alpha (8 B/row) is small: let it live in L2
}
__device__ __forceinline__ double ld_keep_f64(const double* p, uint64_t policy) {   // load that asks L2 to keep the line
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy) : "memory");
  return v;
}
__device__ __forceinline__ void consider(bool nu, int f, double g, int v, Cand (&b)[NCAND], const double* pa) {
  // Branch-free (selects): almost every row leaves the running bests alone, and a branch per test would end the basic
  // block there -- the rows of a batch of tiles could then not be interleaved by the scheduler.
  const bool second = nu && !(f & F_YPOS);
  {
    const double bg = second ? b[2].g : b[0].g;
    const int bi = second ? b[2].idx : b[0].idx;
    const bool take = in_up(f) && better_max(g, v, bg, bi);
    const bool t0 = take && !second, t2 = take && second;
    b[0].g = t0 ? g : b[0].g; b[0].idx = t0 ? v : b[0].idx; b[0].flags = t0 ? f : b[0].flags;
    b[2].g = t2 ? g : b[2].g; b[2].idx = t2 ? v : b[2].idx; b[2].flags = t2 ? f : b[2].flags;
    if (take) touch_alpha(pa);
  }
  {
    const double bg = second ? b[3].g : b[1].g;
    const int bi = second ? b[3].idx : b[1].idx;
    const bool take = in_low(f) && better_min(g, v, bg, bi);
    const bool t1 = take && !second, t3 = take && second;
    b[1].g = t1 ? g : b[1].g; b[1].idx = t1 ? v : b[1].idx; b[1].flags = t1 ? f : b[1].flags;
    b[3].g = t3 ? g : b[3].g; b[3].idx = t3 ? v : b[3].idx; b[3].flags = t3 ? f : b[3].flags;
    if (take) touch_alpha(pa);
  }
}

__device__ __forceinline__ void spin_cycles(int n) {
  if (n > 0) {
    const long long t0 = clock64();
    while (clock64() - t0 < n) {}
  }
}

// compiler fence on one register value: nothing computed from v may be scheduled above this point
__device__ __forceinline__ void opaque(float& v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ void opaque(double& v) { asm volatile("" : "+d"(v)); }
__device__ __forceinline__ void opaque(int& v) { asm volatile("" : "+r"(v)); }

constexpr int WARM_U = 4;   // tiles whose loads a warp keeps in flight in a warm (cached-column) sweep

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
  V* xs = reinterpret_cast<V*>(const_cast<unsigned*>(gen) + MAX_STAGES);      // [NCAND][CH]: the winning row of every candidate type
  Shared<T>* sh = reinterpret_cast<Shared<T>*>(xs + NCAND * CH);

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
    for (int r = 0; r < BOX_RANKS; ++r) sh->peer[r] = (P.world > 1 && r < P.world) ? const_cast<uint4*>(P.peer_box[r]) : nullptr;
    LeaderState L0;
    L0.di = L0.dj = 0.0;
    L0.ri = L0.rj = L0.si = L0.sj = -1;
    L0.n_cached = (int)P.cache_used;
    L0.hand = (int)P.cache_hand;
    L0.ei = L0.ej = -1;
    L0.converged = 0;
    L0.warm_cnt = L0.stored_cnt = L0.evict_cnt = 0;
    sh->leader = L0;
  }
  // shared memory is not cleared between launches: an alpha unit parked by an earlier solve carries a valid-looking
  // tag of the same epoch number, so the hand-over slots start out empty
  for (int e = threadIdx.x; e < NW * NCAND; e += blockDim.x) (&sh->wa[0][0])[e] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();

  // ============================== producer warp ==============================================
  if (warp == NW) {
    if (NM > 0 && P.kind != K_PRECOMPUTED) {   // a precomputed kernel matrix is only ever read through the cache path
      const unsigned char* Xb = reinterpret_cast<const unsigned char*>(P.X);   // this GPU's rows
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
  const T* __restrict__ norms = reinterpret_cast<const T*>(P.norms);   // this GPU's rows
  const T* __restrict__ kdiag = reinterpret_cast<const T*>(P.kdiag);
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
  auto stamp_val = [&](int slot, long long v) {
    if (P.timeline && cta == 0 && threadIdx.x == 0 && epoch >= 1 && epoch <= (uint32_t)TL_ITERS)
      P.timeline[(epoch - 1) * TL_SLOTS + slot] = v;
  };
#else
  auto stamp = [&](int) {};
  auto stamp_val = [&](int, long long) {};
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
  auto load_row_state = [&](int t, bool with_norm = true) {
    RowState<T> r;
    r.g0 = r.g1 = 0.0;
    r.f0 = r.f1 = 0;
    r.nt = 0;
    r.ki = r.kj = 0;
    const long long row = row0 + (long long)t * C::BR + rib;
    if (t < NT && lane_active && row < row1) {
      r.g0 = G[row];
      r.f0 = flags[row];
      if (mirror) { r.g1 = G[row + l]; r.f1 = flags[row + l]; }
      if (with_norm) r.nt = norms[row];
      if (CACHE && sweep_warm) {
        r.ki = cache_i[row];
        r.kj = cache_j[row];
      }
    }
    return r;
  };

  // ---- initial selection pass over G / flags only (no X) -------------------------------------
  reset_best();
  const bool forced = P.forced_i >= 0;
  if (!forced) {
    for (long long r = row0 + warp * 32 + lane; r < row1; r += NW * 32) {
      consider(nu, flags[r], G[r], (int)(r + goff), best, alpha + r);
      if (mirror) consider(nu, flags[r + l], G[r + l], (int)(r + goff + lg), best, alpha + r + l);
    }
  }
  RowState<T> pf = load_row_state(warp * SUB);   // first tile of the first sweep (macro tiles 0..NW-1 are static)

  // every consumer warp owns at most one tile: its gradient-side state never leaves the registers
  const bool one_tile = SUB == 1 && NM <= NW;
  // exchange buffers (see ProblemDev): this CTA's private level-1 / level-2 mailboxes, the bulk slots of the box
  const int nc8 = (ncta + 7) & ~7;
  const size_t l1_total = l1_units(ncta);
  const uint4* const box_base = P.box ? P.box + l2_record_units(ncta) : nullptr;

  for (;;) {
    // ---- (B) warp -> CTA candidate reduce, level-1 (GPU) and level-2 (box) exchange, step -------------------
    ++epoch;
    stamp(0);   // this warp's sweep done
    const uint32_t par = epoch & 1u;
    // warp-level winners as integer keys (three REDUX.MAX per type)
    KCand wk[NCAND];
#pragma unroll
    for (int k = 0; k < NCAND; ++k) {
      wk[k] = kc_empty();
      if (k < ntypes) wk[k] = kc_redux(kc_make((k & 1) == 0, best[k]));
    }
    __syncwarp();
    // one lane per type also fetches the candidate's alpha (a row of this warp: same-SM coherent).  The load is
    // only issued here: its value is parked in shared memory after the CTA join, so neither the join nor the
    // candidate record waits for it (it is needed at step time only).
    double cand_alpha = 0.0;
    if (lane < ntypes) {
      KCand m = wk[0];
#pragma unroll
      for (int k = 1; k < NCAND; ++k)
        if (lane == k) m = wk[k];
      if (m.nik) cand_alpha = ld_keep_f64(alpha + local_var(kc_idx(m)), l2_policy_evict_last());
      sh->wc[warp][lane] = m;
    }
    named_bar_sync(1, NW * 32);
    stamp(1);   // all warps of this CTA done
    auto park_alpha = [&]() {
      if (lane < ntypes)
        st_volatile_shared_v4(&sh->wa[warp][lane],
                              make_unit((uint32_t)__double2loint(cand_alpha), (uint32_t)__double2hiint(cand_alpha), epoch));
    };
    if (warp >= ntypes) park_alpha();

    if (warp < ntypes) {
      // =========================== type warp: everything about candidate type k = warp ======================
      const int k = warp;
      const bool is_max = (k & 1) == 0;
      const bool forced_now = forced && epoch == 1;   // single GPU only (checked on the host)
      Cand w;            // the winner of type k: first of this CTA, then of the GPU, then of the box
      w.g = 0.0; w.idx = -1; w.flags = 0;
      int wsrc = 0;      // where it came from: CTA (level 1), then rank (level 2)
      double w_alpha = 0.0;
      V row[QL];
      T nrm = 0, kself = 0;
#pragma unroll
      for (int q = 0; q < QL; ++q) row[q] = CK::zero();
      bool have_row = false, have_alpha = false;
      int status = 0;
      int raw_slot = -1;   // the winner's entry in the (one sweep late) column-cache slot table

      if (!forced_now) {
        // ---- CTA-level winner (full-warp collectives only: two half-warp masks in one instruction compile to a
        // MATCH.ANY / WARPSYNC.EXCLUSIVE emulation that costs ~1000 cycles on this path)
        KCand mine_w = kc_empty();
        if (lane < NW) mine_w = sh->wc[lane][k];
        const KCand m = kc_redux(mine_w);
        const unsigned holders = __ballot_sync(0xffffffffu, m.nik != 0u && mine_w.nik == m.nik);
        const int from_warp = holders ? __ffs(holders) - 1 : 0;
        const int m_idx = kc_idx(m);
        stamp(7);   // CTA winner known
        // ---- level 1: the tagged record goes into the private mailbox of every CTA of this GPU (all-to-all push)
        constexpr int MAXM = MAX_CTAS / 32;
        auto l1_slot = [&](int target) -> uint4* {   // where CTA `cta` writes for CTA `target`
          return P.mbox + (((size_t)target * 2 + par) * NCAND + k) * nc8 + cta;
        };
        const uint4* const l1_mine = P.mbox + (((size_t)cta * 2 + par) * NCAND + k) * nc8;   // what the others wrote for me
        {
          const uint4 rec = make_uint4(m.kl, m.kh, m.nik, epoch);
#pragma unroll
          for (int mm = 0; mm < MAXM; ++mm) {
            const int t = lane + 32 * mm;
            if (t < ncta && t != cta) st_relaxed_v4(l1_slot(t), rec);
          }
        }
        // ---- the poll itself starts after the alpha hand-over below (see there why not earlier); this CTA's own
        // record never leaves its registers
        uint4 h[MAXM];
#pragma unroll
        for (int mm = 0; mm < MAXM; ++mm) h[mm] = make_uint4(0u, 0u, 0u, 0u);
        auto issue_poll = [&]() {   // (only what is still missing: polling traffic competes with the records themselves)
#pragma unroll
          for (int mm = 0; mm < MAXM; ++mm) {
            const int c = lane + 32 * mm;
            if (c < ncta && c != cta && h[mm].w != epoch) h[mm] = ld_relaxed_v4(l1_mine + c);
          }
        };
        // the candidate's row, requested now and used after the poll: real loads when it may have to travel to the
        // other GPUs (world > 1), an L2 prefetch otherwise (the leaders' fetch of the winning row then hits)
        V myrow[QL];
        T mynorm = 0, mykself = 0;
#pragma unroll
        for (int q = 0; q < QL; ++q) myrow[q] = CK::zero();
        if (m_idx >= 0 && P.kind != K_PRECOMPUTED) {
          const long long lrow = ((mirror && m_idx >= lg) ? m_idx - lg : m_idx) - goff;
          const V* src = reinterpret_cast<const V*>(P.X) + lrow * CH;
          if (world > 1) {
#pragma unroll
            for (int q = 0; q < QL; ++q) {
              const int c = lane + 32 * q;
              if (c < CH) myrow[q] = src[c];
            }
            mynorm = norms[lrow];
            mykself = kdiag[lrow];
          } else if (lane * 128 < C::ROWB) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(src) + lane * 128));
          }
        }
        if (k == 0 && lane == 0) sh->ticket = NW;   // next sweep: tiles 0..NW-1 are static, the rest are drawn here
        stamp(2);   // published
#ifdef B2SVM_TIMELINE
        if (P.timeline && threadIdx.x == 0 && epoch == 41u && cta < 160)   // when did every CTA publish (global ns timer)?
          P.timeline[TL_CTA_BASE + 2 * cta] = (long long)globaltimer_ns();
#endif
        // ---- the CTA winner's alpha: forwarded as soon as the warp that owns it has parked it (its load may still be
        // in flight: the check is repeated between the poll passes below instead of being waited for here)
        park_alpha();
        __syncwarp();
        uint4 ua = make_uint4(0u, 0u, 0u, 0u);
        bool alpha_sent = false;
        auto try_alpha = [&]() {
          if (alpha_sent) return;
          ua = ld_volatile_shared_v4(&sh->wa[from_warp][k]);      // (same address in every lane: one broadcast read)
          if (!unit_ok(ua, epoch)) return;
          alpha_sent = true;
          // world == 1: ONE store into this CTA's outbox (its own 128-byte line behind the mailboxes); only the winner's
          // unit is ever read, once per CTA, after the poll -- pushing it all-to-all like the record doubled the store
          // traffic of the exchange for a unit 147 of 148 CTAs never look at.  (world > 1: it travels with the level-2
          // push instead.)  A snapshot, double-buffered by parity like the records: a fast CTA already storing the new
          // alpha in its next sweep cannot race a slow leader.
          if (world == 1 && ncta > 1 && lane == 0)
            st_relaxed_v4(P.mbox + l1_total + ((size_t)cta * 2 + par) * NCAND + k, make_unit(ua.x, ua.y, epoch));
        };
        try_alpha();

        // ---- poll the ncta level-1 records of this type (this read is the grid barrier)
        KCand gw;          // GPU-level winner
        {
          uint64_t t_start = 0;   // (the timer is only read once the poll has been spinning for a while)
          uint32_t rounds = 0;
          stamp(12);   // poll entry
          uint32_t passes = 0;
          // The first poll goes out late: right behind the publish it can only fail (nobody else is there yet), and the
          // loads of 148 pollers get in the way of the records still travelling (measured: polling from the publish on
          // costs 4 % of the iteration rate, two polls in flight another 4 %)
          if (ncta > 1) spin_cycles(P.poll_delay);
          issue_poll();
          for (;;) {
            bool all = true;
            KCand bk = kc_empty();
            int bsrc = 0;
#pragma unroll
            for (int mm = 0; mm < MAXM; ++mm) {
              const int c = lane + 32 * mm;
              const bool own = c == cta;
              KCand rc;
              rc.kl = own ? m.kl : h[mm].x;
              rc.kh = own ? m.kh : h[mm].y;
              rc.nik = own ? m.nik : h[mm].z;
              const bool valid = own | (h[mm].w == epoch);
              all = all & (valid | (c >= ncta));
              const bool take = (c < ncta) & valid & kc_better(rc, bk);
              kc_take(take, rc, bk);
              bsrc = take ? c : bsrc;
            }
            if (passes++ == 0) stamp(13);   // first pass back
            try_alpha();
            if (__all_sync(0xffffffffu, all)) {
              gw = kc_redux(bk);
              const unsigned hold2 = __ballot_sync(0xffffffffu, gw.nik != 0u && bk.nik == gw.nik);
              wsrc = __shfl_sync(0xffffffffu, bsrc, hold2 ? __ffs(hold2) - 1 : 0);
              break;
            }
            if ((++rounds & 63u) == 0) {
              int bad = *(volatile int*)P.abort_flag | sh->status;
              const uint64_t now = globaltimer_ns();
              if (t_start == 0) t_start = now;
              if (!bad && now - t_start > 6000000000ull) {
                *(volatile int*)P.abort_flag = 1;
                bad = 1;
              }
              if (__any_sync(0xffffffffu, bad)) { status = 1; gw = kc_empty(); break; }
            }
            spin_cycles(P.poll_gap);
            issue_poll();
          }
          while (!alpha_sent) try_alpha();
          stamp(15);   // poll loop left
          stamp_val(14, (long long)passes);
#ifdef B2SVM_TIMELINE
          if (P.timeline && threadIdx.x == 0 && epoch == 41u && cta < 160)   // ... and when did it see everybody's record?
            P.timeline[TL_CTA_BASE + 2 * cta + 1] = (long long)globaltimer_ns();
#endif
        }
        stamp(3);   // GPU-level winner known

        if (world > 1 && !status) {
          // ---- level 2.  Records: every CTA forwards its GPU's winner to its counterpart (the CTA with the same
          // index) on every GPU -- one store instruction, lane = destination rank -- so the records of a GPU travel
          // from ncta SMs in parallel instead of queueing behind each other in one SM's NVLink store path (measured:
          // ~85 cycles per peer store instruction; the owner alone needed 64 of them for 8 GPUs).
          if (lane < world)
            st_relaxed_sys_v4(sh->peer[lane] + (((size_t)cta * 2 + par) * NCAND + k) * BOX_RANKS + P.rank,
                              make_uint4(gw.kl, gw.kh, gw.nik, epoch));
          // Bulk: the owner of the GPU-level winner pushes alpha, norm, K(x,x) and the row into the box of every GPU
          const int owner = gw.nik ? wsrc : 0;
          if (cta == owner && gw.nik) {
            // float: one unit {norm, K(x,x)}; double: one unit each
            uint4 un, uk;
            if (sizeof(T) == 8) {
              un = make_unit((uint32_t)__double2loint((double)mynorm), (uint32_t)__double2hiint((double)mynorm), epoch);
              uk = make_unit((uint32_t)__double2loint((double)mykself), (uint32_t)__double2hiint((double)mykself), epoch);
            } else {
              un = make_unit(__float_as_uint((float)mynorm), __float_as_uint((float)mykself), epoch);
              uk = un;
            }
            const size_t slot_off = l2_record_units(ncta) + (((size_t)par * BOX_RANKS + P.rank) * NCAND + k) * BOX_UNITS;
            // the three small units of all destinations in one instruction: lane = 4 * destination + unit
            {
              const int r = lane >> 2, which = lane & 3;
              if (r < world && which >= 1 && (which < 3 || sizeof(T) == 8)) {
                const uint4 u = which == 1 ? make_unit(ua.x, ua.y, epoch) : (which == 2 ? un : uk);
                st_relaxed_sys_v4(sh->peer[r] + slot_off + which, u);
              }
            }
            for (int rr = 0; rr < world; ++rr) {
              const int r = (P.rank + 1 + rr) % world;   // peers first, the own box last
              uint4* slot = sh->peer[r] + slot_off;
#pragma unroll
              for (int q = 0; q < QL; ++q) {
                const int c = lane + 32 * q;
                if (c < CH) {
                  const uint4 bits = *reinterpret_cast<const uint4*>(&myrow[q]);
                  st_relaxed_sys_v4(slot + 4 + 2 * c, make_unit(bits.x, bits.y, epoch));
                  st_relaxed_sys_v4(slot + 4 + 2 * c + 1, make_unit(bits.z, bits.w, epoch));
                }
              }
            }
          }
          // ---- poll the `world` level-2 records of this type in this CTA's private mailbox (the box barrier)
          {
            const uint4* rbase = P.box + (((size_t)cta * 2 + par) * NCAND + k) * BOX_RANKS;
            uint64_t t_start = 0;
            uint32_t rounds = 0;
            spin_cycles(P.poll_delay2);
            for (;;) {
              KCand bk = kc_empty();
              bool ok = true;
              if (rounds) spin_cycles(P.poll_gap2);
              if (lane < world) {
                const uint4 hh = ld_relaxed_sys_v4(rbase + lane);
                ok = hh.w == epoch;
                bk.kl = hh.x; bk.kh = hh.y; bk.nik = ok ? hh.z : 0u;
              }
              if (__all_sync(0xffffffffu, ok)) {
                gw = kc_redux(bk);
                const unsigned hold2 = __ballot_sync(0xffffffffu, gw.nik != 0u && bk.nik == gw.nik);
                wsrc = hold2 ? __ffs(hold2) - 1 : 0;     // lane = rank
                break;
              }
              if ((++rounds & 63u) == 0) {
                int bad = *(volatile int*)P.abort_flag | sh->status;
                const uint64_t now = globaltimer_ns();
                if (t_start == 0) t_start = now;
                if (!bad && now - t_start > 6000000000ull) {
                  *(volatile int*)P.abort_flag = 1;
                  bad = 1;
                }
                if (__any_sync(0xffffffffu, bad)) { status = 1; gw = kc_empty(); break; }
              }
            }
          }
        }
        w.idx = kc_idx(gw);
        w.flags = kc_flags(gw);
        w.g = kc_g(is_max, gw);
        stamp(4);   // box-wide winner known

        // ---- the winner's alpha, norm and row
        if (w.idx >= 0 && !status) {
          if (CACHE && P.cache_slots > 0)   // (requested here so that it travels together with the row below)
            raw_slot = __ldcg(P.slot_of + ((mirror && w.idx >= lg) ? w.idx - lg : w.idx));
          if (world > 1) {
            // from the box slot of the rank it came from; every unit is re-read until it carries this epoch
            const uint4* slot = box_base + (((size_t)par * BOX_RANKS + wsrc) * NCAND + k) * BOX_UNITS;
            uint32_t spins = 0;
            const uint4 none = make_uint4(0u, 0u, 0u, 0u);
            uint4 a4 = none, n4 = none, k4 = none;
            uint4 lo[QL], hi[QL];
#pragma unroll
            for (int q = 0; q < QL; ++q) lo[q] = hi[q] = none;
            for (;;) {
              // (only what has not arrived yet is read again: re-reading all 2 CH + 3 units from every CTA of the GPU
              // would compete with the very stores it waits for)
              bool ok = true;
              if (!unit_ok(a4, epoch)) a4 = ld_relaxed_sys_v4(slot + 1);
              if (!unit_ok(n4, epoch)) n4 = ld_relaxed_sys_v4(slot + 2);
              if (sizeof(T) == 8) { if (!unit_ok(k4, epoch)) k4 = ld_relaxed_sys_v4(slot + 3); }
#pragma unroll
              for (int q = 0; q < QL; ++q) {
                const int c = lane + 32 * q;
                if (c < CH && P.kind != K_PRECOMPUTED) {
                  if (!unit_ok(lo[q], epoch)) lo[q] = ld_relaxed_sys_v4(slot + 4 + 2 * c);
                  if (!unit_ok(hi[q], epoch)) hi[q] = ld_relaxed_sys_v4(slot + 4 + 2 * c + 1);
                }
              }
              if (sizeof(T) != 8) k4 = n4;
              ok = unit_ok(a4, epoch) && unit_ok(n4, epoch) && unit_ok(k4, epoch);
              w_alpha = __hiloint2double((int)a4.y, (int)a4.x);
              if (sizeof(T) == 8) { nrm = (T)__hiloint2double((int)n4.y, (int)n4.x); kself = (T)__hiloint2double((int)k4.y, (int)k4.x); }
              else { nrm = (T)__uint_as_float(n4.x); kself = (T)__uint_as_float(n4.y); }
#pragma unroll
              for (int q = 0; q < QL; ++q) {
                const int c = lane + 32 * q;
                if (c < CH && P.kind != K_PRECOMPUTED) {
                  ok = ok && unit_ok(lo[q], epoch) && unit_ok(hi[q], epoch);
                  const uint4 bits = make_uint4(lo[q].x, lo[q].y, hi[q].x, hi[q].y);
                  row[q] = *reinterpret_cast<const V*>(&bits);
                }
              }
              if (__all_sync(0xffffffffu, ok)) break;
              if ((++spins & 4095u) == 0) {
                const int bad = *(volatile int*)P.abort_flag;
                if (__any_sync(0xffffffffu, bad)) { status = 1; break; }
              }
            }
            have_row = have_alpha = true;
          } else {
            // single GPU: the alpha unit its CTA left in its outbox, the row straight from X
            const uint4* unit = P.mbox + l1_total + ((size_t)wsrc * 2 + par) * NCAND + k;   // the winner's outbox
            uint4 u = wsrc == cta ? make_unit(0u, 0u, epoch) : ld_relaxed_v4(unit);
            if (P.kind != K_PRECOMPUTED) {
              const long long lrow = (mirror && w.idx >= lg) ? w.idx - lg : w.idx;
              const V* src = reinterpret_cast<const V*>(P.X) + lrow * CH;
#pragma unroll
              for (int q = 0; q < QL; ++q) {
                const int c = lane + 32 * q;
                if (c < CH) row[q] = src[c];
              }
              nrm = norms[lrow];
              have_row = true;
            }
            kself = kdiag[(mirror && w.idx >= lg) ? w.idx - lg : w.idx];
            uint32_t spins = 0;
            while (!unit_ok(u, epoch)) {   // (the unit follows its record by one shared-memory hand-over)
              u = ld_relaxed_v4(unit);
              if ((++spins & 4095u) == 0 && *(volatile int*)P.abort_flag) { status = 1; break; }
            }
            if (wsrc == cta) u = ua;      // my own candidate won: its alpha never left this CTA
            w_alpha = __hiloint2double((int)u.y, (int)u.x);
            have_alpha = true;
          }
        }
      } else if (k < 2) {
        // forced pair of the step-wise API (b2svm_update): no selection, no exchange
        w.idx = (int)(k == 0 ? P.forced_i : P.forced_j);
        w.g = __ldcg(G + w.idx);
        w.flags = __ldcg(flags + w.idx);
        w_alpha = __ldcg(alpha + w.idx);
        have_alpha = true;
        if (CACHE && P.cache_slots > 0) raw_slot = __ldcg(P.slot_of + ((mirror && w.idx >= lg) ? w.idx - lg : w.idx));
        if (k == 0 && lane == 0) sh->ticket = NW;
        if (P.kind != K_PRECOMPUTED) {
          const long long lrow = (mirror && w.idx >= lg) ? w.idx - lg : w.idx;
          const V* src = reinterpret_cast<const V*>(P.X) + lrow * CH;
#pragma unroll
          for (int q = 0; q < QL; ++q) {
            const int c = lane + 32 * q;
            if (c < CH) row[q] = src[c];
          }
          nrm = norms[lrow];
          have_row = true;
        }
        kself = kdiag[(mirror && w.idx >= lg) ? w.idx - lg : w.idx];
      }
      (void)have_alpha;
      stamp(8);   // winner's row / alpha on hand
      // ---- hand the winner of this type to the leader: row into shared memory, K(x, x), slot-table entry
      {
        Selected<T> sv;
        sv.g = w.g; sv.alpha = w_alpha; sv.idx = w.idx; sv.flags = w.flags;
        sv.raw_slot = raw_slot; sv.pad = 0;
        sv.norm = nrm; sv.kself = kself;
        if (w.idx >= 0) {
          if (CACHE && raw_slot >= 0 && !one_tile) {
            // a cached column of a winner is most likely read by the coming sweep (a warm sweep reads nothing else from
            // HBM): start pulling this CTA's rows of it into L2 now, a type barrier, a kernel value and a step ahead
            const char* seg = reinterpret_cast<const char*>(reinterpret_cast<const T*>(P.cache) + (size_t)raw_slot * (size_t)l + row0);
            const int bytes = (int)(row1 - row0) * (int)sizeof(T);
            for (int o = lane * 128, n = 0; o < bytes && n < 8; o += 32 * 128, ++n)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(seg + o));
          }
          if (have_row) {
#pragma unroll
            for (int q = 0; q < QL; ++q) {
              const int c = lane + 32 * q;
              if (c < CH) xs[k * CH + c] = row[q];
            }
          }
        }
        if (lane == 0) {
          sh->sel[k] = sv;
          if (status) sh->status = 1;
        }
      }
      named_bar_sync(3, ntypes * 32);   // the type warps meet
      stamp(9);

      if (warp == 0) {
        // =========================== leader: working_set_select + the analytic step =========================
        LeaderState L = sh->leader;
        const int st_bad = sh->status;
        const long long n_iter = (long long)epoch - 1;
        // K(a, b) of two selected types, from the rows in shared memory (or the precomputed kernel matrix)
        auto cross_k = [&](int ta, int tb) -> T {
          const Selected<T>& A = sh->sel[ta];
          const Selected<T>& B = sh->sel[tb];
          if (CACHE && P.kind == K_PRECOMPUTED) {
            const long long ra = (mirror && A.idx >= lg) ? A.idx - lg : A.idx, rb = (mirror && B.idx >= lg) ? B.idx - lg : B.idx;
            return __ldg(reinterpret_cast<const T*>(P.cache) + ra * l + rb);
          }
          T dab = 0;
#pragma unroll
          for (int q = 0; q < QL; ++q) {
            const int c = lane + 32 * q;
            if (c < CH) dab = CK::dot(xs[ta * CH + c], xs[tb * CH + c], dab);
          }
          dab = warp_sum(dab);
          return kernel_value(kp, dab, A.norm, B.norm);
        };
        // ---- decide the next pair (working_set_select) ------------------------------------------
        int ti = 0, tj = 1;   // candidate types the pair comes from
        bool stop_sel = false;
        T Kij = 0;
        bool have_k = false;
        if (forced_now) {
          // pair given by the caller
        } else if (!nu) {
          // solver.py:66-75
          const int si0 = sh->sel[0].idx, sj0 = sh->sel[1].idx;
          stop_sel = si0 < 0 || sj0 < 0 || (sh->sel[0].g - sh->sel[1].g < P.tol);
        } else {
          // solver.py:300-339: no tol test, gain comparison between the class-wise pairs
          const bool pos_ok = sh->sel[0].idx >= 0 && sh->sel[1].idx >= 0;
          const bool neg_ok = sh->sel[2].idx >= 0 && sh->sel[3].idx >= 0;
          int pick = -1;   // 0 = '+' pair, 1 = '-' pair
          if (!pos_ok && !neg_ok) stop_sel = true;
          else if (!pos_ok) pick = 1;
          else if (!neg_ok) pick = 0;
          else {
            const T Kp = cross_k(0, 1), Kn = cross_k(2, 3);
            double quad_p = ((double)sh->sel[0].kself + (double)sh->sel[1].kself) - 2.0 * (double)Kp;
            if (quad_p <= 0) quad_p = 1e-12;
            const double dp = sh->sel[0].g - sh->sel[1].g;
            const double gain_p = -(dp * dp) / quad_p;
            double quad_n = ((double)sh->sel[2].kself + (double)sh->sel[3].kself) - 2.0 * (double)Kn;
            if (quad_n <= 0) quad_n = 1e-12;
            const double dn = sh->sel[2].g - sh->sel[3].g;
            const double gain_n = -(dn * dn) / quad_n;
            if (gain_p < gain_n) { pick = 0; Kij = Kp; }
            else { pick = 1; Kij = Kn; }
            have_k = true;
          }
          if (pick >= 0) { ti = pick == 0 ? 0 : 2; tj = ti + 1; }
        }
        const Selected<T> SI = sh->sel[ti], SJ = sh->sel[tj];
        const int si = stop_sel ? -1 : SI.idx, sj = stop_sel ? -1 : SJ.idx;
#ifdef B2SVM_TIMELINE
        if (P.dbg_steps && lane == 0 && epoch <= (uint32_t)DBG_EPOCHS) {
          double* o = P.dbg_steps + ((size_t)(epoch - 1) * ncta + cta) * DBG_W;
          o[0] = si; o[1] = sj; o[2] = SI.g; o[3] = SJ.g; o[4] = SI.alpha; o[5] = SJ.alpha;
        }
#endif
        if (stop_sel) L.converged = 1;
        const bool exit_now = st_bad || stop_sel || n_iter >= P.max_iter;
        if (exit_now) {
          if (lane == 0) {
            sh->stop = 1;
            if (cta == 0) {
              SolveResult r;
              r.n_iter = n_iter;
              r.last_i = si;
              r.last_j = sj;
              r.cold_sweeps = n_iter - L.warm_cnt;
              r.warm_sweeps = L.warm_cnt;
              r.cols_stored = L.stored_cnt;
              r.cache_used = L.n_cached;
              r.cache_hand = L.hand;
              r.cols_evicted = L.evict_cnt;
              if (CACHE && P.cache_slots > 0 && L.ri >= 0) {   // pending slot-table entries of the last step
                if (L.ei >= 0) P.slot_of[L.ei] = -1;
                if (L.ej >= 0) P.slot_of[L.ej] = -1;
                if (L.si >= 0) { P.slot_of[L.ri] = L.si; if (P.row_of_slot) P.row_of_slot[L.si] = L.ri; }
                if (L.sj >= 0) { P.slot_of[L.rj] = L.sj; if (P.row_of_slot) P.row_of_slot[L.sj] = L.rj; }
              }
              r.di = L.di;
              r.dj = L.dj;
              r.converged = L.converged;
              r.status = st_bad;
              r.pad[0] = r.pad[1] = 0;
              *P.result = r;
            }
          }
        } else {
          // ---- Solver.update / NuSolver.update: the analytic step (solver.py:82-145, 342-373) ----
          if (!have_k) Kij = cross_k(ti, tj);
          stamp(10);
          const int fi = SI.flags, fj = SJ.flags;
          const double yi = (fi & F_YPOS) ? 1.0 : -1.0;
          const double yj = (fj & F_YPOS) ? 1.0 : -1.0;
          const StepOut so = smo_step(nu, (double)SI.kself, (double)SJ.kself, (double)Kij, yi, yj, SI.g, SJ.g, SI.alpha, SJ.alpha, Cbox);
          stamp(11);
          // ---- Q-column cache: which of the two columns are already resident, which get stored now --------
          const int gri = (int)((mirror && si >= lg) ? si - lg : si), grj = (int)((mirror && sj >= lg) ? sj - lg : sj);
          int slot_i = -1, slot_j = -1, store_i = 0, store_j = 0, ev_i = -1, ev_j = -1;
          if (CACHE && P.cache_slots > 0) {
            // a row evicted by the previous step still shows its old slot in the (one sweep late) table
            const int raw_i = (gri == L.ei || gri == L.ej) ? -1 : SI.raw_slot;
            const int raw_j = (grj == L.ei || grj == L.ej) ? -1 : SJ.raw_slot;
            slot_i = gri == L.ri ? L.si : (gri == L.rj ? L.sj : raw_i);
            slot_j = grj == L.ri ? L.si : (grj == L.rj ? L.sj : raw_j);
            // a free slot while there are any; afterwards the oldest column that neither this nor the previous pair
            // uses is replaced (FIFO: the reference's LRU, solver.py:207,227-229, evicts too -- a cache that only ever
            // filled would stop adapting once full).  Every leader of every GPU takes the same decision.
            auto take_slot = [&](int avoid, int& evicted_row) -> int {
              if (L.n_cached < P.cache_slots) return L.n_cached++;
              if (P.cache_slots < 256 || P.kind == K_PRECOMPUTED) return -1;   // (a handful of slots would only thrash: they stay fill-only)
              for (int t = 0; t < 6; ++t) {
                const int v = L.hand;
                L.hand = (v + 1 == (int)P.cache_slots) ? 0 : v + 1;
                if (v == avoid || v == L.si || v == L.sj) continue;
                evicted_row = __ldcg(P.row_of_slot + v);
                ++L.evict_cnt;
                return v;
              }
              return -1;
            };
            if (slot_i < 0) { slot_i = take_slot(slot_j, ev_i); store_i = slot_i >= 0 ? 1 : 0; }
            if (grj == gri) { slot_j = slot_i; }
            else if (slot_j < 0) { slot_j = take_slot(slot_i, ev_j); store_j = slot_j >= 0 ? 1 : 0; }
          }
          const int warm = (CACHE && P.cache_slots > 0 && slot_i >= 0 && slot_j >= 0 && !store_i && !store_j) ? 1 : 0;
          if (lane == 0) {
            sh->slot_i = slot_i; sh->slot_j = slot_j; sh->store_i = store_i; sh->store_j = store_j; sh->warm = warm;
            sh->p_ri = L.ri; sh->p_rj = L.rj; sh->p_si = L.si; sh->p_sj = L.sj;
            sh->p_ei = L.ei; sh->p_ej = L.ej;
          }
          L.ri = gri; L.rj = grj; L.si = slot_i; L.sj = slot_j;
          L.ei = ev_i; L.ej = ev_j;
          if (warm) ++L.warm_cnt;
          L.stored_cnt += store_i + store_j;
          L.di = so.di; L.dj = so.dj;
#ifdef B2SVM_TIMELINE
          if (P.dbg_steps && epoch <= (uint32_t)DBG_EPOCHS) {
            // checksum of both rows as the leader holds them (bit pattern sum over all chunks)
            unsigned long long cs = 0;
#pragma unroll
            for (int q = 0; q < QL; ++q) {
              const int c = lane + 32 * q;
              if (c < CH) {
                const uint4 a = *reinterpret_cast<const uint4*>(&xs[ti * CH + c]);
                const uint4 b = *reinterpret_cast<const uint4*>(&xs[tj * CH + c]);
                cs += (unsigned long long)a.x + 3ull * a.y + 5ull * a.z + 7ull * a.w + 11ull * b.x + 13ull * b.y + 17ull * b.z + 19ull * b.w;
              }
            }
            for (int mm = 16; mm >= 1; mm >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, mm);
            if (lane == 0) {
              double* o = P.dbg_steps + ((size_t)(epoch - 1) * ncta + cta) * DBG_W;
              o[6] = (double)SI.kself; o[7] = (double)SJ.kself; o[8] = (double)Kij; o[9] = yi * so.di; o[10] = yj * so.dj;
              o[11] = (double)(cs & 0xffffffffffffull) + (double)SI.norm * 1e-3 + (double)SJ.norm * 1e-6;
              if (cta == 0) {
                double* o2 = P.dbg_steps + (size_t)DBG_EPOCHS * ncta * DBG_W + (size_t)(epoch - 1) * 2;
                o2[0] = so.ai; o2[1] = so.aj;
              }
            }
          }
#endif
          if (lane == 0) {
            sh->ci = yi * so.di;
            sh->cj = yj * so.dj;
            sh->ai = so.ai;
            sh->aj = so.aj;
            sh->ni = SI.norm;
            sh->nj = SJ.norm;
            sh->i = si;
            sh->j = sj;
            sh->ti = ti;
            sh->tj = tj;
            sh->fi = make_flags(yi > 0, so.ai, Cbox);
            sh->fj = make_flags(yj > 0, so.aj, Cbox);
          }
        }
        __syncwarp();
        if (lane == 0) sh->leader = L;
        stamp(5);   // leader done: step taken
      }
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
    if (!(CACHE && sh->warm != 0)) {   // (a warm sweep never touches the rows)
      const V* xri = xs + sh->ti * CH;
      const V* xrj = xs + sh->tj * CH;
#pragma unroll
      for (int q = 0; q < C::CPL; ++q) {
        xi[q] = xri[lig + q * C::LPR];
        xj[q] = xrj[lig + q * C::LPR];
      }
    }
    reset_best();
#ifdef B2SVM_TIMELINE
    if (P.dbg_steps && warp == NW - 1 && epoch <= (uint32_t)DBG_EPOCHS) {
      unsigned long long cs = 0;
#pragma unroll
      for (int q = 0; q < C::CPL; ++q) {
        const uint4 a = *reinterpret_cast<const uint4*>(&xi[q]);
        const uint4 b = *reinterpret_cast<const uint4*>(&xj[q]);
        cs += (unsigned long long)a.x + 3ull * a.y + 5ull * a.z + 7ull * a.w + 11ull * b.x + 13ull * b.y + 17ull * b.z + 19ull * b.w;
      }
      for (int mm = 16; mm >= 1; mm >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, mm);
      if (lane == 0) {
        double* o = P.dbg_steps + ((size_t)(epoch - 1) * ncta + cta) * DBG_W;
        o[12] = ci; o[13] = cj;
        o[14] = (double)(cs & 0xffffffffffffull) + (double)ni * 1e-3 + (double)nj * 1e-6;
        o[15] = (double)wi * 1e6 + (double)wj;
      }
    }
#endif
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
      if (sh->p_ei >= 0) P.slot_of[sh->p_ei] = -1;
      if (sh->p_ej >= 0) P.slot_of[sh->p_ej] = -1;
      if (sh->p_si >= 0) { P.slot_of[sh->p_ri] = sh->p_si; if (P.row_of_slot) P.row_of_slot[sh->p_si] = sh->p_ri; }
      if (sh->p_sj >= 0) { P.slot_of[sh->p_rj] = sh->p_sj; if (P.row_of_slot) P.row_of_slot[sh->p_sj] = sh->p_rj; }
      __threadfence();
    }
    // gradient update and candidate search of one row (and of its mirror variable), given its two kernel values
    auto finish_row = [&](RowState<T>& cur, long long row, T kti, T ktj) {
      // y_t * (delta_i Q_i[t] + delta_j Q_j[t]) = (y_i delta_i) K_ti + (y_j delta_j) K_tj   (solver.py:146)
      const double c = ci * (double)kti + cj * (double)ktj;
      {
        const int v = (int)(row + goff);
        const double g = cur.g0 - c;
        G[row] = g;
        int f = cur.f0;
        if (v == wi) { f = nfi; flags[row] = (uint8_t)nfi; alpha[row] = nai; }
        else if (v == wj) { f = nfj; flags[row] = (uint8_t)nfj; alpha[row] = naj; }
        consider(nu, f, g, v, best, alpha + row);
        cur.g0 = g;
        cur.f0 = f;
      }
      if (mirror) {
        const int v = (int)(row + goff + lg);
        const double g = cur.g1 - c;
        G[row + l] = g;
        int f = cur.f1;
        if (v == wi) { f = nfi; flags[row + l] = (uint8_t)nfi; alpha[row + l] = nai; }
        else if (v == wj) { f = nfj; flags[row + l] = (uint8_t)nfj; alpha[row + l] = naj; }
        consider(nu, f, g, v, best, alpha + row + l);
        cur.g1 = g;
        cur.f1 = f;
      }
    };
    int m = warp, sub = 0;              // macro tile (first one static: its gradient state is already in pf), tile in it
    if (CACHE && warm && !one_tile) {
      // ---- warm sweep: two cached columns and the gradient, no arithmetic to hide a load behind.  One tile ahead
      // (the prefetch depth of the recomputing path) makes every tile cost a full memory latency, so the loads of
      // WARM_U tiles are in flight together; tiles are handed out statically (macro tiles warp, warp + NW, ...).
      const int per_warp = warp < NM ? ((NM - 1 - warp) / NW + 1) * SUB : 0;
#ifdef B2SVM_TIMELINE
      const bool wdbg = P.timeline && cta == 0 && threadIdx.x == 0 && epoch == 10u;
      long long* const wo = P.timeline + TL_ITERS * TL_SLOTS + 12 * TL_WARP_SLOTS;
      if (wdbg) wo[3] = clock64();   // warm path entered (absolute)
#endif
      for (int j0 = 0; j0 < per_warp; j0 += WARM_U) {
        RowState<T> rs[WARM_U];
        int ts[WARM_U];
#pragma unroll
        for (int u = 0; u < WARM_U; ++u) {
          const int j = j0 + u;
          ts[u] = j >= per_warp ? NT : (SUB == 1 ? warp + j * NW : (warp + (j / SUB) * NW) * SUB + (j % SUB));
          rs[u] = load_row_state(ts[u], false);
        }
#ifdef B2SVM_TIMELINE
        if (wdbg && j0 == 0) wo[0] = clock64();   // loads of the first batch issued
#endif
        // (without the fences the float -> double conversion of a tile's kernel values is hoisted to right behind its
        // loads, and the in-order issue then waits out one memory latency per tile instead of one per batch)
        asm volatile("" ::: "memory");
#pragma unroll
        for (int u = 0; u < WARM_U; ++u) {
          opaque(rs[u].ki); opaque(rs[u].kj); opaque(rs[u].g0); opaque(rs[u].f0);
          if (mirror) { opaque(rs[u].g1); opaque(rs[u].f1); }
          const long long row = row0 + (long long)ts[u] * C::BR + rib;
          if (ts[u] < NT && lane_active && row < row1) finish_row(rs[u], row, rs[u].ki, rs[u].kj);
#ifdef B2SVM_TIMELINE
          if (wdbg && j0 == 0 && u == 0) wo[1] = clock64();   // first tile done
#endif
        }
      }
#ifdef B2SVM_TIMELINE
      if (wdbg) wo[2] = clock64();   // all tiles done
#endif
      m = NM;   // (the tile loop below is the recomputing path)
    } else if (warm) {   // the first tile's state was prefetched before this sweep's plan was known
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
      RowState<T> cur = pf;
      if (!one_tile) pf = load_row_state(mn * SUB + subn);

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
        if (dbg && threadIdx.x == 0) P.timeline[TL_ITERS * TL_SLOTS + 12 * TL_WARP_SLOTS + 3] = clock64() - dbg_start;   // tile loop entered (sweep set-up done)
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
#ifdef B2SVM_TIMELINE
        if (dbg && threadIdx.x == 0) P.timeline[TL_ITERS * TL_SLOTS + 12 * TL_WARP_SLOTS + 0] = clock64() - dbg_start;   // dot products done
#endif
        TransposeReduce<T, C::NV, C::LPR / 2>::run(acc, lig);
#ifdef B2SVM_TIMELINE
        if (dbg && threadIdx.x == 0) P.timeline[TL_ITERS * TL_SLOTS + 12 * TL_WARP_SLOTS + 1] = clock64() - dbg_start;   // transposed reduce done
#endif
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
        finish_row(cur, row, kti, ktj);
      }
#ifdef B2SVM_TIMELINE
      if (dbg && threadIdx.x == 0) P.timeline[TL_ITERS * TL_SLOTS + 12 * TL_WARP_SLOTS + 2] = clock64() - dbg_start;   // row epilogue done
#endif
      if (one_tile) pf = cur;   // the warp's only tile: gradient, flags and norm stay in registers for the next sweep
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
    if (!one_tile) pf = load_row_state(warp * SUB);   // first tile of the next sweep (after this thread's own stores to it)
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
