// cemc_batch_kernel.cuh -- speculative batch evaluation of one Markov chain.
//
// A single chain is a dependent sequence of trial moves, so with few replicas
// (64 on 148 SMs) a B200 is latency bound: every cycle of one move's evaluation
// is exposed.  This kernel keeps the chain EXACTLY sequential but takes the
// expensive part off the critical path:
//
//   * B warps of a CTA evaluate the CF increments of the next B trial moves
//     concurrently, one WARP per move, all against the current state (phases
//     P0-P2b of cemc_kernels.cuh plus the per-ECI quotients; warp-level syncs only), screen
//     their move's Metropolis test and note which earlier moves of the batch would invalidate
//     the evaluation (conflict mask from a per-warp site bitmap);
//   * after ONE barrier EVERY warp decides the moves IN ORDER from those records (three
//     ballots, no loop) and applies the accepted changes itself -- all warps store identical
//     values to identical addresses, so nobody waits for a deciding warp.  A move whose
//     inputs were touched by an earlier accepted move of the same batch -- one of its
//     gathered sites (or, for swaps, one of its list slots) changed -- is not decided: the
//     batch ends there and the next batch re-evaluates it from the committed state.  The
//     Philox stream is keyed by the step index, so the restart changes nothing;
//   * the observer warp does the exact bookkeeping of a decided batch (CF vector, ordered
//     energy dot products, observer sums) while the evaluation warps work on the next one.
//
// EV_TAB32 (opt-in fp32 variant, cemc_set_precision): the tables and the sums over the
// sub-clusters are single precision (half the shared-memory traffic, FADD instead of DADD
// chains); quotients, CF vector, energies and observer sums stay fp64.  Same decisions as the
// fp64 path unless |dE + kT ln u| is below fp32 rounding; CFs / energies within 1e-5 (tested).
//
// Results are bit-identical to the one-move-at-a-time kernels (tested): the
// arithmetic of each phase is the same code path, operation for operation.
//
// CTA clusters (template parameter C = 2): the two CTAs of a thread-block cluster work on ONE
// chain: twice the evaluation throughput for one chain, the exact sequential Markov chain is
// kept.  Each CTA evaluates B moves of the batch (2 B <= 32 moves per batch) and BOTH decide:
// the CTAs exchange their records over an async DSMEM protocol (st.async + mbarrier
// transaction bytes, see kAsync below) and each commits its own copy of the occupations (or,
// for supercells that do not fit, the same global-memory copy).
//
// Template parameters of batch_kernel:
//   MODE        MODE_SGC (one-site flips) | MODE_CANONICAL (swaps: two changed sites per move)
//   kTree       4-way interleaved sums over sub-clusters (CEMC_ORDER_TREE) instead of the
//               reference's sequential order
//   B           evaluation warps per CTA (15 / 7 / 3; plus one observer / bookkeeper warp)
//   kStateSmem  occupations and site lists in shared memory (else global memory)
//   C           CTAs per chain (1 | 2)
//   EV          EV_PRODUCT | EV_SPIN | EV_TAB | EV_TAB32: how one move is evaluated
//   M           moves per evaluation warp and batch (1 | 2; 2 = interleaved spin evaluations, variant 6)
//   kSplit      site split (swaps, C = 2): both CTAs evaluate the same moves, one changed site each
//   kWide       spin / table evaluation with 32..63 translation columns (two per lane)
//   E           ECIs per lane (1 | 2: up to 64 ECIs)
//   kX          replay / lattice arithmetic / observer boundaries compiled in
//
// Used when the CF vector fits one warp (<= 32 E ECIs) and K <= 31 translation columns (<= 63
// for the spin and table evaluations); several symmetry groups: table evaluation only;
// everything else runs mc_kernel.
#pragma once
#include <cooperative_groups.h>
#include <type_traits>

#include "cemc_kernels.cuh"
#include "cemc_spin_kernel.cuh"

namespace cemc {

// Table evaluation (EV_TAB).  With S species and clusters of n <= 4 sites, the
// left-to-right product of spin_product_one_atom (ce_updater.cpp:271-281) of one
// sub-cluster under one decoration takes one of S^n values.  They are tabulated once
// per (family, decoration) -- with the reference's own multiplication order, so every entry
// is the bit pattern the reference computes -- and a move is evaluated by (1) packing
// the occupations of each sub-cluster into a base-S code (once per sub-cluster, shared
// by all decorations, old and new species of the changed site side by side) and (2)
// one lane per (ECI, decoration) summing table entries in the reference's sub-cluster
// order.  No products, no per-decoration gathers: ~3x fewer instructions per move.
struct TabTables {
  int n_sub;          // sub-clusters of one changed site, families back to back, each padded to 8
  int n_rounds;       // ceil(n_sub / 32), <= 4
  int n_tab;          // doubles in `tab`
  const uint2 *desc;  // [n_rounds*32] x: col0 | col1<<8 | col2<<16 | n_deco<<24; y: w0 | w1<<8 | w2<<16 | wref<<24 (w = S^position)
  const int4 *task;   // [n_tasks] {byte offset of the task's column in its family's table, first sub-cluster, M, 0}
  const double *tab;  // [n_tab] per family [code][decoration] product tables
  const float *tab32; // the same tables rounded to fp32 (EV_TAB32, the fp32 variant)
};

enum BatchEval : int { EV_PRODUCT = 0, EV_SPIN = 1, EV_TAB = 2, EV_TAB32 = 3 };

struct BatchSmem {
  double *V, *PO, *PN, *diff, *sq, *bf, *Pm, *Ch, *obE, *tab, *pub, *qtab;
  uint4 *items;
  int2 *task_sum;
  int4 *ttask;
  uint32_t *codes;          // [B][NJ][n_sub] old code*8 | new code*8 << 16
  uint4 *ring;              // [32][2]: proposal; uniform + Metropolis threshold
  int32_t *prop;            // [B][8] decoded proposal
  int2 *scr;                // [2][BT] {conflict mask (bit k: move b reads a site move k changes), screen verdict (1 accept, 2 inconclusive)}
  uint32_t *bmap;           // [B][ceil(N / 32)] per evaluation warp: the sites its move gathers (conflict masks), 0 words for large cells
  int4 *rec;                // [2][2][BT] site split: {conflict mask, -, dE} of changed site 0 / 1 (summed by every warp when it decides)
  int32_t *ctl;             // cluster protocol: [0..1] mailbox of the exact decision (CTA 1), [4..5] landing pad of the other CTA's token
  int32_t *list;
  int8_t *occ;
  uint64_t *mbar;           // [0] TMA staging copies; async cluster protocol: [1], [2] results of a batch complete (even / odd batches), [3] exact decision arrived (CTA 1)
};

// words of one evaluation warp's site bitmap (conflict masks); 0 = cell too large, shuffle / vote scan instead
__host__ __device__ inline int bmap_words(int N) { return N <= 16384 ? (N + 31) / 32 : 0; }

// B = moves evaluated by this CTA, BT = moves per batch over the whole cluster
// E = ECIs per lane (ECI i lives in lane i % 32, slot i / 32): rows of per-ECI data are 32 E wide
template <int B, int BT = B, int E = 1>
__host__ __device__ inline size_t batch_smem_layout(BatchSmem *s, unsigned char *base,
                                                    const DeviceTables &t, bool canonical,
                                                    bool state_in_smem = true,
                                                    const TabTables *tb = nullptr,
                                                    bool tab_fp32 = false, int spin_wq = 0) {
  size_t o = 0;
#define CEMC_TAKE(field, type, count)                                   \
  do {                                                                  \
    o = align_up(o, sizeof(type) < 16 ? sizeof(type) : 16);             \
    if (s) s->field = reinterpret_cast<type *>(base + o);               \
    o += sizeof(type) * at_least_1((int)(count));                       \
  } while (0)
  // destinations of TMA bulk copies: 16-byte aligned, size rounded up to 16 bytes
#define CEMC_TAKE16(field, type, count)                                 \
  do {                                                                  \
    o = align_up(o, 16);                                                \
    if (s) s->field = reinterpret_cast<type *>(base + o);               \
    o += align_up(sizeof(type) * at_least_1((int)(count)), 16);         \
  } while (0)
  const int nj = canonical ? 2 : 1;
  CEMC_TAKE(V, double, tb ? 0 : B * nj * t.VS);
  CEMC_TAKE(PO, double, tb ? 0 : B * nj * t.max_slots);
  CEMC_TAKE(PN, double, tb ? 0 : B * nj * t.max_slots);
  CEMC_TAKE(diff, double, B * nj * t.max_tasks);
  CEMC_TAKE16(tab, double, tb ? (tab_fp32 ? (tb->n_tab + 1) / 2 : tb->n_tab) : 0);
  CEMC_TAKE16(ttask, int4, tb ? t.n_tasks_total : 0);
  o = align_up(o, 16);                     // code words are read four at a time
  CEMC_TAKE(codes, uint32_t, tb ? B * nj * tb->n_sub : 0);
  CEMC_TAKE(sq, double, 2 * BT * 2 * 32 * E);  // double buffered: the bookkeeper reads batch k during batch k+1
  CEMC_TAKE(pub, double, 2 * (32 * E + 2));   // double buffered (batch parity)
  CEMC_TAKE(qtab, double, spin_wq * 64);      // spin evaluation: quotient table [new species][count][ECI lane]
  CEMC_TAKE(Pm, double, BT * (32 * E + 1));
  CEMC_TAKE(Ch, double, BT * 32 * E);
  CEMC_TAKE(obE, double, BT);
  CEMC_TAKE(bf, double, t.D * t.S);
  CEMC_TAKE16(items, uint4, tb ? 0 : t.n_items_total);
  CEMC_TAKE16(task_sum, int2, tb ? 0 : t.n_tasks_total);
  CEMC_TAKE(ring, uint4, 128 * 2);
  CEMC_TAKE(prop, int32_t, 2 * BT * 8);
  CEMC_TAKE(scr, int2, 2 * BT);
  CEMC_TAKE(rec, int4, 2 * 2 * BT);
  CEMC_TAKE(bmap, uint32_t, B * bmap_words(t.N));
  o = align_up(o, 8);
  CEMC_TAKE(ctl, int32_t, 8);
  CEMC_TAKE(mbar, uint64_t, 4);
  if (state_in_smem) {
    if (canonical) CEMC_TAKE16(list, int32_t, t.N);
    CEMC_TAKE16(occ, int8_t, t.N + 32);      // + the misalignment of the replica's row in global memory (TMA staging)
  }
#undef CEMC_TAKE16
#undef CEMC_TAKE
  return align_up(o, 16);
}

// offs[sp] for a runtime sp without spilling the 9-entry array to local memory
__device__ __forceinline__ int offs_of(const int (&offs)[9], int sp) {
  int v = 0;
#pragma unroll
  for (int q = 0; q < 9; q++) if (q == sp) v = offs[q];
  return v;
}

// one rounding per operation in either precision (no contraction: -fmad=false and the intrinsics)
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }

// One WARP evaluates one trial move; B warps = B moves per batch.
// kStateSmem = false: occupations / site lists stay in global memory (L2): supercells
// whose occupations do not fit in shared memory (64^3); the batch hides the latency.
// The CTA has B evaluation warps plus one OBSERVER warp (CTA 0 only does work in
// it): it folds the decided moves of batch k into the Averager / SGCObserver sums
// while the evaluation warps are already busy with batch k+1, which takes the
// per-move observer arithmetic off the evaluation / decision path.
// kSpin: binary +-1 basis -- the evaluation of a move is the XOR / ballot / popcount
// scheme of cemc_spin_kernel.cuh instead of fp64 products (same quotients, bit for bit).
// M: moves per evaluation warp and batch (spin evaluation: interleaved instruction streams): M = 2
// doubles the batch at the same number of warps, so the per-batch costs (the barrier, the decision
// pass) are shared by twice as many moves; the price is more speculation lost on hot chains.
// kSplit (canonical, C = 2): SITE SPLIT -- both CTAs of the cluster evaluate the SAME B moves of a
// batch, CTA q the CF change of changed site q of every swap (update_cf is called once per changed
// site, ce_updater.cpp:845-852, and the two calls only meet in the sum of their increments).  A
// swap then costs an evaluation warp what a one-site flip costs, instead of twice that.
// kWide (spin evaluation only): 32 <= K <= 63 translation columns -- every lane gathers two columns
// (lane, lane + 32), the occupation mask has 64 bits.
// E (table / product evaluation, C = 1): ECIs per lane -- up to 32 E ECIs (quaternary systems,
// ternary systems with many families); ECI i is slot i / 32 of lane i % 32.
// kX: the rarely used sources / sinks are compiled in -- replay of recorded proposals, translation by
// index arithmetic, observer boundaries.  Their mere presence costs the plain kernels ~10 % (code
// size and registers: 114 -> 126 on the bench kernel, measured A/B), so the flavours the benchmarks
// run exist in both forms and the host picks kX = true only when one of the three is in use.
template <int MODE, bool kTree, int B, bool kStateSmem, int C, int EV, int M = 1, bool kSplit = false, bool kWide = false, int E = 1, bool kX = true>
__global__ void __launch_bounds__((B + 1) * 32, (EV != EV_PRODUCT && B <= 8 && E == 1) ? 2 : 1)
batch_kernel(DeviceTables t, ReplicaState st, RunArgs a, int acc_stride, SpinTables sp, TabTables tb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  namespace cg = cooperative_groups;
  constexpr bool kSpin = (EV == EV_SPIN), kTab32 = (EV == EV_TAB32), kTab = (EV == EV_TAB) || kTab32;
  using TR = typename std::conditional<kTab32, float, double>::type;      // table / sub-cluster-sum type
  constexpr int TSH = kTab32 ? 2 : 3;                                       // log2(sizeof(TR))
  constexpr bool kCanon = (MODE == MODE_CANONICAL);
  constexpr int NJ = kCanon ? 2 : 1;
  static_assert(!kSplit || (C == 2 && MODE == MODE_CANONICAL && M == 1), "site split: swaps on a 2-CTA cluster");
  constexpr int BW = kSplit ? B : B * C;         // moves evaluated concurrently by the chain's warps
  constexpr int BT = BW * M;                     // moves per batch over the whole cluster
  constexpr int NJE = kSplit ? 1 : NJ;           // changed sites one warp evaluates
  static_assert(BT <= 32, "one decision lane per move");
  static_assert(E == 1 || (C == 1 && EV != EV_SPIN && M == 1), "several ECIs per lane: one CTA per chain, table / product evaluation");
  constexpr int LW = 32 * E;                     // width of a row of per-ECI values
  static_assert(!kWide || EV != EV_PRODUCT, "two columns per lane: spin and table evaluation");
  const int crank = C > 1 ? (int)cg::this_cluster().block_rank() : 0;
  const int r = a.order ? a.order[blockIdx.x / C] : (int)(blockIdx.x / C);
  const int tid = threadIdx.x, nthr = (B + 1) * 32;
  const int lane = tid & 31, lwarp = tid >> 5;
  const bool is_obs = (lwarp == B);              // the observer warp (works in CTA 0 only)
  const int warp = is_obs ? 1000 : (kSplit ? lwarp : crank * B + lwarp);   // move index of an evaluation warp
  const int jb = kSplit ? crank : 0;             // first changed site this warp evaluates
  const bool is_decider = (warp == 0) && (!kSplit || crank == 0);   // the ONE warp that counts accepted moves, writes g_loc, mails exact decisions
  auto csync = [&]() { if (C > 1) cg::this_cluster().sync(); else __syncthreads(); };
  const int N = t.N, K = t.K, KP = t.KP, S = t.S, D = t.D, VS = t.VS, n_eci = t.n_eci;
  const int RB = D * KP;
  const int bmw = bmap_words(N);
  const int max_slots = t.max_slots, max_tasks = t.max_tasks;
  const int n_items = t.item_base[1] - t.item_base[0];
  const int n_tasks = t.task_base[1] - t.task_base[0];

  int8_t *g_occ = st.occ + (size_t)r * N;
  int32_t *g_list = st.list + (size_t)r * N;
  int32_t *g_loc = st.loc + (size_t)r * N;
  BatchSmem s;
  batch_smem_layout<B, BT, E>(&s, smem_raw, t, kCanon, kStateSmem, kTab ? &tb : nullptr, kTab32, kSpin ? sp.wq : 0);
  // (the shared-memory occupations start at the row's offset inside its first 16-byte block: TMA staging below)
  const uint32_t occ_mis = kStateSmem ? (uint32_t)(reinterpret_cast<uintptr_t>(g_occ) & 15u) : 0u;
  if (!kStateSmem) { s.occ = g_occ; s.list = g_list; } else s.occ += occ_mis;
  // ---- async cluster protocol (C = 2).  Both CTAs keep a copy of the
  // chain's state and BOTH decide: every evaluation warp writes the screen record of its move
  // into its own CTA's shared memory and sends it to the other CTA with st.async (SASS STAS),
  // whose bytes complete the destination's mbarrier; every warp of either CTA waits for (local
  // arrivals, remote bytes) of the batch, derives the same decision record with ballots and
  // applies the accepted changes to its CTA's copy itself (identical values to identical
  // addresses).  No barrier.cluster (with its MEMBAR.ALL.GPU / L1 invalidation), no block
  // barrier, no decision round trip inside the batch loop.  CTA 1 additionally sends the
  // per-ECI quotients to CTA 0, whose observer warp keeps the books; the rare inconclusive
  // screen is settled by CTA 0 (it has the published CF vector) and mailed to CTA 1.
  // Everything is double buffered by the batch parity -- two mbarriers as well, so that bytes
  // of batch k + 1 can never land in the phase of batch k --, and the observer warps send a
  // token per batch so that no CTA runs more than one batch ahead of ANY warp of the other.
  constexpr bool kAsync = (C == 2);
  const bool remote = kAsync && crank == 1;
  // One CTA per chain: ONE barrier per batch.  Every evaluation warp screens its own move and
  // publishes {conflict mask, verdict}; after the barrier EVERY warp derives the same decision
  // record from those words (ballots, no loop) and applies the accepted changes itself -- the
  // warps store identical values to identical addresses, and each sees its own stores -- so
  // nobody waits for a deciding warp.  Everything a late warp still reads of batch k while an
  // early one writes batch k + 1 is double buffered by the batch parity.
  static_assert(C == 1 || C == 2, "one CTA or a cluster of two per chain");
  // bytes one evaluation warp sends to the other CTA per batch: its screen record and its proposal
  // (site split: a 16-byte partial record; both CTAs derive every proposal themselves); CTA 1 also
  // sends the quotients of its changed site(s)
  constexpr uint32_t kTxRec = kSplit ? 16u : (8u + 32u);
  constexpr uint32_t kTxSq = (kSplit || !kCanon) ? 256u : 512u;        // one 256-byte quotient row per changed site
  uint32_t r_S0 = 0, r_S1 = 0, r_sq = 0, r_scr = 0, r_rec = 0, r_prop = 0, r_X = 0, r_ctl = 0, r_tok = 0, phX = 0;
  if (kAsync) {
    const uint32_t other = (uint32_t)(crank ^ 1);
    r_S0 = mapa_u32(smem_u32(s.mbar + 1), other);
    r_S1 = mapa_u32(smem_u32(s.mbar + 2), other);
    r_sq = mapa_u32(smem_u32(s.sq), 0);
    r_scr = mapa_u32(smem_u32(s.scr), other);
    r_rec = mapa_u32(smem_u32(s.rec), other);
    r_prop = mapa_u32(smem_u32(s.prop), other);
    r_X = mapa_u32(smem_u32(s.mbar + 3), 1);
    r_ctl = mapa_u32(smem_u32(s.ctl), 1);
    r_tok = mapa_u32(smem_u32(s.ctl + 4), other);
  }
  // results of one evaluated move (b = move of the batch, pp = batch parity)
  auto put_sq = [&](int pp, int b, int half, double q, int e = 0) {   // per-ECI quotient, (lane, e) = ECI e * 32 + lane
    const int idx = pp * (BT * 2 * LW) + b * 2 * LW + half * LW + e * 32 + lane;
    if (remote) st_async_f64(r_sq + (uint32_t)idx * 8u, q, pp ? r_S1 : r_S0);
    else s.sq[idx] = q;
  };
  auto put_prop = [&](int pp, int b, int4 p0, int4 p1) {           // lane 0
    const int idx = pp * (BT * 8) + b * 8;
    if (kAsync) {                                                  // own copy; the other CTA's unless it derives the proposal itself
      *reinterpret_cast<int4 *>(s.prop + idx) = p0; *reinterpret_cast<int4 *>(s.prop + idx + 4) = p1;
      if (!kSplit) {
        const uint32_t rS = pp ? r_S1 : r_S0;
        st_async_v4b32(r_prop + (uint32_t)idx * 4u, p0, rS); st_async_v4b32(r_prop + (uint32_t)idx * 4u + 16u, p1, rS);
      }
      return;
    }
    *reinterpret_cast<int4 *>(s.prop + idx) = p0;
    if (kCanon) *reinterpret_cast<int4 *>(s.prop + idx + 4) = p1;    // old species / list slots: swaps only
  };
  // the screen record of move b (lane 0)
  auto put_scr = [&](int pp, int b, uint32_t m, int verdict, double dE) {
    if (kSplit) {                                                  // partial record of this CTA's changed site
      const int idx = (pp * 2 + crank) * BT + b;
      const int4 v = make_int4((int)m, 0, __double2loint(dE), __double2hiint(dE));
      s.rec[idx] = v;
      st_async_v4b32(r_rec + (uint32_t)idx * 16u, v, pp ? r_S1 : r_S0);
    } else {
      const int idx = pp * BT + b;
      s.scr[idx] = make_int2((int)m, verdict);
      if (kAsync) st_async_v2b32(r_scr + (uint32_t)idx * 8u, m, (uint32_t)verdict, pp ? r_S1 : r_S0);
    }
  };

  // ---- stage: the TMA engine copies the read-only tables and the replica's occupations / site
  // lists into shared memory (cp.async.bulk -> mbarrier transaction bytes); the threads only
  // fill what has no global image (V).  A replica's occupation row starts anywhere (N bytes per
  // replica): the copy covers the 16-byte blocks around it and the shared-memory array starts at
  // the same offset inside its first block (the occupation buffer is padded by 16 bytes).
  const bool occ_tma = kStateSmem;
  const bool list_tma = kStateSmem && kCanon && tma_aligned(g_list, (size_t)N * 4);
  if (tid == 0) {
    mbar_init(s.mbar, 1);
    if (kAsync) { mbar_init(s.mbar + 1, B + 1); mbar_init(s.mbar + 2, B + 1); mbar_init(s.mbar + 3, 1); }   // every warp of the CTA; the mailbox poster
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t b_tab = kTab ? (uint32_t)align_up((size_t)tb.n_tab * sizeof(TR), 16) : 0u;
    const uint32_t b_ttask = kTab ? (uint32_t)t.n_tasks_total * 16u : 0u;
    const uint32_t b_items = kTab ? 0u : (uint32_t)t.n_items_total * 16u;
    const uint32_t b_tsum = kTab ? 0u : (uint32_t)align_up((size_t)t.n_tasks_total * 8, 16);
    const uint32_t b_occ = occ_tma ? (uint32_t)align_up((size_t)N + occ_mis, 16) : 0u, b_list = list_tma ? (uint32_t)N * 4u : 0u;
    mbar_expect_tx(s.mbar, b_tab + b_ttask + b_items + b_tsum + b_occ + b_list);
    if (b_tab) tma_bulk_g2s(s.tab, kTab32 ? (const void *)tb.tab32 : (const void *)tb.tab, b_tab, s.mbar);
    if (b_ttask) tma_bulk_g2s(s.ttask, tb.task, b_ttask, s.mbar);
    if (b_items) tma_bulk_g2s(s.items, t.items4, b_items, s.mbar);
    if (b_tsum) tma_bulk_g2s(s.task_sum, t.task_sum, b_tsum, s.mbar);
    if (b_occ) tma_bulk_g2s(s.occ - occ_mis, g_occ - occ_mis, b_occ, s.mbar);
    if (b_list) tma_bulk_g2s(s.list, g_list, b_list, s.mbar);
  }
  for (int i = tid; i < D * S; i += nthr) s.bf[i] = t.bf[i];
  for (int i = tid; i < B * bmw; i += nthr) s.bmap[i] = 0u;
  if (!kTab)
    for (int i = tid; i < B * NJ * VS; i += nthr) s.V[i] = 1.0;      // V[K] is the constant 1.0
  if (kStateSmem) {
    if (!occ_tma) for (int i = tid; i < N; i += nthr) s.occ[i] = g_occ[i];
    if (kCanon && !list_tma) for (int i = tid; i < N; i += nthr) s.list[i] = g_list[i];
  }
  // canonical: species present and their list ranges (constant during a launch)
  int n_present = 0, present[8], offs[9];
#pragma unroll
  for (int q = 0; q < 9; q++) offs[q] = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) present[q] = 0;
  if (kCanon) {
#pragma unroll
    for (int sp = 0; sp < 9; sp++) if (sp <= S) offs[sp] = st.off[(size_t)r * (S + 1) + sp];
#pragma unroll
    for (int sp = 0; sp < 8; sp++)
      if (sp < S && offs[sp + 1] > offs[sp]) {
#pragma unroll
        for (int q = 0; q < 8; q++) if (q == n_present) present[q] = sp;
        n_present++;
      }
  }
  mbar_wait(s.mbar, 0);                        // every staged byte has landed
  if (kCanon && n_present < 2 && (!kX || a.rp_sites == nullptr)) {   // TooFewElementsError, montecarlo.py:310
    if (tid == 0) st.status[r] = 2;
    return;
  }
  csync();

  // ---- lane i owns ECI i (every warp: per-ECI quotients; warp 0: CF vector) -------
  const double dN = (double)(unsigned)N;
  int f_kind[E], f_d[E], f_t0[E], f_nd[E], my_singlet = -1;
  double f_scale[E], f_den[E], f_rden[E], eci_reg[E], cf_reg[E];
  double aE0 = 0.0, aE1 = 0.0, aE2 = 0.0, aS0 = 0.0, aS1 = 0.0, aS2 = 0.0;
#pragma unroll
  for (int e = 0; e < E; e++) {
    f_kind[e] = 0; f_d[e] = 0; f_t0[e] = 0; f_nd[e] = 0;
    f_scale[e] = 0.0; f_den[e] = 1.0; f_rden[e] = 1.0; eci_reg[e] = 0.0; cf_reg[e] = 0.0;
    const int i = e * 32 + lane;
    if (i < n_eci) {
      const int4 f = t.fin_i[i];
      f_kind[e] = f.x; f_d[e] = f.y; f_t0[e] = f.z; f_nd[e] = f.w - f.z;
      const double2 fd = t.fin_d[i];
      f_scale[e] = fd.x;
      f_den[e] = (f_kind[e] == 1) ? dN : fd.y;
      f_rden[e] = __ddiv_rn(1.0, f_den[e]);
      eci_reg[e] = st.eci[(size_t)r * n_eci + i];
      cf_reg[e] = st.cf[(size_t)r * n_eci + i];
    }
  }
  // Several translational symmetry groups (table evaluation only): the per-ECI constants depend
  // on the changed site's group and are read per evaluation; f_any = the ECI takes part in some
  // group (a group that lacks the family contributes a quotient of +0.0)
  const bool multi = kTab && t.n_symm > 1;
  bool f_any[E];
#pragma unroll
  for (int e = 0; e < E; e++) {
    f_any[e] = f_kind[e] > 0;
    if (multi && e * 32 + lane < n_eci)
      for (int g = 1; g < t.n_symm; g++) f_any[e] = f_any[e] || (t.fin_i[g * n_eci + e * 32 + lane].x > 0);
  }
  // singlets are the ECIs right after c0 in name order (c0 < c1_* < c2_*): slot 0 of their lanes
  for (int d = 0; d < t.n_singlets; d++) if (t.singlet_idx[d] == lane) my_singlet = d;
  if (is_obs && crank == 0) {
    const double *ag = st.acc + (size_t)r * acc_stride;
    if (my_singlet >= 0) { aS0 = ag[3 + 3 * my_singlet]; aS1 = ag[4 + 3 * my_singlet]; aS2 = ag[5 + 3 * my_singlet]; }
    aE0 = ag[0]; aE1 = ag[1]; aE2 = ag[2];
  }
  int bk_nd = 0, par = 0;              // bookkeeper: decided moves of the previous batch; batch parity
  uint32_t bk_am = 0u;
  int bk_base = 0;
  double e_cur = st.e_cur[r];
  const double kT = st.kT[r];
  const double rkT = __ddiv_rn(1.0, kT);
  const double ref = st.ref[r];
  const double rref = __ddiv_rn(1.0, ref);
  const bool ref_is_one = (ref == 1.0);
  const unsigned long long step0 = st.step[r];
  unsigned long long n_acc = 0;
  const uint32_t rep_global = a.replica_offset + (uint32_t)r * a.replica_stride;
  const int n_eci4 = pin_reg((n_eci + 3) & ~3);
  const bool few_eci = (E == 1) && n_eci <= 8;
  const int observe = pin_reg(a.observe);
  const bool tracing = (a.tr_acc != nullptr) || (a.tr_e != nullptr);
  const int n_allowed = t.n_allowed;
  // replay of recorded proposals / uniforms: the ring is filled from
  // rp_sites / rp_news / rp_u instead of the Philox stream; records hold sites, not list slots
  // (mode flags pinned in ONE register: tested in every evaluation, and a kernel parameter would
  // be re-read from the constant bank each time -- LDCU + a dependent uniform branch)
  const int mflags = !kX ? 0 : pin_reg((a.rp_sites != nullptr ? 1 : 0) | (t.lat_ok != 0 ? 2 : 0) |
                                       ((a.obs_interval > 0 && a.obs_interval < (1ll << 30)) ? 4 : 0));
  const bool replay = (mflags & 1) != 0;
  // observer boundaries (cemc_set_device_observers): batches end on them, so that the bookkeeper
  // sees the chain's state of the boundary step (the occupations only change in the D phase)
  // (32-bit countdowns: intervals >= 2^30 steps never fire inside one launch segment anyway)
  // ONE countdown register: moves until the next boundary -- as seen by the deciding / evaluation
  // warps (they cut the batch there), and, in the observer warp, the bookkeeper's copy that lags
  // one batch behind (that warp never needs the batch length); the interval itself is re-read
  // from the kernel parameters on the rare update.
  int to_ob = (mflags & 4) ? (int)(a.obs_interval - a.obs_origin % a.obs_interval) : 0x7fffffff;
  // translation-invariant lattice: lane c derives T(site, c) from the site index (no table gather)
  const bool lat_ok = (mflags & 2) != 0;
  const uint32_t my_shift = (lat_ok && lane < K) ? t.col_shift[lane] : 0u;
  const uint32_t my_shift2 = (lat_ok && kWide && lane + 32 < K) ? t.col_shift[lane + 32] : 0u;
  auto neighbour = [&](int site, int col, uint32_t shift) -> int {
    return lat_ok ? lattice_neighbour(t, site, shift) : __ldg(&t.trans[(size_t)site * K + col]);   // :264
  };
  // absolute slack of the Metropolis screen: rounding noise of the two ordered dot
  // products the reference subtracts, N * sum_i |eci_i| * max|cf| * O(n_eci * eps)
  double etol;
  {
    double sa = 0.0;
#pragma unroll
    for (int e = 0; e < E; e++) sa += fabs(eci_reg[e]) * fmax(1.0, fabs(cf_reg[e]));      // 0 beyond n_eci
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sa += __shfl_xor_sync(0xffffffffu, sa, o);
    etol = 1e-13 * dN * sa * 16.0 * a.screen_slack;
  }
  const double c_rel = 1e-9 * a.screen_slack;          // relative part of the screen's band

  // kSpin: this lane's sub-clusters (decoded once) and its ECI's ballot masks
  int sca[4], scb[4], scc[4];
  uint32_t smb[4], smc[4], smv[4], smask[4];
  int s_coef = 0, s_msub = 1;
  const int s_rounds = kSpin ? sp.n_rounds : 0;
  if (kSpin) {
    s_coef = sp.coef[lane]; s_msub = sp.msub[lane];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int qi = q * 32 + lane;
      const bool valid = qi < sp.n_items;
      const uint32_t w = valid ? sp.items[qi] : 0x00ffff00u;
      sca[q] = (int)(w & 0xffu);
      const uint32_t xb = (w >> 8) & 0xffu, xc = (w >> 16) & 0xffu;
      scb[q] = xb != 0xffu ? (int)xb : sca[q];
      scc[q] = xc != 0xffu ? (int)xc : sca[q];
      smb[q] = xb != 0xffu ? 1u : 0u;
      smc[q] = xc != 0xffu ? 1u : 0u;
      smv[q] = valid ? 1u : 0u;
      smask[q] = sp.masks[lane * 4 + q];
    }
    // Every quotient the evaluation can produce, once per launch: the numerator
    // n (sigma_new - sigma_old)(M - 2 count) takes 2 (M + 1) integer values per ECI, so the
    // exact division (ce_updater.cpp:402) becomes one shared-memory load per evaluated site.
    if (lwarp == 0) {
      for (int nw = 0; nw < 2; nw++)
        for (int cnt = 0; cnt < sp.wq; cnt++) {
          const int dsig = 2 * sp.b0 * (1 - 2 * nw);                 // old = 1 - new
          const int num = s_coef * dsig * (s_msub - 2 * cnt);
          s.qtab[(nw * sp.wq + cnt) * 32 + lane] = cnt <= s_msub ? exact_div((double)num, f_den[0], f_rden[0]) : 0.0;
        }
    }
  }

  // kTab: this lane's sub-cluster descriptors (columns of the other sites, code weights)
  uint32_t tdx[4], tdy[4];
  const int t_rounds = kTab ? tb.n_rounds : 0;
  const int n_sub = kTab ? tb.n_sub : 0;
  if (kTab) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint2 d = q < t_rounds ? tb.desc[q * 32 + lane] : make_uint2(0u, 0u);
      tdx[q] = d.x; tdy[q] = d.y;
    }
  }

  // ---- proposal ring: 128 records, produced 32 at a time by the observer warp of every
  // CTA while the evaluation warps work (montecarlo.py:890-908, sgc_montecarlo.py:62-76)
  auto produce32 = [&](int first) {
    const unsigned long long stp = step0 + (unsigned long long)first + lane;
    uint32_t c0 = (uint32_t)stp, c1 = (uint32_t)(stp >> 32), c2 = rep_global, c3 = 0;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    uint4 rec0;
    double u;
    if (replay) {
      // recorded step (or, past the end of the run, a harmless dummy nobody decides)
      const long long q = first + lane;
      int s0 = 0, s1 = kCanon ? 1 : -1, n0 = 0, n1 = 0;
      u = 0.5;
      if (q < a.n_steps) {
        const size_t g = (size_t)r * (size_t)a.n_steps + (size_t)q;
        s0 = a.rp_sites[2 * g]; n0 = a.rp_news[2 * g];
        if (kCanon) { s1 = a.rp_sites[2 * g + 1]; n1 = a.rp_news[2 * g + 1]; }
        u = a.rp_u[g];
      }
      rec0 = make_uint4((uint32_t)s0, (uint32_t)s1, (uint32_t)n0, (uint32_t)n1);
    } else if (!kCanon) {
      const uint32_t ia = __umulhi(c0, (uint32_t)t.n_active);       // sgc_montecarlo.py:69
      rec0 = make_uint4(ia, c1, 0u, 0u);
      u = u53(c2, c3);
    } else {                                                         // montecarlo.py:899-907
      uint32_t d0 = (uint32_t)stp, d1 = (uint32_t)(stp >> 32), d2 = rep_global, d3 = 1;
      philox4x32_10(d0, d1, d2, d3, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
      const int ia = (int)__umulhi(c0, (uint32_t)n_present);
      int ib = (int)__umulhi(c1, (uint32_t)(n_present - 1)); ib += (ib >= ia);
      int sa = 0, sb = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) { if (q == ia) sa = present[q]; if (q == ib) sb = present[q]; }
      const int oa0 = offs_of(offs, sa), oa1 = offs_of(offs, sa + 1);
      const int ob0 = offs_of(offs, sb), ob1 = offs_of(offs, sb + 1);
      const int slot0 = oa0 + (int)__umulhi(c2, (uint32_t)(oa1 - oa0));
      const int slot1 = ob0 + (int)__umulhi(c3, (uint32_t)(ob1 - ob0));
      rec0 = make_uint4((uint32_t)slot0, (uint32_t)slot1, (uint32_t)sb, (uint32_t)sa);
      u = u53(d0, d1);
    }
    // Metropolis threshold: u <= exp(-dE/kT)  <=>  dE <= -kT ln u =: L.  Screen only, so a
    // single-precision logarithm will do: |error| <= kT (6e-8 + 1.2e-7 |ln u|) is covered
    // by the band of the screen, inconclusive moves take the exact expression.  The record holds
    // the two thresholds of the screen with the move-independent part of the band folded in,
    // L -+ (slack 4e-7 (kT + |L|) + etol), rounded outwards to fp32: a move is accepted when
    // dE + c |dE| < L_lo, rejected when dE - c |dE| > L_hi (c = 1e-9 slack), inconclusive between.
    const double L = -kT * (double)logf((float)u);
    const double bb = a.screen_slack * (4e-7 * (kT + fabs(L))) + etol;
    const float L_lo = __double2float_rd(L - bb), L_hi = __double2float_ru(L + bb);
    const int slot = (int)((first + lane) & 127);
    s.ring[slot * 2] = rec0;
    s.ring[slot * 2 + 1] = make_uint4((uint32_t)__double2loint(u), (uint32_t)__double2hiint(u),
                                      __float_as_uint(L_lo), __float_as_uint(L_hi));
  };
#ifdef CEMC_PHASE_TIMING
  unsigned long long tph[24] = {0};
  long long tlast = clock64();
#endif
  // ---- bookkeeper (observer warp of CTA 0): the exact bookkeeping of a decided batch, in the
  // reference's order, off the critical path (it runs while the evaluation warps work on the
  // next batch): CF increments per accepted move (:404), ordered energy dot products
  // (:236-242, one lane per move), trace records, Averager / SGCObserver sums
  // (montecarlo.py:811-814, mc_observers.py:264-270).  The deciding warps only need the CF
  // vector and the energy for an inconclusive screen; they are published in s.pub.
  auto bookkeep = [&](int nd, uint32_t accmask, int base, int pp) {
#ifdef CEMC_PHASE_TIMING
    long long tb0 = clock64();
#define CEMC_OTICK(slot) do { const long long n_ = clock64(); if (lane == 0) tph[slot] += (unsigned long long)(n_ - tb0); tb0 = n_; } while (0)
#else
#define CEMC_OTICK(slot) do { } while (0)
#endif
    const double *sqp = s.sq + pp * (BT * 2 * LW);
#pragma unroll
    for (int e = 0; e < E; e++) {
      double c = cf_reg[e];
      // fully unrolled over the batch: compile-time addresses and mask bits, loads of a
      // group of G moves up front; the only serial chain is the DADDs of accepted moves
      constexpr int G = (BT % 5 == 0) ? 5 : (BT % 7 == 0) ? 7 : 3;
      static_assert(BT % G == 0, "group size must divide the batch");
#pragma unroll
      for (int b0 = 0; b0 < BT; b0 += G) {
        if (b0 < nd) {
          double qa[G], qb[G];
#pragma unroll
          for (int x = 0; x < G; x++) {
            qa[x] = sqp[(b0 + x) * 2 * LW + e * 32 + lane];
            qb[x] = kCanon ? sqp[(b0 + x) * 2 * LW + LW + e * 32 + lane] : 0.0;
          }
#pragma unroll
          for (int x = 0; x < G; x++) {
            if (accmask & (1u << (b0 + x))) {                  // warp-uniform
              if (f_any[e]) {                                  // kinds 0 / -1: copied (:360,:382)
                c = __dadd_rn(c, qa[x]);                       // :404
                if (kCanon) c = __dadd_rn(c, qb[x]);
              }
              s.Pm[(b0 + x) * (LW + 1) + e * 32 + lane] = __dmul_rn(eci_reg[e], c);
            }
            s.Ch[(b0 + x) * LW + e * 32 + lane] = c;           // entries >= nd are never read
          }
        }
      }
      if (accmask) cf_reg[e] = c;
      s.pub[(pp ^ 1) * (LW + 2) + e * 32 + lane] = c;                // read by the decisions of the batch after pp
    }
    __syncwarp();
    CEMC_OTICK(16);
    // exact energies of the accepted moves, one lane per move (ordered dot, :236-242)
    const bool my_acc = lane < nd && ((accmask >> lane) & 1u);
    double E_l = 0.0;
    if (my_acc) {
      const double *pm = s.Pm + lane * (LW + 1);
      double e = 0.0;
      // groups of four, loads up front: the entries of lanes >= n_eci are +0.0 products
      // (eci = cf = 0), and adding +0.0 leaves every partial sum bit for bit (the sum
      // starts at +0.0, so it is never -0.0)
      for (int i = 0; i < n_eci4; i += 4) {
        double v[4];
#pragma unroll
        for (int x = 0; x < 4; x++) v[x] = pm[i + x];
#pragma unroll
        for (int x = 0; x < 4; x++) e = __dadd_rn(e, v[x]);
      }
      E_l = __dmul_rn(e, dN);
    }
    // energy after move b = energy of the last accepted move <= b (else the old one)
    double E_after;
    {
      const uint32_t upto = accmask & (lane >= 31 ? 0xffffffffu : ((2u << lane) - 1u));
      const int src = upto ? 31 - __clz(upto) : 0;
      const double Es = __shfl_sync(0xffffffffu, E_l, src);
      E_after = upto ? Es : e_cur;
    }
    if (accmask) {
      const int last = 31 - __clz(accmask);
      e_cur = __shfl_sync(0xffffffffu, E_l, last);
    }
    if (lane == 0) s.pub[(pp ^ 1) * (LW + 2) + LW] = e_cur;
    if (lane < nd) s.obE[lane] = E_after;
    if (tracing && lane < nd && base + lane < a.tr_capacity) {
      const int4 pa = *reinterpret_cast<const int4 *>(s.prop + pp * (BT * 8) + lane * 8);
      const uint4 rec1 = s.ring[(int)((base + lane) & 127) * 2 + 1];
      const size_t q = (size_t)r * a.tr_capacity + (size_t)(base + lane);
      if (a.tr_sites) { a.tr_sites[2 * q] = pa.x; a.tr_sites[2 * q + 1] = pa.y; }
      if (a.tr_news) { a.tr_news[2 * q] = (int8_t)pa.z; a.tr_news[2 * q + 1] = (int8_t)pa.w; }
      if (a.tr_u) a.tr_u[q] = __hiloint2double((int)rec1.y, (int)rec1.x);
      if (a.tr_acc) a.tr_acc[q] = my_acc ? 1 : 0;
      if (a.tr_e) a.tr_e[q] = E_after;
    }
    __syncwarp();
    CEMC_OTICK(17);
    if (mflags & 4) {
      to_ob -= nd;                       // (observer warp: the lagging copy)
      if (to_ob <= 0) {                  // the batch ended on an observer boundary
        to_ob += (int)a.obs_interval;
        observer_snapshot(a, r, lane, n_eci, N, s.Ch + (nd - 1) * LW, e_cur, s.occ);
      }
    }
    if (!observe) return;
    if (ref_is_one) {            // Averager reference value 1: value / ref is the value itself
      // groups of GO moves: loads and squares of the whole group up front (entries >= nd are
      // in bounds and ignored), then six interleaved DADD chains, the only serial part
      constexpr int GO = (BT % 5 == 0) ? 5 : (BT % 7 == 0) ? 7 : 3;
      aE0 = __dadd_rn(aE0, (double)nd);          // a count: nd additions of 1.0 give the same (exact) bits
#pragma unroll 1
      for (int b0 = 0; b0 < nd; b0 += GO) {
        double Eb[GO], cb[GO], e2[GO], c2[GO], ce[GO];
#pragma unroll
        for (int x = 0; x < GO; x++) { Eb[x] = s.obE[b0 + x]; cb[x] = s.Ch[(b0 + x) * LW + lane]; }
#pragma unroll
        for (int x = 0; x < GO; x++) {
          e2[x] = __dmul_rn(Eb[x], Eb[x]); c2[x] = __dmul_rn(cb[x], cb[x]); ce[x] = __dmul_rn(cb[x], Eb[x]);
        }
#pragma unroll
        for (int x = 0; x < GO; x++) {
          if (b0 + x < nd) {
            aE1 = __dadd_rn(aE1, Eb[x]);
            aE2 = __dadd_rn(aE2, e2[x]);
            aS0 = __dadd_rn(aS0, cb[x]);
            aS1 = __dadd_rn(aS1, c2[x]);
            aS2 = __dadd_rn(aS2, ce[x]);
          }
        }
      }
    } else {
      for (int b = 0; b < nd; b++) {
        const double Eb = s.obE[b], cb = s.Ch[b * LW + lane];
        const double e2 = __dmul_rn(Eb, Eb);
        aE0 = __dadd_rn(aE0, 1.0);
        aE1 = __dadd_rn(aE1, exact_div(Eb, ref, rref));
        aE2 = __dadd_rn(aE2, exact_div(e2, ref, rref));
        aS0 = __dadd_rn(aS0, cb);
        aS1 = __dadd_rn(aS1, __dmul_rn(cb, cb));
        aS2 = __dadd_rn(aS2, __dmul_rn(cb, Eb));
      }
    }
    CEMC_OTICK(18);
  };

  int fill_end = 0;                    // records of steps [sdone, fill_end) are in the ring
  if (is_obs) { produce32(0); produce32(32); produce32(64); }
  if (is_obs && crank == 0) {
#pragma unroll
    for (int e = 0; e < E; e++) { s.pub[e * 32 + lane] = cf_reg[e]; s.pub[(LW + 2) + e * 32 + lane] = cf_reg[e]; }
    if (lane == 0) { s.pub[LW] = e_cur; s.pub[(LW + 2) + LW] = e_cur; }
  }
  fill_end = 96;
  csync();

  // 32-bit step counters inside a launch (the host splits runs into launches of at most 2^30 moves)
  const int n_steps = (int)a.n_steps;
  int sdone = 0;                       // moves decided so far
  int kb = 0;                          // batches done (par = kb & 1; phase of this parity's mbarrier = (kb >> 1) & 1)
#ifdef CEMC_WARP_TIMING
  long long wtW = 0, wtB = 0, wtD = 0;  // per warp: work before the barrier, barrier wait, decision (scripts/warp_timing.py)
#endif

  while (sdone < n_steps) {
    int nb = (n_steps - sdone) < BT ? (n_steps - sdone) : BT;
    {
      int tb = to_ob;
      if (is_obs && crank == 0 && (mflags & 4)) { tb -= bk_nd; if (tb <= 0) tb += (int)a.obs_interval; }   // its copy lags one batch
      if (tb < nb) nb = tb;
    }
    CEMC_TICK(0);
#ifdef CEMC_WARP_TIMING
    const long long wt0 = clock64();
#endif

#ifdef CEMC_PHASE_TIMING
    const long long tob0 = clock64();
#endif
    // ---- observer warp: next 32 proposal records, then the observer sums -----------------
    const bool refill = (fill_end - sdone) <= 64;
    if (is_obs && refill) produce32(fill_end);        // visible after the E1 barrier
#ifdef CEMC_PHASE_TIMING
    if (is_obs && lane == 0) tph[19] += (unsigned long long)(clock64() - tob0);
#endif
    if (refill) fill_end += 32;
    // ---- bookkeeper: exact CF vector, energies, trace and observer sums of the previous batch
    if (is_obs && crank == 0 && bk_nd > 0) bookkeep(bk_nd, bk_am, bk_base, par ^ 1);
#ifdef CEMC_PHASE_TIMING
    const long long tob1 = clock64();
#endif
    // ---- E1: warp b evaluates move sdone + b against the current state --------------
    // ---- which earlier moves of the batch would invalidate an evaluation?  Lane k holds the
    // site(s) move k changes if accepted (state independent for SGC; for swaps the current
    // list entries: had an accepted move changed them, move k would be invalid itself and
    // the batch would end before it)
    auto changed_sites = [&](int bmax, int &sk0, int &sk1) {
      sk0 = -2; sk1 = -2;
      if (lane < bmax) {
        const uint4 rk = s.ring[(int)((sdone + lane) & 127) * 2];
        if (replay) { sk0 = (int)rk.x; if (kCanon) sk1 = (int)rk.y; }
        else if (!kCanon) sk0 = t.active ? t.active[rk.x] : (int)rk.x;
        else { sk0 = s.list[rk.x]; sk1 = s.list[rk.y]; }
      }
    };
    // gsx[j]: lanes 0..K hold the K neighbours and the changed site itself.  With few
    // earlier moves, the site(s) of each one are broadcast and compared by the lanes holding
    // gathered sites (one vote per move); with many, every gathered site is broadcast and
    // compared by the lanes holding the moves.  Both are short independent shuffle / compare
    // / vote sequences (MATCH.ANY over 32 distinct values costs ~400 cycles on sm_100).
    auto conflict_mask = [&](int b, const int (&gsx)[2], int sk0, int sk1, const int (&gsy)[2]) -> uint32_t {
      uint32_t m = 0;
      if (b == 0) return 0u;
      if (bmw > 0) {
        // site bitmap of this warp: the gathering lanes set the bits of their sites, lane k tests
        // the site(s) move k changes, one ballot, the gathering lanes clear their words again
        uint32_t *bm = s.bmap + lwarp * bmw;
#pragma unroll
        for (int j = 0; j < NJE; j++) {
          if (gsx[j] >= 0) atomicOr(bm + (gsx[j] >> 5), 1u << (gsx[j] & 31));
          if (kWide && gsy[j] >= 0) atomicOr(bm + (gsy[j] >> 5), 1u << (gsy[j] & 31));
        }
        __syncwarp();
        bool hit = false;
        if (lane < b) {
          hit = ((bm[sk0 >> 5] >> (sk0 & 31)) & 1u) != 0u;
          if (kCanon) hit |= ((bm[sk1 >> 5] >> (sk1 & 31)) & 1u) != 0u;
        }
        m = __ballot_sync(0xffffffffu, hit);
#pragma unroll
        for (int j = 0; j < NJE; j++) {
          if (gsx[j] >= 0) bm[gsx[j] >> 5] = 0u;
          if (kWide && gsy[j] >= 0) bm[gsy[j] >> 5] = 0u;
        }
        __syncwarp();
        return m;
      }
      if (b <= KP) {
        // eight earlier moves at a time, their shuffles / votes independent of each other (lanes
        // >= b of sk0 / sk1 hold -2, which matches no site: no bound check inside a group)
        for (int k0 = 0; k0 < b; k0 += 8) {
#pragma unroll
          for (int x = 0; x < 8; x++) {
            const int k = (k0 + x) & 31;
            const int a0 = __shfl_sync(0xffffffffu, sk0, k);
            bool hit = (gsx[0] == a0);
            if (NJE == 2) hit |= (gsx[1] == a0);
            if (kWide) { hit |= (gsy[0] == a0); if (NJE == 2) hit |= (gsy[1] == a0); }
            if (kCanon) {
              const int a1 = __shfl_sync(0xffffffffu, sk1, k);
              hit |= (gsx[0] == a1);
              if (NJE == 2) hit |= (gsx[1] == a1);
              if (kWide) { hit |= (gsy[0] == a1); if (NJE == 2) hit |= (gsy[1] == a1); }
            }
            if (__any_sync(0xffffffffu, hit)) m |= 1u << k;
          }
        }
      } else {
        // eight gathered sites at a time (lanes past the last column hold -1, which is no site)
        bool hit = false;
#pragma unroll
        for (int j = 0; j < NJE; j++)
          for (int q0 = 0; q0 < KP; q0 += 8) {
#pragma unroll
            for (int x = 0; x < 8; x++) {
              const int q = q0 + x;
              const int g = __shfl_sync(0xffffffffu, (kWide && q >= 32) ? gsy[j] : gsx[j], q & 31);
              hit |= (g == sk0) | (kCanon & (g == sk1));
            }
          }
        m = __ballot_sync(0xffffffffu, hit) & (b >= 32 ? 0xffffffffu : ((1u << b) - 1u));
      }
      return m;
    };

    // the evaluating warp screens its own move against the two thresholds of its proposal record:
    // 1 = accept, 2 = inconclusive, 0 = reject
    auto screen = [&](int b, double dE) -> int {
      const float2 th = *reinterpret_cast<const float2 *>(&s.ring[(int)((sdone + b) & 127) * 2 + 1].z);
      const double m = c_rel * fabs(dE);
      const bool acc = dE + m < (double)th.x;
      const bool bdr = !acc && !(dE - m > (double)th.y);
      return (acc ? 1 : 0) | (bdr ? 2 : 0);
    };
    if constexpr (kSpin) {
      // ---- spin evaluation (cemc_spin_kernel.cuh), the M moves of this warp side by side
      // (independent instruction streams: the warp has few siblings on its scheduler, so
      // latency must be covered inside the warp).  Lane c gathers column c of the changed
      // site's translation-matrix row and its occupation; ONE ballot turns the K occupations
      // into a bit mask, every lane forms the parity of its sub-cluster(s) from that mask,
      // one ballot per 32 sub-clusters packs the parities, ECI lanes take popc(ballot & mask),
      // the exact integer numerator n (sigma_new - sigma_old)(M_sub - 2 popc), one exact
      // division.  Moves past the end of the run are evaluated too (their records exist in
      // the ring; nobody reads the results): no divergence between the M streams.
      if (!is_obs) {
        int site[M][2], oldv[M][2], newv[M][2], gs[M][2], gs2[M][2];
        double qv[M][2];
#pragma unroll
        for (int mi = 0; mi < M; mi++) {
          const int b = warp + mi * BW;           // warp w evaluates moves w, w + BW, ...
          const uint4 rec0 = s.ring[(int)((sdone + b) & 127) * 2];
          int slot0 = -1, slot1 = -1;
          site[mi][1] = -1; oldv[mi][1] = 0; newv[mi][1] = 0;
          if (replay) {                           // recorded sites / new species
            site[mi][0] = (int)rec0.x; newv[mi][0] = (int)rec0.z;
            oldv[mi][0] = s.occ[site[mi][0]];
            if (kCanon) {
              site[mi][1] = (int)rec0.y; newv[mi][1] = (int)rec0.w;
              oldv[mi][1] = site[mi][1] == site[mi][0] ? newv[mi][0] : (int)s.occ[site[mi][1]];
            }
          } else if (!kCanon) {                   // sgc_montecarlo.py:69-75 (all species allowed)
            site[mi][0] = t.active ? t.active[rec0.x] : (int)rec0.x;
            oldv[mi][0] = s.occ[site[mi][0]];
            newv[mi][0] = 1 - oldv[mi][0];         // two species: "the other one" (sgc_montecarlo.py:70-75 draws nothing else)
          } else {
            slot0 = (int)rec0.x; slot1 = (int)rec0.y; newv[mi][0] = (int)rec0.z; newv[mi][1] = (int)rec0.w;
            site[mi][0] = s.list[slot0]; site[mi][1] = s.list[slot1];
            oldv[mi][0] = newv[mi][1]; oldv[mi][1] = newv[mi][0];
          }
          if (lane == 0)                           // every warp commits from this record
            put_prop(par, b, make_int4(site[mi][0], site[mi][1], newv[mi][0], newv[mi][1]),
                     make_int4(oldv[mi][0], oldv[mi][1], slot0, slot1));
        }
#pragma unroll
        for (int mi = 0; mi < M; mi++) {
#pragma unroll
          for (int je = 0; je < 2; je++) {
            gs[mi][je] = -1; gs2[mi][je] = -1; qv[mi][je] = 0.0;
            if (je < NJE) {
              const int j = jb + je;
              // (selects, not a run-time index: with the site split j is the CTA rank)
              const int site_j = j ? site[mi][1] : site[mi][0];
              const int newv_j = j ? newv[mi][1] : newv[mi][0], oldv_j = j ? oldv[mi][1] : oldv[mi][0];
              uint32_t v = 0;
              if (lane < K) {
                const int nbs = neighbour(site_j, lane, my_shift);
                gs[mi][je] = nbs;
                v = (uint32_t)s.occ[nbs];
                if (j == 1 && nbs == site[mi][0]) v = (uint32_t)newv[mi][0];   // change 1 sees change 0 applied (:845-852)
              } else if (lane == K) gs[mi][je] = site_j;
              typename std::conditional<kWide, unsigned long long, uint32_t>::type ob =
                  __ballot_sync(0xffffffffu, (v & 1u) != 0u);
              if (kWide) {                         // columns 32 .. K-1 and the site itself in the upper half
                uint32_t v1 = 0;
                if (lane + 32 < K) {
                  const int nbs = neighbour(site_j, lane + 32, my_shift2);
                  gs2[mi][je] = nbs;
                  v1 = (uint32_t)s.occ[nbs];
                  if (j == 1 && nbs == site[mi][0]) v1 = (uint32_t)newv[mi][0];
                } else if (lane + 32 == K) gs2[mi][je] = site_j;
                ob |= (unsigned long long)__ballot_sync(0xffffffffu, (v1 & 1u) != 0u) << 32;
              }
              int cnt = 0;
#pragma unroll
              for (int q = 0; q < 4; q++) {
                if (q < s_rounds) {
                  const uint32_t bit = (uint32_t)((ob >> sca[q]) ^ ((ob >> scb[q]) & smb[q]) ^ ((ob >> scc[q]) & smc[q])) & smv[q];
                  cnt += __popc(__ballot_sync(0xffffffffu, bit != 0u) & smask[q]);
                }
              }
              qv[mi][je] = s.qtab[(newv_j * sp.wq + cnt) * 32 + lane];   // n dsigma (M - 2 cnt) / den, :393-402
              if (replay && newv_j == oldv_j) qv[mi][je] = 0.0;   // recorded no-op change (:315): the table assumes old != new
            }
          }
        }
#pragma unroll
        for (int mi = 0; mi < M; mi++) {
          const int b = warp + mi * BW;
          // this move's per-ECI quotients [2][32]
          if (!kSplit) { put_sq(par, b, 0, qv[mi][0]); if (kCanon) put_sq(par, b, 1, qv[mi][1]); }
          else put_sq(par, b, jb, qv[mi][0]);
        }
        double de[M];
#pragma unroll
        for (int mi = 0; mi < M; mi++)           // screen only
          de[mi] = f_kind[0] > 0 ? eci_reg[0] * (kCanon ? qv[mi][0] + qv[mi][1] : qv[mi][0]) : 0.0;
        if (few_eci) {                          // ECIs in lanes 0..7 only: lanes 8.. hold +0.0, three rounds give the same bits
#pragma unroll
          for (int o = 4; o > 0; o >>= 1)
#pragma unroll
            for (int mi = 0; mi < M; mi++) de[mi] += __shfl_xor_sync(0xffffffffu, de[mi], o);
        } else {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int mi = 0; mi < M; mi++) de[mi] += __shfl_xor_sync(0xffffffffu, de[mi], o);
        }
        int verdict[M];
#pragma unroll
        for (int mi = 0; mi < M; mi++) {
          de[mi] *= dN;
          verdict[mi] = 0;
          if (!kSplit) verdict[mi] = screen(warp + mi * BW, de[mi]);
        }
        CEMC_TICK(12);
        int sk0, sk1;
        changed_sites(warp + (M - 1) * BW, sk0, sk1);
#pragma unroll
        for (int mi = 0; mi < M; mi++) {
          const int b = warp + mi * BW;
          const uint32_t m = conflict_mask(b, gs[mi], sk0, sk1, gs2[mi]);
          if (lane == 0) put_scr(par, b, m, verdict[mi], de[mi]);
        }
        CEMC_TICK(13);
      }
    } else {
#pragma unroll 1
    for (int mi = 0; mi < M; mi++) {
      const int b = warp + mi * BW;             // warp w evaluates moves w, w + BW, ...
      if (is_obs || (b >= nb && !kAsync)) break;   // async protocol: fixed byte count per batch, evaluate anyway
      int gsx[2] = {-1, -1};      // this lane's gathered sites (conflict check): columns 0..31 (and the site itself)
      int verdict = 0;
      double dEs = 0.0;           // screen value N sum_i eci_i dcf_i of this warp's changed site(s)
      int gsy[2] = {-1, -1};      // kWide: columns 32..K-1 and the site itself
      const uint4 rec0 = s.ring[(int)((sdone + b) & 127) * 2];
      int site0, site1 = -1, new0, new1 = 0, old0, old1 = 0, slot0 = -1, slot1 = -1;
      if (replay) {                                     // recorded sites / new species
        site0 = (int)rec0.x; new0 = (int)rec0.z;
        old0 = s.occ[site0];
        if (kCanon) {
          site1 = (int)rec0.y; new1 = (int)rec0.w;
          old1 = site1 == site0 ? new0 : (int)s.occ[site1];
        }
      } else if (!kCanon) {
        site0 = t.active ? t.active[rec0.x] : (int)rec0.x;
        old0 = s.occ[site0];
        if (t.allowed_identity) {                       // sgc_montecarlo.py:70-75
          int rr = (int)__umulhi(rec0.y, (uint32_t)(S - 1)); rr += (rr >= old0);
          new0 = rr;
        } else {
          const int p = t.allowed_pos[old0];
          int rr;
          if (p >= 0) { rr = (int)__umulhi(rec0.y, (uint32_t)(n_allowed - 1)); rr += (rr >= p); }
          else rr = (int)__umulhi(rec0.y, (uint32_t)n_allowed);
          new0 = t.allowed[rr];
        }
      } else {
        slot0 = (int)rec0.x; slot1 = (int)rec0.y; new0 = (int)rec0.z; new1 = (int)rec0.w;
        site0 = s.list[slot0]; site1 = s.list[slot1];
        old0 = new1; old1 = new0;
      }
      double *Vb = s.V + lwarp * NJ * VS;
      if (lane == 0)                             // every warp commits from this record
        put_prop(par, b, make_int4(site0, site1, new0, new1), make_int4(old0, old1, slot0, slot1));
      if (kTab) {
        CEMC_TICK(5);
        // ---- table evaluation: neighbour occupations stay in registers (lane c = column c),
        // sub-cluster codes by shuffles, sums over sub-clusters from the product tables
        // this warp's changed site(s), picked with selects: with the site split the index is the CTA
        // rank, and a run-time index into a local array would put it into local memory
        int sites[NJE], olds[NJE], news[NJE];
#pragma unroll
        for (int je = 0; je < NJE; je++) {
          const bool second = kSplit ? (crank != 0) : (je != 0);
          sites[je] = second ? site1 : site0; olds[je] = second ? old1 : old0; news[je] = second ? new1 : new0;
        }
        uint32_t *cw = s.codes + lwarp * NJ * n_sub;
        int gsite[2] = {0, 0};                    // multi: symmetry group of the changed site(s)
#pragma unroll
        for (int je = 0; je < NJE; je++) {
          const int j = jb + je;
          if (multi) gsite[je] = __ldg(&t.symm_of_site[sites[je]]);
          int v = 0, v2 = 0;
          if (lane < K) {
            const int nbs = neighbour(sites[je], lane, my_shift);
            gsx[je] = nbs;
            v = s.occ[nbs];
            if (j && nbs == site0) v = new0;      // change 1 sees change 0 applied (:845-852)
          } else if (lane == K) gsx[je] = sites[je];
          if (kWide) {                            // 32 <= K <= 63: a second column per lane
            if (lane + 32 < K) {
              const int nbs = neighbour(sites[je], lane + 32, my_shift2);
              gsy[je] = nbs;
              v2 = s.occ[nbs];
              if (j && nbs == site0) v2 = new0;
            } else if (lane + 32 == K) gsy[je] = sites[je];
          }
          auto occ_of_col = [&](uint32_t col) -> int {
            const int a1 = __shfl_sync(0xffffffffu, v, (int)(col & 31u));
            if (!kWide) return a1;
            const int a2 = __shfl_sync(0xffffffffu, v2, (int)(col & 31u));
            return (col & 32u) ? a2 : a1;
          };
#pragma unroll
          for (int q = 0; q < 4; q++) {
            if (q < t_rounds) {
              uint32_t dx = tdx[q], dy = tdy[q];
              if (multi) {                        // this group's sub-cluster descriptors (L1-resident)
                const uint2 dg = __ldg(&tb.desc[(gsite[je] * t_rounds + q) * 32 + lane]);
                dx = dg.x; dy = dg.y;
              }
              const int va = occ_of_col(dx & 0xffu);
              const int vb = occ_of_col((dx >> 8) & 0xffu);
              const int vc = occ_of_col((dx >> 16) & 0xffu);
              const uint32_t rest = (uint32_t)va * (dy & 0xffu) + (uint32_t)vb * ((dy >> 8) & 0xffu) +
                                    (uint32_t)vc * ((dy >> 16) & 0xffu);
              const uint32_t wr = dy >> 24, nd = dx >> 24;            // nd: decorations per table row
              const uint32_t cO = (rest + (uint32_t)olds[je] * wr) * nd, cN = (rest + (uint32_t)news[je] * wr) * nd;
              uint32_t word = (cO << TSH) | (cN << (16 + TSH));
              if (dy == 0u) { const uint32_t z = (dx & 0xffffu) >> (3 - TSH); word = z | (z << 16); }   // padding: the table's zero row
              if (q * 32 + lane < n_sub) cw[je * n_sub + q * 32 + lane] = word;
            }
          }
        }
        __syncwarp();
        CEMC_TICK(6);
        // sums over the sub-clusters, reference order (:246-282) or 4-way interleaved (TREE);
        // lane = (ECI, decoration), both changed sites side by side
        double *db = s.diff + lwarp * NJ * max_tasks;
        if (multi) {
          // several symmetry groups: the task list of each changed site's own group, sub-clusters
          // summed one after the other in the stored order (the reference's order, :246-282)
#pragma unroll
          for (int je = 0; je < NJE; je++) {
            const int tb0 = __ldg(&t.task_base[gsite[je]]), ntk = __ldg(&t.task_base[gsite[je] + 1]) - tb0;
            for (int tk = lane; tk < ntk; tk += 32) {
              const int4 tt = s.ttask[tb0 + tk];
              const char *tbl = reinterpret_cast<const char *>(s.tab) + (tt.x >> (3 - TSH));
              const uint32_t *cp = cw + je * n_sub + tt.y;
              TR spO = (TR)0, spN = (TR)0;
              for (int m = 0; m < tt.z; m++) {
                const uint32_t w = cp[m];
                spO = add_rn(spO, *reinterpret_cast<const TR *>(tbl + (w & 0xffffu)));
                spN = add_rn(spN, *reinterpret_cast<const TR *>(tbl + (w >> 16)));
              }
              db[je * max_tasks + tk] = (double)sub_rn(spN, spO);     // :397
            }
          }
        } else
        for (int tk = lane; tk < n_tasks; tk += 32) {
          const int4 tt = s.ttask[tk];
          const char *tbl = reinterpret_cast<const char *>(s.tab) + (tt.x >> (3 - TSH));
          const uint32_t *cp = cw + tt.y;
          auto TV = [&](uint32_t off) { return *reinterpret_cast<const TR *>(tbl + off); };
          if (!kTree) {
            TR spO[NJE], spN[NJE];
#pragma unroll
            for (int j = 0; j < NJE; j++) { spO[j] = (TR)0; spN[j] = (TR)0; }
            uint4 w[NJE][2];                            // code words of the group being summed / the next one
#pragma unroll
            for (int j = 0; j < NJE; j++) {
              w[j][0] = *reinterpret_cast<const uint4 *>(cp + j * n_sub);
              w[j][1] = *reinterpret_cast<const uint4 *>(cp + j * n_sub + 4);
            }
#pragma unroll 1
            for (int m = 0; m < tt.z; m += 8) {         // M is padded to a multiple of 8 with zero entries
              TR vo[NJE][8], vn[NJE][8];
#pragma unroll
              for (int j = 0; j < NJE; j++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                  vo[j][4 * h + 0] = TV(w[j][h].x & 0xffffu); vn[j][4 * h + 0] = TV(w[j][h].x >> 16);
                  vo[j][4 * h + 1] = TV(w[j][h].y & 0xffffu); vn[j][4 * h + 1] = TV(w[j][h].y >> 16);
                  vo[j][4 * h + 2] = TV(w[j][h].z & 0xffffu); vn[j][4 * h + 2] = TV(w[j][h].z >> 16);
                  vo[j][4 * h + 3] = TV(w[j][h].w & 0xffffu); vn[j][4 * h + 3] = TV(w[j][h].w >> 16);
                }
              if (m + 8 < tt.z) {                       // the next group's code words, while this group's sums run
#pragma unroll
                for (int j = 0; j < NJE; j++) {
                  w[j][0] = *reinterpret_cast<const uint4 *>(cp + j * n_sub + m + 8);
                  w[j][1] = *reinterpret_cast<const uint4 *>(cp + j * n_sub + m + 12);
                }
              }
#pragma unroll
              for (int x = 0; x < 8; x++)
#pragma unroll
                for (int j = 0; j < NJE; j++) { spO[j] = add_rn(spO[j], vo[j][x]); spN[j] = add_rn(spN[j], vn[j][x]); }
            }
#pragma unroll
            for (int j = 0; j < NJE; j++) db[j * max_tasks + tk] = (double)sub_rn(spN[j], spO[j]);     // :397
          } else {
#pragma unroll
            for (int j = 0; j < NJE; j++) {
              TR o[4] = {(TR)0, (TR)0, (TR)0, (TR)0}, n[4] = {(TR)0, (TR)0, (TR)0, (TR)0};
#pragma unroll 1
              for (int m = 0; m < tt.z; m += 4) {
                const uint4 w = *reinterpret_cast<const uint4 *>(cp + j * n_sub + m);
                o[0] = add_rn(o[0], TV(w.x & 0xffffu)); n[0] = add_rn(n[0], TV(w.x >> 16));
                o[1] = add_rn(o[1], TV(w.y & 0xffffu)); n[1] = add_rn(n[1], TV(w.y >> 16));
                o[2] = add_rn(o[2], TV(w.z & 0xffffu)); n[2] = add_rn(n[2], TV(w.z >> 16));
                o[3] = add_rn(o[3], TV(w.w & 0xffffu)); n[3] = add_rn(n[3], TV(w.w >> 16));
              }
              db[j * max_tasks + tk] = (double)sub_rn(add_rn(add_rn(n[0], n[1]), add_rn(n[2], n[3])),
                                                 add_rn(add_rn(o[0], o[1]), add_rn(o[2], o[3])));
            }
          }
        }
        __syncwarp();
        CEMC_TICK(7);
        // per-ECI quotients (:393-402): lane i = ECI i, this warp's changed site(s)
        {
          double de = 0.0;                                    // screen only
          if (multi) {
#pragma unroll
            for (int e = 0; e < E; e++) {
              const int i = e * 32 + lane;
              double qsum = 0.0;
#pragma unroll
              for (int je = 0; je < NJE; je++) {
                double qj = 0.0;
                if (i < n_eci) {                              // this group's term of ECI i (:379-404)
                  const int4 f = __ldg(&t.fin_i[gsite[je] * n_eci + i]);
                  const double2 fd = __ldg(&t.fin_d[gsite[je] * n_eci + i]);
                  if (f.x == 1) {
                    qj = __ddiv_rn(__dsub_rn(s.bf[f.y * S + news[je]], s.bf[f.y * S + olds[je]]), dN);
                  } else if (f.x == 2) {
                    double num = 0.0;
                    for (int q = f.z; q < f.w; q++) num = __dadd_rn(num, db[je * max_tasks + q]);    // :397
                    qj = __ddiv_rn(__dmul_rn(num, fd.x), fd.y);                                       // :400-402
                  }
                }
                put_sq(par, b, jb + je, qj, e);
                qsum += qj;
              }
              de += eci_reg[e] * qsum;
            }
          } else
#pragma unroll
          for (int e = 0; e < E; e++) {
            double num[NJE];
#pragma unroll
            for (int je = 0; je < NJE; je++) num[je] = 0.0;
            if (f_kind[e] == 1) {                             // :366-371
#pragma unroll
              for (int je = 0; je < NJE; je++)
                num[je] = __dsub_rn(s.bf[f_d[e] * S + news[je]], s.bf[f_d[e] * S + olds[je]]);
            } else if (f_kind[e] == 2) {
              // sum over the decorations in the stored order (:397), four loads in flight at a time
              for (int q0 = 0; q0 < f_nd[e]; q0 += 4) {
                double v[NJE][4];
#pragma unroll
                for (int x = 0; x < 4; x++) {
                  const int q = (q0 + x < f_nd[e]) ? q0 + x : q0;          // (clamped: a valid address)
#pragma unroll
                  for (int je = 0; je < NJE; je++) v[je][x] = db[je * max_tasks + f_t0[e] + q];
                }
#pragma unroll
                for (int x = 0; x < 4; x++)
                  if (q0 + x < f_nd[e]) {
#pragma unroll
                    for (int je = 0; je < NJE; je++) num[je] = __dadd_rn(num[je], v[je][x]);
                  }
              }
#pragma unroll
              for (int je = 0; je < NJE; je++) num[je] = __dmul_rn(num[je], f_scale[e]);   // :400
            }
            double qsum = 0.0;
#pragma unroll
            for (int je = 0; je < NJE; je++) {
              const double qj = exact_div(num[je], f_den[e], f_rden[e]);         // :402
              put_sq(par, b, jb + je, qj, e);
              qsum += qj;
            }
            if (f_kind[e] > 0) de += eci_reg[e] * qsum;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) de += __shfl_xor_sync(0xffffffffu, de, o);
          dEs = de * dN;
          if (!kSplit) verdict = screen(b, dEs);
        }
      } else {
      // P1: gather
#pragma unroll
      for (int x = 0; x < NJ; x++) {
        if (lane < KP) {
          const int j = x, c = lane;
          double *Vj = Vb + j * VS;
          const int sj = j ? site1 : site0;
          if (c < K) {
            const int nbs = neighbour(sj, c, my_shift);
            gsx[x] = nbs;
            int v = s.occ[nbs];
            if (j && nbs == site0) v = new0;    // change 1 sees change 0 applied (:845-852)
            for (int d = 0; d < D; d++) Vj[d * KP + c] = s.bf[d * S + v];
          } else {
            gsx[x] = sj;
            const int oid = j ? old1 : old0, nid = j ? new1 : new0;
            for (int d = 0; d < D; d++) {
              Vj[RB + d] = s.bf[d * S + oid];
              Vj[RB + D + d] = s.bf[d * S + nid];
            }
          }
        }
      }
      __syncwarp();
      // P2a: products
      double *POb = s.PO + lwarp * NJ * max_slots, *PNb = s.PN + lwarp * NJ * max_slots;
      {
        // one product pair per sub-cluster; two items in flight per lane (ILP: the warp
        // has few siblings on its scheduler, so latency must be covered inside the warp)
        auto item_products = [&](int q, double &tO, double &tN, int &slot) {
          const int j = q >= n_items;
          const uint4 it = s.items[j ? q - n_items : q];
          const char *Vj = reinterpret_cast<const char *>(Vb + j * VS);
          const double f0 = *reinterpret_cast<const double *>(Vj + (it.x & 0xffffu));
          const double f1 = *reinterpret_cast<const double *>(Vj + (it.x >> 16));
          const double f2 = *reinterpret_cast<const double *>(Vj + (it.y & 0xffffu));
          const double f3 = *reinterpret_cast<const double *>(Vj + (it.y >> 16));
          const double fr = *reinterpret_cast<const double *>(Vj + it.w);
          slot = (int)(it.z & 0xffffu) + j * max_slots;
          // left-to-right products (:271-281) for the old and the new species of the
          // changed site; its position in the cluster is (mostly) warp-uniform
          tO = __dmul_rn(__dmul_rn(__dmul_rn(f0, f1), f2), f3);
          switch (it.z >> 16) {
            case 0: tN = __dmul_rn(__dmul_rn(__dmul_rn(fr, f1), f2), f3); break;
            case 1: tN = __dmul_rn(__dmul_rn(__dmul_rn(f0, fr), f2), f3); break;
            case 2: tN = __dmul_rn(__dmul_rn(__dmul_rn(f0, f1), fr), f3); break;
            default: tN = __dmul_rn(__dmul_rn(__dmul_rn(f0, f1), f2), fr); break;
          }
        };
        const int n_all = NJ * n_items;
        int q = lane;
        for (; q + 32 < n_all; q += 64) {
          double tO0, tN0, tO1, tN1;
          int sl0, sl1;
          item_products(q, tO0, tN0, sl0);
          item_products(q + 32, tO1, tN1, sl1);
          POb[sl0] = tO0; PNb[sl0] = tN0;
          POb[sl1] = tO1; PNb[sl1] = tN1;
        }
        if (q < n_all) {
          double tO0, tN0;
          int sl0;
          item_products(q, tO0, tN0, sl0);
          POb[sl0] = tO0; PNb[sl0] = tN0;
        }
      }
      __syncwarp();
      // P2b: sums
      double *db = s.diff + lwarp * NJ * max_tasks;
      for (int q = lane; q < NJ * n_tasks; q += 32) {
        const int j = q >= n_tasks;
        const int tk = j ? q - n_tasks : q;
        const int2 ts = s.task_sum[tk];
        const double *po = POb + ts.x + j * max_slots;
        const double *pn = PNb + ts.x + j * max_slots;
        double dv;
        if (!kTree) {
          double spO = 0.0, spN = 0.0;                       // :246, :282
          int m = 0;
          for (; m + 7 < ts.y; m += 8) {
            double o[8], n[8];
#pragma unroll
            for (int x = 0; x < 8; x++) { o[x] = po[m + x]; n[x] = pn[m + x]; }
#pragma unroll
            for (int x = 0; x < 8; x++) { spO = __dadd_rn(spO, o[x]); spN = __dadd_rn(spN, n[x]); }
          }
          for (; m < ts.y; m++) { spO = __dadd_rn(spO, po[m]); spN = __dadd_rn(spN, pn[m]); }
          dv = __dsub_rn(spN, spO);                          // :397
        } else {
          double o0 = 0.0, o1 = 0.0, o2 = 0.0, o3 = 0.0, n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
          int m = 0;
          for (; m + 3 < ts.y; m += 4) {
            o0 = __dadd_rn(o0, po[m]); o1 = __dadd_rn(o1, po[m + 1]);
            o2 = __dadd_rn(o2, po[m + 2]); o3 = __dadd_rn(o3, po[m + 3]);
            n0 = __dadd_rn(n0, pn[m]); n1 = __dadd_rn(n1, pn[m + 1]);
            n2 = __dadd_rn(n2, pn[m + 2]); n3 = __dadd_rn(n3, pn[m + 3]);
          }
          for (; m < ts.y; m++) { o0 = __dadd_rn(o0, po[m]); n0 = __dadd_rn(n0, pn[m]); }
          dv = __dsub_rn(__dadd_rn(__dadd_rn(n0, n1), __dadd_rn(n2, n3)),
                         __dadd_rn(__dadd_rn(o0, o1), __dadd_rn(o2, o3)));
        }
        db[j * max_tasks + tk] = dv;
      }
      __syncwarp();
      // P2c: per-ECI quotients (:393-402): lane i = ECI i, both changed sites
      {
        // state-independent energy change of this move, N * sum_i eci_i (q0_i + q1_i):
        // only used to SCREEN the Metropolis test (any summation order will do)
        double de = 0.0;
#pragma unroll
        for (int e = 0; e < E; e++) {
          double num0 = 0.0, num1 = 0.0;
          if (f_kind[e] == 1) {                               // :366-371
            num0 = __dsub_rn(Vb[RB + D + f_d[e]], Vb[RB + f_d[e]]);
            if (kCanon) num1 = __dsub_rn(Vb[VS + RB + D + f_d[e]], Vb[VS + RB + f_d[e]]);
          } else if (f_kind[e] == 2) {
            for (int q = 0; q < f_nd[e]; q++) {               // :397
              num0 = __dadd_rn(num0, db[f_t0[e] + q]);
              if (kCanon) num1 = __dadd_rn(num1, db[max_tasks + f_t0[e] + q]);
            }
            num0 = __dmul_rn(num0, f_scale[e]);               // :400
            num1 = __dmul_rn(num1, f_scale[e]);
          }
          const double qa = exact_div(num0, f_den[e], f_rden[e]);               // :402
          const double qb = kCanon ? exact_div(num1, f_den[e], f_rden[e]) : 0.0;
          put_sq(par, b, 0, qa, e);
          if (kCanon) put_sq(par, b, 1, qb, e);
          if (f_kind[e] > 0) de += eci_reg[e] * (qa + qb);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) de += __shfl_xor_sync(0xffffffffu, de, o);
        dEs = de * dN;
        if (!kSplit) verdict = screen(b, dEs);
      }
      }
      CEMC_TICK(12);
      {
        int sk0, sk1;
        changed_sites(b, sk0, sk1);
        const uint32_t m = conflict_mask(b, gsx, sk0, sk1, gsy);
        if (lane == 0) put_scr(par, b, m, verdict, dEs);
      }
      CEMC_TICK(13);
      if (M > 1) __syncwarp();                  // the warp's scratch is reused by its next move
    }
    }
#ifdef CEMC_WARP_TIMING
    const long long wt1 = clock64();
#endif
    if (kAsync) {
      // every warp arrives on this batch's mbarrier of its CTA (warp 0 also posts the byte count the
      // other CTA sends) and waits for the phase: local records written, remote records landed
      __syncwarp();
      uint64_t *S = s.mbar + 1 + par;
      if (lane == 0) {
        // observer warps: the token says this warp is done with the buffers of the previous batch
        // (its payload depends on what the warp read from them)
        if (is_obs) st_async_v2b32(r_tok, (uint32_t)bk_nd, (uint32_t)__double2loint(e_cur), par ? r_S1 : r_S0);
        if (lwarp == 0) mbar_expect_tx(S, (uint32_t)B * (kTxRec + (crank == 0 ? kTxSq : 0u)) + 8u);
        else mbar_arrive(S);
      }
      mbar_wait(S, (uint32_t)(kb >> 1) & 1u);
    } else csync();
#ifdef CEMC_PHASE_TIMING
    if (*reinterpret_cast<volatile int32_t *>(s.ctl + 7) == 0x7fffffff) tlast = 0;   // wait for the barrier release
    if (is_obs && crank == 0 && lane == 0) { tph[14] += (unsigned long long)(tob1 - tob0); tph[15] += (unsigned long long)(clock64() - tob1); }
#endif
    CEMC_TICK(1);
#ifdef CEMC_WARP_TIMING
    const long long wt2 = clock64();
#endif

    // ---- D: EVERY warp decides the moves strictly in order -----------------------------
    // The Metropolis outcome of a move depends on the chain state only through
    // E_new - E_cur = N sum_i eci_i dcf_i (+ rounding), i.e. on the move alone, as long
    // as none of its inputs was changed by an earlier accepted move.  So: (1) the evaluating
    // warp has screened its move against its threshold -kT ln u; (2) a loop-free pass (lane =
    // move, three ballots) applies the conflict masks in order; (3) the warp applies the
    // accepted changes to its CTA's copy of the state.  A move whose screen is inconclusive
    // (|dE - L| inside the band) is decided with the exact expression.  The exact bookkeeping
    // (CF vector, ordered energy dot, observer sums) follows in the observer warp of CTA 0.
    int2 ct_red = make_int2(0, 0);       // this warp's copy of the decision record {moves decided, accept mask}
    {
      bool t_acc, t_bdr;
      uint32_t cm_l;
      double u_l = 0.0;
      // the proposal lane b would commit (loaded up front: its latency overlaps the ballots)
      int4 pa = make_int4(0, 0, 0, 0), pb = pa;
      if (lane < nb) {
        pa = *reinterpret_cast<const int4 *>(s.prop + par * (BT * 8) + lane * 8);
        if (kCanon || kAsync) pb = *reinterpret_cast<const int4 *>(s.prop + par * (BT * 8) + lane * 8 + 4);
      }
      if (!kSplit) {
        const int2 sc = lane < nb ? s.scr[par * BT + lane] : make_int2(0, 0);
        t_acc = (sc.y & 1) != 0; t_bdr = (sc.y & 2) != 0;
        cm_l = (uint32_t)sc.x;
      } else {                             // site split: the two CTAs' partial records of every move
        const int l0 = lane < nb ? lane : 0;
        const int4 ra = s.rec[(par * 2) * BT + l0], rb = s.rec[(par * 2 + 1) * BT + l0];
        const float2 th = *reinterpret_cast<const float2 *>(&s.ring[(int)((sdone + l0) & 127) * 2 + 1].z);
        const double dE_l = __hiloint2double(ra.w, ra.z) + __hiloint2double(rb.w, rb.z);
        const double m_l = c_rel * fabs(dE_l);
        t_acc = lane < nb && (dE_l + m_l < (double)th.x);
        t_bdr = lane < nb && !t_acc && !(dE_l - m_l > (double)th.y);
        cm_l = lane < nb ? (uint32_t)(ra.x | rb.x) : 0u;
      }
      const uint32_t tmask = __ballot_sync(0xffffffffu, t_acc);
      const uint32_t bmask = __ballot_sync(0xffffffffu, t_bdr);
      uint32_t tm = tmask;
      if ((bmask & 1u) && kAsync && crank == 1) {
        // move 0 is inconclusive: CTA 0 settles it (it has the published CF vector) and mails the
        // outcome; bmask is the same in both CTAs, so both take this path in the same batches
        if (tid == 0) mbar_expect_tx(s.mbar + 3, 8u);
        mbar_wait(s.mbar + 3, phX);
        phX ^= 1u;
        if (*reinterpret_cast<volatile int32_t *>(s.ctl)) tm |= 1u; else tm &= ~1u;
      } else if (bmask & 1u) {
        // move 0 is inconclusive: exact path (rare) -- ordered dot (named_array.cpp:27-31)
        // and the reference expression (montecarlo.py:951-956) on the current state, which
        // the bookkeeper published after the previous batch
        const double *pubr = s.pub + par * (LW + 2);
        {
          const uint4 rec1 = s.ring[(int)(sdone & 127) * 2 + 1];
          u_l = __hiloint2double((int)rec1.y, (int)rec1.x);
        }
        const double e_old = pubr[LW];
        double p[E];
#pragma unroll
        for (int e = 0; e < E; e++) {
          double cn = pubr[e * 32 + lane];
          if (f_any[e]) {
            cn = __dadd_rn(cn, s.sq[par * (BT * 2 * LW) + e * 32 + lane]);
            if (kCanon) cn = __dadd_rn(cn, s.sq[par * (BT * 2 * LW) + LW + e * 32 + lane]);
          }
          p[e] = __dmul_rn(eci_reg[e], cn);
        }
        double e_new = 0.0;
#pragma unroll
        for (int e = 0; e < E; e++)          // ECI order: slot 0 of lanes 0..31, then slot 1, ...
          for (int i = 0; i < 32 && e * 32 + i < n_eci4; i++) e_new = __dadd_rn(e_new, __shfl_sync(0xffffffffu, p[e], i));
        e_new = __dmul_rn(e_new, dN);
        const double ub = __shfl_sync(0xffffffffu, u_l, 0);
        const bool acc0 = metropolis(e_new, e_old, ub, kT, rkT);
        if (acc0) tm |= 1u; else tm &= ~1u;
        if (kAsync && is_decider && lane == 0) st_async_v2b32(r_ctl, acc0 ? 1u : 0u, 0u, r_X);
      }
      // In-order semantics without a loop: all moves before the first invalid one are
      // decided by their screen bit, so move b is invalid iff one of the moves it
      // conflicts with (cm_l) is accepted among its predecessors.  The batch ends at the
      // first invalid move or at the first inconclusive move other than move 0.
      const uint32_t below = lane ? ((1u << lane) - 1u) : 0u;
      const bool stop_here = lane < nb && (((cm_l & tm & below) != 0u) || (lane > 0 && ((bmask >> lane) & 1u)));
      const uint32_t stops = __ballot_sync(0xffffffffu, stop_here);
      const int ndone = stops ? (__ffs(stops) - 1) : nb;
      const uint32_t accmask = tm & (ndone >= 32 ? 0xffffffffu : ((1u << ndone) - 1u));
      if (is_decider) n_acc += __popc(accmask);
      const bool my_acc = lane < ndone && ((accmask >> lane) & 1u);
      // commits of the decided moves: lane b applies move b (accepted moves of one
      // batch never share a site, so the order among them is irrelevant)
      if (lane < ndone) {
        if (my_acc) {
          // this CTA's copy of the state: every warp stores the same values to the same addresses
          s.occ[pa.x] = (int8_t)pa.z;
          if (kCanon) {                        // swap_move_index_tracker.py:39-59
            s.occ[pa.y] = (int8_t)pa.w;
            if (pb.z >= 0) { s.list[pb.z] = pa.y; s.list[pb.w] = pa.x; }             // replay: no list slots
          }
          if (kCanon && pb.z >= 0 && is_decider) { g_loc[pa.y] = pb.z - offs_of(offs, pa.w); g_loc[pa.x] = pb.w - offs_of(offs, pa.z); }
        }
      }
      // global-memory state: a warp only ever reads what its own lanes wrote (every warp commits every
      // accepted change itself; the __syncwarp below orders the lanes), so a block-level fence will do
      if (!kStateSmem) __threadfence_block();
      __syncwarp();                      // this warp's lanes see each other's commits
      ct_red = make_int2(ndone, (int)accmask);
      if (is_decider) CEMC_TICK(3);
#ifdef CEMC_PHASE_TIMING
      if (tid == 0) { tph[8] += 1; tph[9] += ndone; tph[10] += __popc(accmask); tph[11] += (stops ? 1 : 0); }
#endif
    }
#ifdef CEMC_PHASE_TIMING
    if (*reinterpret_cast<volatile int32_t *>(s.ctl + 7) == 0x7fffffff) tlast = 0;
#endif
    {
      const int2 ct = ct_red;
      bk_nd = ct.x; bk_am = (uint32_t)ct.y; bk_base = sdone;
      sdone += ct.x;
      if ((mflags & 4) && !(is_obs && crank == 0)) { to_ob -= ct.x; if (to_ob <= 0) to_ob += (int)a.obs_interval; }   // (the bookkeeper's copy lags, see bookkeep)
      par ^= 1;
      kb++;
    }
    CEMC_TICK(4);
#ifdef CEMC_WARP_TIMING
    { const long long wt3 = clock64(); wtW += wt1 - wt0; wtB += wt2 - wt1; wtD += wt3 - wt2; }
#endif
  }
#ifdef CEMC_WARP_TIMING
  if (lane == 0 && crank == 0 && lwarp < 8 && a.phase) {
    a.phase[(size_t)r * 24 + lwarp * 3] = (unsigned long long)wtW;
    a.phase[(size_t)r * 24 + lwarp * 3 + 1] = (unsigned long long)wtB;
    a.phase[(size_t)r * 24 + lwarp * 3 + 2] = (unsigned long long)wtD;
  }
#endif
  if (is_obs && crank == 0 && bk_nd > 0) bookkeep(bk_nd, bk_am, bk_base, par ^ 1);

#ifdef CEMC_PHASE_TIMING
  if (tid == 0 && crank == 0 && a.phase)
    for (int i = 0; i < 14; i++) a.phase[(size_t)r * 24 + i] = tph[i];
  if (is_obs && lane == 0 && crank == 0 && a.phase)
    for (int i = 14; i < 24; i++) a.phase[(size_t)r * 24 + i] = tph[i];
#endif
  // ---- write back ------------------------------------------------------------
  if (is_obs && crank == 0) {
    double *aw = st.acc + (size_t)r * acc_stride;
    if (lane == 0) { aw[0] = aE0; aw[1] = aE1; aw[2] = aE2; }
    if (my_singlet >= 0) { aw[3 + 3 * my_singlet] = aS0; aw[4 + 3 * my_singlet] = aS1; aw[5 + 3 * my_singlet] = aS2; }
#pragma unroll
    for (int e = 0; e < E; e++) if (e * 32 + lane < n_eci) st.cf[(size_t)r * n_eci + e * 32 + lane] = cf_reg[e];
    if (lane == 0) st.e_cur[r] = e_cur;
  }
  if (is_decider) {
    if (lane == 0) {
      st.step[r] = step0 + (unsigned long long)a.n_steps;
      st.accepted[r] += n_acc;
    }
  }
  if (kStateSmem && crank == 0) {
    for (int i = tid; i < N; i += nthr) g_occ[i] = s.occ[i];
    if (kCanon) for (int i = tid; i < N; i += nthr) g_list[i] = s.list[i];
  }
}

}  // namespace cemc
