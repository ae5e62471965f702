// cemc_kernels.cuh -- sm_100a kernels of the cluster-expansion Metropolis hot path.
//
// One CTA per replica (independent Markov chain).  Per-replica state lives in
// shared memory for the whole launch: int8 occupations, the running
// correlation-function (CF) vector, the per-species site lists of the
// canonical sampler.  The read-only cluster "program" (items, tasks, per-ECI
// normalisation) is staged once per CTA; the translation matrix stays in
// global memory and is read through the read-only path (L1/L2 resident,
// shared by all replicas).
//
// Work decomposition of one trial move (c = 1 or 2 changed sites):
//   P1 gather   one thread per (site, translation column): T row -> neighbour
//               occupation -> basis-function values V[d][col] in shared memory
//   P2a items   one thread per (site, ECI, decoration, sub-cluster): the
//               left-to-right product of spin_product_one_atom for the old
//               and the new species of the changed site
//   P2b sums    one thread per (site, ECI, decoration): sum over sub-clusters
//               in the reference's order (or 4-way interleaved in TREE mode)
//   P3  warp 0  per-ECI normalisation, CF increment, sequential energy dot,
//               Metropolis test, commit, observers.  With <= 32 ECIs the CF
//               vector, the ECIs and the observer sums live in registers of
//               warp 0 (one lane per ECI) and the ordered dot product runs
//               over warp shuffles.
// Proposals for 32 moves at a time come from one Philox4x32-10 call per lane
// of warp 0; the translation-matrix rows of those 32 moves are prefetched
// into shared memory in the same pass (state-independent for SGC, speculative
// and validated for canonical swaps).
//
// Arithmetic follows the reference's operation order exactly (SURVEY.md
// Appendix A; /root/reference/cpp/src/ce_updater.cpp:244-285, :313-406,
// named_array.cpp:25-33): every multiply/add is an explicit __dmul_rn /
// __dadd_rn so that nvcc can never contract them into FMAs -- the reference is
// compiled for baseline x86-64 and has none.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cemc {

// -DCEMC_PHASE_TIMING: per-phase clock64() accounting of warp 0 (debug builds only;
// scripts/phase_timing.py).  Slots: 0 refill 1 P0 2 P1 3 P2a 4 P2b 5 P3 6 end barrier
#ifdef CEMC_PHASE_TIMING
#define CEMC_TICK(slot)                                              \
  do {                                                               \
    const long long now_ = clock64();                                \
    if (tid == 0) tph[slot] += (unsigned long long)(now_ - tlast);   \
    tlast = now_;                                                    \
  } while (0)
#else
#define CEMC_TICK(slot) do { } while (0)
#endif

enum Mode : int { MODE_REPLAY = 0, MODE_SGC = 1, MODE_CANONICAL = 2 };

// item word: 4 x 12-bit indices into V (sorted cluster positions 0..3; unused
// positions point at the constant 1.0), bits 48-49 = position of the changed site,
// bits 50-63 = product slot.
#define CEMC_ITEM_BITS 12
#define CEMC_ITEM_MASK 0xfffu

struct DeviceTables {    // device pointers, shared by all replicas
  int N, S, D, K, KP, VS, n_eci, n_symm, n_singlets;
  int n_items_total, n_tasks_total, max_tasks, max_items, max_slots;
  const int32_t *trans;        // [N][K]
  const int32_t *symm_of_site; // [N]
  const double *bf;            // [D][S]
  const unsigned long long *items;  // [n_items_total]
  const uint4 *items4;         // [n_items_total] same items, pre-decoded byte offsets (batch kernel)
  const uint16_t *item_slot;   // [n_items_total] product slot (padded, per group)
  const int32_t *item_base;    // [n_symm+1]
  const int32_t *task_base;    // [n_symm+1]
  const int2 *task_sum;        // [n_tasks_total] {first slot, M}
  const int4 *fin_i;           // [n_symm][n_eci] {kind, d, t0, t1}; kind -1 = copy
  const double2 *fin_d;        // [n_symm][n_eci] {scale (:400), div (:402)}
  const int32_t *singlet_idx;  // [n_singlets]
  int uniform_group;           // 1: one symmetry group, no background sites
  int prefetch_rows;           // 1: T rows of the next 32 moves are staged in smem
  int allowed_identity;        // 1: every species may be inserted (allowed[i] == i)
  // SGC proposal support
  int n_active;                // non-background sites
  const int32_t *active;       // [n_active] or nullptr when n_active == N
  int n_allowed;
  const int8_t *allowed;       // [128] species ids the sampler may insert
  const int8_t *allowed_pos;   // [128] position of a species in `allowed`, -1 if absent
  // Translation-invariant lattice (verified against `trans` by cemc_create): site = (i L2 + j) L3 + k
  // and T(site, col) = ((i + di) mod L1, (j + dj) mod L2, (k + dk) mod L3): the neighbour of a
  // site is index arithmetic (0 bytes of table traffic) instead of a gather from L2.
  int lat_ok;
  uint32_t L1, L2, L3, L23;    // L23 = L2 * L3
  uint32_t lat_m23, lat_m3;    // ceil(2^32 / L23), ceil(2^32 / L3): x / d == umulhi(x, m) (checked for all sites)
  const uint32_t *col_shift;   // [K] di | dj << 10 | dk << 20
};

// T(site, col) of a translation-invariant lattice; `shift` = col_shift[col]
__device__ __forceinline__ int lattice_neighbour(const DeviceTables &t, int site, uint32_t shift) {
  const uint32_t i = __umulhi((uint32_t)site, t.lat_m23);
  const uint32_t rem = (uint32_t)site - i * t.L23;
  const uint32_t j = __umulhi(rem, t.lat_m3);
  const uint32_t k = rem - j * t.L3;
  uint32_t ni = i + (shift & 0x3ffu), nj = j + ((shift >> 10) & 0x3ffu), nk = k + (shift >> 20);
  ni -= ni >= t.L1 ? t.L1 : 0u;
  nj -= nj >= t.L2 ? t.L2 : 0u;
  nk -= nk >= t.L3 ? t.L3 : 0u;
  return (int)((ni * t.L2 + nj) * t.L3 + nk);
}

struct ReplicaState {    // device pointers, replica-major
  int8_t *occ;           // [R][N]
  double *cf;            // [R][n_eci]
  double *eci;           // [R][n_eci]
  double *e_cur;         // [R]
  double *kT;            // [R]
  double *acc;           // [R][acc_stride]
  double *ref;           // [R] Averager reference value
  unsigned long long *step;      // [R]
  unsigned long long *accepted;  // [R]
  int32_t *list;         // [R][N] canonical tracker, species-major
  int32_t *loc;          // [R][N]
  int32_t *off;          // [R][S+1]
  int32_t *status;       // [R] 0 ok, else error code
};

struct RunArgs {
  long long n_steps;
  unsigned long long seed;
  uint32_t replica_offset;   // global id of local replica r = replica_offset + r * replica_stride
  uint32_t replica_stride;   // (keys the Philox streams: sharding over GPUs never changes a chain)
  int force_accept;      // trial semantics: commit every step (CEUpdater::calculate)
  int observe;           // accumulate observers
  double screen_slack;   // multiplies the Metropolis screening band (testing; default 1)
  // replay inputs (device)
  const int32_t *rp_sites;   // [R][n][2]
  const int8_t *rp_news;     // [R][n][2]
  const double *rp_u;        // [R][n]
  // optional outputs / trace (device), may be null
  int32_t *tr_sites;
  int8_t *tr_news;
  double *tr_u;
  uint8_t *tr_acc;
  double *tr_e;
  long long tr_capacity;
  unsigned long long *phase;   // [R][24] per-phase cycle counts (CEMC_PHASE_TIMING builds), may be null
  const int32_t *order;        // [R] replica handled by CTA (cluster) i, or null = identity (load balance)
  // ---- device-side state observers (cemc_set_device_observers): every obs_interval steps the
  // warp that does the bookkeeping SNAPSHOTS the chain's state (CF vector, energy, occupations)
  // into a small per-replica ring; observer_fold_kernel, enqueued after the launch, folds the
  // snapshots into the observers' sums in order.  (Folding inside the Metropolis kernels cost the
  // bench kernels 11 registers and 12 % of their speed even with the observers off.)
  long long obs_interval;      // 0 = off
  long long obs_origin;        // observer step counter when this launch starts
  int obs_flags;               // CEMC_OBS_* bits
  int obs_ring;                // snapshots the ring holds (a launch never crosses more boundaries)
  unsigned long long *ob_n;    // [R] boundaries seen (snapshot k sits in slot k % obs_ring)
  double *ob_snap_cf;          // [R][obs_ring][n_eci]
  double *ob_snap_e;           // [R][obs_ring]
  int8_t *ob_snap_occ;         // [R][obs_ring][N] (only with CEMC_OBS_LOWEST / CEMC_OBS_SITE_ORDER)
};

enum : int { OBS_CF_SUMS = 1, OBS_LOWEST = 2, OBS_ENERGY = 4, OBS_SITE_ORDER = 8 };

// One observer boundary inside a Metropolis kernel, executed by ONE WARP while nobody changes the
// chain's state: snapshot `cf` (the CF vector after the boundary step), the energy `e` and, when an
// observer needs them, the occupations (shared or global memory) into slot k % obs_ring.
__device__ __forceinline__ void observer_snapshot(const RunArgs &a, int r, int lane, int n_eci, int N,
                                                  const double *cf, double e, const int8_t *occ) {
  const size_t slot = (size_t)r * a.obs_ring + (size_t)(a.ob_n[r] % (unsigned long long)a.obs_ring);
  __syncwarp();
  for (int i = lane; i < n_eci; i += 32) a.ob_snap_cf[slot * n_eci + i] = cf[i];
  if (a.obs_flags & (OBS_LOWEST | OBS_SITE_ORDER)) {
    int8_t *dst = a.ob_snap_occ + slot * N;
    for (int i = lane; i < N; i += 32) dst[i] = occ[i];
  }
  if (lane == 0) { a.ob_snap_e[slot] = e; a.ob_n[r] += 1; }
  __syncwarp();
}

// Folding of the snapshots into the observers' sums (one warp per replica, in boundary order).
// Reference semantics: PairCorrelationObserver (mc_observers.py:81-136: sum and sum of squares of
// the CFs), LowestEnergyStructure (:138-183: strictly lower energy replaces the stored state),
// EnergyEvolution / EnergyHistogram (:689-761: the energy of every call), SiteOrderParameter
// (:614-686: number of sites that differ from the initial configuration, sum and sum of squares).
struct ObserverSums {
  unsigned long long *folded;  // [R] boundaries folded so far
  double *cf_sum, *cf_sq;      // [R][n_eci]
  double *best;                // [R][1 + n_eci] lowest energy seen on a boundary, its CFs
  int8_t *best_occ;            // [R][N] its occupations
  double *e;                   // [R][capacity] ring of the energies (sample k at k % capacity)
  long long capacity;
  double *order;               // [R][2] sum / sum of squares of #sites differing from occ_ref
  const int8_t *occ_ref;       // [R][N]
};

static __global__ void observer_fold_kernel(RunArgs a, ObserverSums o, int n_eci, int N) {
  const int r = blockIdx.x, lane = threadIdx.x;
  const unsigned long long n = a.ob_n[r];
  for (unsigned long long k = o.folded[r]; k < n; k++) {
    const size_t slot = (size_t)r * a.obs_ring + (size_t)(k % (unsigned long long)a.obs_ring);
    const double *cf = a.ob_snap_cf + slot * n_eci;
    const double e = a.ob_snap_e[slot];
    const int8_t *occ = a.ob_snap_occ ? a.ob_snap_occ + slot * N : nullptr;
    if (a.obs_flags & OBS_CF_SUMS)
      for (int i = lane; i < n_eci; i += 32) {
        const double c = cf[i];
        double *ps = o.cf_sum + (size_t)r * n_eci + i, *pq = o.cf_sq + (size_t)r * n_eci + i;
        *ps = __dadd_rn(*ps, c);
        *pq = __dadd_rn(*pq, __dmul_rn(c, c));
      }
    if ((a.obs_flags & OBS_ENERGY) && lane == 0 && o.capacity > 0)      // ring: the host drains it in time
      o.e[(size_t)r * o.capacity + (size_t)(k % (unsigned long long)o.capacity)] = e;
    if (a.obs_flags & OBS_LOWEST) {
      double *best = o.best + (size_t)r * (1 + n_eci);
      const bool lower = e < best[0];                 // warp-uniform (same address, same value)
      __syncwarp();
      if (lower) {
        if (lane == 0) best[0] = e;
        for (int i = lane; i < n_eci; i += 32) best[1 + i] = cf[i];
        int8_t *dst = o.best_occ + (size_t)r * N;
        for (int i = lane; i < N; i += 32) dst[i] = occ[i];
      }
    }
    if (a.obs_flags & OBS_SITE_ORDER) {
      const int8_t *ref = o.occ_ref + (size_t)r * N;
      int cnt = 0;
      for (int i = lane; i < N; i += 32) cnt += (occ[i] != ref[i]) ? 1 : 0;
#pragma unroll
      for (int w = 16; w > 0; w >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, w);
      if (lane == 0) {
        const double c = (double)cnt;
        o.order[2 * (size_t)r] = __dadd_rn(o.order[2 * (size_t)r], c);
        o.order[2 * (size_t)r + 1] = __dadd_rn(o.order[2 * (size_t)r + 1], __dmul_rn(c, c));
      }
    }
    __syncwarp();
  }
  if (lane == 0) o.folded[r] = n;
}

// ---------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2,
                                              uint32_t &c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// Correctly rounded a / b from y = RN(1/b) with FMAs (Markstein 1990: with a
// correctly rounded reciprocal and a faithful quotient, q + (a - b q) y rounds
// to RN(a/b)).  Two refinement steps: the first makes the quotient faithful, the
// second makes it exact.  Bit-identical to IEEE division, ~5 dependent FMAs
// instead of the ~30-instruction DDIV sequence; tests/test_exact_div.py and the
// cemc_selftest_division entry point check it against true division.
__device__ __forceinline__ double exact_div(double a, double b, double y) {
  const double q0 = __dmul_rn(a, y);
  const double r0 = __fma_rn(-b, q0, a);
  const double q1 = __fma_rn(r0, y, q0);
  const double r1 = __fma_rn(-b, q1, a);
  const double q2 = __fma_rn(r1, y, q1);
  const double m = fabs(a);
  // zero is returned as is (+-0 / b == +-0); the IEEE routine only for operands
  // whose intermediate products could leave the normal range (never in practice)
  if (m != 0.0 && !(m > 1e-250 && m < 1e250)) return __ddiv_rn(a, b);
  return m == 0.0 ? a : q2;
}

// Keep a loop-invariant kernel parameter in a register: without this ptxas
// re-reads it from the constant bank (LDCU) inside the per-move loop, and with one
// warp per replica that latency is fully exposed.
__device__ __forceinline__ int pin_reg(int x) { asm volatile("" : "+r"(x)); return x; }

// exp(x) for -60 < x <= 0 to ~3e-6 relative: only used to pre-screen the
// Metropolis test; borderline cases fall through to the exact expression.
__device__ __forceinline__ double fast_exp_screen(double x) {
  const float t = (float)(x * 1.4426950408889634);
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  return (double)r;
}

__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP): the copy engine stages the read-only
// tables (and, when 16-byte aligned, the occupations / site lists of the replica) from HBM
// into shared memory; completion is signalled on an mbarrier by transaction bytes, so no
// thread spends issue slots on the staging loop.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bytes: multiple of 16; dst / src 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// ---- distributed shared memory without fences: st.async + mbarrier transaction bytes --------
// A remote store that signals the destination CTA's mbarrier with the bytes it wrote: the
// consumer waits for (arrivals, bytes) of the phase and then reads plain shared memory.  No
// barrier.cluster, no MEMBAR, no L1 invalidation on the critical path.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_async_f64(uint32_t dst, double v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               ::"r"(dst), "l"(__double_as_longlong(v)), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t dst, uint32_t v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(dst), "r"(v), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_async_v2b32(uint32_t dst, uint32_t a, uint32_t b, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
               ::"r"(dst), "r"(a), "r"(b), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_async_v4b32(uint32_t dst, int4 v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool tma_aligned(const void *p, size_t bytes) {
  return ((reinterpret_cast<uintptr_t>(p) | bytes) & 15u) == 0 && bytes > 0;
}

struct Smem {            // carved from dynamic shared memory
  double *cf, *cfn, *eci, *prod, *V, *PO, *PN, *diff, *bf, *acc;
  double2 *fin_d;
  unsigned long long *items;
  int4 *fin_i;
  int2 *task_sum;
  int32_t *item_base, *task_base, *singlet_idx, *off, *present;
  uint4 *rng;            // [32][2]: one 32-byte proposal record per move
  int32_t *tnb;          // [32][2][K] prefetched translation-matrix rows
  int32_t *spec;         // [32][2] sites the prefetched rows belong to
  int8_t *allowed, *allowed_pos;
  int32_t *list;         // or global
  int8_t *occ;           // or global
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline size_t at_least_1(int x) { return x > 0 ? (size_t)x : 1; }

// Walks the layout; with base == nullptr only the size is computed.
template <bool kStateSmem>
__host__ __device__ inline size_t smem_layout(Smem *s, unsigned char *base, const DeviceTables &t,
                                              int acc_stride, bool canonical, int8_t *g_occ,
                                              int32_t *g_list) {
  size_t o = 0;
#define CEMC_TAKE(field, type, count)                                   \
  do {                                                                  \
    o = align_up(o, sizeof(type) < 16 ? sizeof(type) : 16);             \
    if (s) s->field = reinterpret_cast<type *>(base + o);               \
    o += sizeof(type) * at_least_1((int)(count));                       \
  } while (0)
  CEMC_TAKE(fin_d, double2, t.n_symm * t.n_eci);
  CEMC_TAKE(cf, double, t.n_eci);
  CEMC_TAKE(cfn, double, t.n_eci);
  CEMC_TAKE(eci, double, t.n_eci);
  CEMC_TAKE(prod, double, t.n_eci);
  CEMC_TAKE(V, double, 2 * t.VS);
  CEMC_TAKE(PO, double, 2 * t.max_slots);
  CEMC_TAKE(PN, double, 2 * t.max_slots);
  CEMC_TAKE(diff, double, 2 * t.max_tasks);
  CEMC_TAKE(bf, double, t.D * t.S);
  CEMC_TAKE(acc, double, acc_stride);
  CEMC_TAKE(items, unsigned long long, t.n_items_total);
  CEMC_TAKE(fin_i, int4, t.n_symm * t.n_eci);
  CEMC_TAKE(task_sum, int2, t.n_tasks_total);
  CEMC_TAKE(item_base, int32_t, t.n_symm + 1);
  CEMC_TAKE(task_base, int32_t, t.n_symm + 1);
  CEMC_TAKE(singlet_idx, int32_t, t.n_singlets);
  CEMC_TAKE(off, int32_t, t.S + 1);
  CEMC_TAKE(present, int32_t, t.S + 1);
  CEMC_TAKE(rng, uint4, 32 * 2);
  CEMC_TAKE(tnb, int32_t, t.prefetch_rows ? 32 * 2 * t.K : 1);
  CEMC_TAKE(spec, int32_t, 64);
  CEMC_TAKE(allowed, int8_t, 128);
  CEMC_TAKE(allowed_pos, int8_t, 128);
  if (kStateSmem) {
    if (canonical) CEMC_TAKE(list, int32_t, t.N);
    else if (s) s->list = g_list;
    o = align_up(o, 16);
    if (s) s->occ = reinterpret_cast<int8_t *>(base + o);
    o += align_up((size_t)t.N, 16);
  } else if (s) {
    s->list = g_list;
    s->occ = g_occ;
  }
#undef CEMC_TAKE
  return align_up(o, 16);
}

inline size_t smem_bytes(const DeviceTables &t, int acc_stride, bool canonical, bool state_in_smem) {
  return state_in_smem ? smem_layout<true>(nullptr, nullptr, t, acc_stride, canonical, nullptr, nullptr)
                       : smem_layout<false>(nullptr, nullptr, t, acc_stride, canonical, nullptr, nullptr);
}

// u <= exp(-(e_new - e_cur)/kT), decided without the division and the libm exp
// whenever a 1e-5-wide screen around a fast exp() already settles it
// (cemc/mcmc/montecarlo.py:951-956).
__device__ __forceinline__ bool metropolis(double e_new, double e_cur, double u, double kT,
                                           double rkT) {
  if (e_new < e_cur) return true;                               // :951, no uniform consumed
  const double diff = __dsub_rn(e_new, e_cur);                  // :954
  const double xs = -diff * rkT;
  if (xs < -40.0) return u == 0.0;             // exp(x) < 2^-53: only u == 0 passes
  const double ps = fast_exp_screen(xs);
  if (u < ps * (1.0 - 1e-5)) return true;
  if (u > ps * (1.0 + 1e-5)) return false;
  return u <= exp(exact_div(-diff, kT, rkT));                   // :955-956
}

template <int MODE, bool kStateSmem, bool kTree, bool kFast>
__global__ void __launch_bounds__(256, 1)
mc_kernel(DeviceTables t, ReplicaState st, RunArgs a, int acc_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int r = blockIdx.x;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int N = t.N, K = t.K, KP = t.KP, S = t.S, D = t.D, VS = t.VS, n_eci = t.n_eci;
  const int RB = D * KP;                       // first "changed site" slot of V
  constexpr bool kCanon = (MODE == MODE_CANONICAL);
  constexpr bool kSgc = (MODE == MODE_SGC);
  const bool pref = (MODE != MODE_REPLAY) && t.prefetch_rows;

  int8_t *g_occ = st.occ + (size_t)r * N;
  int32_t *g_list = st.list ? st.list + (size_t)r * N : nullptr;
  int32_t *g_loc = st.loc ? st.loc + (size_t)r * N : nullptr;
  Smem s;
  smem_layout<kStateSmem>(&s, smem_raw, t, acc_stride, kCanon, g_occ, g_list);

  // ---- stage per-replica state and the cluster program -------------------
  for (int i = tid; i < n_eci; i += nthr) {
    s.cf[i] = st.cf[(size_t)r * n_eci + i];
    s.eci[i] = st.eci[(size_t)r * n_eci + i];
  }
  for (int i = tid; i < D * S; i += nthr) s.bf[i] = t.bf[i];
  for (int i = tid; i < acc_stride; i += nthr) s.acc[i] = st.acc[(size_t)r * acc_stride + i];
  for (int i = tid; i < t.n_symm * n_eci; i += nthr) { s.fin_i[i] = t.fin_i[i]; s.fin_d[i] = t.fin_d[i]; }
  for (int i = tid; i < t.n_items_total; i += nthr) s.items[i] = t.items[i];
  for (int i = tid; i < t.n_tasks_total; i += nthr) s.task_sum[i] = t.task_sum[i];
  for (int i = tid; i <= t.n_symm; i += nthr) { s.item_base[i] = t.item_base[i]; s.task_base[i] = t.task_base[i]; }
  for (int i = tid; i < t.n_singlets; i += nthr) s.singlet_idx[i] = t.singlet_idx[i];
  for (int i = tid; i < 128; i += nthr) { s.allowed[i] = t.allowed[i]; s.allowed_pos[i] = t.allowed_pos[i]; }
  for (int i = tid; i < 2 * VS; i += nthr) s.V[i] = 1.0;       // V[K] is the constant 1.0
  if (kStateSmem) {
    const int nvec = N / 16;                   // 16-byte vectorised copy of the int8 occupations
    if ((reinterpret_cast<size_t>(g_occ) & 15) == 0) {
      const int4 *src = reinterpret_cast<const int4 *>(g_occ);
      int4 *dst = reinterpret_cast<int4 *>(s.occ);
      for (int i = tid; i < nvec; i += nthr) dst[i] = src[i];
      for (int i = nvec * 16 + tid; i < N; i += nthr) s.occ[i] = g_occ[i];
    } else {
      for (int i = tid; i < N; i += nthr) s.occ[i] = g_occ[i];
    }
  }
  int n_present = 0;
  if (kCanon) {
    for (int i = tid; i <= S; i += nthr) s.off[i] = st.off[(size_t)r * (S + 1) + i];
    if (kStateSmem)
      for (int i = tid; i < N; i += nthr) s.list[i] = g_list[i];
  }
  __syncthreads();
  if (kCanon) {
    if (tid == 0) {
      int np = 0;
      for (int sp = 0; sp < S; sp++)
        if (s.off[sp + 1] > s.off[sp]) s.present[np++] = sp;
      s.present[S] = np;
    }
    __syncthreads();
    n_present = s.present[S];
    if (n_present < 2) {                       // TooFewElementsError, montecarlo.py:310
      if (tid == 0) st.status[r] = 2;
      return;
    }
  }

  double e_cur = st.e_cur[r];
  const double kT = st.kT[r];
  const double rkT = __ddiv_rn(1.0, kT);
  const double ref = st.ref[r];
  const double rref = __ddiv_rn(1.0, ref);
  const double dN = (double)(unsigned)N;
  const double rN = __ddiv_rn(1.0, dN);
  const unsigned long long step0 = st.step[r];
  unsigned long long n_acc = 0;
  const uint32_t rep_global = a.replica_offset + (uint32_t)r * a.replica_stride;
  const int n_allowed = t.n_allowed;
  int err = 0;

  // ---- kFast: lane i of warp 0 owns ECI i (registers for the whole launch) --
  int f_kind = 0, f_d = 0, f_t0 = 0, f_t1 = 0, my_singlet = -1;
  double f_scale = 0.0, f_den = 1.0, f_rden = 1.0, eci_reg = 0.0, cf_reg = 0.0;
  double aE0 = 0.0, aE1 = 0.0, aE2 = 0.0, aS0 = 0.0, aS1 = 0.0, aS2 = 0.0;
  if (kFast && warp == 0) {
    if (lane < n_eci) {
      const int4 f = s.fin_i[lane];
      f_kind = f.x; f_d = f.y; f_t0 = f.z; f_t1 = f.w;
      const double2 fd = s.fin_d[lane];
      f_scale = fd.x;
      f_den = (f_kind == 1) ? dN : fd.y;
      f_rden = __ddiv_rn(1.0, f_den);
      eci_reg = s.eci[lane];
      cf_reg = s.cf[lane];
      for (int d = 0; d < t.n_singlets; d++) if (s.singlet_idx[d] == lane) my_singlet = d;
      if (my_singlet >= 0) {
        aS0 = s.acc[3 + 3 * my_singlet]; aS1 = s.acc[4 + 3 * my_singlet]; aS2 = s.acc[5 + 3 * my_singlet];
      }
    }
    aE0 = s.acc[0]; aE1 = s.acc[1]; aE2 = s.acc[2];
  }

#ifdef CEMC_PHASE_TIMING
  unsigned long long tph[16] = {0};
  long long tlast = clock64();
#endif
  for (long long it0 = 0; it0 < a.n_steps && !err; it0 += 32) {
    const int nblk = (int)((a.n_steps - it0) < 32 ? (a.n_steps - it0) : 32);
    if (MODE != MODE_REPLAY) {
      // ---- refill: Philox proposals (+ translation-matrix rows) for 32 moves
      if (warp == 0) {
        const unsigned long long stp = step0 + (unsigned long long)it0 + lane;
        uint32_t c0 = (uint32_t)stp, c1 = (uint32_t)(stp >> 32), c2 = rep_global, c3 = 0;
        philox4x32_10(c0, c1, c2, c3, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        int sp0, sp1 = -1;
        uint4 rec0, rec1;
        if (kSgc) {
          // sgc_montecarlo.py:69: site uniform over the active sites
          const uint32_t ia = __umulhi(c0, (uint32_t)t.n_active);
          sp0 = t.active ? t.active[ia] : (int)ia;
          const double u = u53(c2, c3);
          rec0 = make_uint4((uint32_t)sp0, c1, 0u, 0u);
          rec1 = make_uint4((uint32_t)__double2loint(u), (uint32_t)__double2hiint(u), 0u, 0u);
        } else {
          // montecarlo.py:899-907: species pair uniform (a != b), slot uniform per species
          uint32_t d0 = (uint32_t)stp, d1 = (uint32_t)(stp >> 32), d2 = rep_global, d3 = 1;
          philox4x32_10(d0, d1, d2, d3, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
          const int ia = (int)__umulhi(c0, (uint32_t)n_present);
          int ib = (int)__umulhi(c1, (uint32_t)(n_present - 1)); ib += (ib >= ia);
          const int sa = s.present[ia], sb = s.present[ib];
          const int slot0 = s.off[sa] + (int)__umulhi(c2, (uint32_t)(s.off[sa + 1] - s.off[sa]));
          const int slot1 = s.off[sb] + (int)__umulhi(c3, (uint32_t)(s.off[sb + 1] - s.off[sb]));
          const double u = u53(d0, d1);
          rec0 = make_uint4((uint32_t)slot0, (uint32_t)slot1, (uint32_t)sb, (uint32_t)sa);
          rec1 = make_uint4((uint32_t)__double2loint(u), (uint32_t)__double2hiint(u), 0u, 0u);
          sp0 = s.list[slot0]; sp1 = s.list[slot1];     // speculative: validated at use
        }
        s.rng[lane * 2] = rec0; s.rng[lane * 2 + 1] = rec1;
        s.spec[lane * 2] = sp0; s.spec[lane * 2 + 1] = sp1;
        if (pref) {
          int32_t *dst = s.tnb + lane * 2 * K;
          const int32_t *src0 = t.trans + (size_t)sp0 * K;
          for (int c = 0; c < K; c++) dst[c] = __ldg(src0 + c);
          if (kCanon) {
            const int32_t *src1 = t.trans + (size_t)sp1 * K;
            for (int c = 0; c < K; c++) dst[K + c] = __ldg(src1 + c);
          }
        }
      }
      __syncthreads();
    }
    CEMC_TICK(0);

    for (int ib_ = 0; ib_ < nblk; ib_++) {
      const long long it = it0 + ib_;
      // ---- P0: proposal (every thread, redundantly) ------------------------
      int site0, site1 = -1, new0, new1 = 0, old0, old1 = 0, slot0 = 0, slot1 = 0;
      bool ch0 = true, ch1 = false, pf0 = pref, pf1 = pref;
      double u;
      if (MODE == MODE_REPLAY) {
        const size_t q = (size_t)r * a.n_steps + it;
        site0 = a.rp_sites[2 * q]; site1 = a.rp_sites[2 * q + 1];
        new0 = a.rp_news[2 * q]; new1 = a.rp_news[2 * q + 1];
        u = a.rp_u[q];
        if (site0 < 0 || site0 >= N || site1 >= N || new0 < 0 || new0 >= S ||
            (site1 >= 0 && (new1 < 0 || new1 >= S))) { err = 3; break; }
        old0 = s.occ[site0];
        old1 = site1 >= 0 ? (site1 == site0 ? new0 : (int)s.occ[site1]) : 0;
        ch0 = (old0 != new0);                              // ce_updater.cpp:315
        ch1 = (site1 >= 0) && (old1 != new1);
      } else {
        const uint4 rec0 = s.rng[ib_ * 2], rec1 = s.rng[ib_ * 2 + 1];
        u = __hiloint2double((int)rec1.y, (int)rec1.x);
        if (kSgc) {
          // sgc_montecarlo.py:70-75: new species uniform among the others
          site0 = (int)rec0.x;
          old0 = s.occ[site0];
          if (t.allowed_identity) {
            int rr = (int)__umulhi(rec0.y, (uint32_t)(S - 1)); rr += (rr >= old0);
            new0 = rr;
          } else {
            const int p = s.allowed_pos[old0];
            int rr;
            if (p >= 0) { rr = (int)__umulhi(rec0.y, (uint32_t)(n_allowed - 1)); rr += (rr >= p); }
            else rr = (int)__umulhi(rec0.y, (uint32_t)n_allowed);
            new0 = s.allowed[rr];
          }
        } else {
          slot0 = (int)rec0.x; slot1 = (int)rec0.y; new0 = (int)rec0.z; new1 = (int)rec0.w;
          site0 = s.list[slot0]; site1 = s.list[slot1];
          old0 = new1; old1 = new0;            // lists are species-consistent
          ch1 = true;
          pf0 = pref && (site0 == s.spec[ib_ * 2]);
          pf1 = pref && (site1 == s.spec[ib_ * 2 + 1]);
        }
      }
      int g0 = 0, g1 = 0;
      if (!t.uniform_group) {
        g0 = t.symm_of_site[site0];
        g1 = site1 >= 0 ? t.symm_of_site[site1] : 0;
        if ((ch0 && g0 < 0) || (ch1 && g1 < 0)) { err = 1; break; }   // :330 background atom
      }

      CEMC_TICK(1);
      // ---- P1: gather neighbour occupations -> basis-function values -------
      for (int q = tid; q < 2 * KP; q += nthr) {
        const int j = q >= KP, c = j ? q - KP : q;
        if (!(j ? ch1 : ch0)) continue;
        double *Vj = s.V + j * VS;
        if (c < K) {
          const int sj = j ? site1 : site0;
          const int nb = (j ? pf1 : pf0) ? s.tnb[(ib_ * 2 + j) * K + c]
                                         : __ldg(&t.trans[(size_t)sj * K + c]);   // :264
          int v = s.occ[nb];
          if (j && nb == site0) v = new0;      // change 1 sees change 0 applied (:845-852)
          for (int d = 0; d < D; d++) Vj[d * KP + c] = s.bf[d * S + v];
        } else {                               // the changed site: old and new (:273-276)
          const int oid = j ? old1 : old0, nid = j ? new1 : new0;
          for (int d = 0; d < D; d++) {
            Vj[RB + d] = s.bf[d * S + oid];
            Vj[RB + D + d] = s.bf[d * S + nid];
          }
        }
      }
      __syncthreads();
      CEMC_TICK(2);

      // ---- P2a: one product per (site, task, sub-cluster), old and new -----
      const int ib0 = s.item_base[g0], ib1 = s.item_base[g1];
      const int ni0 = ch0 ? (s.item_base[g0 + 1] - ib0) : 0;
      const int ni1 = ch1 ? (s.item_base[g1 + 1] - ib1) : 0;
      for (int q = tid; q < ni0 + ni1; q += nthr) {
        const int j = q >= ni0;
        const unsigned long long w = s.items[j ? ib1 + (q - ni0) : ib0 + q];
        const double *Vj = s.V + j * VS;
        const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
        const int i0 = lo & CEMC_ITEM_MASK, i1 = (lo >> 12) & CEMC_ITEM_MASK;
        const int i2 = (uint32_t)(w >> 24) & CEMC_ITEM_MASK, i3 = (hi >> 4) & CEMC_ITEM_MASK;
        const int kref = (hi >> 16) & 3;
        const int slot = (int)(hi >> 18) + j * t.max_slots;
        const double f0 = Vj[i0], f1 = Vj[i1], f2 = Vj[i2], f3 = Vj[i3];
        const int iref = kref == 0 ? i0 : kref == 1 ? i1 : kref == 2 ? i2 : i3;
        const double fr = Vj[iref + D];
        // left-to-right product (:271-281); 1.0 * f == f and f * 1.0 == f exactly
        const double tO = __dmul_rn(__dmul_rn(__dmul_rn(f0, f1), f2), f3);
        const double tN = __dmul_rn(__dmul_rn(__dmul_rn(kref == 0 ? fr : f0, kref == 1 ? fr : f1),
                                              kref == 2 ? fr : f2), kref == 3 ? fr : f3);
        s.PO[slot] = tO;
        s.PN[slot] = tN;
      }
      __syncthreads();
      CEMC_TICK(3);

      // ---- P2b: sum over sub-clusters per (site, task) ----------------------
      const int tb0 = s.task_base[g0], tb1 = s.task_base[g1];
      const int nt0 = ch0 ? (s.task_base[g0 + 1] - tb0) : 0;
      const int nt1 = ch1 ? (s.task_base[g1 + 1] - tb1) : 0;
      for (int q = tid; q < nt0 + nt1; q += nthr) {
        const int j = q >= nt0;
        const int tk = j ? q - nt0 : q;
        const int2 ts = s.task_sum[(j ? tb1 : tb0) + tk];
        const double *po = s.PO + ts.x + j * t.max_slots;
        const double *pn = s.PN + ts.x + j * t.max_slots;
        double dv;
        if (!kTree) {
          double spO = 0.0, spN = 0.0;                       // :246, :282
#pragma unroll 4
          for (int m = 0; m < ts.y; m++) { spO = __dadd_rn(spO, po[m]); spN = __dadd_rn(spN, pn[m]); }
          dv = __dsub_rn(spN, spO);                          // :397
        } else {
          // order-free variant: exact whenever every product is an integer
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
        s.diff[j * t.max_tasks + tk] = dv;
      }
      __syncthreads();
      CEMC_TICK(4);

      // ---- P3 (warp 0): per-ECI increments, energy, Metropolis, commit ------
      if (warp == 0) {
        double e_new;
        double c = 0.0;
        if (kFast) {
          // lane i: numerators of both sites first, then both exact divisions
          c = cf_reg;
          double num0 = 0.0, num1 = 0.0;
          if (f_kind == 1) {                                  // :366-371
            if (ch0) num0 = __dsub_rn(s.bf[f_d * S + new0], s.bf[f_d * S + old0]);
            if (ch1) num1 = __dsub_rn(s.bf[f_d * S + new1], s.bf[f_d * S + old1]);
          } else if (f_kind == 2) {
            if (ch0) {
              for (int q = f_t0; q < f_t1; q++) num0 = __dadd_rn(num0, s.diff[q]);                 // :397
              num0 = __dmul_rn(num0, f_scale);                                                    // :400
            }
            if (ch1) {
              for (int q = f_t0; q < f_t1; q++) num1 = __dadd_rn(num1, s.diff[t.max_tasks + q]);
              num1 = __dmul_rn(num1, f_scale);
            }
          }
          if (f_kind > 0) {                                   // kinds 0 / -1: copied (:360,:382)
            if (ch0) c = __dadd_rn(c, exact_div(num0, f_den, f_rden));                            // :402-404
            if (ch1) c = __dadd_rn(c, exact_div(num1, f_den, f_rden));
          }
          const double p = __dmul_rn(eci_reg, c);             // 0 for lanes >= n_eci
          e_new = 0.0;                                        // named_array.cpp:27-31, in order
          for (int i = 0; i < n_eci; i++) e_new = __dadd_rn(e_new, __shfl_sync(0xffffffffu, p, i));
        } else {
          for (int i = lane; i < n_eci; i += 32) {
            double cc = s.cf[i];
#pragma unroll
            for (int j = 0; j < 2; j++) {
              if (!(j ? ch1 : ch0)) continue;
              const int fi = (j ? g1 : g0) * n_eci + i;
              const int4 f = s.fin_i[fi];
              const int oid = j ? old1 : old0, nid = j ? new1 : new0;
              if (f.x == 1) {                                       // :366-371
                const double dl = exact_div(__dsub_rn(s.bf[f.y * S + nid], s.bf[f.y * S + oid]), dN, rN);
                cc = __dadd_rn(cc, dl);
              } else if (f.x == 2) {
                const double2 fd = s.fin_d[fi];
                double delta = 0.0;
                const double *df = s.diff + j * t.max_tasks;
                for (int q = f.z; q < f.w; q++) delta = __dadd_rn(delta, df[q]);     // :397
                delta = __dmul_rn(delta, fd.x);                     // :400
                delta = __ddiv_rn(delta, fd.y);                     // :402
                cc = __dadd_rn(cc, delta);                          // :404
              }                                                     // else: copied (:360,:382)
            }
            s.cfn[i] = cc;
            s.prod[i] = __dmul_rn(s.eci[i], cc);
          }
          __syncwarp();
          e_new = 0.0;                                              // named_array.cpp:27-31
          for (int i = 0; i < n_eci; i++) e_new = __dadd_rn(e_new, s.prod[i]);
        }
        e_new = __dmul_rn(e_new, dN);                               // ce_updater.cpp:241
        const bool accept = a.force_accept ? true : metropolis(e_new, e_cur, u, kT, rkT);
        if (accept) {
          if (kFast) cf_reg = c;
          else for (int i = lane; i < n_eci; i += 32) s.cf[i] = s.cfn[i];
          e_cur = e_new;
          n_acc++;
          if (lane == 0) {
            if (ch0) s.occ[site0] = (int8_t)new0;
            if (ch1) s.occ[site1] = (int8_t)new1;
            if (kCanon) {                      // swap_move_index_tracker.py:39-59
              s.list[slot0] = site1; s.list[slot1] = site0;
              g_loc[site1] = slot0 - s.off[new1]; g_loc[site0] = slot1 - s.off[new0];
            }
          }
        }
        if (a.observe) {                                            // montecarlo.py:811-814,
          if (kFast) {                                              // mc_observers.py:264-270
            const double e2 = __dmul_rn(e_cur, e_cur);
            aE0 = __dadd_rn(aE0, 1.0);
            aE1 = __dadd_rn(aE1, ref == 1.0 ? e_cur : exact_div(e_cur, ref, rref));
            aE2 = __dadd_rn(aE2, ref == 1.0 ? e2 : exact_div(e2, ref, rref));
            aS0 = __dadd_rn(aS0, cf_reg);
            aS1 = __dadd_rn(aS1, __dmul_rn(cf_reg, cf_reg));
            aS2 = __dadd_rn(aS2, __dmul_rn(cf_reg, e_cur));
          } else {
            __syncwarp();
            if (lane == 0) {
              s.acc[0] = __dadd_rn(s.acc[0], 1.0);
              const double e2 = __dmul_rn(e_cur, e_cur);
              s.acc[1] = __dadd_rn(s.acc[1], ref == 1.0 ? e_cur : exact_div(e_cur, ref, rref));
              s.acc[2] = __dadd_rn(s.acc[2], ref == 1.0 ? e2 : exact_div(e2, ref, rref));
            }
            for (int d = lane; d < t.n_singlets; d += 32) {
              const double sv = s.cf[s.singlet_idx[d]];
              double *ad = s.acc + 3 + 3 * d;
              ad[0] = __dadd_rn(ad[0], sv);
              ad[1] = __dadd_rn(ad[1], __dmul_rn(sv, sv));
              ad[2] = __dadd_rn(ad[2], __dmul_rn(sv, e_cur));
            }
          }
        }
        if (a.obs_interval > 0 && ((a.obs_origin + it + 1) % a.obs_interval) == 0) {
          // state observers on their boundary (warp 0 is the only writer of the chain's state)
          if (kFast) { __syncwarp(); if (lane < n_eci) s.cf[lane] = cf_reg; }
          __syncwarp();
          observer_snapshot(a, r, lane, n_eci, N, s.cf, e_cur, s.occ);
        }
        if (lane == 0 && (a.tr_acc || a.tr_e) && it < a.tr_capacity) {
          const size_t q = (size_t)r * a.tr_capacity + it;
          if (a.tr_sites) { a.tr_sites[2 * q] = site0; a.tr_sites[2 * q + 1] = site1; }
          if (a.tr_news) { a.tr_news[2 * q] = (int8_t)new0; a.tr_news[2 * q + 1] = (int8_t)new1; }
          if (a.tr_u) a.tr_u[q] = u;
          if (a.tr_acc) a.tr_acc[q] = accept ? 1 : 0;
          if (a.tr_e) a.tr_e[q] = e_cur;
        }
      }
      CEMC_TICK(5);
      __syncthreads();
      CEMC_TICK(6);
    }
  }
#ifdef CEMC_PHASE_TIMING
  if (tid == 0 && a.phase)
    for (int i = 0; i < 16; i++) a.phase[(size_t)r * 24 + i] = tph[i];
#endif

  // ---- write back --------------------------------------------------------
  __syncthreads();
  if (kFast && warp == 0) {
    if (lane < n_eci) s.cf[lane] = cf_reg;
    if (lane == 0) { s.acc[0] = aE0; s.acc[1] = aE1; s.acc[2] = aE2; }
    if (my_singlet >= 0) {
      s.acc[3 + 3 * my_singlet] = aS0; s.acc[4 + 3 * my_singlet] = aS1; s.acc[5 + 3 * my_singlet] = aS2;
    }
  }
  __syncthreads();
  if (err) { if (tid == 0) st.status[r] = err; }
  for (int i = tid; i < n_eci; i += nthr) st.cf[(size_t)r * n_eci + i] = s.cf[i];
  for (int i = tid; i < acc_stride; i += nthr) st.acc[(size_t)r * acc_stride + i] = s.acc[i];
  if (kStateSmem) {
    for (int i = tid; i < N; i += nthr) g_occ[i] = s.occ[i];
    if (kCanon) for (int i = tid; i < N; i += nthr) g_list[i] = s.list[i];
  }
  if (tid == 0) {
    st.e_cur[r] = e_cur;
    if (!a.force_accept) {                 // trial changes are not Monte Carlo steps
      st.step[r] = step0 + (unsigned long long)(err ? 0 : a.n_steps);
      st.accepted[r] += n_acc;
    }
  }
}

// cemc_selftest_division: exact_div against IEEE division on random operands
static __global__ void exact_div_selftest_kernel(unsigned long long seed, int iters, const double *dens,
                                          int n_dens, unsigned long long *mismatches) {
  const unsigned long long gid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long bad = 0;
  for (int k = 0; k < iters; k++) {
    uint32_t c0 = (uint32_t)gid, c1 = (uint32_t)(gid >> 32), c2 = (uint32_t)k, c3 = 7;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    // numerator: random mantissa, exponent in [-40, 40]; denominator: table or random
    const double m = 1.0 + u53(c0, c1);
    double aa = ldexp(m, (int)(c2 % 81u) - 40);
    if (c3 & 1u) aa = -aa;
    double b;
    if (n_dens > 0 && (c3 & 2u)) b = dens[(c3 >> 2) % (uint32_t)n_dens];
    else b = ldexp(1.0 + u53(c2, c3), (int)((c3 >> 8) % 61u) - 30);
    if ((c3 & 12u) == 12u) aa = __dmul_rn(aa, b);    // near-exact quotients
    const double y = __ddiv_rn(1.0, b);
    if (exact_div(aa, b, y) != __ddiv_rn(aa, b)) bad++;
  }
  if (bad) atomicAdd(mismatches, bad);
}

}  // namespace cemc
