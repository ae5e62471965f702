// cemc_b200.cu -- C ABI (include/cemc_b200.h) over the sm_100a kernels.
//
// Host side of the drop-in boundary: turns the flattened tables into the
// device "cluster program", owns the per-replica device state, and launches
// the kernels of cemc_kernels.cuh on the handle's stream.  No CPU fallback:
// every entry point either runs on the GPU or returns an error.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cemc_b200.h"
#include "cemc_kernels.cuh"
#include "cemc_spin_kernel.cuh"
#include "cemc_batch_launch.cuh"

using namespace cemc;

static thread_local std::string g_err;
static int fail(const std::string &m, int code = 1) { g_err = m; return code; }

#define CU(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess)                                                         \
      return fail(std::string(#x) + ": " + cudaGetErrorString(e_), 100);           \
  } while (0)

template <class T>
static cudaError_t upload(T **dst, const T *src, size_t n) {
  // padded to a multiple of 16 bytes (zero filled): the kernels stage the tables with
  // TMA bulk copies (cp.async.bulk), whose sizes are multiples of 16 bytes
  const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 15) / 16 * 16;
  cudaError_t e = cudaMalloc((void **)dst, bytes);
  if (e != cudaSuccess) return e;
  e = cudaMemset(*dst, 0, bytes);
  if (e != cudaSuccess) return e;
  if (n) e = cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}

struct TrialEntry { int32_t site; int8_t old_sp; };

struct cemc_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int R = 0, replica_offset = 0, replica_stride = 1;
  int acc_stride = 0;
  int order_mode = CEMC_ORDER_REFERENCE;
  uint64_t seed = 0;
  uint64_t launches = 0;
  int max_smem_optin = 0;
  DeviceTables t{};
  ReplicaState st{};
  std::vector<void *> owned;            // device allocations to free
  bool tracker_dirty = true;
  // trial history (CFHistoryTracker semantics, cf_history_tracker.cpp)
  std::vector<std::vector<TrialEntry>> trial_log;
  double *cf_committed = nullptr;       // [R][n_eci]
  double *e_committed = nullptr;        // [R]
  // scratch
  int32_t *d_sites = nullptr; int8_t *d_news = nullptr; double *d_u = nullptr;
  uint8_t *d_acc = nullptr; double *d_e = nullptr; long long scratch_steps = 0;
  long long trace_capacity = 0;
  int32_t *tr_sites = nullptr; int8_t *tr_news = nullptr; double *tr_u = nullptr;
  uint8_t *tr_acc = nullptr; double *tr_e = nullptr;
  double *cf_partial = nullptr;         // [R][n_jobs]
  double *cf_slots = nullptr;           // [R][n_jobs][cf_n_slots] (table-based recompute)
  int cf_n_slots = 0;
  int32_t *pt_scratch = nullptr; int pt_scratch_n = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;     // cemc_timer_start / _stop
  cudaEvent_t tv0 = nullptr, tv1 = nullptr;     // the autotuner's own pair
  // pinned staging buffers of the setters: the caller's buffer is consumed when the call
  // returns, the copy itself is stream-ordered (no host synchronisation per setter)
  struct Staging { void *host = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; };
  Staging stg_occ, stg_eci, stg_kT, stg_cf, stg_ref;
  Staging stg_out;                     // getters: device -> pinned (one synchronisation) -> caller's buffer
  // host copies needed by the API
  std::vector<int32_t> symm_of_site;
  std::vector<int8_t> allowed;
  int n_jobs = 0;
  bool integer_bf = false;
  int block_threads = 0;              // 0 = auto
  bool force_generic = false;         // testing: disable the register-resident P3 and the spin kernel
  bool no_spin = false;               // testing: skip the binary spin kernel
  double screen_slack = 1.0;          // testing: widen the Metropolis screening band
  bool autotune = true;               // pick the fastest kernel variant on long runs
  int tuned_sgc = -1, tuned_can = -1;  // variant chosen by the autotuner
  // device-side state observers (cemc_set_device_observers)
  long long obs_interval = 0, obs_step = 0, obs_capacity = 0;
  int obs_flags = 0;
  int obs_ring = 0;                   // snapshots per replica between two folds
  unsigned long long *ob_n = nullptr, *ob_folded = nullptr;
  double *ob_snap_cf = nullptr, *ob_snap_e = nullptr;
  int8_t *ob_snap_occ = nullptr;
  double *ob_cf_sum = nullptr, *ob_cf_sq = nullptr, *ob_best = nullptr, *ob_e = nullptr, *ob_order = nullptr;
  int8_t *ob_best_occ = nullptr, *ob_occ_ref = nullptr;
  bool lat_verified = false;          // `trans` is a periodic shift table: index arithmetic usable
  bool observe = true;                // accumulate the Averager / SGCObserver sums during run_*
  int last_variant = -1;              // variant of the most recent Metropolis launch (cemc_last_variant)
  // tuning across short launches: next variant to time, ms per move of the timed ones
  int lt_next[2] = {0, 0}, lt_best[2] = {-1, -1};     // tuning on long launches, resumable across calls
  float lt_best_ms[2] = {1e30f, 1e30f};
  int xt_next[2] = {0, 0};
  float xt_ms[2][16];
  int cluster = 0;                    // CTAs per chain in the batch kernel (0 = auto, 1, 2)
  int n_sms = 148;
  int batch = 0;                      // moves evaluated speculatively per batch (0 = auto)
  bool spin_ok = false;               // binary +-1 basis: warp-per-replica spin kernel usable
  SpinTables spin{};
  bool tab_ok = false;                // product tables fit: table evaluation in the batch kernel
  bool no_tab = false;                // testing: keep the fp64 product evaluation
  bool fp32 = false;                  // cemc_set_precision(32): single-precision tables / sums
  int32_t *d_order = nullptr;         // CTA -> replica map of the batch kernel (load balance), or null
  int32_t *d_order_buf = nullptr;
  int32_t *d_order_auto = nullptr;    // hottest-first order computed on the device (order_kernel)
  bool order_dirty = true;            // temperatures changed since d_order_auto was computed
  TabTables tab{};
  unsigned long long *d_phase = nullptr;   // CEMC_PHASE_TIMING builds
};

static int h2d_staged(cemc_handle *h, cemc_handle::Staging &st, void *dst, const void *src, size_t bytes) {
  if (st.cap < bytes) {
    if (st.host) { CU(cudaEventSynchronize(st.ev)); CU(cudaFreeHost(st.host)); st.host = nullptr; st.cap = 0; }
    CU(cudaMallocHost(&st.host, bytes));
    st.cap = bytes;
    if (!st.ev) CU(cudaEventCreateWithFlags(&st.ev, cudaEventDisableTiming));
  } else {
    CU(cudaEventSynchronize(st.ev));          // the previous copy out of this buffer has finished
  }
  memcpy(st.host, src, bytes);
  CU(cudaMemcpyAsync(dst, st.host, bytes, cudaMemcpyHostToDevice, h->stream));
  CU(cudaEventRecord(st.ev, h->stream));
  return 0;
}

static void reset_tuning(cemc_handle *h) {
  h->tuned_sgc = h->tuned_can = -1;
  h->xt_next[0] = h->xt_next[1] = 0;
  h->lt_next[0] = h->lt_next[1] = 0;
}

// ---------------------------------------------------------------------------
// small kernels

// E = N * sum_i eci_i * cf_i, sequential (ce_updater.cpp:236-242, named_array.cpp:25-33)
__global__ void energy_kernel(int R, int n_eci, int N, const double *eci, const double *cf,
                              double *e_cur) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  double e = 0.0;
  for (int i = 0; i < n_eci; i++)
    e = __dadd_rn(e, __dmul_rn(eci[(size_t)r * n_eci + i], cf[(size_t)r * n_eci + i]));
  e_cur[r] = __dmul_rn(e, (double)(unsigned)N);
}

__global__ void set_sites_kernel(int8_t *occ, int n, const int32_t *sites, const int8_t *vals) {
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int i = 0; i < n; i++) occ[sites[i]] = vals[i];
}

// SwapMoveIndexTracker.init_tracker (swap_move_index_tracker.py:22-36): per-species
// site lists in ascending site order.  One CTA per replica, stable counting sort.
__global__ void tracker_init_kernel(int N, int S, const int32_t *symm_of_site, const int8_t *occ,
                                    int32_t *list, int32_t *loc, int32_t *off) {
  const int r = blockIdx.x;
  const int8_t *o = occ + (size_t)r * N;
  int32_t *ls = list + (size_t)r * N, *lc = loc + (size_t)r * N, *of = off + (size_t)r * (S + 1);
  __shared__ int cnt[129];
  for (int i = threadIdx.x; i <= S; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  for (int a = threadIdx.x; a < N; a += blockDim.x)
    if (symm_of_site[a] >= 0) atomicAdd(&cnt[o[a]], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int sp = 0; sp < S; sp++) { of[sp] = run; run += cnt[sp]; }
    of[S] = run;
  }
  __syncthreads();
  // one thread per species walks the sites in ascending order (deterministic)
  for (int sp = threadIdx.x; sp < S; sp += blockDim.x) {
    int k = 0;
    const int base = of[sp];
    for (int a = 0; a < N; a++)
      if (symm_of_site[a] >= 0 && o[a] == sp) { ls[base + k] = a; lc[a] = k; k++; }
  }
  for (int a = threadIdx.x; a < N; a += blockDim.x)
    if (symm_of_site[a] < 0) lc[a] = -1;
}

// Brute-force CF partial sums (definition in SURVEY.md 8c): one CTA per (job, replica).
// jobs [0, n_tasks_total) are cluster tasks; the rest are the D singlet basis functions.
__global__ void cf_partial_kernel(DeviceTables t, const int8_t *occ, double *partial, int n_jobs) {
  const int job = blockIdx.x, r = blockIdx.y;
  const int8_t *o = occ + (size_t)r * t.N;
  const int RB = t.D * t.KP;
  double acc = 0.0;
  if (job < t.n_tasks_total) {
    int g = 0;
    while (job >= t.task_base[g + 1]) g++;
    // items of this task are contiguous: find the first one through its slot
    const int2 ts = t.task_sum[job];
    int q0 = t.item_base[g];
    while (t.item_slot[q0] != ts.x) q0++;
    // the task's factors decoded once per CTA (no integer division in the site loop):
    // x = row of bf (basis function), y = translation column, or -1 = the site itself
    __shared__ short2 s_dc[1024];
    const bool staged = ts.y * 4 <= 1024;
    if (staged) {
      for (int e = threadIdx.x; e < ts.y * 4; e += blockDim.x) {
        const unsigned long long w = t.items[q0 + (e >> 2)];
        const int idx = (int)((w >> (CEMC_ITEM_BITS * (e & 3))) & CEMC_ITEM_MASK);
        short2 v;
        if (idx == t.K) v = make_short2(-1, 0);                           // unused position
        else if (idx >= RB) v = make_short2((short)(idx - RB), -1);
        else v = make_short2((short)(idx / t.KP), (short)(idx % t.KP));
        s_dc[e] = v;
      }
      __syncthreads();
    }
    for (int a = threadIdx.x; a < t.N; a += blockDim.x) {
      if (t.symm_of_site[a] != g) continue;
      const int me = o[a];
      const int32_t *row = t.trans + (size_t)a * t.K;
      double sp = 0.0;
      if (staged) {
        for (int m = 0; m < ts.y; m++) {
          double tt = 1.0;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const short2 dc = s_dc[m * 4 + k];
            if (dc.x < 0) continue;
            const int sp_id = dc.y < 0 ? me : (int)o[__ldg(row + dc.y)];
            tt *= t.bf[dc.x * t.S + sp_id];
          }
          sp += tt;
        }
      } else {
        for (int m = 0; m < ts.y; m++) {
          const unsigned long long w = t.items[q0 + m];
          double tt = 1.0;
          for (int k = 0; k < 4; k++) {
            const int idx = (int)((w >> (CEMC_ITEM_BITS * k)) & CEMC_ITEM_MASK);
            if (idx == t.K) continue;                       // unused position
            double f;
            if (idx >= RB) f = t.bf[(idx - RB) * t.S + me];
            else { const int d = idx / t.KP, c = idx % t.KP; f = t.bf[d * t.S + o[t.trans[(size_t)a * t.K + c]]]; }
            tt *= f;
          }
          sp += tt;
        }
      }
      acc += sp;
    }
  } else {
    const int d = job - t.n_tasks_total;   // basis function number
    for (int a = threadIdx.x; a < t.N; a += blockDim.x)
      if (t.symm_of_site[a] >= 0) acc += t.bf[d * t.S + o[a]];
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(size_t)r * n_jobs + job] = red[0];
}

// Brute-force CF partial sums with the product tables of the batch kernel (TabTables): per tile
// of CF_TILE sites, (1) two threads per site pack the occupations of every sub-cluster of the
// site into table offsets (the codes are shared by all decorations of a family), (2) one thread
// per (task, site subset) adds table entries.  ~6x fewer instructions than one product per
// (site, task, sub-cluster, factor) and no dependent global gathers in the inner loop: the
// occupations of the replica and the tables sit in shared memory.  Deterministic: every (chunk,
// group, task) partial sum has its own slot; cf_final_kernel adds the slots in a fixed order.
#define CF_TILE 128
__global__ void __launch_bounds__(256)
cf_tab_kernel(DeviceTables t, TabTables tb, const int8_t *occ_all, double *partial, int n_jobs,
              int n_slots, int chunks, int tasks_pad, int occ_in_smem) {
  extern __shared__ __align__(16) unsigned char cf_smem[];
  const int r = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  const int N = t.N, K = t.K, n_sub = tb.n_sub, stride = tb.n_rounds * 32;
  double *tab_s = reinterpret_cast<double *>(cf_smem);
  unsigned short *codes = reinterpret_cast<unsigned short *>(tab_s + tb.n_tab);     // [CF_TILE][n_sub]
  int8_t *occ_s = reinterpret_cast<int8_t *>(codes + (size_t)CF_TILE * n_sub);
  const int8_t *g_occ = occ_all + (size_t)r * N;
  for (int i = tid; i < tb.n_tab; i += 256) tab_s[i] = tb.tab[i];
  if (occ_in_smem) for (int i = tid; i < N; i += 256) occ_s[i] = g_occ[i];
  const int8_t *o = occ_in_smem ? occ_s : g_occ;
  __syncthreads();
  const int per = (N + chunks - 1) / chunks, a_begin = chunk * per, a_end = min(N, a_begin + per);
  const int groups = 256 / tasks_pad, my_task = tid % tasks_pad, my_grp = tid / tasks_pad;
  int task_group = 0;                                   // symmetry group my task belongs to
  int4 tt = make_int4(0, 0, 0, 0);
  if (my_task < t.n_tasks_total) {
    while (my_task >= t.task_base[task_group + 1]) task_group++;
    tt = tb.task[my_task];
  }
  double acc = 0.0;
  for (int a0 = a_begin; a0 < a_end; a0 += CF_TILE) {
    // (1) codes: thread pair (2 s, 2 s + 1) shares site a0 + s, each takes every other sub-cluster
    {
      const int sl = tid >> 1, a = a0 + sl;
      if (a < a_end) {
        const int g = t.symm_of_site[a];
        const int me = o[a];
        const int32_t *row = t.trans + (size_t)a * K;
        for (int q = tid & 1; q < n_sub; q += 2) {
          unsigned short w = 0;
          if (g >= 0) {
            const uint2 d = tb.desc[g * stride + q];
            if (d.y == 0u) w = (unsigned short)(d.x & 0xffffu);                  // padding: the zero row
            else {
              const uint32_t rest = (uint32_t)o[__ldg(row + (d.x & 0xffu))] * (d.y & 0xffu) +
                                    (uint32_t)o[__ldg(row + ((d.x >> 8) & 0xffu))] * ((d.y >> 8) & 0xffu) +
                                    (uint32_t)o[__ldg(row + ((d.x >> 16) & 0xffu))] * ((d.y >> 16) & 0xffu);
              w = (unsigned short)(((rest + (uint32_t)me * (d.y >> 24)) * (d.x >> 24)) << 3);
            }
          }
          codes[sl * n_sub + q] = w;
        }
      }
    }
    __syncthreads();
    // (2) sums: my task over the sites of my group of threads
    if (my_task < t.n_tasks_total) {
      const char *tbl = reinterpret_cast<const char *>(tab_s) + tt.x;
      const int n_here = min(CF_TILE, a_end - a0);
      for (int sl = my_grp; sl < n_here; sl += groups) {
        if (t.symm_of_site[a0 + sl] != task_group) continue;
        const unsigned short *cp = codes + sl * n_sub + tt.y;
        double sp = 0.0;
        for (int m = 0; m < tt.z; m++) sp += *reinterpret_cast<const double *>(tbl + cp[m]);
        acc += sp;
      }
    }
    __syncthreads();
  }
  if (my_task < t.n_tasks_total)
    partial[((size_t)r * n_jobs + my_task) * n_slots + chunk * groups + my_grp] = acc;
  // singlet jobs: sum of the basis function over the active sites of this chunk
  __shared__ double red[256];
  for (int d = 0; d < t.D; d++) {
    double sv = 0.0;
    for (int a = a_begin + tid; a < a_end; a += 256)
      if (t.symm_of_site[a] >= 0) sv += t.bf[d * t.S + o[a]];
    red[tid] = sv;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) { if (tid < w) red[tid] += red[tid + w]; __syncthreads(); }
    if (tid < groups) partial[((size_t)r * n_jobs + t.n_tasks_total + d) * n_slots + chunk * groups + tid] = tid == 0 ? red[0] : 0.0;
    __syncthreads();
  }
}

// out[q] = sum of the n_slots partial sums of (replica, job) q, in ascending slot order
__global__ void cf_slot_sum_kernel(const double *slots, double *out, int n, int n_slots) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  double v = 0.0;
  for (int k = 0; k < n_slots; k++) v += slots[(size_t)q * n_slots + k];
  out[q] = v;
}

__global__ void cf_final_kernel(DeviceTables t, const double *partial, int n_jobs, double *cf) {
  const int r = blockIdx.x;
  for (int i = threadIdx.x; i < t.n_eci; i += blockDim.x) {
    double v = 0.0;
    bool is_cluster = false, any = false;
    for (int g = 0; g < t.n_symm; g++) {
      const int4 f = t.fin_i[g * t.n_eci + i];
      const double2 fd = t.fin_d[g * t.n_eci + i];
      if (f.x == 0) { v = 1.0; break; }
      if (f.x == 1) { v = partial[(size_t)r * n_jobs + t.n_tasks_total + f.y] / (double)t.N; break; }
      is_cluster = true;
      if (f.x == 2 && f.w > f.z) {
        double sg = 0.0;
        for (int q = f.z; q < f.w; q++) sg += partial[(size_t)r * n_jobs + t.task_base[g] + q];
        v += sg / ((double)(f.w - f.z) * fd.y);
        any = true;
      }
    }
    if (is_cluster && !any) v = 0.0;
    cf[(size_t)r * t.n_eci + i] = v;
  }
}

// ParallelTempering._perform_exchange_move (parallel_tempering.py:153-175): the
// slot<->replica map is permuted instead of copying configurations (:146-151).
__global__ void pt_exchange_kernel(int n_total, const double *energies, int32_t *slot_of_replica,
                                   const double *kT_of_slot, int direction, unsigned long long seed,
                                   unsigned long long round, int32_t *rep_of_slot, double *kT_local,
                                   int offset, int stride, int R, int32_t *n_accepted) {
  // `energies` is in all-gather order (rank-major).  Contiguous sharding (stride 1): that is the
  // global replica order; round-robin sharding (replica g on rank g % stride as local g / stride):
  // replica g sits at (g % stride) * (n_total / stride) + g / stride
  auto e_of = [&](int g) { return stride > 1 ? energies[(g % stride) * (n_total / stride) + g / stride] : energies[g]; };
  __shared__ int s_acc;
  if (threadIdx.x == 0) s_acc = 0;
  if (direction < 0) {
    // "up" or "down" per cycle (random.choice, parallel_tempering.py:191) from the counter
    // stream: Philox(seed; round, replica 0, stream 3) -- identical on every rank, restartable
    uint32_t c0 = (uint32_t)round, c1 = (uint32_t)(round >> 32), c2 = 0, c3 = 3;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    direction = (int)(c0 >> 31);
  }
  for (int g = threadIdx.x; g < n_total; g += blockDim.x) rep_of_slot[slot_of_replica[g]] = g;
  __syncthreads();
  const int n_pairs = n_total / 2;
  for (int p = threadIdx.x; p < n_pairs; p += blockDim.x) {
    const int i = direction == 0 ? 2 * p : n_total - 1 - 2 * p;
    const int j = direction == 0 ? i + 1 : i - 1;
    if (j < 0 || j >= n_total) continue;
    const int r1 = rep_of_slot[i], r2 = rep_of_slot[j];
    const double dE = __dsub_rn(e_of(r1), e_of(r2));                  // :139
    const double b1 = __ddiv_rn(1.0, kT_of_slot[i]);                  // :140
    const double b2 = __ddiv_rn(1.0, kT_of_slot[j]);                  // :141
    const double pr = exp(__dmul_rn(__dsub_rn(b1, b2), dE));          // :143
    uint32_t c0 = (uint32_t)round, c1 = (uint32_t)(round >> 32), c2 = (uint32_t)i, c3 = 2;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    if (u53(c0, c1) < pr) {                                           // :166
      rep_of_slot[i] = r2; rep_of_slot[j] = r1;
      atomicAdd(&s_acc, 1);
    }
  }
  __syncthreads();
  for (int s = threadIdx.x; s < n_total; s += blockDim.x) slot_of_replica[rep_of_slot[s]] = s;
  __syncthreads();
  for (int r = threadIdx.x; r < R; r += blockDim.x) kT_local[r] = kT_of_slot[slot_of_replica[offset + r * stride]];
  if (threadIdx.x == 0 && n_accepted) { n_accepted[0] = s_acc; n_accepted[1] += s_acc; }
}

// Energy autocorrelation of the traced window, on the device (the trace never leaves it):
// mean, variance and the first lag k with  sum_i d_i d_{i+k} / (n var) < 1/2,  d = E - mean
// (the quantity Montecarlo._estimate_correlation_time derives its correlation time from,
// cemc/mcmc/montecarlo.py:461-511).  One CTA per replica; lags are scanned in blocks of
// blockDim.x until one falls below 1/2.  out[r] = {mean, var, first lag or -1, min ACF seen}.
__global__ void autocorr_kernel(const double *trace, long long capacity, int n, double *out) {
  const int r = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const double *e = trace + (size_t)r * capacity;
  __shared__ double red[256];
  __shared__ int s_first;
  __shared__ double s_min;
  auto block_sum = [&](double v) {
    red[tid] = v;
    __syncthreads();
    for (int s = nt / 2; s > 0; s >>= 1) { if (tid < s) red[tid] += red[tid + s]; __syncthreads(); }
    const double t = red[0];
    __syncthreads();
    return t;
  };
  double acc = 0.0;
  for (int i = tid; i < n; i += nt) acc += e[i];
  const double mean = block_sum(acc) / (double)n;
  acc = 0.0;
  for (int i = tid; i < n; i += nt) { const double d = e[i] - mean; acc += d * d; }
  const double var = block_sum(acc) / (double)n;
  if (tid == 0) { s_first = 0x7fffffff; s_min = 1.0; }
  __syncthreads();
  if (var > 0.0) {
    const double norm = 1.0 / ((double)n * var);
    for (int k0 = 0; k0 < n; k0 += nt) {
      const int k = k0 + tid;
      if (k < n) {
        double c = 0.0;
        for (int i = 0; i + k < n; i++) c += (e[i] - mean) * (e[i + k] - mean);
        c *= norm;
        if (c < 0.5) atomicMin(&s_first, k);
        // min over the scanned lags (only used for the "window too short" message)
        unsigned long long *pm = reinterpret_cast<unsigned long long *>(&s_min);
        unsigned long long old = *pm;
        while (__longlong_as_double((long long)old) > c) {
          const unsigned long long prev = atomicCAS(pm, old, (unsigned long long)__double_as_longlong(c));
          if (prev == old) break;
          old = prev;
        }
      }
      __syncthreads();
      if (s_first != 0x7fffffff) break;
    }
  }
  if (tid == 0) {
    out[4 * r] = mean; out[4 * r + 1] = var;
    out[4 * r + 2] = s_first == 0x7fffffff ? -1.0 : (double)s_first;
    out[4 * r + 3] = s_min;
  }
}

// ---------------------------------------------------------------------------
template <class T>
static int dalloc(cemc_handle *h, T **p, size_t n, bool zero = true) {
  CU(cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)));
  h->owned.push_back(*p);
  if (zero) CU(cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T)));
  return 0;
}

template <class T>
static int dupload(cemc_handle *h, const T **p, const std::vector<T> &v) {
  T *d = nullptr;
  CU(upload(&d, v.data(), v.size()));
  h->owned.push_back(d);
  *p = d;
  return 0;
}

static int ensure_tracker(cemc_handle *h) {
  if (!h->tracker_dirty) return 0;
  tracker_init_kernel<<<h->R, 128, 0, h->stream>>>(h->t.N, h->t.S, h->t.symm_of_site, h->st.occ,
                                                   h->st.list, h->st.loc, h->st.off);
  h->launches++;
  CU(cudaGetLastError());
  h->tracker_dirty = false;
  return 0;
}

static int upload_allowed(cemc_handle *h) {
  int8_t al[128], ap[128];
  memset(al, 0, sizeof al);
  memset(ap, -1, sizeof ap);
  for (size_t i = 0; i < h->allowed.size(); i++) { al[i] = h->allowed[i]; ap[h->allowed[i]] = (int8_t)i; }
  CU(cudaMemcpy((void *)h->t.allowed, al, 128, cudaMemcpyHostToDevice));
  CU(cudaMemcpy((void *)h->t.allowed_pos, ap, 128, cudaMemcpyHostToDevice));
  h->t.n_allowed = (int)h->allowed.size();
  h->t.allowed_identity = ((int)h->allowed.size() == h->t.S) ? 1 : 0;
  for (size_t i = 0; i < h->allowed.size(); i++) if (h->allowed[i] != (int8_t)i) h->t.allowed_identity = 0;
  return 0;
}

static int check_status(cemc_handle *h) {
  std::vector<int32_t> stt(h->R);
  CU(cudaMemcpyAsync(stt.data(), h->st.status, sizeof(int32_t) * h->R, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (int r = 0; r < h->R; r++) {
    if (stt[r]) {
      CU(cudaMemsetAsync(h->st.status, 0, sizeof(int32_t) * h->R, h->stream));
      char buf[160];
      const char *why = stt[r] == 1 ? "Attempting to move a background atom!"
                        : stt[r] == 2 ? "There is only one element in the given atoms object!"
                                      : "proposal out of range";
      snprintf(buf, sizeof buf, "replica %d: %s", r, why);
      return fail(buf, 10 + stt[r]);
    }
  }
  return 0;
}

static int status_error(cemc_handle *h, const int32_t *stt) {
  for (int r = 0; r < h->R; r++) {
    if (stt[r]) {
      CU(cudaMemsetAsync(h->st.status, 0, sizeof(int32_t) * h->R, h->stream));
      char buf[160];
      const char *why = stt[r] == 1 ? "Attempting to move a background atom!"
                        : stt[r] == 2 ? "There is only one element in the given atoms object!"
                                      : "proposal out of range";
      snprintf(buf, sizeof buf, "replica %d: %s", r, why);
      return fail(buf, 10 + stt[r]);
    }
  }
  return 0;
}

// Getter: device -> pinned staging buffer (a pageable destination would make the driver stage the
// copy itself, several times slower), the kernels' status words ride along, ONE stream
// synchronisation, then a host copy into the caller's buffer.
static int d2h_staged(cemc_handle *h, void *dst, const void *src, size_t bytes, bool with_status) {
  cemc_handle::Staging &st = h->stg_out;
  const size_t sbytes = sizeof(int32_t) * (size_t)h->R, need = ((bytes + 15) / 16) * 16 + sbytes;
  if (st.cap < need) {
    if (st.host) { CU(cudaFreeHost(st.host)); st.host = nullptr; st.cap = 0; }
    CU(cudaMallocHost(&st.host, need));
    st.cap = need;
  }
  char *pin = static_cast<char *>(st.host);
  int32_t *pstat = reinterpret_cast<int32_t *>(pin + ((bytes + 15) / 16) * 16);
  CU(cudaMemcpyAsync(pin, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (with_status) CU(cudaMemcpyAsync(pstat, h->st.status, sbytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  memcpy(dst, pin, bytes);
  return with_status ? status_error(h, pstat) : 0;
}

extern "C" {

const char *cemc_last_error(void) { return g_err.c_str(); }
int cemc_version(void) { return 100; }

int cemc_create(const cemc_tables *tb, int n_replicas, int replica_offset, int device, void *stream,
                cemc_handle **out) {
  if (!tb || !out) return fail("null argument");
  if (n_replicas < 1) return fail("n_replicas must be >= 1");
  const int N = tb->n_sites, S = tb->n_species, D = tb->n_bf, K = tb->n_cols, n_eci = tb->n_eci;
  if (N < 1 || S < 1 || S > 127 || D < 1 || K < 1 || n_eci < 1 || tb->n_symm < 1)
    return fail("invalid table sizes");
  if (K > 254) return fail("more than 254 translation-matrix columns are not supported");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail("no such CUDA device");
  CU(cudaSetDevice(device));

  // validate tables the way the reference's Python layer does
  for (int f = 0; f < tb->n_fam; f++)
    if (tb->fam_size[f] < 2 || tb->fam_size[f] > CEMC_MAX_CLUSTER_SIZE)
      return fail("Only cluster sizes 2, 3 and 4 are supported!");          // cluster.cpp:168
  for (size_t q = 0; q < (size_t)N * K; q++)
    if (tb->trans[q] < 0 || tb->trans[q] >= N) return fail("translation matrix entry out of range");
  {
    std::vector<char> col_used(K, 0);
    for (int f = 0; f < tb->n_fam; f++)
      for (int q = tb->fam_pos_off[f]; q < tb->fam_pos_off[f + 1]; q++) {
        const int p = tb->fam_pos[q];
        if (p != CEMC_POS_REF && (p < 0 || p >= K)) return fail("fam_pos entry out of range");
        if (p >= 0) col_used[p] = 1;
      }
    for (int s = 0; s < N; s++) {
      if (tb->symm_of_site[s] < 0) continue;
      for (int c = 0; c < K; c++)
        if (col_used[c] && tb->trans[(size_t)s * K + c] == s)
          return fail("The simulation cell is so small that the same site is present multiple "
                      "times within one cluster. Increase the size of the simulation cell.");
    }
  }

  cemc_handle *h = new cemc_handle();
  h->device = device;
  h->R = n_replicas;
  h->replica_offset = replica_offset;
  if (stream) h->stream = (cudaStream_t)stream;
  else { CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
  CU(cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  CU(cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, device));
  CU(cudaEventCreate(&h->ev0));
  CU(cudaEventCreate(&h->ev1));
  CU(cudaEventCreate(&h->tv0));
  CU(cudaEventCreate(&h->tv1));

  // ---- build the cluster program ------------------------------------------
  DeviceTables &t = h->t;
  t.N = N; t.S = S; t.D = D; t.K = K; t.KP = K + 1; t.VS = D * (K + 1) + 2 * D;
  t.n_eci = n_eci; t.n_symm = tb->n_symm;
  if (t.VS > (int)CEMC_ITEM_MASK) return fail("too many (basis function, column) pairs for the item encoding");
  const int RB = D * t.KP;
  std::vector<unsigned long long> items;
  std::vector<uint4> items4;
  std::vector<uint16_t> item_slot;
  std::vector<int32_t> item_base(tb->n_symm + 1, 0), task_base(tb->n_symm + 1, 0);
  std::vector<int2> task_sum;
  std::vector<int4> fin_i((size_t)tb->n_symm * n_eci);
  std::vector<double2> fin_d((size_t)tb->n_symm * n_eci);
  std::vector<int32_t> singlet_idx;
  int max_tasks = 0, max_items = 0, max_slots = 0;
  for (int g = 0; g < tb->n_symm; g++) {
    task_base[g] = (int32_t)task_sum.size();
    item_base[g] = (int32_t)items.size();
    int slot = 0;
    for (int i = 0; i < n_eci; i++) {
      int4 &f = fin_i[(size_t)g * n_eci + i];
      double2 &fd = fin_d[(size_t)g * n_eci + i];
      f.x = tb->eci_kind[i]; f.y = tb->eci_bf[i]; f.z = f.w = 0; fd.x = 0.0; fd.y = 1.0;
      if (f.x == CEMC_ECI_SINGLET && (f.y < 0 || f.y >= D)) return fail("singlet decoration out of range");
      if (f.x != CEMC_ECI_CLUSTER) continue;
      const int term = g * n_eci + i;
      const int fam = tb->term_fam[term];
      if (fam < 0) { f.x = -1; continue; }
      if (fam >= tb->n_fam) return fail("family id out of range");
      const int d0 = tb->term_deco_off[term], d1 = tb->term_deco_off[term + 1];
      if (d1 <= d0) return fail("cluster ECI without decorations");
      const int n = tb->fam_size[fam], M = tb->fam_nsub[fam];
      const int32_t *pos = tb->fam_pos + tb->fam_pos_off[fam];
      f.z = (int32_t)task_sum.size() - task_base[g];
      for (int e = d0; e < d1; e++) {
        const int tk = (int)task_sum.size() - task_base[g];
        while (slot % 16 != tk % 16) slot++;        // bank-conflict-free sums (P2b)
        task_sum.push_back(make_int2(slot, M));
        for (int m = 0; m < M; m++) {
          unsigned long long w = 0;
          int kref = -1;
          for (int k = 0; k < 4; k++) {
            int idx = K;                              // constant 1.0
            if (k < n) {
              const int dk = tb->deco[4 * e + k];
              if (dk < 0 || dk >= D) return fail("decoration number out of range");
              const int p = pos[m * n + k];
              if (p == CEMC_POS_REF) {
                if (kref >= 0) return fail("a sub-cluster lists the changed site twice");
                kref = k; idx = RB + dk;
              } else idx = dk * t.KP + p;
            }
            w |= (unsigned long long)idx << (CEMC_ITEM_BITS * k);
          }
          if (kref < 0) return fail("a sub-cluster does not contain the changed site (bad `order`)");
          w |= (unsigned long long)kref << 48;
          if (slot >= (1 << 14)) return fail("cluster program too large (product slots)");
          w |= (unsigned long long)slot << 50;
          items.push_back(w);
          {   // byte offsets into V for the batch kernel: x = off0|off1<<16, y = off2|off3<<16,
              // z = slot | kref<<16, w = offset of the NEW value of the changed site
            uint32_t off[4];
            for (int k = 0; k < 4; k++) off[k] = (uint32_t)((w >> (CEMC_ITEM_BITS * k)) & CEMC_ITEM_MASK) * 8u;
            uint4 it;
            it.x = off[0] | (off[1] << 16); it.y = off[2] | (off[3] << 16);
            it.z = (uint32_t)slot | ((uint32_t)kref << 16);
            it.w = off[kref] + (uint32_t)D * 8u;
            items4.push_back(it);
          }
          item_slot.push_back((uint16_t)slot++);
        }
      }
      f.w = (int32_t)task_sum.size() - task_base[g];
      fd.x = (double)n / (double)(d1 - d0);                                      // :400
      fd.y = (double)(tb->term_count[term] * tb->symm_count[g]);                 // :402
    }
    max_tasks = std::max(max_tasks, (int)task_sum.size() - task_base[g]);
    max_items = std::max(max_items, (int)items.size() - item_base[g]);
    max_slots = std::max(max_slots, slot);
  }
  task_base[tb->n_symm] = (int32_t)task_sum.size();
  item_base[tb->n_symm] = (int32_t)items.size();
  for (int i = 0; i < n_eci; i++)
    if (tb->eci_kind[i] == CEMC_ECI_SINGLET) singlet_idx.push_back(i);
  t.n_items_total = (int)items.size(); t.n_tasks_total = (int)task_sum.size();
  t.max_tasks = max_tasks; t.max_items = max_items; t.max_slots = max_slots;
  t.n_singlets = (int)singlet_idx.size();
  h->acc_stride = CEMC_ACC_STRIDE(t.n_singlets);
  h->n_jobs = t.n_tasks_total + D;
  // order-free summation is bit-exact when every product is a small integer
  h->integer_bf = true;
  for (int q = 0; q < D * S; q++) {
    const double v = tb->bf[q];
    if (v != std::floor(v) || std::fabs(v) > 1024.0) h->integer_bf = false;
  }

  // ---- binary +-1 basis: tables of the warp-per-replica spin kernel ------------
  std::vector<uint32_t> sp_items, sp_masks(32 * 4, 0u);
  std::vector<int32_t> sp_coef(32, 0), sp_msub(32, 1);
  {
    bool ok = (S == 2 && D == 1 && tb->n_symm == 1 && n_eci <= 32 &&
               std::fabs(tb->bf[0]) == 1.0 && tb->bf[1] == -tb->bf[0]);
    for (int s = 0; s < N && ok; s++) if (tb->symm_of_site[s] != 0) ok = false;
    const int b0 = ok ? (int)tb->bf[0] : 1;
    for (int i = 0; i < n_eci && ok; i++) {
      const int kind = tb->eci_kind[i];
      if (kind == CEMC_ECI_SINGLET) { sp_coef[i] = 1; sp_msub[i] = 1; continue; }
      if (kind != CEMC_ECI_CLUSTER) continue;
      const int fam = tb->term_fam[i];
      if (fam < 0) continue;                                   // copied
      if (tb->term_deco_off[i + 1] - tb->term_deco_off[i] != 1) { ok = false; break; }
      const int n = tb->fam_size[fam], M = tb->fam_nsub[fam];
      const int32_t *pos = tb->fam_pos + tb->fam_pos_off[fam];
      sp_coef[i] = n * ((n - 1) % 2 == 0 ? 1 : b0);            // n * b0^(n-1)
      sp_msub[i] = M;
      for (int m = 0; m < M; m++) {
        uint32_t w = 0; int nn = 0;
        for (int k = 0; k < n; k++) {
          const int p = pos[m * n + k];
          if (p == CEMC_POS_REF) continue;
          w |= (uint32_t)p << (8 * nn++);
        }
        for (; nn < 3; nn++) w |= 0xffu << (8 * nn);
        const int q = (int)sp_items.size();
        if (q >= 128) { ok = false; break; }
        sp_masks[i * 4 + q / 32] |= 1u << (q % 32);
        sp_items.push_back(w);
      }
    }
    h->spin_ok = ok && !sp_items.empty();
    h->spin.n_items = (int)sp_items.size();
    h->spin.n_rounds = ((int)sp_items.size() + 31) / 32;
    h->spin.b0 = b0;
    h->spin.wq = *std::max_element(sp_msub.begin(), sp_msub.end()) + 1;
  }

  // ---- product tables of the batch kernel's table evaluation (TabTables) --------
  // Per family one table [code][decoration]: code = sum_k occ_k S^k over the sorted
  // positions, entry = the left-to-right product of ce_updater.cpp:271-281 for those
  // occupations, multiplied in the reference's order (plain IEEE double products: the
  // bits the reference computes).  Decorations are the fastest index, so the lanes of
  // one family (one lane per decoration) read one contiguous row.
  // Several translational symmetry groups (crystals with a basis): one descriptor block and one
  // task list per group -- the evaluation picks them by the changed site's group
  // (ce_updater.cpp:379-384); family ids are already unique per (group, prefix).
  std::vector<uint2> tb_desc;
  std::vector<int4> tb_task;
  std::vector<double> tb_tab;
  {
    bool ok = (n_eci <= 64 && K <= 63 && S <= 9 &&                        // K > 31: two columns per lane;
               (n_eci <= 32 || K <= 31));                                 // n_eci > 32: two ECIs per lane
    auto power = [&](int e) { int v = 1; for (int q = 0; q < e; q++) v *= S; return v; };
    std::vector<std::vector<std::vector<int>>> fam_decos(tb->n_fam);     // distinct decorations per family
    std::vector<int> fam_group(tb->n_fam, -1);
    std::vector<std::pair<int, int>> task_fd;                            // task -> (family, decoration index), task order of fin_i
    for (int g = 0; g < tb->n_symm && ok; g++)
      for (int i = 0; i < n_eci && ok; i++) {
        if (tb->eci_kind[i] != CEMC_ECI_CLUSTER) continue;
        const int term = g * n_eci + i;
        const int fam = tb->term_fam[term];
        if (fam < 0) continue;
        if (fam_group[fam] >= 0 && fam_group[fam] != g) { ok = false; break; }     // one table per (group, family)
        fam_group[fam] = g;
        const int n = tb->fam_size[fam];
        for (int e = tb->term_deco_off[term]; e < tb->term_deco_off[term + 1]; e++) {
          std::vector<int> key;
          for (int k = 0; k < n; k++) key.push_back(tb->deco[4 * e + k]);
          auto &fd = fam_decos[fam];
          int idx = (int)(std::find(fd.begin(), fd.end(), key) - fd.begin());
          if (idx == (int)fd.size()) fd.push_back(key);
          task_fd.push_back(std::make_pair(fam, idx));
        }
      }
    std::vector<int> sub_base(tb->n_fam, 0), tab_base(tb->n_fam, 0);
    std::vector<std::vector<uint2>> desc_g(tb->n_symm);
    for (int fam = 0; fam < tb->n_fam && ok; fam++) {
      const int nd = (int)fam_decos[fam].size();
      if (!nd) continue;
      std::vector<uint2> &dg = desc_g[fam_group[fam]];
      const int n = tb->fam_size[fam], M = tb->fam_nsub[fam];
      const int32_t *pos = tb->fam_pos + tb->fam_pos_off[fam];
      const int n_codes = power(n);
      if (nd > 255 || (n_codes + 1) * nd * 8 > 65536) { ok = false; break; }
      sub_base[fam] = (int)dg.size();
      for (int m = 0; m < M; m++) {
        uint32_t cols = (uint32_t)nd << 24, wts = 0; int nn = 0;
        for (int k = 0; k < n; k++) {
          const int p = pos[m * n + k];
          if (p == CEMC_POS_REF) wts |= (uint32_t)power(k) << 24;
          else { cols |= (uint32_t)p << (8 * nn); wts |= (uint32_t)power(k) << (8 * nn); nn++; }
        }
        dg.push_back(make_uint2(cols, wts));
      }
      // padding to a multiple of 8 sub-clusters: weights 0 mark the entry, x = byte offset of
      // the table's all-zero row (adding +0.0 never changes a sum that started at +0.0)
      while (dg.size() % 8) dg.push_back(make_uint2((uint32_t)(n_codes * nd * 8), 0u));
      tab_base[fam] = (int)tb_tab.size();
      for (int code = 0; code < n_codes; code++)
        for (int e = 0; e < nd; e++) {
          double v = 1.0;
          int c = code;
          for (int k = 0; k < n; k++, c /= S) {
            volatile double prod = v * tb->bf[(size_t)fam_decos[fam][e][k] * S + (c % S)];   // :279, one rounding per factor
            v = prod;
          }
          tb_tab.push_back(v);
        }
      for (int e = 0; e < nd; e++) tb_tab.push_back(0.0);          // the zero row
    }
    for (auto &fd : task_fd)
      tb_task.push_back(make_int4((tab_base[fd.first] + fd.second) * 8, sub_base[fd.first], (tb->fam_nsub[fd.first] + 7) & ~7, 0));
    size_t max_sub = 0;
    for (auto &dg : desc_g) max_sub = std::max(max_sub, dg.size());
    if (max_sub > 128 || tb_tab.size() * 8 > 96 * 1024 || tb_task.empty()) ok = false;
    if (ok && (int)tb_task.size() != (int)task_sum.size()) ok = false;      // same task numbering as fin_i
    h->tab_ok = ok;
    h->tab.n_sub = (int)max_sub;
    h->tab.n_rounds = ((int)max_sub + 31) / 32;
    h->tab.n_tab = (int)tb_tab.size();
    const size_t stride = (size_t)std::max(1, h->tab.n_rounds) * 32;          // descriptors of group g at g * stride
    tb_desc.assign(stride * tb->n_symm, make_uint2(0u, 0u));                  // (entries past a group's count: never read by a task)
    for (int g = 0; g < tb->n_symm; g++) std::copy(desc_g[g].begin(), desc_g[g].end(), tb_desc.begin() + g * stride);
  }

  std::vector<int32_t> trans(tb->trans, tb->trans + (size_t)N * K);
  std::vector<int32_t> symm(tb->symm_of_site, tb->symm_of_site + N);
  std::vector<double> bf(tb->bf, tb->bf + (size_t)D * S);
  h->symm_of_site = symm;
  std::vector<int32_t> active;
  for (int s = 0; s < N; s++) if (symm[s] >= 0) active.push_back(s);
  t.n_active = (int)active.size();
  if (t.n_active == 0) return fail("no active sites");
  h->allowed.resize(S);
  for (int s = 0; s < S; s++) h->allowed[s] = (int8_t)s;
  t.n_allowed = S;
  int rc;
  if ((rc = dupload(h, &t.trans, trans))) return rc;
  if ((rc = dupload(h, &t.symm_of_site, symm))) return rc;
  if ((rc = dupload(h, &t.bf, bf))) return rc;
  if ((rc = dupload(h, &t.items, items))) return rc;
  if ((rc = dupload(h, &t.items4, items4))) return rc;
  if ((rc = dupload(h, &t.item_slot, item_slot))) return rc;
  if ((rc = dupload(h, &t.item_base, item_base))) return rc;
  if ((rc = dupload(h, &t.task_base, task_base))) return rc;
  if ((rc = dupload(h, &t.task_sum, task_sum))) return rc;
  if ((rc = dupload(h, &t.fin_i, fin_i))) return rc;
  if ((rc = dupload(h, &t.fin_d, fin_d))) return rc;
  if ((rc = dupload(h, &t.singlet_idx, singlet_idx))) return rc;
  if (h->spin_ok) {
    if ((rc = dupload(h, &h->spin.items, sp_items))) return rc;
    if ((rc = dupload(h, &h->spin.masks, sp_masks))) return rc;
    if ((rc = dupload(h, &h->spin.coef, sp_coef))) return rc;
    if ((rc = dupload(h, &h->spin.msub, sp_msub))) return rc;
  }
  if (h->tab_ok) {
    if ((rc = dupload(h, &h->tab.desc, tb_desc))) return rc;
    if ((rc = dupload(h, &h->tab.task, tb_task))) return rc;
    if ((rc = dupload(h, &h->tab.tab, tb_tab))) return rc;
    std::vector<float> tb_tab32(tb_tab.begin(), tb_tab.end());      // each entry rounded once
    if ((rc = dupload(h, &h->tab.tab32, tb_tab32))) return rc;
  }
#if defined(CEMC_PHASE_TIMING) || defined(CEMC_WARP_TIMING)
  if ((rc = dalloc(h, &h->d_phase, (size_t)n_replicas * 24))) return rc;
#endif
  if (t.n_active != N) { if ((rc = dupload(h, &t.active, active))) return rc; }
  else t.active = nullptr;
  t.uniform_group = (tb->n_symm == 1 && t.n_active == N) ? 1 : 0;
  // ---- translation-invariant lattice?  (hint from the host, verified entry by entry) ----
  t.lat_ok = 0;
  if (tb->lattice_dims && t.n_active == N) {
    const long long L1 = tb->lattice_dims[0], L2 = tb->lattice_dims[1], L3 = tb->lattice_dims[2];
    bool ok = L1 >= 1 && L2 >= 1 && L3 >= 1 && L1 < 1024 && L2 < 1024 && L3 < 1024 && L1 * L2 * L3 == N;
    std::vector<uint32_t> shift(K, 0u);
    for (int c = 0; c < K && ok; c++) {
      const int s0 = tb->trans[c];                         // T(site 0, c): the shift itself
      const int di = s0 / (int)(L2 * L3), dj = (s0 / (int)L3) % (int)L2, dk = s0 % (int)L3;
      shift[c] = (uint32_t)di | ((uint32_t)dj << 10) | ((uint32_t)dk << 20);
      for (int s = 0; s < N && ok; s++) {
        const int i = s / (int)(L2 * L3), j = (s / (int)L3) % (int)L2, k = s % (int)L3;
        const int want = (((i + di) % (int)L1) * (int)L2 + (j + dj) % (int)L2) * (int)L3 + (k + dk) % (int)L3;
        if (tb->trans[(size_t)s * K + c] != want) ok = false;
      }
    }
    if (ok) {
      const uint32_t L23 = (uint32_t)(L2 * L3);
      const uint32_t m23 = (uint32_t)((0x100000000ull + L23 - 1) / L23), m3 = (uint32_t)((0x100000000ull + L3 - 1) / (uint32_t)L3);
      // the multiply-high division must be exact for every operand the kernels can form
      for (uint32_t s = 0; s < (uint32_t)N && ok; s++) if ((uint32_t)(((unsigned long long)s * m23) >> 32) != s / L23) ok = false;
      for (uint32_t r = 0; r < L23 && ok; r++) if ((uint32_t)(((unsigned long long)r * m3) >> 32) != r / (uint32_t)L3) ok = false;
      if (L23 == 1) ok = false;       // m would not fit 32 bits
      if (L3 == 1) ok = false;
      if (ok) {
        h->lat_verified = true;
        // default: on for tables of several MB (measured on B200: fcc 64^3, 19 MB, gains 5 %; fcc 20^3,
        // 576 KB, gains 1 % -- less than the ~5 % the kernels with the optional code paths compiled in
        // (kX) lose; tables <= 128 KB -- fcc 10^3, 12^3 -- are L1 hits and lose ~1 %)
        t.lat_ok = ((size_t)N * K * sizeof(int32_t) > ((size_t)4 << 20)) ? 1 : 0; t.L1 = (uint32_t)L1; t.L2 = (uint32_t)L2; t.L3 = (uint32_t)L3; t.L23 = L23;
        t.lat_m23 = m23; t.lat_m3 = m3;
        if ((rc = dupload(h, &t.col_shift, shift))) return rc;
      }
    }
  }
  t.prefetch_rows = ((size_t)32 * 2 * K * sizeof(int32_t) <= 24 * 1024) ? 1 : 0;
  {
    int8_t *al = nullptr, *ap = nullptr;
    if ((rc = dalloc(h, &al, 128))) return rc;
    if ((rc = dalloc(h, &ap, 128))) return rc;
    t.allowed = al; t.allowed_pos = ap;
    if ((rc = upload_allowed(h))) return rc;
  }

  // ---- per-replica state ---------------------------------------------------
  const size_t R = (size_t)n_replicas;
  ReplicaState &st = h->st;
  if ((rc = dalloc(h, &st.occ, R * N + 32))) return rc;       // + padding: the kernels' TMA staging reads whole 16-byte blocks
  if ((rc = dalloc(h, &st.cf, R * n_eci))) return rc;
  if ((rc = dalloc(h, &st.eci, R * n_eci))) return rc;
  if ((rc = dalloc(h, &st.e_cur, R))) return rc;
  if ((rc = dalloc(h, &st.kT, R))) return rc;
  if ((rc = dalloc(h, &st.acc, R * h->acc_stride))) return rc;
  if ((rc = dalloc(h, &st.ref, R))) return rc;
  if ((rc = dalloc(h, &st.step, R))) return rc;
  if ((rc = dalloc(h, &st.accepted, R))) return rc;
  if ((rc = dalloc(h, &st.list, R * N))) return rc;
  if ((rc = dalloc(h, &st.loc, R * N))) return rc;
  if ((rc = dalloc(h, &st.off, R * (S + 1)))) return rc;
  if ((rc = dalloc(h, &st.status, R))) return rc;
  if ((rc = dalloc(h, &h->cf_committed, R * n_eci))) return rc;
  if ((rc = dalloc(h, &h->e_committed, R))) return rc;
  if ((rc = dalloc(h, &h->cf_partial, R * h->n_jobs))) return rc;
  h->trial_log.resize(R);
  {
    std::vector<double> ones(R, 1.0), eci(R * n_eci);
    for (size_t r = 0; r < R; r++) memcpy(&eci[r * n_eci], tb->eci, sizeof(double) * n_eci);
    CU(cudaMemcpy(st.kT, ones.data(), sizeof(double) * R, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(st.ref, ones.data(), sizeof(double) * R, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(st.eci, eci.data(), sizeof(double) * R * n_eci, cudaMemcpyHostToDevice));
  }
  *out = h;
  return 0;
}

int cemc_destroy(cemc_handle *h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (void *p : h->owned) cudaFree(p);
  void *extra[] = {h->d_sites, h->d_news, h->d_u, h->d_acc, h->d_e, h->tr_sites, h->tr_news,
                   h->tr_u, h->tr_acc, h->tr_e, h->pt_scratch, h->ob_n, h->ob_folded, h->ob_snap_cf,
                   h->ob_snap_e, h->ob_snap_occ, h->ob_cf_sum, h->ob_cf_sq,
                   h->ob_best, h->ob_e, h->ob_order, h->ob_best_occ, h->ob_occ_ref, h->cf_slots};
  for (void *p : extra) if (p) cudaFree(p);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (cemc_handle::Staging *st : {&h->stg_occ, &h->stg_eci, &h->stg_kT, &h->stg_cf, &h->stg_ref, &h->stg_out}) {
    if (st->ev) { cudaEventSynchronize(st->ev); cudaEventDestroy(st->ev); }
    if (st->host) cudaFreeHost(st->host);
  }
  if (h->tv0) cudaEventDestroy(h->tv0);
  if (h->tv1) cudaEventDestroy(h->tv1);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int cemc_set_replica_stride(cemc_handle *h, int stride) {
  if (!h) return fail("null handle");
  if (stride < 1) return fail("replica stride must be >= 1");
  h->replica_stride = stride;
  return 0;
}

int cemc_set_stream(cemc_handle *h, void *stream) {
  if (!h) return fail("null handle");
  CU(cudaStreamSynchronize(h->stream));
  if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
  if (stream) h->stream = (cudaStream_t)stream;
  else { CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
  return 0;
}

int cemc_synchronize(cemc_handle *h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  return check_status(h);
}

int cemc_set_order_mode(cemc_handle *h, int mode) {
  if (!h) return fail("null handle");
  if (mode != CEMC_ORDER_REFERENCE && mode != CEMC_ORDER_TREE) return fail("unknown order mode");
  h->order_mode = mode;
  return 0;
}

// Load balance of the batch kernel when there are more chains than SMs (one CTA per chain,
// two CTAs per SM): CTA i works on the chain with the i-th highest temperature.  The block
// scheduler fills the SMs breadth first, so the hottest chains (most accepted moves, most
// bookkeeping: the slowest CTAs) end up sharing an SM with the coldest ones and the launch,
// which ends with its slowest CTA, gets shorter (config 2: 209.6 -> 202.4 ns/move).  Results
// never depend on the order (chains are keyed by replica id).
__global__ void order_kernel(int R, const double *kT, int32_t *order) {
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const double k = kT[r];
    int rank = 0;
    for (int j = 0; j < R; j++) {
      const double kj = kT[j];
      rank += (kj > k) || (kj == k && j < r);
    }
    order[rank] = r;
  }
}

static int refresh_energy(cemc_handle *h) {
  energy_kernel<<<(h->R + 127) / 128, 128, 0, h->stream>>>(h->R, h->t.n_eci, h->t.N, h->st.eci,
                                                           h->st.cf, h->st.e_cur);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

static void drop_trials(cemc_handle *h) { for (auto &l : h->trial_log) l.clear(); }

int cemc_set_occupancy(cemc_handle *h, const int8_t *occ) {
  if (!h || !occ) return fail("null argument");
  CU(cudaSetDevice(h->device));
  const size_t n = (size_t)h->R * h->t.N;
  {   // range check without an early exit, so that it vectorises (256 KB per bench step)
    unsigned bad = 0;
    const unsigned S = (unsigned)h->t.S;
    for (size_t q = 0; q < n; q++) bad |= (unsigned)((unsigned char)occ[q] >= S);
    if (bad) return fail("occupancy value out of range");
  }
  { const int rc = h2d_staged(h, h->stg_occ, h->st.occ, occ, n); if (rc) return rc; }
  h->tracker_dirty = true;
  drop_trials(h);
  return 0;
}

int cemc_get_occupancy(cemc_handle *h, int8_t *occ) {
  if (!h || !occ) return fail("null argument");
  CU(cudaSetDevice(h->device));
  return d2h_staged(h, occ, h->st.occ, (size_t)h->R * h->t.N, false);
}

int cemc_set_cf(cemc_handle *h, const double *cf) {
  if (!h || !cf) return fail("null argument");
  CU(cudaSetDevice(h->device));
  { const int rc = h2d_staged(h, h->stg_cf, h->st.cf, cf, sizeof(double) * h->R * h->t.n_eci); if (rc) return rc; }
  drop_trials(h);
  return refresh_energy(h);
}

int cemc_get_cf(cemc_handle *h, double *cf) {
  if (!h || !cf) return fail("null argument");
  CU(cudaSetDevice(h->device));
  return d2h_staged(h, cf, h->st.cf, sizeof(double) * h->R * h->t.n_eci, false);
}

int cemc_recompute_cf(cemc_handle *h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->device));
  bool done = false;
  if (h->tab_ok && h->t.n_tasks_total <= 256 && !h->force_generic) {
    // product tables available: codes per sub-cluster + table sums (cf_tab_kernel)
    int tasks_pad = 1;
    while (tasks_pad < h->t.n_tasks_total) tasks_pad *= 2;
    const int groups = 256 / tasks_pad;
    int chunks = std::max(1, std::min(16, 2 * h->n_sms / std::max(1, h->R)));
    chunks = std::min(chunks, (h->t.N + CF_TILE - 1) / CF_TILE);
    const int n_slots = chunks * groups;
    size_t sm = (size_t)h->tab.n_tab * sizeof(double) + (size_t)CF_TILE * h->tab.n_sub * sizeof(unsigned short);
    const int occ_in_smem = (sm + (size_t)h->t.N + 16 <= (size_t)h->max_smem_optin) ? 1 : 0;
    if (occ_in_smem) sm += (size_t)h->t.N;
    sm = (sm + 15) / 16 * 16;
    if (sm <= (size_t)h->max_smem_optin) {
      if (h->cf_n_slots < n_slots) {
        CU(cudaStreamSynchronize(h->stream));
        if (h->cf_slots) cudaFree(h->cf_slots);
        h->cf_slots = nullptr; h->cf_n_slots = 0;
        CU(cudaMalloc((void **)&h->cf_slots, sizeof(double) * (size_t)h->R * h->n_jobs * n_slots));
        h->cf_n_slots = n_slots;
      }
      CU(cudaFuncSetAttribute(cf_tab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      cf_tab_kernel<<<dim3(chunks, h->R), 256, sm, h->stream>>>(h->t, h->tab, h->st.occ, h->cf_slots, h->n_jobs,
                                                                n_slots, chunks, tasks_pad, occ_in_smem);
      const int nq = h->R * h->n_jobs;
      cf_slot_sum_kernel<<<(nq + 127) / 128, 128, 0, h->stream>>>(h->cf_slots, h->cf_partial, nq, n_slots);
      h->launches += 2;
      done = true;
    }
  }
  if (!done) {
    dim3 grid(h->n_jobs, h->R);
    cf_partial_kernel<<<grid, 256, 0, h->stream>>>(h->t, h->st.occ, h->cf_partial, h->n_jobs);
    h->launches++;
  }
  cf_final_kernel<<<h->R, 64, 0, h->stream>>>(h->t, h->cf_partial, h->n_jobs, h->st.cf);
  h->launches++;
  CU(cudaGetLastError());
  drop_trials(h);
  return refresh_energy(h);
}

int cemc_set_ecis(cemc_handle *h, const double *eci, int per_replica) {
  if (!h || !eci) return fail("null argument");
  CU(cudaSetDevice(h->device));
  const int n = h->t.n_eci;
  std::vector<double> tmp;
  const double *src = eci;
  if (!per_replica) {
    tmp.resize((size_t)h->R * n);
    for (int r = 0; r < h->R; r++) memcpy(&tmp[(size_t)r * n], eci, sizeof(double) * n);
    src = tmp.data();
  }
  { const int rc = h2d_staged(h, h->stg_eci, h->st.eci, src, sizeof(double) * h->R * n); if (rc) return rc; }
  // the reference always reports dot(ecis, cf) (ce_updater.cpp:236-242): with trial changes
  // pending, the energy an undo restores must be the committed CFs under the NEW ECIs
  for (int r = 0; r < h->R; r++)
    if (!h->trial_log[r].empty()) {
      energy_kernel<<<1, 32, 0, h->stream>>>(1, n, h->t.N, h->st.eci + (size_t)r * n,
                                             h->cf_committed + (size_t)r * n, h->e_committed + r);
      h->launches++;
    }
  CU(cudaGetLastError());
  return refresh_energy(h);
}

int cemc_get_ecis(cemc_handle *h, double *eci) {
  if (!h || !eci) return fail("null argument");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(eci, h->st.eci, sizeof(double) * h->R * h->t.n_eci, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cemc_get_energy(cemc_handle *h, double *energy) {
  if (!h || !energy) return fail("null argument");
  CU(cudaSetDevice(h->device));
  return d2h_staged(h, energy, h->st.e_cur, sizeof(double) * h->R, true);
}

int cemc_set_kT(cemc_handle *h, const double *kT) {
  if (!h || !kT) return fail("null argument");
  CU(cudaSetDevice(h->device));
  for (int r = 0; r < h->R; r++) if (!(kT[r] > 0.0)) return fail("kT must be positive");
  h->order_dirty = true;
  return h2d_staged(h, h->stg_kT, h->st.kT, kT, sizeof(double) * h->R);
}

int cemc_get_kT(cemc_handle *h, double *kT) {
  if (!h || !kT) return fail("null argument");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(kT, h->st.kT, sizeof(double) * h->R, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cemc_debug_phase_cycles(cemc_handle *h, uint64_t *out8) {
  if (!h || !out8) return fail("null argument");
#if defined(CEMC_PHASE_TIMING) || defined(CEMC_WARP_TIMING)
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpy(out8, h->d_phase, (size_t)h->R * 24 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return 0;
#else
  return fail("library built without -DCEMC_PHASE_TIMING");
#endif
}

int cemc_set_autotune(cemc_handle *h, int on) {
  if (!h) return fail("null handle");
  h->autotune = on != 0;
  reset_tuning(h);
  return 0;
}

static const int kNumVariants = 10;  // see launch_variant

int cemc_set_variant(cemc_handle *h, int sgc, int canonical) {
  if (!h) return fail("null handle");
  if (sgc < -1 || sgc >= kNumVariants || canonical < -1 || canonical >= kNumVariants) return fail("no such kernel variant");
  h->tuned_sgc = sgc;
  h->tuned_can = canonical;
  return 0;
}

int cemc_set_lattice_arithmetic(cemc_handle *h, int on) {
  if (!h) return fail("null handle");
  h->t.lat_ok = (on != 0 && h->lat_verified) ? 1 : 0;
  reset_tuning(h);
  return 0;
}

int cemc_get_lattice_arithmetic(cemc_handle *h, int *on) {
  if (!h || !on) return fail("null argument");
  *on = h->t.lat_ok;
  return 0;
}

int cemc_set_observe(cemc_handle *h, int on) {
  if (!h) return fail("null handle");
  h->observe = on != 0;
  return 0;
}

static bool batch_applicable(const cemc_handle *h) {
  const bool spin_eval = h->spin_ok && !h->no_spin && h->t.allowed_identity && h->spin.n_rounds <= 4;
  const bool tab_eval = h->tab_ok && (h->fp32 || !h->no_tab);
  // K <= 31 translation columns (one per lane); the spin and table evaluations also take 32..63
  // up to 32 ECIs (one per lane); the table / product evaluations also take 33..64 (two per lane, K <= 31)
  if (h->t.n_eci > 32 && (h->t.n_eci > 64 || spin_eval || h->t.KP > 32)) return false;
  // several translational symmetry groups: table evaluation only; background sites: generic kernel
  if (h->t.n_active != h->t.N || (h->t.n_symm > 1 && !tab_eval)) return false;
  return !(h->force_generic || h->t.S > 8 || h->batch < 0 ||
           h->t.KP > ((spin_eval || tab_eval) ? 64 : 32));
}

int cemc_batch_applicable(cemc_handle *h, int *yes) {
  if (!h || !yes) return fail("null argument");
  *yes = batch_applicable(h) ? 1 : 0;
  return 0;
}

int cemc_last_variant(cemc_handle *h, int *variant) {
  if (!h || !variant) return fail("null argument");
  *variant = h->last_variant;
  return 0;
}

int cemc_get_variant(cemc_handle *h, int *sgc, int *canonical) {
  if (!h) return fail("null handle");
  if (sgc) *sgc = h->tuned_sgc;
  if (canonical) *canonical = h->tuned_can;
  return 0;
}

int cemc_set_cluster(cemc_handle *h, int c) {
  if (!h) return fail("null handle");
  if (c < 0 || c > 2) return fail("cluster size must be 0 (auto), 1 or 2");
  reset_tuning(h);
  h->cluster = c;
  return 0;
}

int cemc_set_spin_kernel(cemc_handle *h, int on) {
  if (!h) return fail("null handle");
  h->no_spin = (on == 0);
  reset_tuning(h);
  return 0;
}

int cemc_set_table_eval(cemc_handle *h, int on) {
  if (!h) return fail("null handle");
  h->no_tab = (on == 0);
  reset_tuning(h);
  return 0;
}

int cemc_set_replica_order(cemc_handle *h, const int32_t *order) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->device));
  if (!order) { h->d_order = nullptr; return 0; }
  std::vector<char> seen(h->R, 0);
  for (int i = 0; i < h->R; i++) {
    if (order[i] < 0 || order[i] >= h->R || seen[order[i]]) return fail("replica order must be a permutation of 0..R-1");
    seen[order[i]] = 1;
  }
  if (!h->d_order_buf) { int rc = dalloc(h, &h->d_order_buf, h->R); if (rc) return rc; }
  CU(cudaMemcpyAsync(h->d_order_buf, order, sizeof(int32_t) * h->R, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->d_order = h->d_order_buf;
  return 0;
}

int cemc_set_precision(cemc_handle *h, int bits) {
  if (!h) return fail("null handle");
  if (bits != 32 && bits != 64) return fail("precision must be 32 or 64");
  const bool spin = h->spin_ok && h->t.allowed_identity && h->spin.n_rounds <= 4;
  if (bits == 32 && !spin && !h->tab_ok)
    return fail("the fp32 variant needs the table evaluation (<= 32 ECIs, one symmetry group, tables in shared memory)");
  h->fp32 = (bits == 32);
  reset_tuning(h);
  return 0;
}

int cemc_get_batch_eval(cemc_handle *h, int *ev) {
  if (!h || !ev) return fail("null argument");
  if (h->spin_ok && !h->no_spin && h->t.allowed_identity && h->spin.n_rounds <= 4) *ev = EV_SPIN;
  else if (h->tab_ok && h->fp32) *ev = EV_TAB32;
  else if (h->tab_ok && !h->no_tab) *ev = EV_TAB;
  else *ev = EV_PRODUCT;
  return 0;
}

int cemc_set_screen_slack(cemc_handle *h, double factor) {
  if (!h) return fail("null handle");
  if (!(factor >= 1.0)) return fail("screen slack must be >= 1");
  h->screen_slack = factor;
  return 0;
}

int cemc_set_batch(cemc_handle *h, int b) {
  if (!h) return fail("null handle");
  if (!(b == -1 || b == 0 || b == 4 || b == 8 || b == 16)) return fail("batch must be -1 (off), 0 (auto), 4, 8 or 16");
  reset_tuning(h);
  h->batch = b;
  return 0;
}

int cemc_set_generic_path(cemc_handle *h, int on) {
  if (!h) return fail("null handle");
  h->force_generic = on != 0;
  reset_tuning(h);
  return 0;
}

int cemc_selftest_division(cemc_handle *h, uint64_t seed, int n_blocks, int iters, uint64_t *mismatches) {
  if (!h || !mismatches) return fail("null argument");
  CU(cudaSetDevice(h->device));
  // denominators the kernels really use: count * N_g of every term, N, and some kT values
  std::vector<double> dens;
  std::vector<double2> fd((size_t)h->t.n_symm * h->t.n_eci);
  CU(cudaMemcpy(fd.data(), h->t.fin_d, fd.size() * sizeof(double2), cudaMemcpyDeviceToHost));
  for (auto &f : fd) dens.push_back(f.y);
  dens.push_back((double)h->t.N);
  for (int k = 1; k <= 16; k++) dens.push_back(8.617330337217213e-05 * 100.0 * k);
  double *d_dens; unsigned long long *d_bad;
  CU(cudaMalloc((void **)&d_dens, dens.size() * sizeof(double)));
  CU(cudaMalloc((void **)&d_bad, sizeof(unsigned long long)));
  CU(cudaMemcpy(d_dens, dens.data(), dens.size() * sizeof(double), cudaMemcpyHostToDevice));
  CU(cudaMemset(d_bad, 0, sizeof(unsigned long long)));
  exact_div_selftest_kernel<<<n_blocks, 256, 0, h->stream>>>(seed, iters, d_dens, (int)dens.size(), d_bad);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  unsigned long long bad = 0;
  CU(cudaMemcpy(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost));
  cudaFree(d_dens); cudaFree(d_bad);
  *mismatches = bad;
  return 0;
}

int cemc_set_block_threads(cemc_handle *h, int n) {
  if (!h) return fail("null handle");
  if (n != 0 && (n < 32 || n > 256 || n % 32)) return fail("block threads must be 0 (auto) or a multiple of 32 in [32, 256]");
  h->block_threads = n;
  return 0;
}

int cemc_seed(cemc_handle *h, uint64_t seed) {
  if (!h) return fail("null handle");
  h->seed = seed;
  return 0;
}

int cemc_set_sgc_species(cemc_handle *h, int n_allowed, const int8_t *allowed) {
  if (!h || !allowed) return fail("null argument");
  if (n_allowed < 2 || n_allowed > h->t.S) return fail("At least 2 symbols have to be specified");
  for (int i = 0; i < n_allowed; i++)
    if (allowed[i] < 0 || allowed[i] >= h->t.S) return fail("species id out of range");
  CU(cudaSetDevice(h->device));
  for (int i = 0; i < n_allowed; i++)
    for (int j = 0; j < i; j++) if (allowed[i] == allowed[j]) return fail("duplicate species");
  CU(cudaStreamSynchronize(h->stream));
  h->allowed.assign(allowed, allowed + n_allowed);
  return upload_allowed(h);
}

int cemc_get_counters(cemc_handle *h, uint64_t *steps, uint64_t *accepted) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->device));
  if (steps) CU(cudaMemcpyAsync(steps, h->st.step, sizeof(uint64_t) * h->R, cudaMemcpyDeviceToHost, h->stream));
  if (accepted) CU(cudaMemcpyAsync(accepted, h->st.accepted, sizeof(uint64_t) * h->R, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cemc_reset_counters(cemc_handle *h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->device));
  CU(cudaMemsetAsync(h->st.accepted, 0, sizeof(uint64_t) * h->R, h->stream));
  return 0;
}

int cemc_set_step(cemc_handle *h, const uint64_t *steps) {
  if (!h || !steps) return fail("null argument");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(h->st.step, steps, sizeof(uint64_t) * h->R, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

}  // extern "C"  (templates below need C++ linkage)

// ---------------------------------------------------------------------------
static int n_threads_for(const cemc_handle *h, int sites_changed) {
  if (h->block_threads > 0) return h->block_threads;
  // one pass over the products of a move, but never fewer threads than gather columns
  const int items = std::max({sites_changed * h->t.KP, sites_changed * h->t.max_items, 32});
  return std::min(256, (items + 31) / 32 * 32);
}

template <int MODE, bool kSmem, bool kTree, bool kFast>
static int launch_one(cemc_handle *h, const ReplicaState &st, const RunArgs &a, int n_rep, int nthr, size_t sm) {
  CU(cudaFuncSetAttribute(mc_kernel<MODE, kSmem, kTree, kFast>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  mc_kernel<MODE, kSmem, kTree, kFast><<<n_rep, nthr, sm, h->stream>>>(h->t, st, a, h->acc_stride);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

template <int MODE, bool kSmem, bool kTree>
static int launch_fast(cemc_handle *h, const ReplicaState &st, const RunArgs &a, int n_rep, int nthr, size_t sm) {
  // register-resident CF vector: one lane of warp 0 per ECI
  const bool fast = h->t.n_eci <= 32 && h->t.uniform_group && !h->force_generic;
  return fast ? launch_one<MODE, kSmem, kTree, true>(h, st, a, n_rep, nthr, sm)
              : launch_one<MODE, kSmem, kTree, false>(h, st, a, n_rep, nthr, sm);
}

template <int MODE>
static int launch_mc(cemc_handle *h, const RunArgs &a, int first_replica, int n_rep) {
  const bool canonical = (MODE == MODE_CANONICAL);
  const size_t sm_state = smem_bytes(h->t, h->acc_stride, canonical, true);
  const bool in_smem = sm_state <= (size_t)h->max_smem_optin;
  const size_t sm = in_smem ? sm_state : smem_bytes(h->t, h->acc_stride, canonical, false);
  if (sm > (size_t)h->max_smem_optin) return fail("cluster program does not fit in shared memory");
  const int nthr = n_threads_for(h, MODE == MODE_SGC ? 1 : 2);
  // TREE sums are used when asked for, or when they are provably bit-identical
  const bool tree = (h->order_mode == CEMC_ORDER_TREE) || h->integer_bf;
  ReplicaState st = h->st;
  if (first_replica) {      // sub-range launch (trial API): offset the replica-major pointers
    const size_t r0 = first_replica;
    st.occ += r0 * h->t.N; st.cf += r0 * h->t.n_eci; st.eci += r0 * h->t.n_eci; st.e_cur += r0;
    st.kT += r0; st.acc += r0 * h->acc_stride; st.ref += r0; st.step += r0; st.accepted += r0;
    st.list += r0 * h->t.N; st.loc += r0 * h->t.N; st.off += r0 * (h->t.S + 1); st.status += r0;
  }
  if (in_smem) return tree ? launch_fast<MODE, true, true>(h, st, a, n_rep, nthr, sm)
                           : launch_fast<MODE, true, false>(h, st, a, n_rep, nthr, sm);
  return tree ? launch_fast<MODE, false, true>(h, st, a, n_rep, nthr, sm)
              : launch_fast<MODE, false, false>(h, st, a, n_rep, nthr, sm);
}

template <int MODE, int NR>
static int launch_spin_nr(cemc_handle *h, const RunArgs &a, size_t sm) {
  CU(cudaFuncSetAttribute(spin_kernel<MODE, NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  spin_kernel<MODE, NR><<<h->R, 32, sm, h->stream>>>(h->spin, h->t, h->st, a, h->acc_stride);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

// returns -1 when the spin kernel is not applicable (caller falls back to the generic kernel)
template <int MODE>
static int launch_spin(cemc_handle *h, const RunArgs &a) {
  if (!h->spin_ok || h->force_generic || h->no_spin || !h->t.allowed_identity) return -1;
  const size_t N = (size_t)h->t.N;
  const size_t sm = 1024 + 512 * (size_t)h->spin.wq + 4 * (size_t)((h->spin.n_items + 3) & ~3) +
                    (MODE == MODE_CANONICAL ? 4 * ((N + 3) & ~(size_t)3) : 0) + ((N + 15) & ~(size_t)15);
  if (sm > (size_t)h->max_smem_optin) return -1;
  const int nr = h->spin.n_rounds * 1;
  if (nr <= 1) return launch_spin_nr<MODE, 1>(h, a, sm);
  if (nr <= 2) return launch_spin_nr<MODE, 2>(h, a, sm);
  if (nr <= 4) return launch_spin_nr<MODE, 4>(h, a, sm);
  return -1;
}

static RunArgs run_args(cemc_handle *h, long long n_steps) {
  RunArgs a{};
  a.n_steps = n_steps; a.seed = h->seed; a.replica_offset = (uint32_t)h->replica_offset; a.replica_stride = (uint32_t)h->replica_stride;
  a.observe = h->observe ? 1 : 0;
  a.screen_slack = h->screen_slack;
  a.phase = h->d_phase;
  a.order = h->d_order;
  if (h->obs_interval > 0) {
    a.obs_interval = h->obs_interval; a.obs_origin = h->obs_step; a.obs_flags = h->obs_flags;
    a.obs_ring = h->obs_ring; a.ob_n = h->ob_n; a.ob_snap_cf = h->ob_snap_cf; a.ob_snap_e = h->ob_snap_e;
    a.ob_snap_occ = h->ob_snap_occ;
  }
  if (h->trace_capacity > 0) {
    a.tr_sites = h->tr_sites; a.tr_news = h->tr_news; a.tr_u = h->tr_u; a.tr_acc = h->tr_acc;
    a.tr_e = h->tr_e; a.tr_capacity = h->trace_capacity;
  }
  return a;
}

// speculative batch kernel with B moves per CTA and C CTAs per chain; -1 when not applicable
template <int MODE>
static int launch_batch(cemc_handle *h, const RunArgs &a, int B, int C, int M = 1, int split = 0) {
  // K <= 31 translation columns (one per lane); the spin evaluation also takes 32..63 (two per lane)
  if (!batch_applicable(h)) return -1;
  BatchLaunch L{};
  L.mode = MODE; L.B = B; L.C = C; L.M = M; L.split = split; L.R = h->R; L.max_smem_optin = h->max_smem_optin;
  L.extras = (a.rp_sites != nullptr || h->t.lat_ok || a.obs_interval > 0) ? 1 : 0;
  L.tree = ((h->order_mode == CEMC_ORDER_TREE) || h->integer_bf) ? 1 : 0;
  L.stream = h->stream; L.t = h->t; L.st = h->st; L.a = a; L.acc_stride = h->acc_stride;
  L.sp = h->spin; L.tb = h->tab;
  if (!h->d_order && C == 1 && h->R > h->n_sms && h->R <= 4096) {      // more chains than SMs: hottest first
    if (!h->d_order_auto) { const int rc0 = dalloc(h, &h->d_order_auto, h->R); if (rc0) return rc0; h->order_dirty = true; }
    if (h->order_dirty) {
      order_kernel<<<1, 256, 0, h->stream>>>(h->R, h->st.kT, h->d_order_auto);
      h->launches++;
      CU(cudaGetLastError());
      h->order_dirty = false;
    }
    L.a.order = h->d_order_auto;
  }
  int rc;
  if (h->spin_ok && !h->no_spin && h->t.allowed_identity && h->spin.n_rounds <= 4)
    rc = batch_launch_spin(L);          // binary +-1 basis: spin evaluation
  else if (h->tab_ok && h->fp32)
    rc = batch_launch_tab32(L);         // fp32 product tables (opt-in)
  else if (h->tab_ok && !h->no_tab)
    rc = batch_launch_tab(L);           // product tables
  else
    rc = batch_launch_product(L);       // fp64 products
  if (rc == 0) h->launches++;
  if (rc >= 1000) return fail(std::string("batch kernel launch: ") + cudaGetErrorString((cudaError_t)(rc - 1000)), 100);
  return rc;
}

// Kernel variants of one sampler.  All of them produce the same trajectory bit for
// bit, so the choice is a pure performance knob: 0 spin, 1..4 batch (B,C) =
// (16,2) (16,1) (8,1) (4,1), 5 one move at a time (mc_kernel, always applicable),
// 6..7 retired, 8 / 9 batch (16,2) / (8,2) with the two changed sites of a swap split over the
// two CTAs of a cluster (canonical only; 9 = short batches for hot chains on small cells).

template <int MODE>
static int launch_variant_raw(cemc_handle *h, const RunArgs &a, int v);

template <int MODE>
static int launch_variant(cemc_handle *h, const RunArgs &a, int v) {
  // the kernels count the moves of one launch in 32 bits
  const long long kMaxLaunch = 1ll << 30;
  if (a.obs_interval <= 0) {
    if (a.n_steps > kMaxLaunch && (a.rp_sites || a.tr_acc || a.tr_e || a.tr_sites || a.tr_news || a.tr_u))
      return fail("traced / replayed runs are limited to 2^30 moves per call");
    long long done = 0;
    while (done < a.n_steps) {
      RunArgs part = a;
      part.n_steps = std::min<long long>(a.n_steps - done, kMaxLaunch);
      const int rc = launch_variant_raw<MODE>(h, part, v);
      if (rc) return rc;
      h->last_variant = v;
      done += part.n_steps;
    }
    return 0;
  }
  // device observers: a launch must not cross more boundaries than the snapshot ring holds; after
  // every launch the (tiny) fold kernel turns the snapshots into the observers' sums, in order
  const long long cap = std::min<long long>((long long)h->obs_ring * a.obs_interval, kMaxLaunch);
  long long done = 0;
  while (done < a.n_steps) {
    RunArgs part = a;
    part.n_steps = std::min<long long>(a.n_steps - done, cap);
    part.obs_origin = h->obs_step;
    const int rc = launch_variant_raw<MODE>(h, part, v);
    if (rc) return rc;
    h->last_variant = v;
    h->obs_step += part.n_steps;                           // observer step counter (boundaries span launches)
    done += part.n_steps;
    ObserverSums o{h->ob_folded, h->ob_cf_sum, h->ob_cf_sq, h->ob_best, h->ob_best_occ, h->ob_e,
                   h->obs_capacity, h->ob_order, h->ob_occ_ref};
    observer_fold_kernel<<<h->R, 32, 0, h->stream>>>(part, o, h->t.n_eci, h->t.N);
    h->launches++;
    CU(cudaGetLastError());
  }
  return 0;
}

template <int MODE>
static int launch_variant_raw(cemc_handle *h, const RunArgs &a, int v) {
  switch (v) {
    case 0: return launch_spin<MODE>(h, a);
    case 1: return (2 * h->R <= h->n_sms || h->cluster == 2) ? launch_batch<MODE>(h, a, 16, 2) : -1;
    case 2: return launch_batch<MODE>(h, a, 16, 1);
    case 3: return launch_batch<MODE>(h, a, 8, 1);
    case 4: return launch_batch<MODE>(h, a, 4, 1);
    case 6: return launch_batch<MODE>(h, a, 8, 1, 2);      // (8,1) with two moves per evaluation warp (spin evaluation)
    case 7: return -1;              // retired ((16,1) with two moves per warp: never the fastest)
    case 8: return (MODE == MODE_CANONICAL && (2 * h->R <= h->n_sms || h->cluster == 2))
                       ? launch_batch<MODE>(h, a, 16, 2, 1, 1) : -1;
    case 9: return (MODE == MODE_CANONICAL && (2 * h->R <= h->n_sms || h->cluster == 2))
                       ? launch_batch<MODE>(h, a, 8, 2, 1, 1) : -1;
    default: return launch_mc<MODE>(h, a, 0, h->R);
  }
}

static bool variant_allowed(const cemc_handle *h, int v) {
  // fp32 variant of a non-binary system: only the batch kernels evaluate in fp32; mixing in
  // the fp64 kernels would make the result depend on the tuner's timing
  if (h->fp32 && !h->spin_ok && (v == 0 || v == 5)) return false;
  if (v == 0 && h->batch > 0) return false;       // an explicit batch size asks for the batch kernel
  if (v == 0 && h->obs_interval > 0) return false;   // the warp-per-replica kernel has no observer boundaries
  if ((v >= 1 && v <= 4) || v >= 6) {
    static const int Bs[10] = {0, 16, 16, 8, 4, 0, 8, 16, 16, 8}, Cs[10] = {0, 2, 1, 1, 1, 0, 1, 1, 2, 2};
    if (h->batch > 0 && h->batch != Bs[v]) return false;
    if (h->cluster > 0 && h->cluster != Cs[v]) return false;
  }
  return true;
}

// Run n_steps with the preferred variant; on long runs without a trace, time the
// applicable variants on short segments of the run itself (productive work, the
// trajectory does not depend on the variant) and keep the fastest.
template <int MODE>
static int run_tuned(cemc_handle *h, long long n_steps) {
  int &best = (MODE == MODE_SGC) ? h->tuned_sgc : h->tuned_can;
  const int md = (MODE == MODE_SGC) ? 0 : 1;
  const long long seg = 2048, warm = seg / 4;
  long long done = 0;
  int provisional = -1;
  // Long launches: every applicable variant gets an untimed warm-up launch (first-use costs) and
  // two timed segments of which the faster one counts (one sample alone is at the mercy of a
  // noisy neighbour).  The tuning moves are part of the run -- never more than n_steps: what
  // does not fit this call continues in the next one (lt_next), the rest of this call runs on
  // the fastest variant so far.
  if (best < 0 && h->autotune && h->trace_capacity == 0 && n_steps >= warm + seg) {
    int &next = h->lt_next[md];
    if (next == 0) { h->lt_best_ms[md] = 1e30f; h->lt_best[md] = -1; }
    while (next < kNumVariants && n_steps - done >= warm + seg) {
      const int v = next++;
      if (!variant_allowed(h, v)) continue;
      int rc = launch_variant<MODE>(h, run_args(h, warm), v);
      if (rc == -1) continue;
      if (rc) return rc;
      done += warm;
      float ms = 1e30f;
      for (int rep_ = 0; rep_ < 2 && n_steps - done >= seg; rep_++) {
        CU(cudaEventRecord(h->tv0, h->stream));
        rc = launch_variant<MODE>(h, run_args(h, seg), v);
        if (rc) return rc;
        CU(cudaEventRecord(h->tv1, h->stream));
        CU(cudaEventSynchronize(h->tv1));
        float m1 = 0.f;
        CU(cudaEventElapsedTime(&m1, h->tv0, h->tv1));
        ms = std::min(ms, m1);
        done += seg;
      }
      if (ms < h->lt_best_ms[md]) { h->lt_best_ms[md] = ms; h->lt_best[md] = v; }
    }
    if (next >= kNumVariants) best = h->lt_best[md];
    else provisional = h->lt_best[md];
  }
  // short launches (e.g. the 1728-move legs between parallel-tempering exchanges): tune
  // across calls -- every call runs one untested variant (a quarter of the moves untimed
  // as warm-up, the rest timed); when all are timed the fastest is kept
  if (best < 0 && done == 0 && h->autotune && h->trace_capacity == 0 && n_steps >= 256 && n_steps < warm + seg) {
    int &next = h->xt_next[MODE == MODE_SGC ? 0 : 1];
    float *xms = h->xt_ms[MODE == MODE_SGC ? 0 : 1];
    while (next < kNumVariants) {
      const int v = next++;
      xms[v] = 1e30f;
      if (!variant_allowed(h, v)) continue;
      const long long n1 = n_steps / 4;
      int rc = launch_variant<MODE>(h, run_args(h, n1), v);
      if (rc == -1) continue;
      if (rc) return rc;
      CU(cudaEventRecord(h->tv0, h->stream));
      rc = launch_variant<MODE>(h, run_args(h, n_steps - n1), v);
      if (rc) return rc;
      CU(cudaEventRecord(h->tv1, h->stream));
      CU(cudaEventSynchronize(h->tv1));
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, h->tv0, h->tv1));
      xms[v] = ms / (float)(n_steps - n1);
      done = n_steps;
      break;
    }
    if (next >= kNumVariants) {
      float bm = 1e30f;
      for (int v = 0; v < kNumVariants; v++) if (xms[v] < bm) { bm = xms[v]; best = v; }
    }
  }
  if (done >= n_steps) return 0;
  RunArgs a = run_args(h, n_steps - done);
  if (best < 0 && provisional >= 0) {                // tuning continues in the next call
    const int rc = launch_variant<MODE>(h, a, provisional);
    if (rc != -1) return rc;
  }
  if (best >= 0 && variant_allowed(h, best)) {       // (a tuned / pinned variant can become inapplicable:
    const int rc = launch_variant<MODE>(h, a, best);  //  e.g. the spin kernel once device observers are on)
    if (rc != -1) return rc;
  }
  for (int v = 0; v < kNumVariants; v++) {          // default preference order
    if (!variant_allowed(h, v)) continue;
    const int rc = launch_variant<MODE>(h, a, v);
    if (rc != -1) return rc;
  }
  return fail("no kernel variant applicable");
}

static int ensure_scratch(cemc_handle *h, long long n_steps) {
  if (n_steps <= h->scratch_steps) return 0;
  void *old[] = {h->d_sites, h->d_news, h->d_u, h->d_acc, h->d_e};
  CU(cudaStreamSynchronize(h->stream));
  for (void *p : old) if (p) cudaFree(p);
  // a failed allocation below must not leave dangling pointers for cemc_destroy
  h->d_sites = nullptr; h->d_news = nullptr; h->d_u = nullptr; h->d_acc = nullptr; h->d_e = nullptr;
  h->scratch_steps = 0;
  const size_t n = (size_t)h->R * n_steps;
  CU(cudaMalloc((void **)&h->d_sites, n * 2 * sizeof(int32_t)));
  CU(cudaMalloc((void **)&h->d_news, n * 2));
  CU(cudaMalloc((void **)&h->d_u, n * sizeof(double)));
  CU(cudaMalloc((void **)&h->d_acc, n));
  CU(cudaMalloc((void **)&h->d_e, n * sizeof(double)));
  h->scratch_steps = n_steps;
  return 0;
}

extern "C" {

int cemc_replay(cemc_handle *h, int n_steps, const int32_t *sites, const int8_t *news,
                const double *uniforms, uint8_t *accepted_out, double *e_after_out) {
  if (!h || !sites || !news || !uniforms) return fail("null argument");
  if (n_steps <= 0) return 0;
  CU(cudaSetDevice(h->device));
  int rc;
  if ((rc = ensure_scratch(h, n_steps))) return rc;
  const size_t n = (size_t)h->R * n_steps;
  CU(cudaMemcpyAsync(h->d_sites, sites, n * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_news, news, n * 2, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_u, uniforms, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  RunArgs a{};
  a.n_steps = n_steps; a.seed = h->seed; a.replica_offset = (uint32_t)h->replica_offset; a.replica_stride = (uint32_t)h->replica_stride;
  a.observe = 1;
  a.screen_slack = h->screen_slack;
  a.phase = h->d_phase;
  a.rp_sites = h->d_sites; a.rp_news = h->d_news; a.rp_u = h->d_u;
  a.tr_acc = h->d_acc; a.tr_e = h->d_e; a.tr_capacity = n_steps;
  drop_trials(h);
  // The recorded trajectory runs through the SAME kernels the samplers use (spin / batch
  // variants; the pinned one, else the default preference order) when every step has the
  // same shape -- all one-site (SGC kernels) or all two-site (canonical kernels) -- and is
  // in range; everything else (mixed records, bad input, background sites, pinned variant 5)
  // takes the generic one-move-at-a-time kernel, which also reports the errors.
  bool all1 = true, all2 = true, valid = true;
  for (size_t q = 0; q < n && valid; q++) {
    const int s0 = sites[2 * q], s1 = sites[2 * q + 1];
    if (s0 < 0 || s0 >= h->t.N || s1 >= h->t.N || news[2 * q] < 0 || news[2 * q] >= h->t.S) valid = false;
    if (s1 >= 0) { all1 = false; if (news[2 * q + 1] < 0 || news[2 * q + 1] >= h->t.S) valid = false; }
    else all2 = false;
  }
  int launched = -1;
  if (valid && (all1 || all2) && h->t.n_active == h->t.N) {
    if (all2) { if ((rc = ensure_tracker(h))) return rc; }      // offsets read at kernel start
    const int pinned = all1 ? h->tuned_sgc : h->tuned_can;
    for (int k = -1; k < kNumVariants && launched < 0; k++) {
      const int v = k < 0 ? pinned : k;
      if (v < 0 || v == 5 || (k >= 0 && pinned >= 0)) continue;
      if (!variant_allowed(h, v)) continue;
      rc = all1 ? launch_variant<MODE_SGC>(h, a, v) : launch_variant<MODE_CANONICAL>(h, a, v);
      if (rc == 0) launched = v;
      else if (rc != -1) return rc;
    }
  }
  h->tracker_dirty = true;
  if (launched < 0) {
    if ((rc = launch_mc<MODE_REPLAY>(h, a, 0, h->R))) return rc;
    h->last_variant = 5;
  }
  if (accepted_out) CU(cudaMemcpyAsync(accepted_out, h->d_acc, n, cudaMemcpyDeviceToHost, h->stream));
  if (e_after_out) CU(cudaMemcpyAsync(e_after_out, h->d_e, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return check_status(h);
}


int cemc_run_sgc(cemc_handle *h, int64_t n_steps) {
  if (!h) return fail("null handle");
  if (n_steps <= 0) return 0;
  CU(cudaSetDevice(h->device));
  drop_trials(h);
  h->tracker_dirty = true;
  return run_tuned<MODE_SGC>(h, n_steps);
}

int cemc_run_canonical(cemc_handle *h, int64_t n_steps) {
  if (!h) return fail("null handle");
  if (n_steps <= 0) return 0;
  CU(cudaSetDevice(h->device));
  drop_trials(h);
  { const int rc0 = ensure_tracker(h); if (rc0) return rc0; }
  return run_tuned<MODE_CANONICAL>(h, n_steps);
}

// Checkpoint support: the per-species site lists of the canonical sampler
// (SwapMoveIndexTracker.tracker, swap_move_index_tracker.py:8-36) are chain state.
int cemc_get_tracker(cemc_handle *h, int32_t *list, int32_t *off) {
  if (!h || !list || !off) return fail("null argument");
  CU(cudaSetDevice(h->device));
  int rc = ensure_tracker(h);
  if (rc) return rc;
  CU(cudaMemcpyAsync(list, h->st.list, sizeof(int32_t) * h->R * h->t.N, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(off, h->st.off, sizeof(int32_t) * h->R * (h->t.S + 1), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cemc_set_tracker(cemc_handle *h, const int32_t *list) {
  if (!h || !list) return fail("null argument");
  CU(cudaSetDevice(h->device));
  const int N = h->t.N, S = h->t.S;
  std::vector<int8_t> occ((size_t)h->R * N);
  CU(cudaMemcpyAsync(occ.data(), h->st.occ, occ.size(), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  std::vector<int32_t> loc((size_t)h->R * N, -1), off((size_t)h->R * (S + 1), 0);
  for (int r = 0; r < h->R; r++) {
    const int8_t *o = &occ[(size_t)r * N];
    int32_t *of = &off[(size_t)r * (S + 1)];
    std::vector<int> cnt(S, 0);
    int n_act = 0;
    for (int a = 0; a < N; a++) if (h->symm_of_site[a] >= 0) { cnt[o[a]]++; n_act++; }
    for (int sp = 0; sp < S; sp++) of[sp + 1] = of[sp] + cnt[sp];
    const int32_t *ls = list + (size_t)r * N;
    for (int sp = 0; sp < S; sp++)
      for (int k = of[sp]; k < of[sp + 1]; k++) {
        const int a = ls[k];
        if (a < 0 || a >= N || h->symm_of_site[a] < 0 || o[a] != sp || loc[(size_t)r * N + a] != -1)
          return fail("The atom position tracker does not match the current state");
        loc[(size_t)r * N + a] = k - of[sp];
      }
    (void)n_act;
  }
  CU(cudaMemcpyAsync(h->st.list, list, sizeof(int32_t) * h->R * N, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->st.loc, loc.data(), sizeof(int32_t) * h->R * N, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->st.off, off.data(), sizeof(int32_t) * h->R * (S + 1), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->tracker_dirty = false;
  return 0;
}

int cemc_set_trace(cemc_handle *h, int64_t capacity) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  void *old[] = {h->tr_sites, h->tr_news, h->tr_u, h->tr_acc, h->tr_e};
  for (void *p : old) if (p) cudaFree(p);
  h->tr_sites = nullptr; h->tr_news = nullptr; h->tr_u = nullptr; h->tr_acc = nullptr; h->tr_e = nullptr;
  h->trace_capacity = 0;
  if (capacity <= 0) return 0;
  const size_t n = (size_t)h->R * capacity;
  CU(cudaMalloc((void **)&h->tr_sites, n * 2 * sizeof(int32_t)));
  CU(cudaMalloc((void **)&h->tr_news, n * 2));
  CU(cudaMalloc((void **)&h->tr_u, n * sizeof(double)));
  CU(cudaMalloc((void **)&h->tr_acc, n));
  CU(cudaMalloc((void **)&h->tr_e, n * sizeof(double)));
  h->trace_capacity = capacity;
  return 0;
}

int cemc_get_trace(cemc_handle *h, int64_t n_steps, int32_t *sites, int8_t *news, double *u,
                   uint8_t *accepted, double *e_after) {
  if (!h) return fail("null handle");
  if (n_steps > h->trace_capacity) return fail("trace capacity exceeded");
  CU(cudaSetDevice(h->device));
  const size_t cap = (size_t)h->trace_capacity;
  for (int r = 0; r < h->R; r++) {
    const size_t so = (size_t)r * cap, d_o = (size_t)r * n_steps;
    if (sites) CU(cudaMemcpyAsync(sites + 2 * d_o, h->tr_sites + 2 * so, n_steps * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (news) CU(cudaMemcpyAsync(news + 2 * d_o, h->tr_news + 2 * so, n_steps * 2, cudaMemcpyDeviceToHost, h->stream));
    if (u) CU(cudaMemcpyAsync(u + d_o, h->tr_u + so, n_steps * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (accepted) CU(cudaMemcpyAsync(accepted + d_o, h->tr_acc + so, n_steps, cudaMemcpyDeviceToHost, h->stream));
    if (e_after) CU(cudaMemcpyAsync(e_after + d_o, h->tr_e + so, n_steps * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int cemc_energy_autocorrelation(cemc_handle *h, int64_t n_steps, double *out) {
  if (!h || !out) return fail("null argument");
  if (n_steps < 2 || n_steps > h->trace_capacity) return fail("the trace does not hold that many steps");
  if (n_steps > 0x7fffffff) return fail("window too long");
  CU(cudaSetDevice(h->device));
  double *d_out = nullptr;
  CU(cudaMalloc((void **)&d_out, sizeof(double) * 4 * h->R));
  autocorr_kernel<<<h->R, 256, 0, h->stream>>>(h->tr_e, h->trace_capacity, (int)n_steps, d_out);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, d_out, sizeof(double) * 4 * h->R, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  cudaFree(d_out);
  return 0;
}

// ---- the reference's per-call surface ---------------------------------------
int cemc_trial_changes(cemc_handle *h, int replica, int n_changes, const int32_t *sites,
                       const int8_t *old_species, const int8_t *new_species, double *energy_out) {
  if (!h) return fail("null handle");
  if (replica < 0 || replica >= h->R) return fail("replica out of range");
  CU(cudaSetDevice(h->device));
  const int n_eci = h->t.n_eci, N = h->t.N;
  auto &log = h->trial_log[replica];
  if (n_changes > 0) {
    if (!sites || !new_species) return fail("null argument");
    if (log.size() + (size_t)n_changes > 999)     // ring of 1000, cf_history_tracker.hpp:56
      return fail("Can't store more trial changes than the history buffer holds");
    // device occupancy is the truth; old_species (if given) must agree with it
    std::vector<int8_t> cur(n_changes);
    for (int i = 0; i < n_changes; i++) {
      if (sites[i] < 0 || sites[i] >= N) return fail("site index out of range");
      if (new_species[i] < 0 || new_species[i] >= h->t.S) return fail("species out of range");
    }
    CU(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n_changes; i++) {
      CU(cudaMemcpy(&cur[i], h->st.occ + (size_t)replica * N + sites[i], 1, cudaMemcpyDeviceToHost));
      for (int j = 0; j < i; j++) if (sites[j] == sites[i]) cur[i] = new_species[j];
      if (old_species && old_species[i] != cur[i])
        return fail("The atom position tracker does not match the current state");
    }
    if (log.empty()) {
      CU(cudaMemcpyAsync(h->cf_committed + (size_t)replica * n_eci, h->st.cf + (size_t)replica * n_eci,
                         sizeof(double) * n_eci, cudaMemcpyDeviceToDevice, h->stream));
      CU(cudaMemcpyAsync(h->e_committed + replica, h->st.e_cur + replica, sizeof(double),
                         cudaMemcpyDeviceToDevice, h->stream));
    }
    // each change is one forced-accept one-site step (ce_updater.cpp:845-852)
    std::vector<int32_t> s2(2 * n_changes);
    std::vector<int8_t> n2(2 * n_changes);
    std::vector<double> uu(n_changes, 0.0);
    for (int i = 0; i < n_changes; i++) {
      s2[2 * i] = sites[i]; s2[2 * i + 1] = -1; n2[2 * i] = new_species[i]; n2[2 * i + 1] = 0;
    }
    int32_t *ds; int8_t *dn; double *du;
    CU(cudaMalloc((void **)&ds, s2.size() * sizeof(int32_t)));
    CU(cudaMalloc((void **)&dn, n2.size()));
    CU(cudaMalloc((void **)&du, uu.size() * sizeof(double)));
    CU(cudaMemcpyAsync(ds, s2.data(), s2.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dn, n2.data(), n2.size(), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(du, uu.data(), uu.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    RunArgs a{};
    a.n_steps = n_changes; a.seed = h->seed; a.force_accept = 1; a.observe = 0; a.replica_stride = 1;
    a.rp_sites = ds; a.rp_news = dn; a.rp_u = du;
    int rc = launch_mc<MODE_REPLAY>(h, a, replica, 1);   // rp arrays hold this replica only
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(ds); cudaFree(dn); cudaFree(du);
    if (rc) return rc;
    if ((rc = check_status(h))) return rc;
    for (int i = 0; i < n_changes; i++)
      if (cur[i] != new_species[i]) log.push_back(TrialEntry{sites[i], cur[i]});
    h->tracker_dirty = true;
  }
  if (energy_out) {
    CU(cudaMemcpyAsync(energy_out, h->st.e_cur + replica, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

int cemc_undo_changes(cemc_handle *h, int replica) {
  if (!h) return fail("null handle");
  if (replica < 0 || replica >= h->R) return fail("replica out of range");
  CU(cudaSetDevice(h->device));
  auto &log = h->trial_log[replica];
  if (log.empty()) return 0;
  const int n = (int)log.size(), n_eci = h->t.n_eci;
  std::vector<int32_t> s(n);
  std::vector<int8_t> v(n);
  for (int i = 0; i < n; i++) { s[i] = log[n - 1 - i].site; v[i] = log[n - 1 - i].old_sp; }  // pop order
  int32_t *ds; int8_t *dv;
  CU(cudaMalloc((void **)&ds, n * sizeof(int32_t)));
  CU(cudaMalloc((void **)&dv, n));
  CU(cudaMemcpyAsync(ds, s.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(dv, v.data(), n, cudaMemcpyHostToDevice, h->stream));
  set_sites_kernel<<<1, 32, 0, h->stream>>>(h->st.occ + (size_t)replica * h->t.N, n, ds, dv);
  h->launches++;
  CU(cudaMemcpyAsync(h->st.cf + (size_t)replica * n_eci, h->cf_committed + (size_t)replica * n_eci,
                     sizeof(double) * n_eci, cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaMemcpyAsync(h->st.e_cur + replica, h->e_committed + replica, sizeof(double),
                     cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  cudaFree(ds); cudaFree(dv);
  log.clear();
  h->tracker_dirty = true;
  return 0;
}

int cemc_clear_history(cemc_handle *h, int replica) {
  if (!h) return fail("null handle");
  if (replica < 0 || replica >= h->R) return fail("replica out of range");
  h->trial_log[replica].clear();
  return 0;
}

// ---- observers -----------------------------------------------------------------
int cemc_reset_accumulators(cemc_handle *h, const double *ref) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->device));
  CU(cudaMemsetAsync(h->st.acc, 0, sizeof(double) * h->R * h->acc_stride, h->stream));
  if (ref) {
    for (int r = 0; r < h->R; r++) if (ref[r] == 0.0) return fail("Averager reference value must be non-zero");
    const int rc = h2d_staged(h, h->stg_ref, h->st.ref, ref, sizeof(double) * h->R);
    if (rc) return rc;
  }
  return 0;
}

int cemc_get_accumulators(cemc_handle *h, double *acc) {
  if (!h || !acc) return fail("null argument");
  CU(cudaSetDevice(h->device));
  return d2h_staged(h, acc, h->st.acc, sizeof(double) * h->R * h->acc_stride, true);
}

// ---- device-side state observers ---------------------------------------------------
static void free_device_observers(cemc_handle *h) {
  void **ps[] = {(void **)&h->ob_n, (void **)&h->ob_folded, (void **)&h->ob_snap_cf, (void **)&h->ob_snap_e,
                 (void **)&h->ob_snap_occ, (void **)&h->ob_cf_sum, (void **)&h->ob_cf_sq, (void **)&h->ob_best,
                 (void **)&h->ob_e, (void **)&h->ob_order, (void **)&h->ob_best_occ, (void **)&h->ob_occ_ref};
  for (void **p : ps) { if (*p) cudaFree(*p); *p = nullptr; }
  h->obs_interval = 0; h->obs_flags = 0; h->obs_capacity = 0; h->obs_step = 0; h->obs_ring = 0;
}

int cemc_reset_device_observers(cemc_handle *h, const int8_t *occ_ref) {
  if (!h) return fail("null handle");
  if (h->obs_interval <= 0) return fail("device observers are not enabled");
  CU(cudaSetDevice(h->device));
  const size_t R = (size_t)h->R, n = (size_t)h->t.n_eci, N = (size_t)h->t.N;
  CU(cudaMemsetAsync(h->ob_n, 0, sizeof(unsigned long long) * R, h->stream));
  CU(cudaMemsetAsync(h->ob_folded, 0, sizeof(unsigned long long) * R, h->stream));
  CU(cudaMemsetAsync(h->ob_cf_sum, 0, sizeof(double) * R * n, h->stream));
  CU(cudaMemsetAsync(h->ob_cf_sq, 0, sizeof(double) * R * n, h->stream));
  CU(cudaMemsetAsync(h->ob_order, 0, sizeof(double) * R * 2, h->stream));
  CU(cudaMemsetAsync(h->ob_e, 0, sizeof(double) * R * (size_t)std::max<long long>(h->obs_capacity, 1), h->stream));
  std::vector<double> best(R * (1 + n), 0.0);
  for (size_t r = 0; r < R; r++) best[r * (1 + n)] = INFINITY;      // LowestEnergyStructure.reset (:150)
  CU(cudaMemcpyAsync(h->ob_best, best.data(), sizeof(double) * best.size(), cudaMemcpyHostToDevice, h->stream));
  if (occ_ref) CU(cudaMemcpyAsync(h->ob_occ_ref, occ_ref, R * N, cudaMemcpyHostToDevice, h->stream));
  else CU(cudaMemcpyAsync(h->ob_occ_ref, h->st.occ, R * N, cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaMemcpyAsync(h->ob_best_occ, h->st.occ, R * N, cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));      // `best` / occ_ref are host buffers of this call
  h->obs_step = 0;
  return 0;
}

int cemc_set_device_observers(cemc_handle *h, int64_t interval, int flags, int64_t capacity) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  free_device_observers(h);
  if (interval <= 0 || flags == 0) return 0;
  if (flags & ~15) return fail("unknown observer flag");
  if (capacity < 0) return fail("negative sample capacity");
  const size_t R = (size_t)h->R, n = (size_t)h->t.n_eci, N = (size_t)h->t.N;
  // snapshot ring: 32 boundaries per launch (fewer when the occupation snapshots would get large)
  const bool need_occ = (flags & (CEMC_OBS_LOWEST | CEMC_OBS_SITE_ORDER)) != 0;
  int ring = 32;
  while (ring > 2 && need_occ && (size_t)ring * R * N > ((size_t)512 << 20)) ring /= 2;
  h->obs_ring = ring;
  CU(cudaMalloc((void **)&h->ob_n, sizeof(unsigned long long) * R));
  CU(cudaMalloc((void **)&h->ob_folded, sizeof(unsigned long long) * R));
  CU(cudaMalloc((void **)&h->ob_snap_cf, sizeof(double) * R * ring * n));
  CU(cudaMalloc((void **)&h->ob_snap_e, sizeof(double) * R * ring));
  if (need_occ) CU(cudaMalloc((void **)&h->ob_snap_occ, R * ring * N));
  CU(cudaMalloc((void **)&h->ob_cf_sum, sizeof(double) * R * n));
  CU(cudaMalloc((void **)&h->ob_cf_sq, sizeof(double) * R * n));
  CU(cudaMalloc((void **)&h->ob_best, sizeof(double) * R * (1 + n)));
  CU(cudaMalloc((void **)&h->ob_e, sizeof(double) * R * (size_t)std::max<int64_t>(capacity, 1)));
  CU(cudaMalloc((void **)&h->ob_order, sizeof(double) * R * 2));
  CU(cudaMalloc((void **)&h->ob_best_occ, R * N));
  CU(cudaMalloc((void **)&h->ob_occ_ref, R * N));
  h->obs_interval = interval; h->obs_flags = flags; h->obs_capacity = capacity;
  return cemc_reset_device_observers(h, nullptr);
}

int cemc_get_device_observers(cemc_handle *h, uint64_t *n_samples, double *cf_sum, double *cf_sq,
                              double *best_energy, double *best_cf, int8_t *best_occ,
                              double *site_order, double *energies, int64_t n_energies) {
  if (!h) return fail("null handle");
  if (h->obs_interval <= 0) return fail("device observers are not enabled");
  CU(cudaSetDevice(h->device));
  const size_t R = (size_t)h->R, n = (size_t)h->t.n_eci, N = (size_t)h->t.N;
  if (n_energies > h->obs_capacity) return fail("more energy samples requested than the capacity");
  std::vector<double> best;
  if (n_samples) CU(cudaMemcpyAsync(n_samples, h->ob_n, sizeof(uint64_t) * R, cudaMemcpyDeviceToHost, h->stream));
  if (cf_sum) CU(cudaMemcpyAsync(cf_sum, h->ob_cf_sum, sizeof(double) * R * n, cudaMemcpyDeviceToHost, h->stream));
  if (cf_sq) CU(cudaMemcpyAsync(cf_sq, h->ob_cf_sq, sizeof(double) * R * n, cudaMemcpyDeviceToHost, h->stream));
  if (best_energy || best_cf) {
    best.resize(R * (1 + n));
    CU(cudaMemcpyAsync(best.data(), h->ob_best, sizeof(double) * best.size(), cudaMemcpyDeviceToHost, h->stream));
  }
  if (best_occ) CU(cudaMemcpyAsync(best_occ, h->ob_best_occ, R * N, cudaMemcpyDeviceToHost, h->stream));
  if (site_order) CU(cudaMemcpyAsync(site_order, h->ob_order, sizeof(double) * R * 2, cudaMemcpyDeviceToHost, h->stream));
  if (energies && n_energies > 0)
    CU(cudaMemcpy2DAsync(energies, sizeof(double) * (size_t)n_energies, h->ob_e, sizeof(double) * (size_t)h->obs_capacity,
                         sizeof(double) * (size_t)n_energies, R, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (size_t r = 0; r < R && !best.empty(); r++) {
    if (best_energy) best_energy[r] = best[r * (1 + n)];
    if (best_cf) memcpy(best_cf + r * n, &best[r * (1 + n) + 1], sizeof(double) * n);
  }
  return check_status(h);
}

// ---- parallel tempering -----------------------------------------------------------
int cemc_pt_exchange(cemc_handle *h, int n_total, const double *energies_dev,
                     int32_t *slot_of_replica_dev, const double *kT_of_slot_dev, int direction,
                     uint64_t round, int32_t *n_accepted_dev) {
  if (!h || !energies_dev || !slot_of_replica_dev || !kT_of_slot_dev) return fail("null argument");
  if (h->replica_offset + (h->R - 1) * h->replica_stride + 1 > n_total) return fail("local replicas exceed n_total");
  if (h->replica_stride > 1 && h->R * h->replica_stride != n_total)
    return fail("round-robin sharding: n_total must be n_replicas * replica_stride");
  CU(cudaSetDevice(h->device));
  if (h->pt_scratch_n < n_total) {
    if (h->pt_scratch) { CU(cudaStreamSynchronize(h->stream)); cudaFree(h->pt_scratch); }
    h->pt_scratch = nullptr; h->pt_scratch_n = 0;
    CU(cudaMalloc((void **)&h->pt_scratch, sizeof(int32_t) * n_total));
    h->pt_scratch_n = n_total;
  }
  h->order_dirty = true;
  pt_exchange_kernel<<<1, 256, 0, h->stream>>>(n_total, energies_dev, slot_of_replica_dev, kT_of_slot_dev,
                                               direction, h->seed, round, h->pt_scratch, h->st.kT,
                                               h->replica_offset, h->replica_stride, h->R, n_accepted_dev);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

int cemc_energy_dev(cemc_handle *h, double **ptr) {
  if (!h || !ptr) return fail("null argument");
  *ptr = h->st.e_cur;
  return 0;
}

// ---- timing ------------------------------------------------------------------------
int cemc_timer_start(cemc_handle *h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->device));
  CU(cudaEventRecord(h->ev0, h->stream));
  return 0;
}

int cemc_timer_stop(cemc_handle *h, float *ms) {
  if (!h || !ms) return fail("null argument");
  CU(cudaSetDevice(h->device));
  CU(cudaEventRecord(h->ev1, h->stream));
  CU(cudaEventSynchronize(h->ev1));
  CU(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return 0;
}

int cemc_launch_count(cemc_handle *h, uint64_t *n) {
  if (!h || !n) return fail("null argument");
  *n = h->launches;
  return 0;
}

}  // extern "C"
