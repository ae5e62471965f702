// cemc_kernels.cuh -- sm_100a kernels of the cluster-expansion Metropolis hot path.
//
// One CTA per replica (independent Markov chain).  Per-replica state lives in
// shared memory for the whole launch: int8 occupations, the running
// correlation-function (CF) vector, the per-species site lists of the
// canonical sampler.  The read-only cluster "program" (families, decorations,
// per-ECI normalisation) is staged once per CTA; the translation matrix stays
// in global memory and is read through the read-only path (L1/L2 resident,
// shared by all replicas).
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

enum Mode : int { MODE_REPLAY = 0, MODE_SGC = 1, MODE_CANONICAL = 2 };

struct Task {            // one (ECI, equivalent decoration) spin-product sum
  uint32_t deco;         // 4 x u8 decoration numbers
  uint16_t fam;          // family id
  uint16_t eci;          // ECI index
};
struct Fam {
  uint16_t n;            // cluster size 2..4
  uint16_t M;            // sub-clusters
  uint32_t pos_off;      // first packed position word
};
struct Fin {             // per (symmetry group, ECI)
  int32_t kind;          // cemc_eci_kind, or -1: cluster not in this group (copy)
  int32_t d;             // singlet decoration number
  int32_t t0, t1;        // task range (relative to the group's first task)
  double scale;          // (double)n / |E|            ce_updater.cpp:400
  double div;            // (double)(count * N_g)      ce_updater.cpp:402
};

struct DeviceTables {    // device pointers, shared by all replicas
  int N, S, D, K, KP, n_eci, n_symm, n_fam, n_pos_words, n_tasks_total, max_tasks;
  const int32_t *trans;        // [N][K]
  const int32_t *symm_of_site; // [N]
  const double *bf;            // [D][S]
  const Task *tasks;           // [n_tasks_total]
  const int32_t *task_base;    // [n_symm+1]
  const Fam *fams;             // [n_fam]
  const uint32_t *pos;         // packed positions, one word per sub-cluster
  const Fin *fin;              // [n_symm][n_eci]
  const int32_t *singlet_idx;  // [n_singlets]
  int n_singlets;
  // SGC proposal support
  int n_active;                // non-background sites
  const int32_t *active;       // [n_active] or nullptr when n_active == N
  int n_allowed;
  const int8_t *allowed;       // [n_allowed]
};

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
  uint32_t replica_offset;
  int force_accept;      // trial semantics: commit every step (CEUpdater::calculate)
  int observe;           // accumulate observers
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
};

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

__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

struct Smem {            // carved from dynamic shared memory
  double *cf, *cfn, *eci, *prod, *A, *diff, *bf, *acc;
  Fin *fin;
  Task *tasks;
  Fam *fams;
  uint32_t *pos;
  int32_t *task_base, *singlet_idx, *off, *present;
  uint32_t *rng;         // [32][8]
  int32_t *list;         // or global
  int8_t *occ;           // or global
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Shared-memory footprint; state_in_smem selects occ/list residency.
inline size_t smem_bytes(const DeviceTables &t, int acc_stride, bool canonical,
                         bool state_in_smem) {
  size_t b = 0;
  b += sizeof(double) * (size_t)t.n_eci * 4;               // cf cfn eci prod
  b += sizeof(double) * (size_t)2 * t.D * t.KP;            // A
  b += sizeof(double) * (size_t)2 * (t.max_tasks > 0 ? t.max_tasks : 1);  // diff
  b += sizeof(double) * (size_t)t.D * t.S;                 // bf
  b += sizeof(double) * (size_t)acc_stride;                // acc
  b += sizeof(Fin) * (size_t)t.n_symm * t.n_eci;
  b = align_up(b, 8);
  b += sizeof(Task) * (size_t)(t.n_tasks_total > 0 ? t.n_tasks_total : 1);
  b += sizeof(Fam) * (size_t)(t.n_fam > 0 ? t.n_fam : 1);
  b += sizeof(uint32_t) * (size_t)(t.n_pos_words > 0 ? t.n_pos_words : 1);
  b += sizeof(int32_t) * (size_t)(t.n_symm + 1);
  b += sizeof(int32_t) * (size_t)(t.n_singlets > 0 ? t.n_singlets : 1);
  b += sizeof(int32_t) * (size_t)(t.S + 1) * 2;            // off, present
  b += sizeof(uint32_t) * 32 * 8;                          // rng ring
  b = align_up(b, 16) + 16;
  if (state_in_smem) {
    if (canonical) b += sizeof(int32_t) * (size_t)t.N;
    b += align_up((size_t)t.N, 16);
  }
  return align_up(b, 16);
}

template <bool kStateSmem>
__device__ __forceinline__ Smem carve(unsigned char *base, const DeviceTables &t, int acc_stride,
                                      bool canonical, int8_t *g_occ, int32_t *g_list) {
  Smem s;
  double *d = reinterpret_cast<double *>(base);
  s.cf = d; d += t.n_eci;
  s.cfn = d; d += t.n_eci;
  s.eci = d; d += t.n_eci;
  s.prod = d; d += t.n_eci;
  s.A = d; d += 2 * t.D * t.KP;
  s.diff = d; d += 2 * (t.max_tasks > 0 ? t.max_tasks : 1);
  s.bf = d; d += t.D * t.S;
  s.acc = d; d += acc_stride;
  s.fin = reinterpret_cast<Fin *>(d);
  unsigned char *p = reinterpret_cast<unsigned char *>(s.fin + (size_t)t.n_symm * t.n_eci);
  p = reinterpret_cast<unsigned char *>(align_up(reinterpret_cast<size_t>(p), 8));
  s.tasks = reinterpret_cast<Task *>(p); p += sizeof(Task) * (t.n_tasks_total > 0 ? t.n_tasks_total : 1);
  s.fams = reinterpret_cast<Fam *>(p); p += sizeof(Fam) * (t.n_fam > 0 ? t.n_fam : 1);
  s.pos = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * (t.n_pos_words > 0 ? t.n_pos_words : 1);
  s.task_base = reinterpret_cast<int32_t *>(p); p += sizeof(int32_t) * (t.n_symm + 1);
  s.singlet_idx = reinterpret_cast<int32_t *>(p); p += sizeof(int32_t) * (t.n_singlets > 0 ? t.n_singlets : 1);
  s.off = reinterpret_cast<int32_t *>(p); p += sizeof(int32_t) * (t.S + 1);
  s.present = reinterpret_cast<int32_t *>(p); p += sizeof(int32_t) * (t.S + 1);
  s.rng = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * 32 * 8;
  p = reinterpret_cast<unsigned char *>(align_up(reinterpret_cast<size_t>(p), 16));
  if (kStateSmem) {
    if (canonical) { s.list = reinterpret_cast<int32_t *>(p); p += sizeof(int32_t) * (size_t)t.N; }
    else s.list = g_list;
    s.occ = reinterpret_cast<int8_t *>(p);
  } else {
    s.list = g_list;
    s.occ = g_occ;
  }
  return s;
}

// One (ECI, decoration) task: sp_new - sp_ref of ce_updater.cpp:395-397 with
// spin_product_one_atom (:244-285) evaluated for old and new in one pass.
template <int N_>
__device__ __forceinline__ double task_diff(const uint32_t *__restrict__ pos, int M,
                                            const double *__restrict__ A, int KP, int K,
                                            uint32_t deco, const double *__restrict__ bf, int S,
                                            int old_id, int new_id) {
  int aoff[N_];
  double rO[N_], rN[N_];
#pragma unroll
  for (int k = 0; k < N_; k++) {
    const int dk = (deco >> (8 * k)) & 0xff;
    aoff[k] = dk * KP;
    rO[k] = bf[dk * S + old_id];
    rN[k] = bf[dk * S + new_id];
  }
  double spO = 0.0, spN = 0.0;
#pragma unroll 4
  for (int m = 0; m < M; m++) {
    const uint32_t pp = pos[m];
    double tO = 0.0, tN = 0.0;
#pragma unroll
    for (int k = 0; k < N_; k++) {
      const int p = (pp >> (8 * k)) & 0xff;
      const double f = A[aoff[k] + p];
      const bool isref = (p == K);
      const double fO = isref ? rO[k] : f;
      const double fN = isref ? rN[k] : f;
      if (k == 0) { tO = fO; tN = fN; }        // 1.0 * f == f exactly (:255,:275)
      else { tO = __dmul_rn(tO, fO); tN = __dmul_rn(tN, fN); }
    }
    spO = __dadd_rn(spO, tO);                  // :282
    spN = __dadd_rn(spN, tN);
  }
  return __dsub_rn(spN, spO);                  // :397
}

template <int MODE, bool kStateSmem>
__global__ void __launch_bounds__(256)
mc_kernel(DeviceTables t, ReplicaState st, RunArgs a, int acc_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int r = blockIdx.x;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int N = t.N, K = t.K, KP = t.KP, S = t.S, D = t.D, n_eci = t.n_eci;
  constexpr bool kCanon = (MODE == MODE_CANONICAL);

  int8_t *g_occ = st.occ + (size_t)r * N;
  int32_t *g_list = st.list ? st.list + (size_t)r * N : nullptr;
  int32_t *g_loc = st.loc ? st.loc + (size_t)r * N : nullptr;
  Smem s = carve<kStateSmem>(smem_raw, t, acc_stride, kCanon, g_occ, g_list);

  // ---- stage per-replica state and the cluster program -------------------
  for (int i = tid; i < n_eci; i += nthr) {
    s.cf[i] = st.cf[(size_t)r * n_eci + i];
    s.eci[i] = st.eci[(size_t)r * n_eci + i];
  }
  for (int i = tid; i < D * S; i += nthr) s.bf[i] = t.bf[i];
  for (int i = tid; i < acc_stride; i += nthr) s.acc[i] = st.acc[(size_t)r * acc_stride + i];
  for (int i = tid; i < t.n_symm * n_eci; i += nthr) s.fin[i] = t.fin[i];
  for (int i = tid; i < t.n_tasks_total; i += nthr) s.tasks[i] = t.tasks[i];
  for (int i = tid; i < t.n_fam; i += nthr) s.fams[i] = t.fams[i];
  for (int i = tid; i < t.n_pos_words; i += nthr) s.pos[i] = t.pos[i];
  for (int i = tid; i <= t.n_symm; i += nthr) s.task_base[i] = t.task_base[i];
  for (int i = tid; i < t.n_singlets; i += nthr) s.singlet_idx[i] = t.singlet_idx[i];
  for (int i = tid; i < 2 * D * KP; i += nthr) s.A[i] = 1.0;   // column K stays 1.0
  if (kStateSmem) {
    // 16-byte vectorised copy of the int8 occupations
    const int nvec = N / 16;
    const int4 *src = reinterpret_cast<const int4 *>(g_occ);
    int4 *dst = reinterpret_cast<int4 *>(s.occ);
    if ((reinterpret_cast<size_t>(g_occ) & 15) == 0) {
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
  const double ref = st.ref[r];
  const double dN = (double)(unsigned)N;
  unsigned long long step0 = st.step[r];
  unsigned long long n_acc = 0;
  const uint32_t rep_global = a.replica_offset + (uint32_t)r;
  int err = 0;

  for (long long it = 0; it < a.n_steps; it++) {
    // ---- P0: proposal ------------------------------------------------------
    int site0, site1 = -1, new0, new1 = 0, slot0 = 0, slot1 = 0;
    double u;
    if (MODE == MODE_REPLAY) {
      const size_t q = (size_t)r * a.n_steps + it;
      site0 = a.rp_sites[2 * q]; site1 = a.rp_sites[2 * q + 1];
      new0 = a.rp_news[2 * q]; new1 = a.rp_news[2 * q + 1];
      u = a.rp_u[q];
      if (site0 < 0 || site0 >= N || site1 >= N || new0 < 0 || new0 >= S ||
          (site1 >= 0 && (new1 < 0 || new1 >= S))) { err = 3; break; }
    } else {
      if ((it & 31) == 0) {
        __syncthreads();                       // everyone done with the old ring
        if (warp == 0) {
          const unsigned long long stp = step0 + (unsigned long long)it + lane;
          uint32_t c0 = (uint32_t)stp, c1 = (uint32_t)(stp >> 32), c2 = rep_global, c3 = 0;
          philox4x32_10(c0, c1, c2, c3, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
          uint32_t *w = s.rng + lane * 8;
          w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
          if (kCanon) {
            c0 = (uint32_t)stp; c1 = (uint32_t)(stp >> 32); c2 = rep_global; c3 = 1;
            philox4x32_10(c0, c1, c2, c3, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
            w[4] = c0; w[5] = c1;
          }
        }
        __syncthreads();
      }
      const uint32_t *w = s.rng + (it & 31) * 8;
      if (MODE == MODE_SGC) {
        // sgc_montecarlo.py:69-75: site uniform, new species uniform among others
        const uint32_t ia = __umulhi(w[0], (uint32_t)t.n_active);
        site0 = t.active ? t.active[ia] : (int)ia;
        const int old = s.occ[site0];
        int p = -1;
        for (int q = 0; q < t.n_allowed; q++) if (t.allowed[q] == old) p = q;
        int rr;
        if (p >= 0) { rr = (int)__umulhi(w[1], (uint32_t)(t.n_allowed - 1)); rr += (rr >= p); }
        else rr = (int)__umulhi(w[1], (uint32_t)t.n_allowed);
        new0 = t.allowed[rr];
        u = u53(w[2], w[3]);
      } else {
        // montecarlo.py:899-907: species pair uniform (a != b), site uniform per species
        const int ia = (int)__umulhi(w[0], (uint32_t)n_present);
        int ib = (int)__umulhi(w[1], (uint32_t)(n_present - 1)); ib += (ib >= ia);
        const int sa = s.present[ia], sb = s.present[ib];
        slot0 = s.off[sa] + (int)__umulhi(w[2], (uint32_t)(s.off[sa + 1] - s.off[sa]));
        slot1 = s.off[sb] + (int)__umulhi(w[3], (uint32_t)(s.off[sb + 1] - s.off[sb]));
        site0 = s.list[slot0]; site1 = s.list[slot1];
        new0 = sb; new1 = sa;
        u = u53(w[4], w[5]);
      }
    }
    const int old0 = s.occ[site0];
    const int old1 = site1 >= 0 ? (site1 == site0 ? new0 : (int)s.occ[site1]) : 0;
    const bool ch0 = (old0 != new0);                       // ce_updater.cpp:315
    const bool ch1 = (site1 >= 0) && (old1 != new1);
    const int g0 = t.symm_of_site[site0];
    const int g1 = site1 >= 0 ? t.symm_of_site[site1] : 0;
    if ((ch0 && g0 < 0) || (ch1 && g1 < 0)) { err = 1; break; }   // :330 background atom

    // ---- P1: gather neighbour occupations -> basis-function values ---------
    for (int q = tid; q < 2 * K; q += nthr) {
      const int j = q >= K, c = j ? q - K : q;
      if (j ? ch1 : ch0) {
        const int sj = j ? site1 : site0;
        const int nb = __ldg(&t.trans[(size_t)sj * K + c]);       // :264
        int v = s.occ[nb];
        if (j && nb == site0) v = new0;        // change 1 sees change 0 applied (:845-852)
        double *Aj = s.A + (size_t)j * D * KP;
        for (int d = 0; d < D; d++) Aj[d * KP + c] = s.bf[d * S + v];
      }
    }
    __syncthreads();

    // ---- P2: spin-product sums, one thread per (changed site, task) --------
    {
      const int nt0 = ch0 ? (s.task_base[g0 + 1] - s.task_base[g0]) : 0;
      const int nt1 = ch1 ? (s.task_base[g1 + 1] - s.task_base[g1]) : 0;
      for (int q = tid; q < nt0 + nt1; q += nthr) {
        const int j = q >= nt0, tk = j ? q - nt0 : q;
        const int g = j ? g1 : g0;
        const Task T = s.tasks[s.task_base[g] + tk];
        const Fam F = s.fams[T.fam];
        const double *Aj = s.A + (size_t)j * D * KP;
        const int oid = j ? old1 : old0, nid = j ? new1 : new0;
        double dv;
        if (F.n == 2) dv = task_diff<2>(s.pos + F.pos_off, F.M, Aj, KP, K, T.deco, s.bf, S, oid, nid);
        else if (F.n == 3) dv = task_diff<3>(s.pos + F.pos_off, F.M, Aj, KP, K, T.deco, s.bf, S, oid, nid);
        else dv = task_diff<4>(s.pos + F.pos_off, F.M, Aj, KP, K, T.deco, s.bf, S, oid, nid);
        s.diff[j * t.max_tasks + tk] = dv;
      }
    }
    __syncthreads();

    // ---- P3/P4 (warp 0): per-ECI increments, energy, Metropolis, commit ----
    if (warp == 0) {
      for (int i = lane; i < n_eci; i += 32) {
        double c = s.cf[i];
#pragma unroll
        for (int j = 0; j < 2; j++) {
          if (!(j ? ch1 : ch0)) continue;
          const Fin f = s.fin[(j ? g1 : g0) * n_eci + i];
          const int oid = j ? old1 : old0, nid = j ? new1 : new0;
          if (f.kind == 1) {                                    // :366-371
            const double dl = __ddiv_rn(__dsub_rn(s.bf[f.d * S + nid], s.bf[f.d * S + oid]), dN);
            c = __dadd_rn(c, dl);
          } else if (f.kind == 2) {
            double delta = 0.0;
            const double *df = s.diff + j * t.max_tasks;
            for (int q = f.t0; q < f.t1; q++) delta = __dadd_rn(delta, df[q]);   // :397
            delta = __dmul_rn(delta, f.scale);                  // :400
            delta = __ddiv_rn(delta, f.div);                    // :402
            c = __dadd_rn(c, delta);                            // :404
          }                                                     // else: copied (:360,:382)
        }
        s.cfn[i] = c;
        s.prod[i] = __dmul_rn(s.eci[i], c);
      }
      __syncwarp();
      double e_new = 0.0;                                       // named_array.cpp:27-31
      for (int i = 0; i < n_eci; i++) e_new = __dadd_rn(e_new, s.prod[i]);
      e_new = __dmul_rn(e_new, dN);                             // ce_updater.cpp:241
      bool accept;
      if (a.force_accept) accept = true;
      else if (e_new < e_cur) accept = true;                    // montecarlo.py:951
      else accept = (u <= exp(__ddiv_rn(-__dsub_rn(e_new, e_cur), kT)));   // :953-956
      if (accept) {
        for (int i = lane; i < n_eci; i += 32) s.cf[i] = s.cfn[i];
        e_cur = e_new;
        n_acc++;
        if (lane == 0) {
          if (ch0) s.occ[site0] = (int8_t)new0;
          if (ch1) s.occ[site1] = (int8_t)new1;
          if (kCanon) {                        // swap_move_index_tracker.py:39-59
            s.list[slot0] = site1; s.list[slot1] = site0;
            g_loc[site1] = slot0 - s.off[new1]; g_loc[site0] = slot1 - s.off[new0];
          } else if (MODE == MODE_REPLAY && g_list != nullptr && site1 >= 0 && ch0 && ch1) {
            const int l1 = g_loc[site0], l2 = g_loc[site1];
            const int32_t *off = st.off + (size_t)r * (S + 1);
            s.list[off[old0] + l1] = site1; g_loc[site1] = l1;
            s.list[off[old1] + l2] = site0; g_loc[site0] = l2;
          }
        }
      }
      __syncwarp();
      if (a.observe) {                                          // montecarlo.py:811-814,
        if (lane == 0) {                                        // mc_observers.py:264-270
          s.acc[0] = __dadd_rn(s.acc[0], 1.0);
          const double e2 = __dmul_rn(e_cur, e_cur);
          s.acc[1] = __dadd_rn(s.acc[1], ref == 1.0 ? e_cur : __ddiv_rn(e_cur, ref));
          s.acc[2] = __dadd_rn(s.acc[2], ref == 1.0 ? e2 : __ddiv_rn(e2, ref));
        }
        for (int d = lane; d < t.n_singlets; d += 32) {
          const double sv = s.cf[s.singlet_idx[d]];
          double *ad = s.acc + 3 + 3 * d;
          ad[0] = __dadd_rn(ad[0], sv);
          ad[1] = __dadd_rn(ad[1], __dmul_rn(sv, sv));
          ad[2] = __dadd_rn(ad[2], __dmul_rn(sv, e_cur));
        }
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
    __syncthreads();
  }

  // ---- write back --------------------------------------------------------
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

}  // namespace cemc
