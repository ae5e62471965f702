/*
 * ce_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded restatement of the reference's cluster-expansion
 * Metropolis hot path, used only as the checker in tests/, in
 * __graft_entry__.smoke() and in bench.py's cpu_baseline leg.  Nothing under
 * cemc_b200/ may link or call this file.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py::test_oracle_vs_compiled_reference drives this file
 * and the reference's own compiled CEUpdater (oracle/_ref, built from
 * /root/reference by oracle/build_ref.sh) on the same inputs and requires
 * bit-identical CFs / energies / accept sequences; the committed fixtures in
 * tests/golden/ were produced by that compiled reference
 * (tests/golden/make_golden.py).
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference).  Floating point: compile with -O2 -ffp-contract=off so
 * that a*b+c is never fused -- the reference is built for baseline x86-64
 * (no FMA), setup.py:34.
 *
 * The proposal generators (oracle_run_sgc / oracle_run_canonical) follow the
 * reference's proposal DISTRIBUTIONS (montecarlo.py:890-908,
 * sgc_montecarlo.py:62-76) but draw from a Philox4x32-10 counter stream
 * defined by this project (DESIGN.md "Random streams"), because the
 * reference's MT19937 + Python `random` streams are hash-order dependent
 *  and cannot be reproduced on any other implementation.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/cemc_b200.h"

/* ------------------------------------------------------------------ */
/* A.1 step 5: CEUpdater::spin_product_one_atom, ce_updater.cpp:244-285 */
static double spin_product_one_atom(const cemc_tables *t, const int8_t *occ,
                                    int ref, int fam, const int8_t *deco,
                                    int ref_id)
{
  const int n = t->fam_size[fam];
  const int M = t->fam_nsub[fam];
  const int32_t *pos = t->fam_pos + t->fam_pos_off[fam];
  const int S = t->n_species, K = t->n_cols;
  double sp = 0.0;
  for (int i = 0; i < M; i++) {                       /* :253 */
    double sp_temp = 1.0;                             /* :255 */
    for (int k = 0; k < n; k++) {                     /* :271 */
      int p = pos[i * n + k];
      int id;
      if (p == CEMC_POS_REF) {
        id = ref_id;                                  /* :273-276 */
      } else {
        int site = t->trans[(size_t)ref * K + p];     /* :264 */
        /* the reference tests the SITE index (:273); a neighbour equal to
         * ref is excluded by SelfInteractionError (ce_calculator.py:596) */
        id = (site == ref) ? ref_id : occ[site];      /* :279 */
      }
      sp_temp *= t->bf[deco[k] * S + id];
    }
    sp += sp_temp;                                    /* :282 */
  }
  return sp;
}

/* A.1: CEUpdater::update_cf(SymbolChange&), ce_updater.cpp:313-406.
 * occ is mutated (occ[site] = new_sp) BEFORE the products, as at :334-336. */
int oracle_update_cf(const cemc_tables *t, int8_t *occ, const double *cf_cur,
                     double *cf_next, int site, int new_sp)
{
  const int S = t->n_species;
  const int old_sp = occ[site];
  if (old_sp == new_sp) {                             /* :315-318 */
    if (cf_next != cf_cur) memcpy(cf_next, cf_cur, sizeof(double) * t->n_eci);
    return 0;
  }
  const int g = t->symm_of_site[site];
  if (g < 0) return 1;                                /* :330 background atom */
  occ[site] = (int8_t)new_sp;                         /* :335 */
  for (int i = 0; i < t->n_eci; i++) {                /* :353 */
    if (t->eci_kind[i] == CEMC_ECI_EMPTY) {           /* :357-362 */
      cf_next[i] = cf_cur[i];
      continue;
    }
    if (t->eci_kind[i] == CEMC_ECI_SINGLET) {         /* :366-371 */
      int dec = t->eci_bf[i];
      cf_next[i] = cf_cur[i] +
          (t->bf[dec * S + new_sp] - t->bf[dec * S + old_sp]) /
              (double)(unsigned)t->n_sites;
      continue;
    }
    const int term = g * t->n_eci + i;
    const int fam = t->term_fam[term];
    if (fam < 0) {                                    /* :380-384 */
      cf_next[i] = cf_cur[i];
      continue;
    }
    const int size = t->fam_size[fam];
    const int d0 = t->term_deco_off[term], d1 = t->term_deco_off[term + 1];
    double delta_sp = 0.0;
    for (int e = d0; e < d1; e++) {                   /* :393-398 */
      const int8_t *deco = t->deco + 4 * e;
      double sp_ref = spin_product_one_atom(t, occ, site, fam, deco, old_sp);
      double sp_new = spin_product_one_atom(t, occ, site, fam, deco, new_sp);
      delta_sp += sp_new - sp_ref;
    }
    delta_sp *= ((double)size / (double)(d1 - d0));   /* :400 */
    delta_sp /= (double)(t->term_count[term] * t->symm_count[g]); /* :402 */
    cf_next[i] = cf_cur[i] + delta_sp;                /* :404 */
  }
  return 0;
}

/* A.3: CEUpdater::get_energy :236-242 + NamedArray::dot named_array.cpp:25-33 */
double oracle_energy(const cemc_tables *t, const double *eci, const double *cf)
{
  double dot_prod = 0.0;
  for (int i = 0; i < t->n_eci; i++) dot_prod += eci[i] * cf[i];
  return dot_prod * (double)(unsigned)t->n_sites;
}

/* CF definition implied by ce_updater.cpp:393-404.  This is
 * the from-scratch evaluator the reference delegates to ase.clease
 * CorrFunction (ce_calculator.py:169-175); summation order is ours. */
void oracle_full_cf(const cemc_tables *t, const int8_t *occ, double *cf)
{
  const int S = t->n_species, N = t->n_sites;
  for (int i = 0; i < t->n_eci; i++) {
    if (t->eci_kind[i] == CEMC_ECI_EMPTY) { cf[i] = 1.0; continue; }
    if (t->eci_kind[i] == CEMC_ECI_SINGLET) {
      double s = 0.0;
      for (int a = 0; a < N; a++)
        if (t->symm_of_site[a] >= 0) s += t->bf[t->eci_bf[i] * S + occ[a]];
      cf[i] = s / (double)N;
      continue;
    }
    double tot = 0.0;
    int any = 0;
    for (int g = 0; g < t->n_symm; g++) {
      const int term = g * t->n_eci + i;
      const int fam = t->term_fam[term];
      if (fam < 0) continue;
      any = 1;
      const int d0 = t->term_deco_off[term], d1 = t->term_deco_off[term + 1];
      double sg = 0.0;
      for (int a = 0; a < N; a++) {
        if (t->symm_of_site[a] != g) continue;
        for (int e = d0; e < d1; e++)
          sg += spin_product_one_atom(t, occ, a, fam, t->deco + 4 * e, occ[a]);
      }
      tot += sg / ((double)(d1 - d0) *
                   (double)(t->term_count[term] * t->symm_count[g]));
    }
    cf[i] = any ? tot : 0.0;
  }
}

/* ------------------------------------------------------------------ */
/* observers: Averager (cemc/mcmc/averager.py:21-23) as used at
 * montecarlo.py:811-814 and SGCObserver.__call__ mc_observers.py:264-270 */
typedef struct {
  int n_singlets;
  const int32_t *singlet_idx; /* indices of c1_* in the CF vector */
  double ref;                 /* Averager ref_value */
  double *acc;                /* [CEMC_ACC_STRIDE(n_singlets)] */
} oracle_obs;

static void observe(const oracle_obs *o, const double *cf, double E)
{
  if (!o || !o->acc) return;
  double *a = o->acc;
  a[CEMC_ACC_COUNT] += 1.0;
  a[CEMC_ACC_E] += E / o->ref;
  a[CEMC_ACC_E2] += (E * E) / o->ref;
  for (int d = 0; d < o->n_singlets; d++) {
    double s = cf[o->singlet_idx[d]];
    a[CEMC_ACC_SINGLET0 + 3 * d + 0] += s;
    a[CEMC_ACC_SINGLET0 + 3 * d + 1] += s * s;
    a[CEMC_ACC_SINGLET0 + 3 * d + 2] += s * E;
  }
}

/* One Metropolis trial: _mc_step + _accept, montecarlo.py:910-1038 (A.4/A.5).
 * occ/cf/e_cur are the committed state; scratch is n_eci doubles x 2. */
static int trial_move(const cemc_tables *t, const double *eci, int8_t *occ,
                      double *cf, double *e_cur, double kT, int n_changes,
                      const int32_t *sites, const int8_t *news, double u,
                      double *scratch, int *err)
{
  double *c1 = scratch, *c2 = scratch + t->n_eci;
  int8_t olds[2];
  const double *cur = cf;
  double *nxt = c1;
  for (int j = 0; j < n_changes; j++) {               /* ce_updater.cpp:845-852 */
    olds[j] = occ[sites[j]];
    if (oracle_update_cf(t, occ, cur, nxt, sites[j], news[j])) { *err = 1; return 0; }
    cur = nxt;
    nxt = (nxt == c1) ? c2 : c1;
  }
  const double e_new = oracle_energy(t, eci, cur);
  int accept;
  if (e_new < *e_cur) {                               /* montecarlo.py:951-952 */
    accept = 1;
  } else {
    double energy_diff = e_new - *e_cur;              /* :954 */
    double probability = exp(-energy_diff / kT);      /* :955 */
    accept = (u <= probability);                      /* :956 */
  }
  if (accept) {                                       /* clear_history */
    memcpy(cf, cur, sizeof(double) * t->n_eci);
    *e_cur = e_new;
  } else {                                            /* undo_changes :414-448 */
    for (int j = n_changes - 1; j >= 0; j--) occ[sites[j]] = olds[j];
  }
  return accept;
}

/* Appendix D replay.  sites[2*s+1] < 0 marks a one-site step. */
int oracle_replay(const cemc_tables *t, const double *eci, int8_t *occ,
                  double *cf, double *e_cur, double kT, int n_steps,
                  const int32_t *sites, const int8_t *news, const double *u,
                  uint8_t *accepted_out, double *e_after_out,
                  const oracle_obs *obs, uint64_t *n_accepted)
{
  double *scratch = (double *)malloc(sizeof(double) * 2 * t->n_eci);
  int err = 0;
  for (int s = 0; s < n_steps && !err; s++) {
    int nch = sites[2 * s + 1] < 0 ? 1 : 2;
    int acc = trial_move(t, eci, occ, cf, e_cur, kT, nch, sites + 2 * s,
                         news + 2 * s, u[s], scratch, &err);
    if (accepted_out) accepted_out[s] = (uint8_t)acc;
    if (e_after_out) e_after_out[s] = *e_cur;
    if (n_accepted) *n_accepted += (uint64_t)acc;
    observe(obs, cf, *e_cur);
  }
  free(scratch);
  return err;
}

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11), the project's counter-based stream. */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

void oracle_philox(uint64_t seed, uint64_t step, uint32_t replica,
                   uint32_t stream, uint32_t out[4])
{
  out[0] = (uint32_t)step; out[1] = (uint32_t)(step >> 32);
  out[2] = replica; out[3] = stream;
  philox4x32_10(out, (uint32_t)seed, (uint32_t)(seed >> 32));
}

static inline uint32_t mulhi(uint32_t w, uint32_t n)
{
  return (uint32_t)(((uint64_t)w * n) >> 32);
}

/* 53-bit uniform in [0,1), numpy random_sample construction */
static inline double u53(uint32_t a, uint32_t b)
{
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

/* SGC flip chain: proposal distribution of sgc_montecarlo.py:62-76 (site
 * uniform over active sites, new species uniform among allowed != old).
 * trace_* (optional) record the proposals for replay on other engines. */
int oracle_run_sgc(const cemc_tables *t, const double *eci, int8_t *occ,
                   double *cf, double *e_cur, double kT, uint64_t seed,
                   uint32_t replica, uint64_t step0, int64_t n_steps,
                   int n_active, const int32_t *active, int n_allowed,
                   const int8_t *allowed, const oracle_obs *obs,
                   uint64_t *n_accepted, int32_t *trace_sites,
                   int8_t *trace_news, double *trace_u, uint8_t *trace_acc,
                   double *trace_e)
{
  double *scratch = (double *)malloc(sizeof(double) * 2 * t->n_eci);
  int err = 0;
  for (int64_t s = 0; s < n_steps && !err; s++) {
    uint32_t w[4];
    oracle_philox(seed, step0 + (uint64_t)s, replica, 0, w);
    int32_t site = active[mulhi(w[0], (uint32_t)n_active)];
    int old = occ[site];
    int p = -1;
    for (int a = 0; a < n_allowed; a++) if (allowed[a] == old) p = a;
    int r;
    if (p >= 0) { r = (int)mulhi(w[1], (uint32_t)(n_allowed - 1)); r += (r >= p); }
    else r = (int)mulhi(w[1], (uint32_t)n_allowed);
    int32_t sites[2] = {site, -1};
    int8_t news[2] = {allowed[r], 0};
    double u = u53(w[2], w[3]);
    int acc = trial_move(t, eci, occ, cf, e_cur, kT, 1, sites, news, u, scratch, &err);
    if (n_accepted) *n_accepted += (uint64_t)acc;
    observe(obs, cf, *e_cur);
    if (trace_sites) { trace_sites[2 * s] = site; trace_sites[2 * s + 1] = -1; }
    if (trace_news) { trace_news[2 * s] = news[0]; trace_news[2 * s + 1] = 0; }
    if (trace_u) trace_u[s] = u;
    if (trace_acc) trace_acc[s] = (uint8_t)acc;
    if (trace_e) trace_e[s] = *e_cur;
  }
  free(scratch);
  return err;
}

/* SwapMoveIndexTracker.init_tracker, swap_move_index_tracker.py:22-36:
 * per-species site lists in ascending site order + index_loc. list is laid
 * out species-major: species sp occupies [off[sp], off[sp+1]). */
void oracle_tracker_init(const cemc_tables *t, const int8_t *occ, int32_t *list,
                         int32_t *loc, int32_t *off /*[S+1]*/)
{
  const int S = t->n_species, N = t->n_sites;
  int *cnt = (int *)calloc((size_t)S + 1, sizeof(int));
  for (int a = 0; a < N; a++) if (t->symm_of_site[a] >= 0) cnt[occ[a]]++;
  off[0] = 0;
  for (int sp = 0; sp < S; sp++) off[sp + 1] = off[sp] + cnt[sp];
  memset(cnt, 0, sizeof(int) * (size_t)S);
  for (int a = 0; a < N; a++) {
    if (t->symm_of_site[a] < 0) { loc[a] = -1; continue; }
    int sp = occ[a];
    loc[a] = cnt[sp];
    list[off[sp] + cnt[sp]++] = a;
  }
  free(cnt);
}

/* SwapMoveIndexTracker.update_swap_move, swap_move_index_tracker.py:39-59 */
static void tracker_swap(int32_t *list, int32_t *loc, const int32_t *off,
                         int indx1, int indx2, int symb1, int symb2)
{
  int loc1 = loc[indx1], loc2 = loc[indx2];
  list[off[symb1] + loc1] = indx2; loc[indx2] = loc1;
  list[off[symb2] + loc2] = indx1; loc[indx1] = loc2;
}

/* Canonical swap chain: proposal distribution of montecarlo.py:890-908
 * (species pair uniform among species present, a != b; then site uniform
 * within each species list). */
int oracle_run_canonical(const cemc_tables *t, const double *eci, int8_t *occ,
                         double *cf, double *e_cur, double kT, uint64_t seed,
                         uint32_t replica, uint64_t step0, int64_t n_steps,
                         int32_t *list, int32_t *loc, const int32_t *off,
                         const oracle_obs *obs, uint64_t *n_accepted,
                         int32_t *trace_sites, int8_t *trace_news,
                         double *trace_u, uint8_t *trace_acc, double *trace_e)
{
  const int S = t->n_species;
  int present[128], np_ = 0;
  for (int sp = 0; sp < S; sp++) if (off[sp + 1] > off[sp]) present[np_++] = sp;
  if (np_ < 2) return 2;                              /* TooFewElementsError */
  double *scratch = (double *)malloc(sizeof(double) * 2 * t->n_eci);
  int err = 0;
  for (int64_t s = 0; s < n_steps && !err; s++) {
    uint32_t w[4], v[4];
    oracle_philox(seed, step0 + (uint64_t)s, replica, 0, w);
    oracle_philox(seed, step0 + (uint64_t)s, replica, 1, v);
    int ia = (int)mulhi(w[0], (uint32_t)np_);
    int ib = (int)mulhi(w[1], (uint32_t)(np_ - 1)); ib += (ib >= ia);
    int a = present[ia], b = present[ib];
    int site_a = list[off[a] + mulhi(w[2], (uint32_t)(off[a + 1] - off[a]))];
    int site_b = list[off[b] + mulhi(w[3], (uint32_t)(off[b + 1] - off[b]))];
    int32_t sites[2] = {site_a, site_b};
    int8_t news[2] = {(int8_t)b, (int8_t)a};
    double u = u53(v[0], v[1]);
    int acc = trial_move(t, eci, occ, cf, e_cur, kT, 2, sites, news, u, scratch, &err);
    if (acc) tracker_swap(list, loc, off, site_a, site_b, a, b);
    if (n_accepted) *n_accepted += (uint64_t)acc;
    observe(obs, cf, *e_cur);
    if (trace_sites) { trace_sites[2 * s] = site_a; trace_sites[2 * s + 1] = site_b; }
    if (trace_news) { trace_news[2 * s] = news[0]; trace_news[2 * s + 1] = news[1]; }
    if (trace_u) trace_u[s] = u;
    if (trace_acc) trace_acc[s] = (uint8_t)acc;
    if (trace_e) trace_e[s] = *e_cur;
  }
  free(scratch);
  return err;
}

/* A.6: ParallelTempering._perform_exchange_move parallel_tempering.py:153-175
 * with _accept_probability :138-144.  Slots are temperature indices; instead
 * of copying configurations (:146-151) the slot<->replica map is permuted.
 * Uniforms come from Philox(seed, round, slot pair index, stream 2). */
/* direction of one exchange cycle, 0 = "up", 1 = "down" (random.choice per cycle,
 * parallel_tempering.py:191), from this project's counter stream (stream 3) */
int oracle_pt_direction(uint64_t seed, uint64_t round)
{
  uint32_t w[4];
  oracle_philox(seed, round, 0, 3, w);
  return (int)(w[0] >> 31);
}

int oracle_pt_exchange(int n_total, const double *energies /*by replica*/,
                       int32_t *slot_of_replica, const double *kT_of_slot,
                       int direction, uint64_t seed, uint64_t round)
{
  int32_t *rep_of_slot = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_total);
  for (int g = 0; g < n_total; g++) rep_of_slot[slot_of_replica[g]] = g;
  int n_acc = 0;
  int i0 = direction == 0 ? 0 : n_total - 1;
  int step = direction == 0 ? 2 : -2;
  for (int i = i0; direction == 0 ? (i < n_total - 1) : (i > 0); i += step) {
    int j = direction == 0 ? i + 1 : i - 1;           /* move = (i, j) */
    int r1 = rep_of_slot[i], r2 = rep_of_slot[j];
    double dE = energies[r1] - energies[r2];
    double b1 = 1.0 / kT_of_slot[i];
    double b2 = 1.0 / kT_of_slot[j];
    double db = b1 - b2;
    double p = exp(db * dE);
    uint32_t w[4];
    oracle_philox(seed, round, (uint32_t)i, 2, w);
    double u = u53(w[0], w[1]);
    if (u < p) {                                      /* :166 strict */
      rep_of_slot[i] = r2; rep_of_slot[j] = r1;
      n_acc++;
    }
  }
  for (int s = 0; s < n_total; s++) slot_of_replica[rep_of_slot[s]] = s;
  free(rep_of_slot);
  return n_acc;
}
