/*
 * cemc_b200.h -- C ABI of the B200-native cluster-expansion Metropolis hot path.
 *
 * Drop-in boundary for davidkleiven/CEMC's `CEUpdater` (reference paths are
 * relative to /root/reference):
 *
 *   reference interface                                   replaced by
 *   ----------------------------------------------------  -----------------------------
 *   CEUpdater::init           cpp/src/ce_updater.cpp:32    cemc_create (+ cemc_tables)
 *   CEUpdater::update_cf      cpp/src/ce_updater.cpp:313   cemc_trial_changes / cemc_replay / cemc_run_*
 *   CEUpdater::calculate      cpp/src/ce_updater.cpp:471   cemc_trial_changes
 *   CEUpdater::undo_changes   cpp/src/ce_updater.cpp:408   cemc_undo_changes
 *   CEUpdater::clear_history  cpp/src/ce_updater.cpp:552   cemc_clear_history
 *   CEUpdater::get_energy     cpp/src/ce_updater.cpp:236   cemc_get_energy
 *   CEUpdater::get_cf         cpp/src/ce_updater.cpp:570   cemc_get_cf
 *   CEUpdater::get_singlets   cpp/src/ce_updater.cpp:655   cemc_get_cf (+ singlet index list, host side)
 *   CEUpdater::set_ecis       cpp/src/ce_updater.cpp:615   cemc_set_ecis
 *   CEUpdater::set_symbols    cpp/src/ce_updater.cpp:606   cemc_set_occupancy
 *   Cython binding            cemc/cpp_ext/ce_updater.pxd:9-43, pyce_updater.pyx:5-60
 *
 * The per-move Python loop of the reference samplers cannot feed a GPU one
 * launch per move, so the loop bodies move under the boundary as well:
 *
 *   Montecarlo._mc_step/_accept        cemc/mcmc/montecarlo.py:910-1038     cemc_run_canonical
 *   SGCMonteCarlo._get_trial_move      cemc/mcmc/sgc_montecarlo.py:62-76    cemc_run_sgc
 *   SGCObserver.__call__               cemc/mcmc/mc_observers.py:222-270    accumulators (cemc_get_accumulators)
 *   ParallelTempering._perform_exchange_move
 *                                      cemc/mcmc/parallel_tempering.py:153  cemc_pt_exchange
 *
 * Conventions: every function returns 0 on success, non-zero on error;
 * cemc_last_error() gives the message.  Host pointers unless the name says
 * `_dev`.  All arrays are dense, row-major, replica-major ([R][...]).
 * One handle = one CUDA device + one stream; calls are stream-ordered and
 * the handle is not thread-safe.  Setters copy the caller's buffer before they
 * return (pinned staging), the upload itself is stream-ordered; getters and
 * cemc_synchronize wait for the stream.  No torch types appear in this ABI.
 */
#ifndef CEMC_B200_H
#define CEMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CEMC_MAX_CLUSTER_SIZE 4      /* cpp/src/cluster.cpp:131-171: sizes 2..4 */
#define CEMC_POS_REF (-1)            /* fam_pos entry that denotes the changed site itself */

enum cemc_eci_kind {
  CEMC_ECI_EMPTY   = 0,              /* name starts with "c0": CF copied (ce_updater.cpp:357) */
  CEMC_ECI_SINGLET = 1,              /* name starts with "c1" (ce_updater.cpp:366)           */
  CEMC_ECI_CLUSTER = 2               /* n-body term          (ce_updater.cpp:373-404)        */
};

/* Flattened, read-only tables: what CEUpdater::init builds from the Python
 * settings object (cluster_info, trans_matrix, basis_functions, ECIs).      */
typedef struct cemc_tables {
  int32_t n_sites;       /* N, including background sites                          */
  int32_t n_species;     /* S: sorted(unique_elements U symbols present)           */
  int32_t n_bf;          /* D = num_unique_elements - 1                            */
  int32_t n_cols;        /* K: distinct translation-matrix columns used            */
  int32_t n_eci;         /* number of ECIs == CFs, in lexicographic name order     */
  int32_t n_symm;        /* translational symmetry groups                          */
  int32_t n_fam;         /* (symmetry group, cluster family) tables                */
  int32_t n_deco;        /* total rows of `deco`                                   */
  const int32_t *trans;          /* [N][K]   T(site, col)                           */
  const int32_t *symm_of_site;   /* [N]      symmetry group, -1 = background        */
  const int32_t *symm_count;     /* [n_symm] N_g sites per group                    */
  const double  *bf;             /* [D][S]   basis functions                        */
  const double  *eci;            /* [n_eci]  initial ECIs (same for all replicas)   */
  const int32_t *eci_kind;       /* [n_eci]  enum cemc_eci_kind                     */
  const int32_t *eci_bf;         /* [n_eci]  singlets: decoration number            */
  const int32_t *fam_size;       /* [n_fam]  n (2..4)                               */
  const int32_t *fam_nsub;       /* [n_fam]  M sub-clusters                         */
  const int32_t *fam_pos_off;    /* [n_fam+1] offsets into fam_pos                  */
  const int32_t *fam_pos;        /* per family [M][n]: sorted position k holds the
                                    changed site (CEMC_POS_REF) or column 0..K-1   */
  const int32_t *term_fam;       /* [n_symm][n_eci] family id or -1                 */
  const int32_t *term_count;     /* [n_symm][n_eci] cluster_symm_group_count[prefix]*/
  const int32_t *term_deco_off;  /* [n_symm*n_eci + 1] range of rows in `deco`      */
  const int8_t  *deco;           /* [n_deco][4] equivalent decorations, stored order */
  /* optional hint (may be NULL): the supercell is L1 x L2 x L3 primitive cells with
   * site = (i L2 + j) L3 + k.  cemc_create CHECKS that every used translation-matrix
   * column is a periodic shift on that grid; if so the kernels compute T(site, col)
   * by index arithmetic instead of reading `trans` (same values, no table traffic).   */
  const int32_t *lattice_dims;   /* [3] or NULL                                         */
} cemc_tables;

/* per-replica accumulator slots (doubles), see cemc_get_accumulators */
enum cemc_acc_slot {
  CEMC_ACC_COUNT     = 0,   /* number of sampled steps                              */
  CEMC_ACC_E         = 1,   /* sum E/ref      (Averager, cemc/mcmc/averager.py:21)   */
  CEMC_ACC_E2        = 2,   /* sum E*E/ref                                          */
  CEMC_ACC_SINGLET0  = 3    /* then per singlet d: sum s, sum s*s, sum s*E          */
};
#define CEMC_ACC_STRIDE(n_singlets) (3 + 3 * (n_singlets))

/* summation-order policy for the cluster spin products */
enum cemc_order_mode {
  CEMC_ORDER_REFERENCE = 0, /* reference operation order: bit-exact CF and E       */
  CEMC_ORDER_TREE      = 1  /* lanes over sub-clusters + tree reduction; bit-exact
                               only when all basis-function values are small
                               integers (binary +-1), else ~1e-15 relative         */
};

typedef struct cemc_handle cemc_handle;

const char *cemc_last_error(void);
int  cemc_version(void);

/* n_replicas chains on `device`; `stream` is a cudaStream_t or NULL (own stream).
 * replica_offset = global index of local replica 0 (keys the Philox streams so
 * that sharding over GPUs does not change any chain).                        */
int cemc_create(const cemc_tables *tables, int n_replicas, int replica_offset,
                int device, void *stream, cemc_handle **out);
int cemc_destroy(cemc_handle *h);
/* Round-robin sharding (replica g on GPU g mod n_gpus): the global id of
 * local replica r is replica_offset + r * stride (default stride 1 = a contiguous block).
 * With stride > 1 cemc_pt_exchange expects the gathered energies in all-gather (rank-major)
 * order and n_total == n_replicas * stride.                                              */
int cemc_set_replica_stride(cemc_handle *h, int stride);
int cemc_set_stream(cemc_handle *h, void *stream);
int cemc_synchronize(cemc_handle *h);
int cemc_set_order_mode(cemc_handle *h, int mode);
/* testing hooks: force the generic (shared-memory CF) kernel path; compare the
 * FMA-based exact division used by the kernels with IEEE division on
 * n_blocks*256*iters random operand pairs                                    */
int cemc_set_generic_path(cemc_handle *h, int on);
/* trial moves evaluated speculatively per batch by the batch kernel
 * (cemc_batch_kernel.cuh): 0 = auto, 4/8/16, -1 = one move at a time           */
int cemc_set_batch(cemc_handle *h, int b);
/* All kernel variants (spin / batch (B,C) / one move at a time) give the same
 * trajectory bit for bit; on long runs the fastest one is measured on segments
 * of the run itself.  cemc_get_variant: 0 spin, 1..4 batch (warps, CTAs per chain)
 * (16,2) (16,1) (8,1) (4,1), 5 mc_kernel, 6 batch (8,1) with two moves per warp
 * (binary +-1 basis), 8 / 9 batch (16,2) / (8,2) with site split (swaps),
 * -1 not tuned yet.                                                            */
int cemc_set_autotune(cemc_handle *h, int on);
int cemc_get_variant(cemc_handle *h, int *sgc, int *canonical);
/* variant of the most recent Metropolis launch (run_sgc / run_canonical / replay), -1 = none */
int cemc_last_variant(cemc_handle *h, int *variant);
/* 1 when the speculative batch kernel applies to this system (size limits in DESIGN.md) */
int cemc_batch_applicable(cemc_handle *h, int *yes);
/* pin the variant per sampler (-1 = let the autotuner decide); inapplicable variants
 * fall back to the default preference order                                     */
int cemc_set_variant(cemc_handle *h, int sgc, int canonical);
/* CTAs (SMs) of one thread-block cluster that cooperate on ONE chain in the batch
 * kernel: 0 = auto (2 when 2 x replicas still fit the GPU in one wave), 1, 2   */
int cemc_set_cluster(cemc_handle *h, int c);
/* testing hook: 0 = do not use the binary spin kernel (cemc_spin_kernel.cuh)   */
int cemc_set_spin_kernel(cemc_handle *h, int on);
/* testing hook: 0 = keep the fp64 product evaluation in the batch kernel instead of
 * the product tables (cemc_batch_kernel.cuh, EV_TAB); both give the same bits      */
int cemc_set_table_eval(cemc_handle *h, int on);
/* evaluation scheme the batch kernel uses for this system: 0 fp64 products, 1 binary
 * spin (XOR / popcount), 2 product tables, 3 fp32 product tables                    */
int cemc_get_batch_eval(cemc_handle *h, int *ev);
/* Translation by index arithmetic (cemc_tables.lattice_dims, verified at create): on by
 * default when the table is a periodic shift table; 0 = gather T(site, col) from `trans`
 * (same values: a pure performance knob / testing hook).  get: 1 when in use.            */
int cemc_set_lattice_arithmetic(cemc_handle *h, int on);
int cemc_get_lattice_arithmetic(cemc_handle *h, int *on);
/* Precision of the cluster-product sums: 64 (default; bit-identical to the reference's
 * fp64 CEUpdater) or 32 = the fp32 variant: product tables and sums over sub-clusters
 * (spin_product_one_atom, ce_updater.cpp:244-285) in single precision, quotients, CF
 * vector, energies and observer sums in fp64.  Same accept/reject decisions as fp64
 * unless a move lies within fp32 rounding of its Metropolis threshold; CFs and
 * energies within 1e-5 relative.  Needs a system the table evaluation takes (<= 32
 * ECIs, one symmetry group, tables fit shared memory); a binary +-1 basis is exact
 * integer arithmetic in either setting.  Fails otherwise.                            */
int cemc_set_precision(cemc_handle *h, int bits);
/* Load balance of the batch kernel: CTA (cluster) i works on replica order[i] (a permutation of
 * 0..R-1; NULL = identity).  Results never depend on it (chains are keyed by replica id).     */
int cemc_set_replica_order(cemc_handle *h, const int32_t *order);
/* testing hook: widen the band in which the batch kernel's Metropolis screen
 * defers to the exact expression (factor >= 1; 1e30 = always exact)           */
int cemc_set_screen_slack(cemc_handle *h, double factor);
/* debug builds (-DCEMC_PHASE_TIMING) only: clock64() cycles warp 0 of every replica
 * spent per kernel phase in the last launch                                   */
int cemc_debug_phase_cycles(cemc_handle *h, uint64_t *out /*[n_replicas][24]*/);
int cemc_selftest_division(cemc_handle *h, uint64_t seed, int n_blocks, int iters,
                           uint64_t *mismatches);
/* threads per CTA (= per replica): 0 = auto, else a multiple of 32 in [32,256] */
int cemc_set_block_threads(cemc_handle *h, int n);

/* ---- state ---- */
int cemc_set_occupancy(cemc_handle *h, const int8_t *occ /*[R][N]*/);
int cemc_get_occupancy(cemc_handle *h, int8_t *occ /*[R][N]*/);
int cemc_set_cf(cemc_handle *h, const double *cf /*[R][n_eci]*/);
int cemc_get_cf(cemc_handle *h, double *cf /*[R][n_eci]*/);
/* brute-force CFs from the definition, all replicas */
int cemc_recompute_cf(cemc_handle *h);
/* per_replica = 0: eci is [n_eci] broadcast; 1: [R][n_eci].  Energies are
 * re-evaluated (montecarlo.py:203, sgc_montecarlo.py:260).                   */
int cemc_set_ecis(cemc_handle *h, const double *eci, int per_replica);
int cemc_get_ecis(cemc_handle *h, double *eci /*[R][n_eci]*/);
int cemc_get_energy(cemc_handle *h, double *energy /*[R]*/);
int cemc_set_kT(cemc_handle *h, const double *kT /*[R]*/);
int cemc_get_kT(cemc_handle *h, double *kT /*[R]*/);
int cemc_seed(cemc_handle *h, uint64_t seed);
/* Philox counter = per-replica step index; set it to resume a recorded chain */
int cemc_set_step(cemc_handle *h, const uint64_t *steps /*[R]*/);
/* species ids the SGC sampler may insert (SGCMonteCarlo(symbols=...),
 * sgc_montecarlo.py:38-43); default: all species                            */
int cemc_set_sgc_species(cemc_handle *h, int n_allowed, const int8_t *allowed);
/* checkpointing of the canonical sampler: per-species site lists, species-major
 * ([R][N]; species s occupies [off[s], off[s+1]) of a replica's row)           */
int cemc_get_tracker(cemc_handle *h, int32_t *list /*[R][N]*/, int32_t *off /*[R][S+1]*/);
int cemc_set_tracker(cemc_handle *h, const int32_t *list /*[R][N]*/);
int cemc_get_counters(cemc_handle *h, uint64_t *steps /*[R]*/, uint64_t *accepted /*[R]*/);
int cemc_reset_counters(cemc_handle *h);

/* ---- the reference's per-call surface (one replica, trial semantics) ---- */
/* CEUpdater::calculate: apply changes sequentially, keep them undoable.     */
int cemc_trial_changes(cemc_handle *h, int replica, int n_changes,
                       const int32_t *sites, const int8_t *old_species,
                       const int8_t *new_species, double *energy_out);
int cemc_undo_changes(cemc_handle *h, int replica);
int cemc_clear_history(cemc_handle *h, int replica);

/* ---- batched Metropolis ---- */
/* Replay recorded proposals + uniforms.  sites[.][1] < 0
 * marks a one-site (SGC) step.  accepted_out / e_after_out may be NULL.  When all
 * steps are one-site or all are two-site the trajectory runs through the samplers'
 * own kernels (the variant pinned with cemc_set_variant, else the default order;
 * cemc_last_variant tells which); mixed records use the generic kernel.        */
int cemc_replay(cemc_handle *h, int n_steps,
                const int32_t *sites /*[R][n_steps][2]*/,
                const int8_t *new_species /*[R][n_steps][2]*/,
                const double *uniforms /*[R][n_steps]*/,
                uint8_t *accepted_out /*[R][n_steps]*/,
                double *e_after_out /*[R][n_steps]*/);
/* n_steps Metropolis trial moves per replica with on-device Philox proposals. */
int cemc_run_sgc(cemc_handle *h, int64_t n_steps);
int cemc_run_canonical(cemc_handle *h, int64_t n_steps);
/* optional trace of the last run (device proposals): enable before running   */
int cemc_set_trace(cemc_handle *h, int64_t capacity_steps);
int cemc_get_trace(cemc_handle *h, int64_t n_steps, int32_t *sites /*[R][n][2]*/,
                   int8_t *new_species /*[R][n][2]*/, double *uniforms /*[R][n]*/,
                   uint8_t *accepted /*[R][n]*/, double *e_after /*[R][n]*/);

/* Statistics of the energies traced by the last run (cemc_set_trace), computed on the
 * device so that the trace stays there: out[r] = {mean, variance, first lag k with
 * autocorrelation(k) / (n var) < 1/2 or -1, smallest normalised autocorrelation seen}
 * -- what Montecarlo._estimate_correlation_time (cemc/mcmc/montecarlo.py:461-511) needs. */
int cemc_energy_autocorrelation(cemc_handle *h, int64_t n_steps, double *out /*[R][4]*/);

/* ---- observers (Averager / SGCObserver sums) ---- */
/* 0: run_sgc / run_canonical skip the per-step sums (legs whose averages nobody reads:
 * bias probes, equilibration windows, parallel-tempering burn-in); default 1             */
int cemc_set_observe(cemc_handle *h, int on);
int cemc_reset_accumulators(cemc_handle *h, const double *ref /*[R] or NULL (=1.0)*/);
int cemc_get_accumulators(cemc_handle *h, double *acc /*[R][CEMC_ACC_STRIDE(D)]*/);

/* ---- device-side state observers ----
 * Observers of the chain state with fixed semantics, folded on the device every `interval`
 * steps of run_sgc / run_canonical (no host round trip per interval; boundaries are counted
 * across launches from the last reset):
 *   CEMC_OBS_CF_SUMS    sum and sum of squares of every CF    PairCorrelationObserver  cemc/mcmc/mc_observers.py:81-136
 *   CEMC_OBS_LOWEST     lowest energy on a boundary, its CFs
 *                       and occupations (strictly lower wins)   LowestEnergyStructure    :138-183
 *   CEMC_OBS_ENERGY     the energy on every boundary            EnergyEvolution / EnergyHistogram  :689-761
 *   CEMC_OBS_SITE_ORDER number of sites differing from a
 *                       reference configuration, sum / sum^2    SiteOrderParameter       :614-686          */
enum cemc_observer_flag {
  CEMC_OBS_CF_SUMS = 1, CEMC_OBS_LOWEST = 2, CEMC_OBS_ENERGY = 4, CEMC_OBS_SITE_ORDER = 8
};
/* interval <= 0 or flags == 0 switches them off; capacity = length of the per-replica ring of
 * energy samples (sample k is stored at k % capacity: read it back before it wraps)           */
int cemc_set_device_observers(cemc_handle *h, int64_t interval, int flags, int64_t capacity);
/* zero the sums and the boundary counter; occ_ref [R][N] = reference configuration of
 * CEMC_OBS_SITE_ORDER (NULL: the current occupations)                                         */
int cemc_reset_device_observers(cemc_handle *h, const int8_t *occ_ref);
/* any output pointer may be NULL                                                              */
int cemc_get_device_observers(cemc_handle *h, uint64_t *n_samples /*[R]*/, double *cf_sum /*[R][n_eci]*/,
                              double *cf_sq /*[R][n_eci]*/, double *best_energy /*[R]*/,
                              double *best_cf /*[R][n_eci]*/, int8_t *best_occ /*[R][N]*/,
                              double *site_order /*[R][2]*/, double *energies /*[R][n_energies]*/,
                              int64_t n_energies);

/* ---- parallel tempering ---- */
/* One exchange sweep over temperature slots.  slot_of_replica[g] for all
 * n_total global replicas and the matching energies (device pointer, e.g. the
 * output of an NCCL all-gather).  direction 0 = "up", 1 = "down", -1 = drawn on
 * the device from Philox(seed; round, stream 3) (random.choice per cycle,
 * parallel_tempering.py:191; the same on every rank).  Updates
 * slot_of_replica_dev in place (identically on every rank) and the kT of the
 * local replicas from kT_of_slot_dev.  n_accepted_dev (may be NULL) points at
 * int32[2]: [0] = exchanges accepted by this sweep, [1] += the same (running
 * total, so a sync-free round loop reads it back once at the end).  The call
 * only enqueues work on the handle's stream.                                 */
int cemc_pt_exchange(cemc_handle *h, int n_total, const double *energies_dev,
                     int32_t *slot_of_replica_dev, const double *kT_of_slot_dev,
                     int direction, uint64_t round, int32_t *n_accepted_dev);
/* device pointer to the local energies [R] (input of the all-gather)        */
int cemc_energy_dev(cemc_handle *h, double **ptr);

/* ---- timing on the handle's stream (CUDA events) ---- */
int cemc_timer_start(cemc_handle *h);
int cemc_timer_stop(cemc_handle *h, float *ms);
/* kernels launched by this handle since creation (bench gpu_launches)       */
int cemc_launch_count(cemc_handle *h, uint64_t *n);

#ifdef __cplusplus
}
#endif
#endif /* CEMC_B200_H */
