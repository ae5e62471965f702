// cemc_spin_kernel.cuh -- specialised Metropolis kernel for binary systems whose
// single basis function is the spin {+1, -1} (the ase.clease basis of every
// binary alloy; BASELINE configs 1 and 2).
//
// One WARP per replica, no block-level barriers.  With spins +-1 every term of
// spin_product_one_atom (/root/reference/cpp/src/ce_updater.cpp:244-285) is
// +-1, so the sums over sub-clusters are exact integers in any order and the
// reference's result is reproduced bit for bit by
//
//   item lane   XOR of the neighbour occupation bits of one sub-cluster
//   __ballot    one 32-bit word per 32 sub-clusters
//   ECI lane    S = M - 2 popc(ballot & mask)          (sum of neighbour products)
//               I = n * (sigma_new - sigma_old) * S     (exact integer numerator, :397-400)
//               delta = RN(I / (count * N))             (:402, FMA exact division)
//               cf'  = cf + delta                       (:404)
//
// followed by the ordered energy dot product over warp shuffles
// (named_array.cpp:25-33) and the Metropolis test (montecarlo.py:951-956).
// The CF vector, ECIs and observer sums live in registers (lane i <-> ECI i).
#pragma once
#include "cemc_kernels.cuh"

namespace cemc {

struct SpinTables {
  int n_items;                 // sub-clusters of all cluster ECIs (one site change)
  int n_rounds;                // ceil(n_items / 32)
  const uint32_t *items;       // [n_items] col0 | col1<<8 | col2<<16 (0xff = none)
  const uint32_t *masks;       // [32][4] ballot masks of ECI lane i, round r
  const int32_t *coef;         // [32] n * b0^(n-1) (clusters), 1 (singlets), 0 (copied)
  const int32_t *msub;         // [32] M (clusters), 1 (singlets)
  int b0;                      // bf[0][species 0] = +1 or -1
  int wq;                      // max msub + 1: width of the batch kernel's quotient table
};

template <int MODE, int NR>
__global__ void __launch_bounds__(32)
spin_kernel(SpinTables sp, DeviceTables t, ReplicaState st, RunArgs a, int acc_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool kCanon = (MODE == MODE_CANONICAL);
  constexpr int NJ = kCanon ? 2 : 1;
  const int r = blockIdx.x;
  const int lane = threadIdx.x;
  const int N = t.N, K = t.K, n_eci = t.n_eci;

  // ---- shared memory: ring | quotient table | items | [list] | occ ----------
  uint4 *ring = reinterpret_cast<uint4 *>(smem_raw);                   // [32][2]
  double *s_qtab = reinterpret_cast<double *>(ring + 64);              // [2][wq][32]
  uint32_t *s_items = reinterpret_cast<uint32_t *>(s_qtab + sp.wq * 64);   // [n_items]
  int32_t *s_list = reinterpret_cast<int32_t *>(s_items + ((sp.n_items + 3) & ~3));
  int8_t *s_occ = reinterpret_cast<int8_t *>(s_list + (kCanon ? ((N + 3) & ~3) : 0));

  int8_t *g_occ = st.occ + (size_t)r * N;
  int32_t *g_list = st.list + (size_t)r * N;
  int32_t *g_loc = st.loc + (size_t)r * N;
  for (int i = lane; i < sp.n_items; i += 32) s_items[i] = sp.items[i];
  for (int i = lane; i < N; i += 32) s_occ[i] = g_occ[i];
  int off1 = 0;                                   // species 1 starts here in the site lists
  if (kCanon) {
    for (int i = lane; i < N; i += 32) s_list[i] = g_list[i];
    off1 = st.off[(size_t)r * 3 + 1];
    if ((off1 == 0 || off1 == st.off[(size_t)r * 3 + 2]) && a.rp_sites == nullptr) {   // TooFewElementsError
      if (lane == 0) st.status[r] = 2;
      return;
    }
  }
  const int n_tot = kCanon ? st.off[(size_t)r * 3 + 2] : 0;
  __syncwarp();

  // ---- lane i owns ECI i ----------------------------------------------------
  const bool mine = lane < n_eci;
  const int4 f = mine ? t.fin_i[lane] : make_int4(0, 0, 0, 0);
  const int f_kind = f.x;
  const double dN = (double)(unsigned)N;
  const double f_den = (f_kind == 1) ? dN : (mine ? t.fin_d[lane].y : 1.0);
  const double f_rden = __ddiv_rn(1.0, f_den);
  const int coef = sp.coef[lane], msub = sp.msub[lane];
  // every quotient n (sigma_new - sigma_old)(M - 2 count) / den this lane's ECI can take, once per
  // launch: the exact division (ce_updater.cpp:402) becomes one shared-memory load per site
  for (int nw = 0; nw < 2; nw++)
    for (int cnt = 0; cnt < sp.wq; cnt++) {
      const int num = coef * (2 * sp.b0 * (1 - 2 * nw)) * (msub - 2 * cnt);
      s_qtab[(nw * sp.wq + cnt) * 32 + lane] = cnt <= msub ? exact_div((double)num, f_den, f_rden) : 0.0;
    }
  __syncwarp();
  uint32_t mask[NR];
#pragma unroll
  for (int q = 0; q < NR; q++) mask[q] = sp.masks[lane * 4 + q];
  const double eci_reg = mine ? st.eci[(size_t)r * n_eci + lane] : 0.0;
  double cf_reg = mine ? st.cf[(size_t)r * n_eci + lane] : 0.0;
  int my_singlet = -1;
  for (int d = 0; d < t.n_singlets; d++) if (t.singlet_idx[d] == lane) my_singlet = d;
  const double *acc_g = st.acc + (size_t)r * acc_stride;
  double aE0 = acc_g[0], aE1 = acc_g[1], aE2 = acc_g[2];
  double aS0 = 0.0, aS1 = 0.0, aS2 = 0.0;
  if (my_singlet >= 0) { aS0 = acc_g[3 + 3 * my_singlet]; aS1 = acc_g[4 + 3 * my_singlet]; aS2 = acc_g[5 + 3 * my_singlet]; }

  double e_cur = st.e_cur[r];
  const double kT = st.kT[r];
  const double rkT = __ddiv_rn(1.0, kT);
  const double ref = st.ref[r];
  const double rref = __ddiv_rn(1.0, ref);
  const bool ref_is_one = (ref == 1.0);
  const unsigned long long step0 = st.step[r];
  unsigned long long n_acc = 0;
  const uint32_t rep_global = a.replica_offset + (uint32_t)r * a.replica_stride;
  const int32_t *__restrict__ trans = t.trans;
  const int b0 = pin_reg(sp.b0);
  const int Kr = pin_reg(K);
  const int n_eci4 = pin_reg((n_eci + 3) & ~3);
  const int observe = pin_reg(a.observe);
  const bool tracing = (a.tr_acc != nullptr) || (a.tr_e != nullptr);
  // replay of recorded proposals / uniforms: records hold sites, not list slots
  const bool replay = (a.rp_sites != nullptr);

  // ---- this lane's items (sub-clusters), decoded once: branch-free evaluation --
  // unused neighbour slots alias slot a and are masked out of the XOR
  int ca[NR], cb[NR], cc[NR];
  uint32_t mb[NR], mc[NR], mv[NR];
#pragma unroll
  for (int q = 0; q < NR; q++) {
    const int qi = q * 32 + lane;
    const bool valid = qi < sp.n_items;
    const uint32_t w = valid ? s_items[qi] : 0x00ffff00u;
    ca[q] = (int)(w & 0xffu);
    const uint32_t xb = (w >> 8) & 0xffu, xc = (w >> 16) & 0xffu;
    cb[q] = xb != 0xffu ? (int)xb : ca[q];
    cc[q] = xc != 0xffu ? (int)xc : ca[q];
    mb[q] = xb != 0xffu ? 1u : 0u;
    mc[q] = xc != 0xffu ? 1u : 0u;
    mv[q] = valid ? 1u : 0u;
  }

  // neighbour site indices of the NEXT move are fetched one move ahead (the T row
  // of an SGC move depends on the Philox stream only; for swaps the site comes
  // from the species lists and is re-validated when the move is evaluated)
  int pna[NJ][NR], pnb[NJ][NR], pnc[NJ][NR], psite[NJ];
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    psite[j] = -1;
#pragma unroll
    for (int q = 0; q < NR; q++) { pna[j][q] = 0; pnb[j][q] = 0; pnc[j][q] = 0; }
  }

  for (long long it0 = 0; it0 < a.n_steps; it0 += 32) {
    const int nblk = (int)((a.n_steps - it0) < 32 ? (a.n_steps - it0) : 32);
    // ---- refill: Philox proposals for 32 moves, one per lane ----------------
    {
      const unsigned long long stp = step0 + (unsigned long long)it0 + lane;
      uint32_t c0 = (uint32_t)stp, c1 = (uint32_t)(stp >> 32), c2 = rep_global, c3 = 0;
      philox4x32_10(c0, c1, c2, c3, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
      uint4 rec0, rec1;
      if (replay) {
        const long long q = it0 + lane;
        int s0 = 0, s1 = kCanon ? 1 : -1, n0 = 0, n1 = 0;
        double u = 0.5;
        if (q < a.n_steps) {
          const size_t g = (size_t)r * (size_t)a.n_steps + (size_t)q;
          s0 = a.rp_sites[2 * g]; n0 = a.rp_news[2 * g];
          if (kCanon) { s1 = a.rp_sites[2 * g + 1]; n1 = a.rp_news[2 * g + 1]; }
          u = a.rp_u[g];
        }
        rec0 = make_uint4((uint32_t)s0, (uint32_t)s1, (uint32_t)n0, (uint32_t)n1);
        rec1 = make_uint4((uint32_t)__double2loint(u), (uint32_t)__double2hiint(u), 0u, 0u);
      } else if (!kCanon) {
        // sgc_montecarlo.py:69: site uniform; binary: the new species is the other one
        const uint32_t ia = __umulhi(c0, (uint32_t)N);
        const double u = u53(c2, c3);
        rec0 = make_uint4(ia, c1, 0u, 0u);
        rec1 = make_uint4((uint32_t)__double2loint(u), (uint32_t)__double2hiint(u), 0u, 0u);
      } else {
        // montecarlo.py:899-907 with two species present: (a, b) = (0,1) or (1,0)
        uint32_t d0 = (uint32_t)stp, d1 = (uint32_t)(stp >> 32), d2 = rep_global, d3 = 1;
        philox4x32_10(d0, d1, d2, d3, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        const int ia = (int)__umulhi(c0, 2u);
        int ib = (int)__umulhi(c1, 1u); ib += (ib >= ia);          // the other species
        const int cnt_a = ia ? n_tot - off1 : off1, cnt_b = ib ? n_tot - off1 : off1;
        const int slot0 = (ia ? off1 : 0) + (int)__umulhi(c2, (uint32_t)cnt_a);
        const int slot1 = (ib ? off1 : 0) + (int)__umulhi(c3, (uint32_t)cnt_b);
        const double u = u53(d0, d1);
        rec0 = make_uint4((uint32_t)slot0, (uint32_t)slot1, (uint32_t)ib, (uint32_t)ia);
        rec1 = make_uint4((uint32_t)__double2loint(u), (uint32_t)__double2hiint(u), 0u, 0u);
      }
      __syncwarp();
      ring[lane * 2] = rec0; ring[lane * 2 + 1] = rec1;
      __syncwarp();
    }

    for (int ib_ = 0; ib_ < nblk; ib_++) {
      // ---- proposal -----------------------------------------------------------
      const uint4 rec0 = ring[ib_ * 2], rec1 = ring[ib_ * 2 + 1];
      const double u = __hiloint2double((int)rec1.y, (int)rec1.x);
      int site[NJ], newsp[NJ], oldsp[NJ], slot0 = -1, slot1 = -1;
      if (replay) {
        site[0] = (int)rec0.x; newsp[0] = (int)rec0.z;
        oldsp[0] = s_occ[site[0]];
        if (kCanon) {
          site[NJ - 1] = (int)rec0.y; newsp[NJ - 1] = (int)rec0.w;
          oldsp[NJ - 1] = site[NJ - 1] == site[0] ? newsp[0] : (int)s_occ[site[NJ - 1]];
        }
      } else if (!kCanon) {
        site[0] = (int)rec0.x;
        oldsp[0] = s_occ[site[0]];
        newsp[0] = 1 - oldsp[0];
      } else {
        slot0 = (int)rec0.x; slot1 = (int)rec0.y;
        newsp[0] = (int)rec0.z; newsp[NJ - 1] = (int)rec0.w;
        site[0] = s_list[slot0]; site[NJ - 1] = s_list[slot1];
        oldsp[0] = newsp[NJ - 1]; oldsp[NJ - 1] = newsp[0];
      }

      // ---- neighbour sites: prefetched last move, or loaded now -------------------
      int na[NJ][NR], nb[NJ][NR], nc[NJ][NR];
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        if (psite[j] == site[j]) {             // warp-uniform
#pragma unroll
          for (int q = 0; q < NR; q++) { na[j][q] = pna[j][q]; nb[j][q] = pnb[j][q]; nc[j][q] = pnc[j][q]; }
        } else {
          const int32_t *row = trans + (size_t)site[j] * Kr;                      // :264
#pragma unroll
          for (int q = 0; q < NR; q++) {
            na[j][q] = __ldg(row + ca[q]); nb[j][q] = __ldg(row + cb[q]); nc[j][q] = __ldg(row + cc[q]);
          }
        }
      }
      // issue the next move's T-row loads now; they complete behind this move's math
      if (ib_ + 1 < nblk) {
        const uint4 nx = ring[(ib_ + 1) * 2];
#pragma unroll
        for (int j = 0; j < NJ; j++) {
          const int sj = replay ? (j ? (int)nx.y : (int)nx.x) : kCanon ? s_list[j ? (int)nx.y : (int)nx.x] : (int)nx.x;
          psite[j] = sj;
          const int32_t *row = trans + (size_t)sj * Kr;
#pragma unroll
          for (int q = 0; q < NR; q++) {
            pna[j][q] = __ldg(row + ca[q]); pnb[j][q] = __ldg(row + cb[q]); pnc[j][q] = __ldg(row + cc[q]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < NJ; j++) psite[j] = -1;
      }

      // ---- items: XOR of neighbour occupation bits, one ballot per 32 -----------
      uint32_t ball[NJ][NR];
#pragma unroll
      for (int j = 0; j < NJ; j++) {
#pragma unroll
        for (int q = 0; q < NR; q++) {
          uint32_t va = (uint32_t)s_occ[na[j][q]];
          uint32_t vb = (uint32_t)s_occ[nb[j][q]];
          uint32_t vc = (uint32_t)s_occ[nc[j][q]];
          if (kCanon && j == 1) {               // change 1 sees change 0 applied (:845-852)
            if (na[j][q] == site[0]) va = (uint32_t)newsp[0];
            if (nb[j][q] == site[0]) vb = (uint32_t)newsp[0];
            if (nc[j][q] == site[0]) vc = (uint32_t)newsp[0];
          }
          const uint32_t b = (va ^ (vb & mb[q]) ^ (vc & mc[q])) & mv[q];
          ball[j][q] = __ballot_sync(0xffffffffu, b);
        }
      }

      // ---- ECI lanes: exact integer numerators, exact division, CF increment ----
      double c = cf_reg;
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < NR; q++) cnt += __popc(ball[j][q] & mask[q]);
        const double dl = s_qtab[(newsp[j] * sp.wq + cnt) * 32 + lane];   // n dsigma (M - 2 cnt) / den, :393-402
        if (f_kind > 0 && newsp[j] != oldsp[j]) c = __dadd_rn(c, dl);   // :404 (:315: a recorded no-op changes nothing)
      }
      const double p = __dmul_rn(eci_reg, c);
      double e_new = 0.0;                       // named_array.cpp:27-31; lanes >= n_eci add +0.0
      for (int i = 0; i < n_eci4; i += 4) {
        const double p0 = __shfl_sync(0xffffffffu, p, i), p1 = __shfl_sync(0xffffffffu, p, i + 1);
        const double p2 = __shfl_sync(0xffffffffu, p, i + 2), p3 = __shfl_sync(0xffffffffu, p, i + 3);
        e_new = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(e_new, p0), p1), p2), p3);
      }
      e_new = __dmul_rn(e_new, dN);                                // ce_updater.cpp:241
      const bool accept = metropolis(e_new, e_cur, u, kT, rkT);
      __syncwarp();               // every lane's reads of this move precede lane 0's commit
      if (accept) {
        cf_reg = c;
        e_cur = e_new;
        n_acc++;
        if (lane == 0) {
          s_occ[site[0]] = (int8_t)newsp[0];
          if (kCanon) {                        // swap_move_index_tracker.py:39-59
            s_occ[site[NJ - 1]] = (int8_t)newsp[NJ - 1];
            if (slot0 >= 0) {                  // replay: no list slots
              s_list[slot0] = site[NJ - 1]; s_list[slot1] = site[0];
              g_loc[site[NJ - 1]] = slot0 - (newsp[NJ - 1] ? off1 : 0);
              g_loc[site[0]] = slot1 - (newsp[0] ? off1 : 0);
            }
          }
        }
      }
      __syncwarp();
      if (observe) {                                               // montecarlo.py:811-814,
        const double e2 = __dmul_rn(e_cur, e_cur);                 // mc_observers.py:264-270
        aE0 = __dadd_rn(aE0, 1.0);
        aE1 = __dadd_rn(aE1, ref_is_one ? e_cur : exact_div(e_cur, ref, rref));
        aE2 = __dadd_rn(aE2, ref_is_one ? e2 : exact_div(e2, ref, rref));
        aS0 = __dadd_rn(aS0, cf_reg);
        aS1 = __dadd_rn(aS1, __dmul_rn(cf_reg, cf_reg));
        aS2 = __dadd_rn(aS2, __dmul_rn(cf_reg, e_cur));
      }
      if (tracing && lane == 0 && it0 + ib_ < a.tr_capacity) {
        const size_t q = (size_t)r * a.tr_capacity + (size_t)(it0 + ib_);
        if (a.tr_sites) { a.tr_sites[2 * q] = site[0]; a.tr_sites[2 * q + 1] = kCanon ? site[NJ - 1] : -1; }
        if (a.tr_news) { a.tr_news[2 * q] = (int8_t)newsp[0]; a.tr_news[2 * q + 1] = kCanon ? (int8_t)newsp[NJ - 1] : 0; }
        if (a.tr_u) a.tr_u[q] = u;
        if (a.tr_acc) a.tr_acc[q] = accept ? 1 : 0;
        if (a.tr_e) a.tr_e[q] = e_cur;
      }
    }
  }

  // ---- write back ------------------------------------------------------------
  __syncwarp();
  if (mine) st.cf[(size_t)r * n_eci + lane] = cf_reg;
  double *acc_w = st.acc + (size_t)r * acc_stride;
  if (lane == 0) { acc_w[0] = aE0; acc_w[1] = aE1; acc_w[2] = aE2; }
  if (my_singlet >= 0) { acc_w[3 + 3 * my_singlet] = aS0; acc_w[4 + 3 * my_singlet] = aS1; acc_w[5 + 3 * my_singlet] = aS2; }
  for (int i = lane; i < N; i += 32) g_occ[i] = s_occ[i];
  if (kCanon) for (int i = lane; i < N; i += 32) g_list[i] = s_list[i];
  if (lane == 0) {
    st.e_cur[r] = e_cur;
    st.step[r] = step0 + (unsigned long long)a.n_steps;
    st.accepted[r] += n_acc;
  }
}

}  // namespace cemc
