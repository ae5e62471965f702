// cemc_batch_launch.cuh -- host-side launcher of batch_kernel, shared by the three
// translation units that instantiate it (one per evaluation scheme: fp64 products,
// binary spin, product tables), so that they compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "cemc_batch_kernel.cuh"

namespace cemc {

struct BatchLaunch {
  int mode;               // MODE_SGC | MODE_CANONICAL
  int tree;               // TREE summation order
  int B, C, M;            // warps per CTA, CTAs per chain, moves per evaluation warp
  int split;              // site split: the two CTAs of a cluster evaluate the two sites of the same swaps
  int extras;             // replay / lattice arithmetic / observer boundaries in use: the kX = true kernels
  int R;                  // replicas
  int max_smem_optin;
  cudaStream_t stream;
  DeviceTables t;
  ReplicaState st;
  RunArgs a;
  int acc_stride;
  SpinTables sp;
  TabTables tb;
};

// -1: not applicable (does not fit); 0: launched; > 0: cudaError_t + 1000
int batch_launch_product(const BatchLaunch &L);
int batch_launch_spin(const BatchLaunch &L);
int batch_launch_tab(const BatchLaunch &L);
int batch_launch_tab32(const BatchLaunch &L);

template <int MODE, bool kTree, int B, bool kSmem, int C, int EV, int M, bool kSplit = false, bool kWide = false, int E = 1, bool kX = true>
static int batch_launch_kc(const BatchLaunch &L, size_t sm) {
  auto kern = batch_kernel<MODE, kTree, B, kSmem, C, EV, M, kSplit, kWide, E, kX>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return 1000 + (int)e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(L.R * C);
  cfg.blockDim = dim3((B + 1) * 32);
  cfg.dynamicSmemBytes = sm;
  cfg.stream = L.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = C > 1 ? 1 : 0;
  e = cudaLaunchKernelEx(&cfg, kern, L.t, L.st, L.a, L.acc_stride, L.sp, L.tb);
  if (e == cudaSuccess) e = cudaGetLastError();
  return e == cudaSuccess ? 0 : 1000 + (int)e;
}

template <int MODE, bool kTree, int B, int C, int EV, int M = 1>
static int batch_launch_b(const BatchLaunch &L) {
  const TabTables *tb = (EV == EV_TAB || EV == EV_TAB32) ? &L.tb : nullptr;
  constexpr bool f32 = (EV == EV_TAB32);
  const int wq = EV == EV_SPIN ? L.sp.wq : 0;
  size_t sm = batch_smem_layout<B, B * C * M>(nullptr, nullptr, L.t, MODE == MODE_CANONICAL, true, tb, f32, wq);
  const bool in_smem = sm <= (size_t)L.max_smem_optin;
  if (!in_smem) sm = batch_smem_layout<B, B * C * M>(nullptr, nullptr, L.t, MODE == MODE_CANONICAL, false, tb, f32, wq);
  if (sm > (size_t)L.max_smem_optin) return -1;
  if constexpr (M > 1) {            // two moves per warp: shared-memory state only
    if (in_smem && !L.extras) return batch_launch_kc<MODE, kTree, B, true, C, EV, M, false, false, 1, false>(L, sm);
    return in_smem ? batch_launch_kc<MODE, kTree, B, true, C, EV, M>(L, sm) : -1;
  } else {
    if (in_smem && !L.extras && B >= 7) return batch_launch_kc<MODE, kTree, B, true, C, EV, M, false, false, 1, false>(L, sm);
    return in_smem ? batch_launch_kc<MODE, kTree, B, true, C, EV, M>(L, sm)
                   : batch_launch_kc<MODE, kTree, B, false, C, EV, M>(L, sm);
  }
}

// site split (canonical, cluster of 2, shared-memory state, spin / table evaluation)
template <int MODE, bool kTree, int B, int EV>
static int batch_launch_split(const BatchLaunch &L) {
  if constexpr (MODE == MODE_CANONICAL && EV != EV_PRODUCT) {
    const TabTables *tb = (EV == EV_TAB || EV == EV_TAB32) ? &L.tb : nullptr;
    const size_t sm = batch_smem_layout<B, B>(nullptr, nullptr, L.t, true, true, tb, EV == EV_TAB32, EV == EV_SPIN ? L.sp.wq : 0);
    if (sm > (size_t)L.max_smem_optin) return -1;
    if (!L.extras) return batch_launch_kc<MODE, kTree, B, true, 2, EV, 1, true, false, 1, false>(L, sm);
    return batch_launch_kc<MODE, kTree, B, true, 2, EV, 1, true>(L, sm);
  } else {
    return -1;
  }
}

// 32 <= K <= 63 translation columns (spin / table evaluation, one CTA per chain, shared-memory state)
template <int MODE, bool kTree, int B, int EV>
static int batch_launch_wide(const BatchLaunch &L) {
  if constexpr (EV != EV_PRODUCT) {
    const TabTables *tb = (EV == EV_TAB || EV == EV_TAB32) ? &L.tb : nullptr;
    const size_t sm = batch_smem_layout<B, B>(nullptr, nullptr, L.t, MODE == MODE_CANONICAL, true, tb, EV == EV_TAB32,
                                              EV == EV_SPIN ? L.sp.wq : 0);
    if (sm > (size_t)L.max_smem_optin) return -1;
    return batch_launch_kc<MODE, kTree, B, true, 1, EV, 1, false, true>(L, sm);
  } else {
    return -1;
  }
}

// 33..64 ECIs: two per lane (table / product evaluation, one CTA per chain, shared-memory state)
template <int MODE, bool kTree, int B, int EV>
static int batch_launch_e2(const BatchLaunch &L) {
  if constexpr (EV != EV_SPIN) {
    const TabTables *tb = (EV == EV_TAB || EV == EV_TAB32) ? &L.tb : nullptr;
    const size_t sm = batch_smem_layout<B, B, 2>(nullptr, nullptr, L.t, MODE == MODE_CANONICAL, true, tb, EV == EV_TAB32, 0);
    if (sm > (size_t)L.max_smem_optin) return -1;
    return batch_launch_kc<MODE, kTree, B, true, 1, EV, 1, false, false, 2>(L, sm);
  } else {
    return -1;
  }
}

template <int MODE, bool kTree, int EV>
static int batch_launch_bc(const BatchLaunch &L) {
  if (L.t.n_eci > 32) {
    if (L.t.n_eci > 64 || L.split || L.M != 1 || L.C != 1 || L.t.K > 31) return -1;
    if (L.B == 16) return batch_launch_e2<MODE, kTree, 15, EV>(L);
    if (L.B == 8) return batch_launch_e2<MODE, kTree, 7, EV>(L);
    if (L.B == 4) return batch_launch_e2<MODE, kTree, 3, EV>(L);      // large product programs: scratch of 3 moves fits
    return -1;
  }
  if (L.t.K > 31) {
    if (L.split || L.M != 1 || L.C != 1) return -1;
    if (L.B == 16) return batch_launch_wide<MODE, kTree, 15, EV>(L);
    if (L.B == 8) return batch_launch_wide<MODE, kTree, 7, EV>(L);
    return -1;
  }
  if (L.split) {
    if (L.C != 2) return -1;
    if (L.B == 16) return batch_launch_split<MODE, kTree, 15, EV>(L);
    if (L.B == 8) return batch_launch_split<MODE, kTree, 7, EV>(L);      // hot chains on small cells: short batches
    return -1;
  }
  // L.B counts the warps of a CTA: B - 1 evaluation warps (moves per batch) + the observer
  // warp; power-of-two CTAs get the full register budget (512 threads x 128 registers)
  if (L.M == 2) {
    // two moves per evaluation warp, their instruction streams interleaved (14-move batches at 8 warps; spin
    // evaluation only).  Slower than one move per warp in round 1; since every warp decides and the conflict
    // masks cost the same for any batch length it wins on config 2 (173 -> 157 ns/move; M = 3: 190, M = 4: 238)
    if constexpr (EV == EV_SPIN) {
      if (L.B == 8 && L.C == 1) return batch_launch_b<MODE, kTree, 7, 1, EV, 2>(L);
    }
    return -1;
  }
  if (L.M != 1) return -1;
  if (L.B == 16 && L.C == 2) return batch_launch_b<MODE, kTree, 15, 2, EV>(L);
  if (L.B == 16 && L.C == 1) return batch_launch_b<MODE, kTree, 15, 1, EV>(L);
  if (L.B == 8 && L.C == 1) return batch_launch_b<MODE, kTree, 7, 1, EV>(L);
  if (L.B == 4 && L.C == 1) return batch_launch_b<MODE, kTree, 3, 1, EV>(L);
  return -1;
}

// kBothOrders = false: only the TREE instantiation exists (binary +-1 basis: every
// order gives the same bits, the host always asks for TREE)
template <int EV, bool kBothOrders>
static int batch_launch_ev(const BatchLaunch &L) {
  if (L.mode == MODE_SGC) {
    if constexpr (kBothOrders) { if (!L.tree) return batch_launch_bc<MODE_SGC, false, EV>(L); }
    return batch_launch_bc<MODE_SGC, true, EV>(L);
  }
  if (L.mode == MODE_CANONICAL) {
    if constexpr (kBothOrders) { if (!L.tree) return batch_launch_bc<MODE_CANONICAL, false, EV>(L); }
    return batch_launch_bc<MODE_CANONICAL, true, EV>(L);
  }
  return -1;
}

}  // namespace cemc
