// CUDA kernels of the local-BA engine (sm_100a).  All arithmetic is IEEE double like g2o.
// Kernel inventory and the roofline that bounds each: DESIGN.md section 5.
#pragma once
#include "../../../include/ppo_ba.h"
#include "ppo_dense.h"
#include "ppo_device.cuh"
#include "ppo_geom.cuh"

namespace ppo {

#define PPO_EF_LEVEL1_ 1u
#define PPO_EF_ROBUST_ 2u
constexpr unsigned FULL = 0xffffffffu;

// State of the Levenberg-Marquardt controller (OptimizationAlgorithmLevenberg::solve, core/optimization_algorithm_levenberg.cpp:61-164)
// lives on the device: lambda, nu, the chi2 bookkeeping, the accept / reject decision, the trial and iteration counters, the
// termination rules and the per-iteration trace.  Kernels that need the damping read it from here, so one damped trial is a
// fixed sequence of launches that can be captured in a CUDA graph and replayed under conditional WHILE nodes.
struct LmIn {  // written by the host before an optimize() call
  int iters, max_trials;
  double tau, good_upper, good_lower;
  double chi_const;  // robust chi2 of the constant-residual cuboid-plane edges (they take no part in the linearisation)
};
struct LmDev {
  double lambda, ni, currentChi, iniChi, tempChi, rho, chi_const, chi2_initial;
  double tau, good_upper, good_lower;
  int it, iters, qmax, max_trials, nBad, done, term, accepted, total_trials, stop_seen;
  int trial_continue, iter_continue;  // loop flags (mirrored into the conditional handles when the loop runs as a graph)
  ppo_ba_iter trace[PPO_TRACE_MAX];
};
constexpr double NUM_DELTA = 1e-9;                      // base_binary_edge.hpp:232
constexpr double NUM_SCALAR = 1.0 / (2 * NUM_DELTA);    // :233

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// deterministic block sum (fixed tree), result valid in thread 0
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *sm /* THREADS/32 */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x == 0)
    for (int i = 0; i < THREADS / 32; i++) r += sm[i];
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------------
// index mapping: SparseOptimizer::initializeOptimization + buildIndexMapping
// (core/sparse_optimizer.cpp:166-267)
// ---------------------------------------------------------------------------------------------
__global__ void k_mark_active(DevGraph g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.n_pe) {
    const int kf = g.pe_rec[i].kf, pt = g.pe_pt[i];
    if (!(g.pe_flags[i] & PPO_EF_LEVEL1_) && !(g.kf_fixed[kf] && g.pt_fixed[pt])) {
      g.kf_act[kf] = 1;
      g.pt_act[pt] = 1;
    }
  }
  if (i < g.n_ple && !(g.ple_flags[i] & PPO_EF_LEVEL1_)) {
    g.kf_act[g.ple_kf[i]] = 1;
    g.pl_act[g.ple_plane[i]] = 1;
  }
  if (i < g.n_cbe && !(g.cbe_flags[i] & PPO_EF_LEVEL1_)) {
    g.kf_act[g.cbe_kf[i]] = 1;
    g.cu_act[g.cbe_cuboid[i]] = 1;
  }
  if (i < g.n_pce && !(g.pce_flags[i] & PPO_EF_LEVEL1_)) g.cu_act[g.pce_cuboid[i]] = 1;
}
// cuboid-plane edges (constant residual) only make their vertices active; host passes them packed
__global__ void k_mark_active_cpe(DevGraph g, const int *cpe_cuboid, const int *cpe_plane, const uint8_t *cpe_flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.n_cpe && !(cpe_flags[i] & PPO_EF_LEVEL1_)) {
    g.cu_act[cpe_cuboid[i]] = 1;
    g.pl_act[cpe_plane[i]] = 1;
  }
}
// compacts the free active key-frames and the active cuboids into the pose block [key-frames by slot | cuboids]: one warp, ballot + popc
// prefix sums over 32 vertices per step (launched <<<1, 32>>>)
__global__ void k_build_index(DevGraph g) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  const unsigned below = (1u << lane) - 1u;
  int nb = 0;
  for (int i0 = 0; i0 < g.n_kf; i0 += 32) {
    const int i = i0 + lane;
    const bool on = i < g.n_kf && g.kf_act[i] && !g.kf_fixed[i];
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (i < g.n_kf) g.kf_idx[i] = on ? nb + __popc(m & below) : -1;
    nb += __popc(m);
  }
  int off = 6 * nb;
  for (int i0 = 0; i0 < g.n_cu; i0 += 32) {
    const int i = i0 + lane;
    const bool on = i < g.n_cu && g.cu_act[i];
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (i < g.n_cu) g.cu_off[i] = on ? off + 9 * __popc(m & below) : -1;
    off += 9 * __popc(m);
  }
  int nl = 0;
  for (int i0 = 0; i0 < g.n_pl; i0 += 32) {
    const int i = i0 + lane;
    nl += __popc(__ballot_sync(0xffffffffu, i < g.n_pl && g.pl_act[i] != 0));
  }
  if (lane == 0) {
    g.dims[0] = nb;
    g.dims[1] = off;
    g.dims[2] = nl;  // active planes; active points are counted by k_entry_pidx
    g.dims[3] = 0;
    g.dims[4] = 0;
  }
}
__global__ void k_entry_pidx(DevGraph g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.n_slots) g.ent_pidx[i] = -2;
  if (i < g.n_pe) {
    const int kf = g.pe_rec[i].kf, pt = g.pe_pt[i];
    const bool act = !(g.pe_flags[i] & PPO_EF_LEVEL1_) && !(g.kf_fixed[kf] && g.pt_fixed[pt]);
    g.ent_pidx[g.n_slots + i] = act ? g.kf_idx[kf] : -2;
    if (act) atomicAdd(&g.dims[4], 1);
  }
  if (i < g.n_pt && g.pt_act[i] && !g.pt_fixed[i]) atomicAdd(&g.dims[3], 1);
  if (i < g.n_ple && !(g.ple_flags[i] & PPO_EF_LEVEL1_)) atomicAdd(&g.dims[4], 1);
  if (i < g.n_cbe && !(g.cbe_flags[i] & PPO_EF_LEVEL1_)) atomicAdd(&g.dims[4], 1);
  if (i < g.n_pce && !(g.pce_flags[i] & PPO_EF_LEVEL1_)) atomicAdd(&g.dims[4], 1);
}
__global__ void k_slot_pidx(DevGraph g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.n_ple && !(g.ple_flags[i] & PPO_EF_LEVEL1_)) g.ent_pidx[g.ple_slot[i]] = g.kf_idx[g.ple_kf[i]];
}

// ---------------------------------------------------------------------------------------------
// point edges: residual + analytic Jacobians (types_six_dof_expmap.cpp:135-266)
// ---------------------------------------------------------------------------------------------
struct PointEdgeLin {
  double err[3];
  double Jpt[9];   // D x 3
  double Jkf[18];  // D x 6
  int D;
};
PPO_D void point_edge_linearize(const double Rt[12], const double X[3], const float intr[5], const PointEdgeRec &rec,
                                PointEdgeLin &L) {
  double p[3];
  cam_point(Rt, X, p);
  L.D = point_edge_error(p, intr, rec.u, rec.v, rec.ur, L.err);
  const double fx = intr[0], fy = intr[1], bf = intr[4];
  const double x = p[0], y = p[1];
  // one reciprocal instead of the reference's ~15 divisions by z / z^2 (same values to 1 ulp)
  const double iz = 1.0 / p[2], iz2 = iz * iz;
  const double fxz = fx * iz, fyz = fy * iz, xz2 = x * iz2, yz2 = y * iz2;
  // d(residual)/d(camera-frame point), rows 0..2
  const double a0[3] = {-fxz, 0.0, fx * xz2};
  const double a1[3] = {0.0, -fyz, fy * yz2};
#pragma unroll
  for (int j = 0; j < 3; j++) {
    L.Jpt[j] = a0[0] * Rt[j] + a0[2] * Rt[6 + j];
    L.Jpt[3 + j] = a1[1] * Rt[3 + j] + a1[2] * Rt[6 + j];
    L.Jpt[6 + j] = L.Jpt[j] - bf * iz2 * Rt[6 + j];
  }
  L.Jkf[0] = x * yz2 * fx;
  L.Jkf[1] = -(1 + x * xz2) * fx;
  L.Jkf[2] = y * fxz;
  L.Jkf[3] = -fxz;
  L.Jkf[4] = 0;
  L.Jkf[5] = xz2 * fx;
  L.Jkf[6] = (1 + y * yz2) * fy;
  L.Jkf[7] = -x * yz2 * fy;
  L.Jkf[8] = -x * fyz;
  L.Jkf[9] = 0;
  L.Jkf[10] = -fyz;
  L.Jkf[11] = yz2 * fy;
  L.Jkf[12] = L.Jkf[0] - bf * yz2;
  L.Jkf[13] = L.Jkf[1] + bf * xz2;
  L.Jkf[14] = L.Jkf[2];
  L.Jkf[15] = L.Jkf[3];
  L.Jkf[16] = 0;
  L.Jkf[17] = L.Jkf[5] - bf * iz2;
  if (L.D == 2) {
#pragma unroll
    for (int j = 0; j < 3; j++) L.Jpt[6 + j] = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) L.Jkf[12 + j] = 0;
  }
}

constexpr int LIN_WARPS = 4;     // warps per CTA; every warp walks over work units with a grid stride (all control flow is warp-uniform)
constexpr int LIN_CTAS_PER_SM = 5;  // 20 warps per SM at 94 registers
constexpr int STAGE_LD = 18;  // the staged 6x3 blocks are dense (144 bytes each): the run leaves shared memory as ONE TMA bulk store

// The Jacobian / assembly pass over the point edges (computeActiveErrors + linearizeOplus +
// constructQuadraticForm of core/block_solver.hpp:502-560 for EdgeSE3ProjectXYZ /
// EdgeStereoSE3ProjectXYZ, landmark side): one warp owns a run of consecutive points with <= 32
// edges, one lane per edge; Hll / bl come from a segmented warp-shuffle scan, the 6x3 Hpl blocks
// are staged in shared memory and stored as one contiguous coalesced run.
__global__ void __launch_bounds__(LIN_WARPS * 32, LIN_CTAS_PER_SM) k_point_linearize(DevGraph g, DevState s, double *chi_part) {
  __shared__ __align__(16) double stage[LIN_WARPS][32 * STAGE_LD];
  __shared__ double wsum[LIN_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stride = gridDim.x * LIN_WARPS;
  double rho_sum = 0;
  int unit = blockIdx.x * LIN_WARPS + warp;
  // the bounds of the NEXT unit are requested while the current one is processed: one dependent memory round trip less per unit
  int e0n = unit < g.n_units ? g.unit_e0[unit] : 0, e1n = unit < g.n_units ? g.unit_e0[unit + 1] : 0;
  for (; unit < g.n_units; unit += stride) {
    const int e0 = e0n, e1 = e1n;
    {
      const int un = unit + stride;
      e0n = un < g.n_units ? g.unit_e0[un] : 0;
      e1n = un < g.n_units ? g.unit_e0[un + 1] : 0;
    }
    for (int base = e0; base < e1; base += 32) {
      const int e = base + lane;
      const bool valid = e < e1;
      double acc[9];  // Hll xx xy xz yy yz zz, bl x y z
#pragma unroll
      for (int i = 0; i < 9; i++) acc[i] = 0;
      double *st = &stage[warp][lane * STAGE_LD];
      int key = -1;
      bool lm_free = false, have_hpl = false;
      // All loads of the unit are issued up front in TWO dependent levels (ncu: the kernel is bound by the latency of its chain of
      // global loads, not by bandwidth): level 1 everything indexed by the edge, level 2 everything indexed by its key-frame / point.
      // Lanes beyond the end of the unit read the last edge again, so that no load sits behind a branch.
      {
        const int el = min(e, e1 - 1);
        const PointEdgeRec rec = g.pe_rec[el];
        const int pt = g.pe_pt[el];
        const unsigned fl = g.pe_flags[el];
        const float is2f = g.pe_is2[el];
        const int pidx = g.ent_pidx[g.n_slots + el];
        const bool ptfix = g.pt_fixed[pt], kffix = g.kf_fixed[rec.kf];
        double Rt[12], p[3], err[3];
        float intr[5];
#pragma unroll
        for (int i = 0; i < 12; i++) Rt[i] = __ldg(&s.kf_Rt[12 * rec.kf + i]);
#pragma unroll
        for (int i = 0; i < 5; i++) intr[i] = __ldg(&g.kf_intr[5 * rec.kf + i]);
        const double X0 = s.pt[3 * pt], X1 = s.pt[3 * pt + 1], X2 = s.pt[3 * pt + 2];
        p[0] = Rt[0] * X0 + Rt[1] * X1 + Rt[2] * X2 + Rt[9];
        p[1] = Rt[3] * X0 + Rt[4] * X1 + Rt[5] * X2 + Rt[10];
        p[2] = Rt[6] * X0 + Rt[7] * X1 + Rt[8] * X2 + Rt[11];
        if (valid) key = pt;
        lm_free = valid && !ptfix;
        // The branch condition is made to depend on every loaded value (x == x is false only for a NaN and cannot be folded
        // away): otherwise ptxas sinks those loads into the branch, which costs one more dependent memory round trip
        const bool loaded = (is2f == is2f) & (intr[0] == intr[0]) & (intr[1] == intr[1]) & (intr[2] == intr[2]) & (intr[3] == intr[3]) & (intr[4] == intr[4]) &
                            (pidx >= -2) & (p[0] == p[0]) & (p[1] == p[1]) & (p[2] == p[2]);
        const bool active = valid && !(fl & PPO_EF_LEVEL1_) && !(kffix && ptfix) && loaded;
        if (active) {
          const int D = point_edge_error(p, intr, rec.u, rec.v, rec.ur, err);
          const double is2 = (double)is2f;
          const double chi2 = err[0] * (is2 * err[0]) + err[1] * (is2 * err[1]) + err[2] * (is2 * err[2]);
          g.pe_chi2[e] = chi2;
          double rho0 = chi2, w = 1.0;
          if (fl & PPO_EF_ROBUST_) w = huber_w(chi2, D == 2 ? g.huber_mono : g.huber_stereo, &rho0);
          rho_sum += rho0;
          const double ws = w * is2;
          if (!ptfix) {
            const double fx = intr[0], fy = intr[1], bf = D == 3 ? (double)intr[4] : 0.0;
            const double x = p[0], y = p[1];
            // one reciprocal instead of the reference's ~15 divisions by z / z^2 (same values to 1 ulp)
            const double iz = 1.0 / p[2], iz2 = iz * iz;
            const double fxz = fx * iz, fyz = fy * iz, xz2 = x * iz2, yz2 = y * iz2;
            // weighted point Jacobian wj = (w Omega) * Jpt, Jpt = d(residual)/d(point) (types_six_dof_expmap.cpp:147-156,234-244)
            double wj[9], Jpt[9];
            {
              const double a02 = fx * xz2, a12 = fy * yz2, bz = bf * iz2;
#pragma unroll
              for (int j = 0; j < 3; j++) {
                Jpt[j] = a02 * Rt[6 + j] - fxz * Rt[j];
                Jpt[3 + j] = a12 * Rt[6 + j] - fyz * Rt[3 + j];
                Jpt[6 + j] = D == 3 ? Jpt[j] - bz * Rt[6 + j] : 0.0;
              }
            }
#pragma unroll
            for (int i = 0; i < 9; i++) wj[i] = ws * Jpt[i];
            acc[0] = wj[0] * Jpt[0] + wj[3] * Jpt[3] + wj[6] * Jpt[6];
            acc[1] = wj[0] * Jpt[1] + wj[3] * Jpt[4] + wj[6] * Jpt[7];
            acc[2] = wj[0] * Jpt[2] + wj[3] * Jpt[5] + wj[6] * Jpt[8];
            acc[3] = wj[1] * Jpt[1] + wj[4] * Jpt[4] + wj[7] * Jpt[7];
            acc[4] = wj[1] * Jpt[2] + wj[4] * Jpt[5] + wj[7] * Jpt[8];
            acc[5] = wj[2] * Jpt[2] + wj[5] * Jpt[5] + wj[8] * Jpt[8];
#pragma unroll
            for (int a = 0; a < 3; a++) acc[6 + a] = -(wj[a] * err[0] + wj[3 + a] * err[1] + wj[6 + a] * err[2]);
            if (pidx >= 0) {
              have_hpl = true;
              // pose Jacobian one column at a time (types_six_dof_expmap.cpp:158-170,246-265): Hpl row a = Jkf(:,a)^T wj
              const double sm = D == 3 ? 1.0 : 0.0;
              // two rows = six values = three 16-byte stores (lane stride 144 bytes: conflict-free per quarter warp)
#define PPO_HPL_ROW(h, j0, j1, j2)                             \
  {                                                            \
    const double q0 = (j0), q1 = (j1), q2 = sm * (j2);         \
    h[0] = q0 * wj[0] + q1 * wj[3] + q2 * wj[6];               \
    h[1] = q0 * wj[1] + q1 * wj[4] + q2 * wj[7];               \
    h[2] = q0 * wj[2] + q1 * wj[5] + q2 * wj[8];               \
  }
#define PPO_HPL_STORE2(a, h)                                                \
  {                                                                         \
    double2 *d2 = reinterpret_cast<double2 *>(st + 3 * (a));               \
    d2[0] = make_double2(h[0], h[1]);                                       \
    d2[1] = make_double2(h[2], h[3]);                                       \
    d2[2] = make_double2(h[4], h[5]);                                       \
  }
              const double k00 = x * yz2 * fx, k01 = -(1 + x * xz2) * fx, k02 = y * fxz, k03 = -fxz, k05 = xz2 * fx;
              double hh[6];
              double *h0 = hh, *h1 = hh + 3;
              PPO_HPL_ROW(h0, k00, (1 + y * yz2) * fy, k00 - bf * yz2)
              PPO_HPL_ROW(h1, k01, -x * yz2 * fy, k01 + bf * xz2)
              PPO_HPL_STORE2(0, hh)
              PPO_HPL_ROW(h0, k02, -x * fyz, k02)
              PPO_HPL_ROW(h1, k03, 0.0, k03)
              PPO_HPL_STORE2(2, hh)
              PPO_HPL_ROW(h0, 0.0, -fyz, 0.0)
              PPO_HPL_ROW(h1, k05, yz2 * fy, k05 - bf * iz2)
              PPO_HPL_STORE2(4, hh)
#undef PPO_HPL_ROW
#undef PPO_HPL_STORE2
            }
          }
        }
      }
      if (!have_hpl) {
#pragma unroll
        for (int i = 0; i < 9; i++) reinterpret_cast<double2 *>(st)[i] = make_double2(0.0, 0.0);
      }
      // segmented inclusive scan over the lanes of one point
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int ku = __shfl_up_sync(FULL, key, d);
        const bool take = lane >= d && ku == key;
#pragma unroll
        for (int i = 0; i < 9; i++) {
          const double t = __shfl_up_sync(FULL, acc[i], d);
          if (take) acc[i] += t;
        }
      }
      const int kn = __shfl_down_sync(FULL, key, 1);
      const bool seg_end = valid && (lane == 31 || kn != key);
      if (seg_end && lm_free) {
        const int L = g.n_pl + key;
#pragma unroll
        for (int i = 0; i < 6; i++) atomicAdd(&g.Hll[6 * (size_t)L + i], acc[i]);  // one add per value unless the point has > 32 edges
#pragma unroll
        for (int i = 0; i < 3; i++) atomicAdd(&g.bl[3 * (size_t)L + i], acc[6 + i]);
      }
      // Hpl blocks of this run: staged above, written to global memory by ONE TMA bulk store (cp.async.bulk shared -> global)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my generic-proxy stores to the staging buffer, before the async proxy reads it
      __syncwarp();
      if (lane == 0) {
        const unsigned bytes = (unsigned)(min(e1, base + 32) - base) * 144u;
        double *dst = g.Hpl + 18 * (size_t)(g.n_slots + base);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"((unsigned)__cvta_generic_to_shared(&stage[warp][0])), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging buffer may be overwritten (the global writes complete on their own)
      }
      __syncwarp();
    }
  }
  rho_sum = warp_sum(rho_sum);
  if (lane == 0) wsum[warp] = rho_sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < LIN_WARPS; i++) t += wsum[i];
    chi_part[blockIdx.x] = t;
  }
}

// residual-only pass (computeActiveErrors + activeRobustChi2, core/sparse_optimizer.cpp:61-114)
constexpr int RES_THREADS = 256;
__global__ void __launch_bounds__(RES_THREADS) k_point_residual(DevGraph g, DevState s, double *chi_part) {
  __shared__ double sm[RES_THREADS / 32];
  const int e = blockIdx.x * RES_THREADS + threadIdx.x;
  double rho0 = 0;
  if (e < g.n_pe) {
    const PointEdgeRec rec = g.pe_rec[e];
    const int pt = g.pe_pt[e];
    const unsigned fl = g.pe_flags[e];
    if (!(fl & PPO_EF_LEVEL1_) && !(g.kf_fixed[rec.kf] && g.pt_fixed[pt])) {
      double Rt[12], X[3], p[3], err[3];
      float intr[5];
#pragma unroll
      for (int i = 0; i < 12; i++) Rt[i] = __ldg(&s.kf_Rt[12 * rec.kf + i]);
#pragma unroll
      for (int i = 0; i < 3; i++) X[i] = s.pt[3 * pt + i];
#pragma unroll
      for (int i = 0; i < 5; i++) intr[i] = __ldg(&g.kf_intr[5 * rec.kf + i]);
      cam_point(Rt, X, p);
      const int D = point_edge_error(p, intr, rec.u, rec.v, rec.ur, err);
      const double is2 = (double)g.pe_is2[e];
      const double chi2 = err[0] * (is2 * err[0]) + err[1] * (is2 * err[1]) + err[2] * (is2 * err[2]);
      g.pe_chi2[e] = chi2;
      rho0 = chi2;
      if (fl & PPO_EF_ROBUST_) huber_w(chi2, D == 2 ? g.huber_mono : g.huber_stereo, &rho0);
    }
  }
  const double t = block_sum<RES_THREADS>(rho0, sm);
  if (threadIdx.x == 0) chi_part[blockIdx.x] = t;
}

// e->computeError() of the level-1 point edges only (the residual pass above skips them): Optimizer.cc:400-403,431-434
__global__ void k_point_error_level1(DevGraph g, DevState s) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.n_pe || !(g.pe_flags[e] & PPO_EF_LEVEL1_)) return;
  const PointEdgeRec rec = g.pe_rec[e];
  const int pt = g.pe_pt[e];
  double Rt[12], X[3], p[3], err[3];
  float intr[5];
#pragma unroll
  for (int i = 0; i < 12; i++) Rt[i] = s.kf_Rt[12 * rec.kf + i];
#pragma unroll
  for (int i = 0; i < 3; i++) X[i] = s.pt[3 * pt + i];
#pragma unroll
  for (int i = 0; i < 5; i++) intr[i] = g.kf_intr[5 * rec.kf + i];
  cam_point(Rt, X, p);
  point_edge_error(p, intr, rec.u, rec.v, rec.ur, err);
  const double is2 = (double)g.pe_is2[e];
  g.pe_chi2[e] = err[0] * (is2 * err[0]) + err[1] * (is2 * err[1]) + err[2] * (is2 * err[2]);
}

// pose side of the point edges: Hpp_jj += Jkf^T (w Omega) Jkf, bp_j += -Jkf^T (w Omega) r.
// Edges are grouped by key-frame and cut into chunks; every chunk writes a 27-value partial
// (21 upper entries + 6 gradient) that k_pose_reduce sums in a fixed order (deterministic, no atomics).
constexpr int POSE_THREADS = 256;
__global__ void __launch_bounds__(POSE_THREADS) k_pose_accumulate(DevGraph g, DevState s, double *chunk_part) {
  __shared__ double sm[POSE_THREADS / 32][27];
  const int c = blockIdx.x;
  const int kf = g.chunk_kf[c];
  const int idx = g.kf_idx[kf];
  double v[27];
#pragma unroll
  for (int i = 0; i < 27; i++) v[i] = 0;
  const int k = g.chunk_begin[c] + threadIdx.x;
  if (idx >= 0 && k < g.chunk_end[c]) {
    const int e = g.kfe_edge[k];
    const unsigned fl = g.pe_flags[e];
    if (!(fl & PPO_EF_LEVEL1_)) {
      const PointEdgeRec rec = g.pe_rec[e];
      const int pt = g.pe_pt[e];
      double Rt[12], X[3];
      float intr[5];
#pragma unroll
      for (int i = 0; i < 12; i++) Rt[i] = __ldg(&s.kf_Rt[12 * kf + i]);
#pragma unroll
      for (int i = 0; i < 3; i++) X[i] = s.pt[3 * pt + i];
#pragma unroll
      for (int i = 0; i < 5; i++) intr[i] = __ldg(&g.kf_intr[5 * kf + i]);
      PointEdgeLin L;
      point_edge_linearize(Rt, X, intr, rec, L);
      const double is2 = (double)g.pe_is2[e];
      const double chi2 = L.err[0] * (is2 * L.err[0]) + L.err[1] * (is2 * L.err[1]) + L.err[2] * (is2 * L.err[2]);
      double rho0, w = 1.0;
      if (fl & PPO_EF_ROBUST_) w = huber_w(chi2, L.D == 2 ? g.huber_mono : g.huber_stereo, &rho0);
      const double ws = w * is2;
      int q = 0;
#pragma unroll
      for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = a; b < 6; b++) v[q++] = ws * (L.Jkf[a] * L.Jkf[b] + L.Jkf[6 + a] * L.Jkf[6 + b] + L.Jkf[12 + a] * L.Jkf[12 + b]);
#pragma unroll
      for (int a = 0; a < 6; a++) v[21 + a] = -ws * (L.Jkf[a] * L.err[0] + L.Jkf[6 + a] * L.err[1] + L.Jkf[12 + a] * L.err[2]);
    }
  }
  const int lane = threadIdx.x & 31, w_ = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 27; i++) {
    const double t = warp_sum(v[i]);
    if (lane == 0) sm[w_][i] = t;
  }
  __syncthreads();
  if (threadIdx.x < 27) {
    double t = 0;
    for (int i = 0; i < POSE_THREADS / 32; i++) t += sm[i][threadIdx.x];
    chunk_part[27 * (size_t)c + threadIdx.x] = t;
  }
}
// per-key-frame sums of the chunk partials (sharded windows: summed across ranks before k_combine)
// plus the records of this rank's plane edges: pose side of ALL landmark edges of the rank
__global__ void k_chunk_reduce(DevGraph g, const int *kf_chunk_ptr, const double *chunk_part, double *kf_part) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.n_kf * 27) return;
  const int kf = t / 27, i = t % 27;
  double sum = 0;
  for (int c = kf_chunk_ptr[kf]; c < kf_chunk_ptr[kf + 1]; c++) sum += chunk_part[27 * (size_t)c + i];
  for (int q = g.kf_ple_ptr[kf]; q < g.kf_ple_ptr[kf + 1]; q++) {
    const int e = g.kf_ple_idx[q];
    if (!(g.ple_flags[e] & PPO_EF_LEVEL1_)) sum += g.ple_part[54 * (size_t)e + i];
  }
  kf_part[t] = sum;
}
// Fixed-order assembly of everything that is not a point landmark: one thread per scalar of
//   a key-frame block   (27: 21 upper + 6 gradient)  = point-edge chunk partials, then its plane edges, then its cuboid edges
//   a cuboid block      (54: 45 upper + 9 gradient)  = its camera-cuboid edges, then its point-cuboid edges
//   a plane landmark    ( 9: Hll + bl)               = its plane edges
//   a (plane, KF) slot  (18: Hpl block)              = its plane edges
// summed in list order and written with plain stores: no atomics, bit-reproducible, and no clearing of the targets beforehand.
// with_ple = 0 (sharded window): the plane-edge records of the key-frame blocks are already inside the chunk partials (k_chunk_reduce).
__global__ void k_combine(DevGraph g, const int *kf_chunk_ptr, const double *chunk_part, int with_ple) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < g.n_kf * 27) {
    const int kf = t / 27, i = t % 27;
    const int idx = g.kf_idx[kf];
    if (idx < 0) return;
    double sum = 0;
    for (int c = kf_chunk_ptr[kf]; c < kf_chunk_ptr[kf + 1]; c++) sum += chunk_part[27 * (size_t)c + i];
    for (int q = g.kf_ple_ptr[kf]; with_ple && q < g.kf_ple_ptr[kf + 1]; q++) {
      const int e = g.kf_ple_idx[q];
      if (!(g.ple_flags[e] & PPO_EF_LEVEL1_)) sum += g.ple_part[54 * (size_t)e + i];
    }
    for (int q = g.kf_cbe_ptr[kf]; q < g.kf_cbe_ptr[kf + 1]; q++) {
      const int e = g.kf_cbe_idx[q];
      if (!(g.cbe_flags[e] & PPO_EF_LEVEL1_)) sum += g.cbe_part[81 * (size_t)e + i];
    }
    if (i < 21) {
      int a = 0, rem = i;  // unpack upper-triangular index
      while (rem >= 6 - a) {
        rem -= 6 - a;
        a++;
      }
      const int b = a + rem;
      g.Hpp_kf[36 * (size_t)idx + 6 * a + b] = sum;
      g.Hpp_kf[36 * (size_t)idx + 6 * b + a] = sum;
    } else {
      g.bp[6 * idx + (i - 21)] = sum;
    }
    return;
  }
  t -= g.n_kf * 27;
  if (t < g.n_cu * 54) {
    const int cu = t / 54, i = t % 54;
    const int off = g.cu_off[cu];
    if (off < 0) return;
    double sum = 0;
    for (int q = g.cu_cbe_ptr[cu]; q < g.cu_cbe_ptr[cu + 1]; q++) {
      const int e = g.cu_cbe_idx[q];
      if (!(g.cbe_flags[e] & PPO_EF_LEVEL1_)) sum += g.cbe_part[81 * (size_t)e + 27 + i];
    }
    for (int q = g.cu_pce_ptr[cu]; q < g.cu_pce_ptr[cu + 1]; q++) {
      const int e = g.cu_pce_idx[q];
      if (!(g.pce_flags[e] & PPO_EF_LEVEL1_)) sum += g.pce_part[54 * (size_t)e + i];
    }
    if (i < 45) {
      int a = 0, rem = i;
      while (rem >= 9 - a) {
        rem -= 9 - a;
        a++;
      }
      const int b = a + rem;
      g.Hpp_cu[81 * (size_t)cu + 9 * a + b] = sum;
      g.Hpp_cu[81 * (size_t)cu + 9 * b + a] = sum;
    } else {
      g.bp[off + (i - 45)] = sum;
    }
    return;
  }
  t -= g.n_cu * 54;
  if (t < g.n_pl * 9) {
    const int pl = t / 9, i = t % 9;
    double sum = 0;
    for (int q = g.pl_ple_ptr[pl]; q < g.pl_ple_ptr[pl + 1]; q++) {
      const int e = g.pl_ple_idx[q];
      if (!(g.ple_flags[e] & PPO_EF_LEVEL1_)) sum += g.ple_part[54 * (size_t)e + 27 + i];
    }
    if (i < 6) g.Hll[6 * (size_t)pl + i] = sum;
    else g.bl[3 * (size_t)pl + (i - 6)] = sum;
    return;
  }
  t -= g.n_pl * 9;
  if (t < g.n_slots * 18) {
    const int slot = t / 18, i = t % 18;
    double sum = 0;
    for (int q = g.slot_ple_ptr[slot]; q < g.slot_ple_ptr[slot + 1]; q++) {
      const int e = g.slot_ple_idx[q];
      if (!(g.ple_flags[e] & PPO_EF_LEVEL1_) && g.kf_idx[g.ple_kf[e]] >= 0) sum += g.ple_part[54 * (size_t)e + 36 + i];
    }
    g.Hpl[18 * (size_t)slot + i] = sum;
  }
}

// ---------------------------------------------------------------------------------------------
// plane edges (numeric Jacobians, base_binary_edge.hpp:216-320): vertex 0 = plane (3), vertex 1 = KF (6)
// ---------------------------------------------------------------------------------------------
__global__ void k_plane_jac(DevGraph g, DevState s) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int col = t / g.n_ple, e = t - col * g.n_ple;  // column-major: a warp differentiates ONE variable of 32 edges (no branch divergence)
  if (col >= 9) return;
  if (g.ple_flags[e] & PPO_EF_LEVEL1_) return;
  const int kf = g.ple_kf[e], pl = g.ple_plane[e], kind = g.ple_kind[e];
  double meas[4], pc[4], ep[3], em[3];
#pragma unroll
  for (int i = 0; i < 4; i++) meas[i] = g.ple_meas[4 * e + i], pc[i] = s.pl[4 * pl + i];
  double *J = g.ple_J + 27 * (size_t)e + 3 * col;
  if (col < 3) {
    double Rt[12];
#pragma unroll
    for (int i = 0; i < 12; i++) Rt[i] = s.kf_Rt[12 * kf + i];
    double add[3] = {0, 0, 0}, pp[4];
    add[col] = NUM_DELTA;
    plane_oplus(pc, add, pp);
    plane_edge_error(kind, pp, Rt, meas, ep);
    add[col] = -NUM_DELTA;
    plane_oplus(pc, add, pp);
    plane_edge_error(kind, pp, Rt, meas, em);
  } else {
    if (g.kf_fixed[kf]) {
      J[0] = J[1] = J[2] = 0;
      return;
    }
    double pose[7], po[7], Rt[12], add[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 7; i++) pose[i] = s.kf_pose[7 * kf + i];
    add[col - 3] = NUM_DELTA;
    se3_oplus(pose, add, po);
    pose_to_Rt(po, Rt);
    plane_edge_error(kind, pc, Rt, meas, ep);
    add[col - 3] = -NUM_DELTA;
    se3_oplus(pose, add, po);
    pose_to_Rt(po, Rt);
    plane_edge_error(kind, pc, Rt, meas, em);
  }
#pragma unroll
  for (int r = 0; r < 3; r++) J[r] = NUM_SCALAR * (ep[r] - em[r]);
}
// residual (+ optional quadratic form) of the plane edges
constexpr int SMALL_THREADS = 128;
template <bool ASSEMBLE>
__global__ void __launch_bounds__(SMALL_THREADS) k_plane_edges(DevGraph g, DevState s, double *chi_part) {
  __shared__ double sm[SMALL_THREADS / 32];
  const int e = blockIdx.x * SMALL_THREADS + threadIdx.x;
  double rho0 = 0;
  if (e < g.n_ple && !(g.ple_flags[e] & PPO_EF_LEVEL1_)) {
    const int kf = g.ple_kf[e], pl = g.ple_plane[e], kind = g.ple_kind[e];
    double meas[4], pc[4], Rt[12], err[3];
#pragma unroll
    for (int i = 0; i < 4; i++) meas[i] = g.ple_meas[4 * e + i], pc[i] = s.pl[4 * pl + i];
#pragma unroll
    for (int i = 0; i < 12; i++) Rt[i] = s.kf_Rt[12 * kf + i];
    const int D = plane_edge_error(kind, pc, Rt, meas, err);
    double info[3] = {g.ple_info[3 * e], g.ple_info[3 * e + 1], D == 3 ? g.ple_info[3 * e + 2] : 0.0};
    const double chi2 = err[0] * (info[0] * err[0]) + err[1] * (info[1] * err[1]) + err[2] * (info[2] * err[2]);
    g.ple_chi2[e] = chi2;
    rho0 = chi2;
    double w = 1.0;
    if (g.ple_flags[e] & PPO_EF_ROBUST_) w = huber_w(chi2, kind == 0 ? g.huber_plane : g.huber_vp, &rho0);
    if (ASSEMBLE) {
      double J[27];  // [col][row]
#pragma unroll
      for (int i = 0; i < 27; i++) J[i] = g.ple_J[27 * (size_t)e + i];
      if (D == 2) {
#pragma unroll
        for (int c = 0; c < 9; c++) J[3 * c + 2] = 0;
      }
      const double wi[3] = {w * info[0], w * info[1], w * info[2]};
      // the edge's contribution to the normal equations (constructQuadraticForm, base_binary_edge.hpp:54-120) goes to its own
      // record; k_combine sums the records of a vertex in a fixed order (no atomics: bit-reproducible)
      double *out = g.ple_part + 54 * (size_t)e;
      const double *Jk = J + 9;  // key-frame columns
      int q = 0;
#pragma unroll
      for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = a; b < 6; b++) out[q++] = Jk[3 * a] * wi[0] * Jk[3 * b] + Jk[3 * a + 1] * wi[1] * Jk[3 * b + 1] + Jk[3 * a + 2] * wi[2] * Jk[3 * b + 2];
#pragma unroll
      for (int a = 0; a < 6; a++) out[21 + a] = -(Jk[3 * a] * wi[0] * err[0] + Jk[3 * a + 1] * wi[1] * err[1] + Jk[3 * a + 2] * wi[2] * err[2]);
      const int q6[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
#pragma unroll
      for (int i = 0; i < 6; i++) {
        const int a = q6[i][0], b = q6[i][1];
        out[27 + i] = J[3 * a] * wi[0] * J[3 * b] + J[3 * a + 1] * wi[1] * J[3 * b + 1] + J[3 * a + 2] * wi[2] * J[3 * b + 2];
      }
#pragma unroll
      for (int a = 0; a < 3; a++) out[33 + a] = -(J[3 * a] * wi[0] * err[0] + J[3 * a + 1] * wi[1] * err[1] + J[3 * a + 2] * wi[2] * err[2]);
#pragma unroll
      for (int a = 0; a < 6; a++)
#pragma unroll
        for (int c = 0; c < 3; c++) out[36 + 3 * a + c] = Jk[3 * a] * wi[0] * J[3 * c] + Jk[3 * a + 1] * wi[1] * J[3 * c + 1] + Jk[3 * a + 2] * wi[2] * J[3 * c + 2];
    }
  }
  const double t = block_sum<SMALL_THREADS>(rho0, sm);
  if (threadIdx.x == 0) chi_part[blockIdx.x] = t;
}

// ---------------------------------------------------------------------------------------------
// camera-cuboid edges (numeric): vertex 0 = KF (6), vertex 1 = cuboid (9)
// ---------------------------------------------------------------------------------------------
__global__ void k_cuboid_jac(DevGraph g, DevState s) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int col = t / g.n_cbe, e = t - col * g.n_cbe;  // column-major (see k_plane_jac)
  if (col >= 15) return;
  if (g.cbe_flags[e] & PPO_EF_LEVEL1_) return;
  const int kf = g.cbe_kf[e], cu = g.cbe_cuboid[e], kind = g.cbe_kind[e];
  const double *meas = g.cbe_meas + 16 * (size_t)e;
  float intr[5];
#pragma unroll
  for (int i = 0; i < 5; i++) intr[i] = g.kf_intr[5 * kf + i];
  double c[10], ep[16], em[16];
#pragma unroll
  for (int i = 0; i < 10; i++) c[i] = s.cu[10 * cu + i];
  double *J = g.cbe_J + 240 * (size_t)e + 16 * col;
  const int D = kind == 0 ? 4 : (kind == 2 ? 9 : 16);
  // kind 2 (EdgeSE3Cuboid) works on the pose itself, the projection edges on its [R|t] form
  if (col < 6) {
    if (g.kf_fixed[kf]) {
      for (int r = 0; r < 16; r++) J[r] = 0;
      return;
    }
    double pose[7], po[7], Rt[12], add[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 7; i++) pose[i] = s.kf_pose[7 * kf + i];
    add[col] = NUM_DELTA;
    se3_oplus(pose, add, po);
    if (kind == 2) {
      cuboid_se3_error(po, c, meas, ep);
    } else {
      pose_to_Rt(po, Rt);
      cuboid_cam_error(kind, Rt, c, intr, meas, ep);
    }
    add[col] = -NUM_DELTA;
    se3_oplus(pose, add, po);
    if (kind == 2) {
      cuboid_se3_error(po, c, meas, em);
    } else {
      pose_to_Rt(po, Rt);
      cuboid_cam_error(kind, Rt, c, intr, meas, em);
    }
  } else {
    double Rt[12], pose[7], add[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, co[10];
#pragma unroll
    for (int i = 0; i < 12; i++) Rt[i] = s.kf_Rt[12 * kf + i];
#pragma unroll
    for (int i = 0; i < 7; i++) pose[i] = s.kf_pose[7 * kf + i];
    const unsigned cf = g.cu_flags[cu];
    add[col - 6] = NUM_DELTA;
    cuboid_oplus(c, cf, add, co);
    if (kind == 2) cuboid_se3_error(pose, co, meas, ep);
    else cuboid_cam_error(kind, Rt, co, intr, meas, ep);
    add[col - 6] = -NUM_DELTA;
    cuboid_oplus(c, cf, add, co);
    if (kind == 2) cuboid_se3_error(pose, co, meas, em);
    else cuboid_cam_error(kind, Rt, co, intr, meas, em);
  }
  for (int r = 0; r < 16; r++) J[r] = r < D ? NUM_SCALAR * (ep[r] - em[r]) : 0.0;
}
// residual of the camera-cuboid edges; with STORE also the error vector and the robust weight for k_cuboid_assemble
template <bool STORE>
__global__ void __launch_bounds__(SMALL_THREADS) k_cuboid_edges(DevGraph g, DevState s, double *chi_part) {
  __shared__ double sm[SMALL_THREADS / 32];
  const int e = blockIdx.x * SMALL_THREADS + threadIdx.x;
  double rho0 = 0;
  if (e < g.n_cbe && !(g.cbe_flags[e] & PPO_EF_LEVEL1_)) {
    const int kf = g.cbe_kf[e], cu = g.cbe_cuboid[e], kind = g.cbe_kind[e];
    float intr[5];
#pragma unroll
    for (int i = 0; i < 5; i++) intr[i] = g.kf_intr[5 * kf + i];
    double c[10], Rt[12], err[16];
#pragma unroll
    for (int i = 0; i < 10; i++) c[i] = s.cu[10 * cu + i];
#pragma unroll
    for (int i = 0; i < 12; i++) Rt[i] = s.kf_Rt[12 * kf + i];
    int D = 9;
    if (kind == 2) {
      double pose[7];
#pragma unroll
      for (int i = 0; i < 7; i++) pose[i] = s.kf_pose[7 * kf + i];
      cuboid_se3_error(pose, c, g.cbe_meas + 16 * (size_t)e, err);
    } else {
      D = cuboid_cam_error(kind, Rt, c, intr, g.cbe_meas + 16 * (size_t)e, err);
    }
    const double info = g.cbe_info[e];
    double chi2 = 0, nrm = 0;
    for (int r = 0; r < D; r++) chi2 += err[r] * (info * err[r]), nrm += err[r] * err[r];
    g.cbe_chi2[e] = chi2;
    g.cbe_norm[e] = sqrt(nrm);
    rho0 = chi2;
    double w = 1.0;
    if (g.cbe_flags[e] & PPO_EF_ROBUST_) w = huber_w(chi2, kind == 0 ? g.huber_bbox : (kind == 2 ? g.huber_se3 : g.huber_corner), &rho0);
    if (STORE) {
      for (int r = 0; r < 16; r++) g.cbe_err[16 * (size_t)e + r] = r < D ? err[r] : 0.0;
      g.cbe_w[e] = w * info;
    }
  }
  const double t = block_sum<SMALL_THREADS>(rho0, sm);
  if (threadIdx.x == 0) chi_part[blockIdx.x] = t;
}
// quadratic form of the camera-cuboid edges: one thread per (edge, row a of the 15 x 15 block); the key-frame block (21 upper
// entries + 6 gradient) and the cuboid block (45 + 9) go to the edge's own record (summed per vertex by k_combine), the 6 x 9
// off-diagonal block to Hpc
__device__ __forceinline__ int upper_index(int n, int a, int b) { return a * n - a * (a - 1) / 2 + (b - a); }  // b >= a
constexpr int CBA_EDGES = 8;  // edges per CTA of k_cuboid_assemble (15 threads each)
__global__ void __launch_bounds__(CBA_EDGES * 15) k_cuboid_assemble(DevGraph g) {
  // the Jacobians (15 columns x 16 rows) and error vectors of the CTA's edges are contiguous in global memory: staged in shared memory
  // with coalesced loads, row stride 17 so that the 15 threads of an edge (one column each) hit different banks
  __shared__ double sJ[CBA_EDGES][15][17];
  __shared__ double sE[CBA_EDGES][16];
  const int e0 = blockIdx.x * CBA_EDGES;
  const int ne = min(CBA_EDGES, g.n_cbe - e0);
  for (int i = threadIdx.x; i < ne * 240; i += CBA_EDGES * 15) {
    const int le = i / 240, r = i % 240;
    sJ[le][r >> 4][r & 15] = g.cbe_J[240 * (size_t)e0 + i];
  }
  for (int i = threadIdx.x; i < ne * 16; i += CBA_EDGES * 15) sE[i >> 4][i & 15] = g.cbe_err[16 * (size_t)e0 + i];
  __syncthreads();
  const int le = threadIdx.x / 15, a = threadIdx.x % 15;
  const int e = e0 + le;
  if (le >= ne || (g.cbe_flags[e] & PPO_EF_LEVEL1_)) return;
  const double wi = g.cbe_w[e];
  double *out = g.cbe_part + 81 * (size_t)e;
  double ja[16];
#pragma unroll
  for (int r = 0; r < 16; r++) ja[r] = sJ[le][a][r];
  double ga = 0;
#pragma unroll
  for (int r = 0; r < 16; r++) ga += ja[r] * sE[le][r];
  ga *= -wi;
  if (a < 6) out[21 + a] = ga;
  else out[27 + 45 + (a - 6)] = ga;
  for (int b = a; b < 15; b++) {
    double hv = 0;
#pragma unroll
    for (int r = 0; r < 16; r++) hv += ja[r] * sJ[le][b][r];
    hv *= wi;
    if (a < 6 && b < 6) out[upper_index(6, a, b)] = hv;
    else if (a < 6) g.Hpc[54 * (size_t)e + 9 * a + (b - 6)] = hv;
    else out[27 + upper_index(9, a - 6, b - 6)] = hv;
  }
}

// ---------------------------------------------------------------------------------------------
// point-cuboid unary edges (numeric, base_unary_edge.hpp:81-123)
// ---------------------------------------------------------------------------------------------
__global__ void k_ptcu_jac(DevGraph g, DevState s) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = t / 9, col = t % 9;
  if (e >= g.n_pce) return;
  if (g.pce_flags[e] & PPO_EF_LEVEL1_) return;
  const int cu = g.pce_cuboid[e];
  double c[10], co[10], ep[3], em[3], add[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 10; i++) c[i] = s.cu[10 * cu + i];
  const double *pts = g.pce_pts + 3 * (size_t)g.pce_rowptr[e];
  const int n = g.pce_rowptr[e + 1] - g.pce_rowptr[e];
  const unsigned cf = g.cu_flags[cu];
  add[col] = NUM_DELTA;
  cuboid_oplus(c, cf, add, co);
  point_cuboid_error(co, pts, n, g.ptcu_ratio, g.ptcu_prior, ep);
  add[col] = -NUM_DELTA;
  cuboid_oplus(c, cf, add, co);
  point_cuboid_error(co, pts, n, g.ptcu_ratio, g.ptcu_prior, em);
  double *J = g.pce_J + 27 * (size_t)e + 3 * col;
#pragma unroll
  for (int r = 0; r < 3; r++) J[r] = NUM_SCALAR * (ep[r] - em[r]);
}
template <bool ASSEMBLE>
__global__ void __launch_bounds__(SMALL_THREADS) k_ptcu_edges(DevGraph g, DevState s, double *chi_part) {
  __shared__ double sm[SMALL_THREADS / 32];
  const int e = blockIdx.x * SMALL_THREADS + threadIdx.x;
  double rho0 = 0;
  if (e < g.n_pce && !(g.pce_flags[e] & PPO_EF_LEVEL1_)) {
    const int cu = g.pce_cuboid[e];
    double c[10], err[3];
#pragma unroll
    for (int i = 0; i < 10; i++) c[i] = s.cu[10 * cu + i];
    point_cuboid_error(c, g.pce_pts + 3 * (size_t)g.pce_rowptr[e], g.pce_rowptr[e + 1] - g.pce_rowptr[e], g.ptcu_ratio, g.ptcu_prior, err);
    const double chi2 = err[0] * err[0] + err[1] * err[1] + err[2] * err[2];  // information = I (Optimizer.cc:2644-2646)
    g.pce_chi2[e] = chi2;
    rho0 = chi2;
    double w = 1.0;
    if (g.pce_flags[e] & PPO_EF_ROBUST_) w = huber_w(chi2, 1.0, &rho0);
    if (ASSEMBLE) {  // the edge's cuboid block (45 upper entries + 9 gradient) goes to its own record, summed by k_combine
      const double *J = g.pce_J + 27 * (size_t)e;
      double *out = g.pce_part + 54 * (size_t)e;
      int q = 0;
      for (int a = 0; a < 9; a++) {
        out[45 + a] = -w * (J[3 * a] * err[0] + J[3 * a + 1] * err[1] + J[3 * a + 2] * err[2]);
        for (int b = a; b < 9; b++) out[q++] = w * (J[3 * a] * J[3 * b] + J[3 * a + 1] * J[3 * b + 1] + J[3 * a + 2] * J[3 * b + 2]);
      }
    }
  }
  const double t = block_sum<SMALL_THREADS>(rho0, sm);
  if (threadIdx.x == 0) chi_part[blockIdx.x] = t;
}

// ---------------------------------------------------------------------------------------------
// Schur complement (BlockSolver::solve, core/block_solver.hpp:376-431) with lambda on the diagonal of
// Hll as setLambda adds it (:582-587).  S is zero on entry; k_compose adds Hpp afterwards.
// ---------------------------------------------------------------------------------------------
PPO_D void inv_sym3(const double h[6], double lam, double d[6]) {
  const double a = h[0] + lam, b = h[1], c = h[2], e = h[3] + lam, f = h[4], i = h[5] + lam;
  const double c00 = e * i - f * f, c01 = f * c - b * i, c02 = b * f - e * c;
  const double det = a * c00 + b * c01 + c * c02;
  const double id = 1.0 / det;
  d[0] = c00 * id;
  d[1] = (c * f - b * i) * id;
  d[2] = (b * f - c * e) * id;
  d[3] = (a * i - c * c) * id;
  d[4] = (c * b - a * f) * id;
  d[5] = (a * e - b * b) * id;
}
PPO_D bool landmark_active(const DevGraph &g, int L) {
  return L < g.n_pl ? g.pl_act[L] != 0 : (g.pt_act[L - g.n_pl] != 0 && !g.pt_fixed[L - g.n_pl]);
}
// ---------------------------------------------------------------------------------------------
// Pair-major Schur complement (no per-scalar atomics).  At set_graph every landmark emits one contribution per
// pair of its (free key-frame) blocks, keyed by the key-frame pair; the list is radix-sorted by key once.  Per
// damped trial  k_schur_bd  forms Dinv = C C^T and the 6x3 products Y = W C of every block, and  k_schur_pairs  streams the
// sorted list: a warp owns 64 consecutive contributions, lane (r,c) accumulates entry (r,c) of the current
// key-frame pair in a register and flushes 36 REDs only when the key changes.
// ---------------------------------------------------------------------------------------------
// set_graph: packed point-edge records / edge ids (payload of the by-key-frame sort) and the edge -> point map
__global__ void k_pack_point_edges(int n_pe, const int *__restrict__ pe_kf, const float *__restrict__ obs, PointEdgeRec *rec, int *iota) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_pe) return;
  PointEdgeRec r;
  r.kf = pe_kf[e];
  r.u = obs[3 * (size_t)e];
  r.v = obs[3 * (size_t)e + 1];
  r.ur = obs[3 * (size_t)e + 2];
  rec[e] = r;
  iota[e] = e;
}
__global__ void k_fill_pe_pt(int n_pt, const int *__restrict__ rowptr, int *pe_pt) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pt) return;
  for (int e = rowptr[p]; e < rowptr[p + 1]; e++) pe_pt[e] = p;
}
// contributions of landmark L: pairs (i1 <= i2) of its free-key-frame blocks; also the one-observation-per-key-frame check
__global__ void k_pair_count(DevGraph g, int *cnt, int *dup) {
  const int L = blockIdx.x * blockDim.x + threadIdx.x;
  if (L >= g.n_lm) return;
  const int b0 = g.lm_rowptr[L], b1 = g.lm_rowptr[L + 1];
  int kfree = 0;
  bool twice = false;
  for (int i = b0; i < b1; i++) {
    const int s = i < g.n_slots ? g.slot_kf[i] : g.pe_rec[i - g.n_slots].kf;
    kfree += !g.kf_fixed[s];
    for (int j = b0; j < i; j++) twice |= (j < g.n_slots ? g.slot_kf[j] : g.pe_rec[j - g.n_slots].kf) == s;
  }
  cnt[L] = kfree * (kfree + 1) / 2;
  if (twice) *dup = 1;
}
__global__ void k_gen_pairs(DevGraph g, const int *lm_pair_off, unsigned *keys, unsigned long long *vals) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= g.n_lm) return;
  const int b0 = g.lm_rowptr[warp], b1 = g.lm_rowptr[warp + 1];
  int out = lm_pair_off[warp];
  // free-key-frame entries of this landmark, in entry order; pairs (i1 <= i2) are numbered row by row
  for (int i1 = b0; i1 < b1; i1++) {
    const int s1 = i1 < g.n_slots ? g.slot_kf[i1] : g.pe_rec[i1 - g.n_slots].kf;
    if (g.kf_fixed[s1]) continue;  // warp-uniform
    int cnt = 0;
    for (int base = i1; base < b1; base += 32) {
      const int i2 = base + lane;
      int s2 = -1;
      bool ok = false;
      if (i2 < b1) {
        s2 = i2 < g.n_slots ? g.slot_kf[i2] : g.pe_rec[i2 - g.n_slots].kf;
        ok = !g.kf_fixed[s2];
      }
      const unsigned m = __ballot_sync(FULL, ok);
      if (ok) {
        const int pos = out + cnt + __popc(m & ((1u << lane) - 1));
        const bool sw = s1 > s2;
        const unsigned a = sw ? s2 : s1, b = sw ? s1 : s2;
        keys[pos] = a * (unsigned)g.n_kf + b;
        vals[pos] = ((unsigned long long)(unsigned)(sw ? i2 : i1) << 32) | (unsigned)(sw ? i1 : i2);
      }
      cnt += __popc(m);
    }
    out += cnt;
  }
}
constexpr int BD_WARPS = 8;
__global__ void __launch_bounds__(BD_WARPS * 32) k_schur_bd(DevGraph g, const LmDev *lm, int n_p, int ld, int planes_write_S, int n_first) {
  const double lambda = lm->lambda;
  const int lane = threadIdx.x & 31;
  const int L = blockIdx.x * BD_WARPS + (threadIdx.x >> 5);
  if (L >= n_first) return;  // one warp per landmark: used for the planes (tens of blocks each)
  if (!landmark_active(g, L)) {  // inactive landmark: its blocks must not carry products of an earlier trial
    const int a0 = g.lm_rowptr[L] * 18, a1 = g.lm_rowptr[L + 1] * 18;
    for (int it = a0 + lane; it < a1; it += 32) g.BD[it] = 0.0;
    return;
  }
  double h[6], D[6], bl[3];
#pragma unroll
  for (int i = 0; i < 6; i++) h[i] = g.Hll[6 * (size_t)L + i];
#pragma unroll
  for (int i = 0; i < 3; i++) bl[i] = g.bl[3 * (size_t)L + i];
  inv_sym3(h, lambda, D);
  if (lane < 6) g.Dinv[6 * (size_t)L + lane] = D[lane];
  const bool to_S = L >= g.n_pl || planes_write_S;  // sharded window: planes are reduced by rank 0 only
  // Dinv = (Hll + lambda)^-1 = C C^T with C = chol(Hll + lambda)^-T: the Schur product W Dinv W'^T of two blocks of this
  // landmark is then Y Y'^T with Y = W C, so k_schur_pairs gathers from ONE array (half the working set: L2 resident).
  const double l00 = sqrt(h[0] + lambda), i00 = 1.0 / l00;
  const double l10 = h[1] * i00, l20 = h[2] * i00;
  const double l11 = sqrt(fmax(h[3] + lambda - l10 * l10, 1e-300)), i11 = 1.0 / l11;
  const double l21 = (h[4] - l20 * l10) * i11;
  const double i22 = rsqrt(fmax(h[5] + lambda - l20 * l20 - l21 * l21, 1e-300));
  const int b0 = g.lm_rowptr[L], b1 = g.lm_rowptr[L + 1];
  const double *W = g.Hpl + 18 * (size_t)b0;
  double *Y = g.BD + 18 * (size_t)b0;
  const int n = (b1 - b0) * 6;
  for (int it = lane; it < n; it += 32) {  // one row (e, r) of a 6 x 3 block per lane
    const double w0 = W[3 * it], w1 = W[3 * it + 1], w2 = W[3 * it + 2];
    const double y0 = w0 * i00, y1 = (w1 - y0 * l10) * i11, y2 = (w2 - y0 * l20 - y1 * l21) * i22;  // y C^-1 = w
    Y[3 * it] = to_S ? y0 : 0.0;
    Y[3 * it + 1] = to_S ? y1 : 0.0;
    Y[3 * it + 2] = to_S ? y2 : 0.0;
  }
  // z = C^T bl per block: the reduced gradient term  W Dinv bl = Y z  (core/block_solver.hpp:403-405) is accumulated by
  // k_schur_pairs together with the diagonal Schur blocks (no atomics here)
  const double z0 = bl[0] * i00, z1 = (bl[1] - l10 * z0) * i11, z2 = (bl[2] - l20 * z0 - l21 * z1) * i22;
  double *Z = g.Zent + 3 * (size_t)b0;
  for (int it = lane; it < 3 * (b1 - b0); it += 32) {
    const int c = it % 3;
    Z[it] = c == 0 ? z0 : (c == 1 ? z1 : z2);
  }
}
// Point landmarks: one lane per 6x3 block (the work units of k_point_linearize: every lane moves 144 contiguous bytes in and
// out with 16-byte accesses); the 3x3 factorisation of the landmark is recomputed by each of its lanes.
__global__ void __launch_bounds__(32) k_schur_bd_points(DevGraph g, const LmDev *lm) {
  const double lambda = lm->lambda;
  const int lane = threadIdx.x;
  const int e0 = g.unit_e0[blockIdx.x], e1 = g.unit_e0[blockIdx.x + 1];
  for (int e = e0 + lane; e < e1; e += 32) {
    const int pt = g.pe_pt[e];
    const int L = g.n_pl + pt;
    const size_t ent = (size_t)g.n_slots + e;
    double2 *Y = reinterpret_cast<double2 *>(g.BD + 18 * ent);
    double *Z = g.Zent + 3 * ent;
    if (!landmark_active(g, L)) {  // its blocks must not carry products of an earlier trial
#pragma unroll
      for (int i = 0; i < 9; i++) Y[i] = make_double2(0.0, 0.0);
      continue;
    }
    double h[6], bl[3];
#pragma unroll
    for (int i = 0; i < 6; i++) h[i] = g.Hll[6 * (size_t)L + i];
#pragma unroll
    for (int i = 0; i < 3; i++) bl[i] = g.bl[3 * (size_t)L + i];
    if (e == g.pt_rowptr[pt]) {  // first block of the point: publish Dinv for the back-substitution
      double D[6];
      inv_sym3(h, lambda, D);
#pragma unroll
      for (int i = 0; i < 6; i++) g.Dinv[6 * (size_t)L + i] = D[i];
    }
    const double l00 = sqrt(h[0] + lambda), i00 = 1.0 / l00;
    const double l10 = h[1] * i00, l20 = h[2] * i00;
    const double l11 = sqrt(fmax(h[3] + lambda - l10 * l10, 1e-300)), i11 = 1.0 / l11;
    const double l21 = (h[4] - l20 * l10) * i11;
    const double i22 = rsqrt(fmax(h[5] + lambda - l20 * l20 - l21 * l21, 1e-300));
    const double2 *W = reinterpret_cast<const double2 *>(g.Hpl + 18 * ent);
    double w[18];
#pragma unroll
    for (int i = 0; i < 9; i++) {
      const double2 t = W[i];
      w[2 * i] = t.x, w[2 * i + 1] = t.y;
    }
#pragma unroll
    for (int r = 0; r < 6; r++) {  // y C^-1 = w, row by row
      const double y0 = w[3 * r] * i00, y1 = (w[3 * r + 1] - y0 * l10) * i11, y2 = (w[3 * r + 2] - y0 * l20 - y1 * l21) * i22;
      w[3 * r] = y0, w[3 * r + 1] = y1, w[3 * r + 2] = y2;
    }
#pragma unroll
    for (int i = 0; i < 9; i++) Y[i] = make_double2(w[2 * i], w[2 * i + 1]);
    const double z0 = bl[0] * i00, z1 = (bl[1] - l10 * z0) * i11, z2 = (bl[2] - l20 * z0 - l21 * z1) * i22;
    Z[0] = z0, Z[1] = z1, Z[2] = z2;
  }
}
// Every contribution is a 6x3 by 3x6 product Y1 Y2^T: exactly one FP64 tensor-core instruction (mma.m8n8k4, rows/columns
// 6..7 and k = 3 zero padded), accumulated in the instruction's own C registers while the key-frame pair stays the same.
// Column 6 of the B operand carries z of the landmark for the self pairs (Y1 == Y2), so C(:, 6) is the reduced-gradient
// term Y z of the key-frame.  Per contribution a warp issues two predicated 8-byte loads per lane and one DMMA.
constexpr int PAIR_CHUNK = 128;
constexpr int PAIR_WARPS = 8;
PPO_D void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// Writes the accumulated 6 x 6 block (+ reduced-gradient column) of key-frame pair `key` into S: S is zero on entry and every block
// has exactly one writer, so these are plain stores (bit-reproducible); k_compose adds Hpp afterwards.
// upper (6 p1 + row, 6 p2 + col) = lower element (6 p2 + col, 6 p1 + row) of the tiled storage; ld = Tm
PPO_D void schur_store(const DevGraph &g, unsigned key, int lane, double acc0, double acc1, int ld, int grow) {
  const int row = lane >> 2, kk = lane & 3;
  const unsigned ka = key / (unsigned)g.n_kf, kb = key - ka * (unsigned)g.n_kf;
  const int p1 = g.kf_idx[ka], p2 = g.kf_idx[kb];
  if (p1 >= 0 && p2 >= 0 && row < 6) {
    const int c = 6 * p1 + row;
    if (kk < 3) {  // (diagonal blocks: the mirrored half is not stored)
      const int r = 6 * p2 + 2 * kk;
      if (r >= c) g.S[dense_elem_index(ld, r, c)] = -acc0;
      if (r + 1 >= c) g.S[dense_elem_index(ld, r + 1, c)] = -acc1;
    } else if (ka == kb) {
      g.S[dense_elem_index(ld, grow, c)] = -acc0;  // bschur -= W Dinv bl
    }
  }
}
// A warp streams PAIR_CHUNK consecutive records of the sorted list.  A key-frame pair whose records lie inside one chunk is
// finished by that warp.  A pair that crosses a chunk boundary leaves its partial C fragment in the chunk's boundary record
// (slot 0: the chunk's first pair continues from the previous chunk, slot 1: its last pair continues into the next one) and
// k_schur_pairs_fix adds the partials of such a pair in chunk order: no atomics, the result does not depend on scheduling.
// bnd_flag[chunk]: bit 0 slot 0 valid, bit 1 slot 1 valid, bit 2 the whole chunk is one pair that continues on both sides.
template <int U, int CHAINS, int MINB>
__global__ void __launch_bounds__(PAIR_WARPS * 32, MINB) k_schur_pairs(DevGraph g, const unsigned *__restrict__ keys, const unsigned long long *__restrict__ vals,
                                                                       int n_pairs, int ld, int grow, double *bnd, unsigned *bnd_key, int *bnd_flag) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * PAIR_WARPS + (threadIdx.x >> 5);
  const long long c0 = w * PAIR_CHUNK;
  if (c0 >= n_pairs) return;  // warp-uniform
  const int cnt = (int)min((long long)PAIR_CHUNK, (long long)n_pairs - c0);
  // fragment layout: A element (row = lane/4, k = lane%4), B element (k = lane%4, col = lane/4), C elements (row = lane/4, col = 2 (lane%4) + {0,1})
  const int row = lane >> 2, kk = lane & 3;
  const bool in_blk = row < 6 && kk < 3, z_lane = row == 6 && kk < 3;
  const int idx = 3 * row + kk;
  const unsigned PAD = 0xffffffffu;
  const unsigned first_key = keys[c0], last_key = keys[c0 + cnt - 1];
  const bool cont_prev = c0 > 0 && first_key != PAD && keys[c0 - 1] == first_key;
  const bool cont_next = c0 + cnt < n_pairs && last_key != PAD && keys[c0 + cnt] == last_key;
  int flags = 0;
  unsigned cur = PAD;
  double acc0 = 0.0, acc1 = 0.0;
  // CHAINS > 1: the records of a group go round-robin to CHAINS accumulator fragments (independent DMMA chains); they are added
  // in a fixed order when the key-frame pair ends
  double ex0[CHAINS > 1 ? CHAINS - 1 : 1], ex1[CHAINS > 1 ? CHAINS - 1 : 1];
#pragma unroll
  for (int q = 0; q < CHAINS - 1; q++) ex0[q] = ex1[q] = 0.0;
  auto flush = [&]() {
#pragma unroll
    for (int q = 0; q < CHAINS - 1; q++) {
      acc0 += ex0[q], acc1 += ex1[q];
      ex0[q] = ex1[q] = 0.0;
    }
    if (cur != PAD) {
      const bool from_prev = cont_prev && cur == first_key, to_next = cont_next && cur == last_key;
      if (!from_prev && !to_next) {
        schur_store(g, cur, lane, acc0, acc1, ld, grow);
      } else {
        const int slot = from_prev ? 0 : 1;
        double *dst = bnd + ((size_t)(2 * w + slot) * 32 + lane) * 2;
        dst[0] = acc0, dst[1] = acc1;
        if (lane == 0) bnd_key[2 * w + slot] = cur;
        flags |= from_prev ? (to_next ? 5 : 1) : 2;
      }
    }
    acc0 = acc1 = 0.0;
  };
  // Per-lane operand addressing, fixed for the whole chunk: the lanes of the 6 x 3 block read Y (18 doubles per entry) for both operands,
  // the three lanes of row 6 read z (3 doubles per entry) as their B operand -- one load per operand and record, no per-record selects.
  // Column 6 of C (the reduced-gradient term) is only ever stored for the self pairs kf_a == kf_b, which are exactly the records with
  // e1 == e2 (a landmark has at most one block per key-frame: set_graph rejects duplicates), so the z lanes need no e1 == e2 test.
  const double *pa = g.BD + idx;
  const double *pb = in_blk ? g.BD + idx : g.Zent + kk;
  const unsigned sb = in_blk ? 18u : 3u;
  const bool lb = in_blk || z_lane;
  // U contributions have their operand loads in flight together
  for (int base = 0; base < cnt; base += 32) {
    const int c = base + lane;
    const unsigned kl = c < cnt ? keys[c0 + c] : PAD;  // (the list ends with 0xffffffff padding)
    const unsigned long long vl = c < cnt ? vals[c0 + c] : 0ull;
    const unsigned v1 = (unsigned)(vl >> 32), v2 = (unsigned)(vl & 0xffffffffu);
    for (int cb = 0; cb < 32; cb += U) {
      if (base + cb >= cnt) break;  // warp-uniform
      unsigned key[U];
      double av[U], bv[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        key[u] = __shfl_sync(FULL, kl, cb + u);
        const unsigned e1 = __shfl_sync(FULL, v1, cb + u), e2 = __shfl_sync(FULL, v2, cb + u);
        const bool ok = key[u] != PAD;
        av[u] = (ok && in_blk) ? pa[18 * (size_t)e1] : 0.0;
        bv[u] = (ok && lb) ? pb[(size_t)sb * e2] : 0.0;
      }
      if (key[0] == cur && key[U - 1] == cur) {  // the list is sorted: the whole group continues the current key-frame pair (~90 records per pair)
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (CHAINS == 1 || u % CHAINS == 0) dmma884(acc0, acc1, av[u], bv[u]);
          else dmma884(ex0[u % CHAINS - 1], ex1[u % CHAINS - 1], av[u], bv[u]);
        }
        continue;
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (key[u] == PAD) continue;  // warp-uniform
        if (key[u] != cur) {
          flush();
          cur = key[u];
        }
        dmma884(acc0, acc1, av[u], bv[u]);
      }
    }
  }
  flush();
  if (lane == 0) bnd_flag[w] = flags;
}
// One warp per chunk whose LAST pair continues into the next chunk (it is the head of that pair's run of boundary records)
__global__ void __launch_bounds__(PAIR_WARPS * 32) k_schur_pairs_fix(DevGraph g, int n_chunks, int ld, int grow, const double *bnd, const unsigned *bnd_key,
                                                                     const int *bnd_flag) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * PAIR_WARPS + (threadIdx.x >> 5);
  if (w >= n_chunks || !(bnd_flag[w] & 2)) return;
  const unsigned key = bnd_key[2 * w + 1];
  const double *p = bnd + ((size_t)(2 * w + 1) * 32 + lane) * 2;
  double a0 = p[0], a1 = p[1];
  for (int u = w + 1; u < n_chunks; u++) {
    const int f = bnd_flag[u];
    if (!(f & 1) || bnd_key[2 * u] != key) break;  // (cannot happen: the head saw the pair continue)
    const double *q = bnd + ((size_t)(2 * u) * 32 + lane) * 2;
    a0 += q[0], a1 += q[1];
    if (!(f & 4)) break;  // the pair ends in this chunk
  }
  schur_store(g, key, lane, a0, a1, ld, grow);
}

// S += Hpp (+ lambda on the diagonal), rhs column += bp.  One thread per scalar of each block.
__global__ void k_compose(DevGraph g, const LmDev *lm, int n_p, int ld, int grow) {
  const double lambda = lm->lambda;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nkf = g.n_kf * 36, ncu = g.n_cu * 81, nhpc = g.n_cbe * 54;
  if (t < nkf) {
    const int kf = t / 36, a = (t % 36) / 6, b = t % 6;
    const int idx = g.kf_idx[kf];
    if (idx >= 0 && a <= b) {
      double v = g.Hpp_kf[36 * (size_t)idx + 6 * a + b];
      if (a == b) v += lambda;
      atomicAdd(&g.S[dense_elem_index(ld, 6 * idx + b, 6 * idx + a)], v);
    }
  } else if (t < nkf + ncu) {
    const int u = t - nkf, cu = u / 81, a = (u % 81) / 9, b = u % 9;
    const int off = g.cu_off[cu];
    if (off >= 0 && a <= b) {
      double v = g.Hpp_cu[81 * (size_t)cu + 9 * a + b];
      if (a == b) v += lambda;
      atomicAdd(&g.S[dense_elem_index(ld, off + b, off + a)], v);
    }
  } else if (t < nkf + ncu + nhpc) {
    const int u = t - nkf - ncu, e = u / 54, a = (u % 54) / 9, b = u % 9;
    if (!(g.cbe_flags[e] & PPO_EF_LEVEL1_)) {
      const int idx = g.kf_idx[g.cbe_kf[e]], off = g.cu_off[g.cbe_cuboid[e]];
      if (idx >= 0 && off >= 0) atomicAdd(&g.S[dense_elem_index(ld, off + b, 6 * idx + a)], g.Hpc[54 * (size_t)e + 9 * a + b]);
    }
  } else if (t < nkf + ncu + nhpc + n_p) {
    const int j = t - nkf - ncu - nhpc;
    atomicAdd(&g.S[dense_elem_index(ld, grow, j)], g.bp[j]);
  }
}

// ---------------------------------------------------------------------------------------------
// back-substitution  xl = Dinv (bl - Hpl^T xp)   (core/block_solver.hpp:459-481) + landmark part of
// computeScale (levenberg.cpp:182-189).  One warp per landmark, lanes stride the contiguous blocks.
// ---------------------------------------------------------------------------------------------
constexpr int BS_WARPS = 8;
__global__ void __launch_bounds__(BS_WARPS * 32) k_backsub(DevGraph g, const LmDev *lm, double *scale_part, int planes_in_scale, int n_first) {
  __shared__ double wsum[BS_WARPS];
  const double lambda = lm->lambda;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = blockIdx.x * BS_WARPS + warp;
  double sc = 0;
  const bool in_range = L < n_first;  // one warp per landmark: used for the planes (tens of blocks each)
  const bool act = in_range && landmark_active(g, L);
  double c[3] = {0, 0, 0};
  if (act) {
    const int b0 = g.lm_rowptr[L], b1 = g.lm_rowptr[L + 1];
    const int n = (b1 - b0) * 18;
    const double *W = g.Hpl + 18 * (size_t)b0;
    for (int i = lane; i < n; i += 32) {
      const int ent = b0 + i / 18, a = (i % 18) / 3, col = i % 3;
      const int p = g.ent_pidx[ent];
      if (p >= 0) {
        const double t = W[i] * g.xp[6 * p + a];
        if (col == 0) c[0] += t;
        else if (col == 1) c[1] += t;
        else c[2] += t;
      }
    }
  }
  // reductions in warp-uniform control flow (no convergence barriers around the shuffles)
#pragma unroll
  for (int k = 0; k < 3; k++) c[k] = warp_sum(c[k]);
  if (act && lane == 0) {
    double D[6], bl[3], cl[3], x[3];
#pragma unroll
    for (int i = 0; i < 6; i++) D[i] = g.Dinv[6 * (size_t)L + i];
#pragma unroll
    for (int i = 0; i < 3; i++) bl[i] = g.bl[3 * (size_t)L + i], cl[i] = bl[i] - c[i];
    x[0] = D[0] * cl[0] + D[1] * cl[1] + D[2] * cl[2];
    x[1] = D[1] * cl[0] + D[3] * cl[1] + D[4] * cl[2];
    x[2] = D[2] * cl[0] + D[4] * cl[1] + D[5] * cl[2];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      g.xl[3 * (size_t)L + i] = x[i];
      if (L >= g.n_pl || planes_in_scale) sc += x[i] * (lambda * x[i] + bl[i]);
    }
  } else if (in_range && !act && lane < 3) {
    g.xl[3 * (size_t)L + lane] = 0.0;
  }
  if (lane == 0) wsum[warp] = sc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < BS_WARPS; i++) t += wsum[i];
    scale_part[blockIdx.x] = t;
  }
}
// Point landmarks: same work units as k_point_linearize (a run of consecutive points with <= 32 blocks per warp), one
// lane per 6x3 block: the lane reads its 144 contiguous bytes and the 6 pose increments, forms Hpl^T x_p, and a
// segmented warp scan sums the blocks of each point; the last lane of a point applies Dinv.
__global__ void __launch_bounds__(32) k_backsub_points(DevGraph g, const LmDev *lm, double *scale_part) {
  const double lambda = lm->lambda;
  const int lane = threadIdx.x;
  const int unit = blockIdx.x;
  double sc = 0;
  const int e0 = g.unit_e0[unit], e1 = g.unit_e0[unit + 1];
  const bool big = e1 - e0 > 32;  // a single point with more than 32 blocks: partial sums carried from chunk to chunk
  double carry[3] = {0, 0, 0};
  for (int base = e0; base < e1; base += 32) {
    const int e = base + lane;
    const bool valid = e < e1;
    double c[3] = {0, 0, 0};
    int key = -1;
    if (valid) {
      key = g.pe_pt[e];
      const int p = g.ent_pidx[g.n_slots + e];
      if (p >= 0) {
        const double2 *W = reinterpret_cast<const double2 *>(g.Hpl + 18 * (size_t)(g.n_slots + e));
        double w[18], x[6];
#pragma unroll
        for (int i = 0; i < 9; i++) {
          const double2 t = W[i];
          w[2 * i] = t.x, w[2 * i + 1] = t.y;
        }
#pragma unroll
        for (int a = 0; a < 6; a++) x[a] = g.xp[6 * p + a];
#pragma unroll
        for (int a = 0; a < 6; a++) {
          c[0] = fma(w[3 * a], x[a], c[0]);
          c[1] = fma(w[3 * a + 1], x[a], c[1]);
          c[2] = fma(w[3 * a + 2], x[a], c[2]);
        }
      }
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int ku = __shfl_up_sync(FULL, key, d);
      const bool take = lane >= d && ku == key;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double t = __shfl_up_sync(FULL, c[i], d);
        if (take) c[i] += t;
      }
    }
    const int kn = __shfl_down_sync(FULL, key, 1);
    const bool seg_end = valid && (lane == 31 || kn != key);
    const bool final_chunk = base + 32 >= e1;
    if (big) {
      const int last = min(32, e1 - base) - 1;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double t = __shfl_sync(FULL, c[i], last);
        if (!final_chunk) carry[i] += t;
      }
    }
    if (seg_end) {
      const int L = g.n_pl + key;
      if (final_chunk) {  // (units of several points are a single chunk)
        if (landmark_active(g, L)) {
          double D[6], bl[3], cl[3], x[3];
#pragma unroll
          for (int i = 0; i < 6; i++) D[i] = g.Dinv[6 * (size_t)L + i];
#pragma unroll
          for (int i = 0; i < 3; i++) bl[i] = g.bl[3 * (size_t)L + i], cl[i] = bl[i] - (c[i] + carry[i]);
          x[0] = D[0] * cl[0] + D[1] * cl[1] + D[2] * cl[2];
          x[1] = D[1] * cl[0] + D[3] * cl[1] + D[4] * cl[2];
          x[2] = D[2] * cl[0] + D[4] * cl[1] + D[5] * cl[2];
#pragma unroll
          for (int i = 0; i < 3; i++) {
            g.xl[3 * (size_t)L + i] = x[i];
            sc += x[i] * (lambda * x[i] + bl[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 3; i++) g.xl[3 * (size_t)L + i] = 0.0;
        }
      }
    }
  }
  sc = warp_sum(sc);
  if (lane == 0) scale_part[blockIdx.x] = sc;
}

// ---------------------------------------------------------------------------------------------
// SparseOptimizer::update (core/sparse_optimizer.cpp:422-435): trial = cur (+) x
// ---------------------------------------------------------------------------------------------
__global__ void k_update(DevGraph g, DevState cur, DevState tr) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < g.n_kf) {
    double p[7], o[7], Rt[12];
#pragma unroll
    for (int i = 0; i < 7; i++) p[i] = cur.kf_pose[7 * t + i];
    const int idx = g.kf_idx[t];
    if (idx >= 0) {
      double u[6];
#pragma unroll
      for (int i = 0; i < 6; i++) u[i] = g.xp[6 * idx + i];
      se3_oplus(p, u, o);
    } else {
#pragma unroll
      for (int i = 0; i < 7; i++) o[i] = p[i];
    }
    pose_to_Rt(o, Rt);
#pragma unroll
    for (int i = 0; i < 7; i++) tr.kf_pose[7 * t + i] = o[i];
#pragma unroll
    for (int i = 0; i < 12; i++) tr.kf_Rt[12 * t + i] = Rt[i];
    return;
  }
  t -= g.n_kf;
  if (t < g.n_cu) {
    double c[10], o[10];
#pragma unroll
    for (int i = 0; i < 10; i++) c[i] = cur.cu[10 * t + i];
    const int off = g.cu_off[t];
    if (off >= 0) {
      double u[9];
#pragma unroll
      for (int i = 0; i < 9; i++) u[i] = g.xp[off + i];
      cuboid_oplus(c, g.cu_flags[t], u, o);
    } else {
#pragma unroll
      for (int i = 0; i < 10; i++) o[i] = c[i];
    }
#pragma unroll
    for (int i = 0; i < 10; i++) tr.cu[10 * t + i] = o[i];
    return;
  }
  t -= g.n_cu;
  if (t < g.n_pl) {
    double c[4], o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) c[i] = cur.pl[4 * t + i];
    if (g.pl_act[t]) {
      const double v[3] = {g.xl[3 * (size_t)t], g.xl[3 * (size_t)t + 1], g.xl[3 * (size_t)t + 2]};
      plane_oplus(c, v, o);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) o[i] = c[i];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) tr.pl[4 * t + i] = o[i];
    return;
  }
  t -= g.n_pl;
  if (t < g.n_pt) {
    const bool act = g.pt_act[t] && !g.pt_fixed[t];
    const size_t L = (size_t)g.n_pl + t;
#pragma unroll
    for (int i = 0; i < 3; i++) tr.pt[3 * (size_t)t + i] = cur.pt[3 * (size_t)t + i] + (act ? g.xl[3 * L + i] : 0.0);
  }
}
// kf_Rt cache from kf_pose (after set_graph / reset)
__global__ void k_pose_cache(int n_kf, const double *pose, double *Rt) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_kf) return;
  double p[7], r[12];
#pragma unroll
  for (int i = 0; i < 7; i++) p[i] = pose[7 * t + i];
  pose_to_Rt(p, r);
#pragma unroll
  for (int i = 0; i < 12; i++) Rt[12 * t + i] = r[i];
}

// ---------------------------------------------------------------------------------------------
// scalars
// ---------------------------------------------------------------------------------------------
struct Scalars {
  double chi2;      // active robust chi2
  double scale;     // computeScale
  double max_diag;  // for computeLambdaInit
  int not_spd;
  int pad;
};
// sums partial arrays in a fixed order; single block
constexpr int SCAL_THREADS = 1024;  // one CTA, fixed summation order (deterministic)
__global__ void __launch_bounds__(SCAL_THREADS) k_scalars(DevGraph g, Scalars *out, const double *chi_a, int na, const double *chi_b, int nb, const double *chi_c, int nc,
                          const double *chi_d, int nd, const LmDev *lm, int with_const, const double *scale_part, int ns, int n_p,
                          const int *not_spd, double *red /* [chi2, scale] for the cross-rank reduction, may be null */) {
  __shared__ double sm[SCAL_THREADS / 32];
  const double lambda = lm->lambda, chi_const = with_const ? lm->chi_const : 0.0;
  double c = 0, s = 0;
  for (int i = threadIdx.x; i < na; i += SCAL_THREADS) c += chi_a[i];
  for (int i = threadIdx.x; i < nb; i += SCAL_THREADS) c += chi_b[i];
  for (int i = threadIdx.x; i < nc; i += SCAL_THREADS) c += chi_c[i];
  for (int i = threadIdx.x; i < nd; i += SCAL_THREADS) c += chi_d[i];
  for (int i = threadIdx.x; i < ns; i += SCAL_THREADS) s += scale_part[i];
  for (int i = threadIdx.x; i < n_p; i += SCAL_THREADS) s += g.xp[i] * (lambda * g.xp[i] + g.bp[i]);
  c = block_sum<SCAL_THREADS>(c, sm);
  s = block_sum<SCAL_THREADS>(s, sm);
  if (threadIdx.x == 0) {
    out->chi2 = c + chi_const;
    out->scale = s;
    out->not_spd = not_spd ? *not_spd : 0;
    if (red) {
      red[0] = out->chi2;
      red[1] = out->scale;
      red[2] = (double)out->not_spd;
    }
  }
}
// max |diagonal| over the active blocks (computeLambdaInit, levenberg.cpp:166-180).  Grid-stride; the values are
// non-negative, so the maximum of their bit patterns is the maximum of the values (out->max_diag is zeroed before).
__global__ void __launch_bounds__(256) k_max_diag(DevGraph g, Scalars *out) {
  __shared__ double sm[256];
  double m = 0;
  const int nb = g.dims[0];
  const int t0 = blockIdx.x * 256 + threadIdx.x, stride = gridDim.x * 256;
  for (int i = t0; i < nb * 6; i += stride) m = fmax(m, fabs(g.Hpp_kf[36 * (size_t)(i / 6) + 7 * (i % 6)]));
  for (int i = t0; i < g.n_cu * 9; i += stride)
    if (g.cu_off[i / 9] >= 0) m = fmax(m, fabs(g.Hpp_cu[81 * (size_t)(i / 9) + 10 * (i % 9)]));
  for (int i = t0; i < g.n_lm * 3; i += stride) {
    const int L = i / 3, j = i % 3;
    m = fmax(m, fabs(g.Hll[6 * (size_t)L + (j == 0 ? 0 : (j == 1 ? 3 : 5))]));
  }
  sm[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long *>(&out->max_diag), (unsigned long long)__double_as_longlong(sm[0]));
}
// sharded window: the caller's stop flag joins the per-trial all-reduce, so that every rank leaves the LM loop at the same trial
// (red[3]: this rank's view before the reduction, the number of ranks that saw the flag after it)
__global__ void k_stop_to_red(const volatile int *stop, double *red) {
  if (threadIdx.x == 0 && blockIdx.x == 0) red[3] = (stop && *stop) ? 1.0 : 0.0;
}
__global__ void k_stop_from_red(volatile int *stop, const double *red) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && red[3] > 0.0) *stop = 1;
}
// after the cross-rank reductions: copy the reduced scalars back into the Scalars block
__global__ void k_scalars_from_red(Scalars *out, const double *red, int which) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (which == 0 || which == 2) {
      out->chi2 = red[0];
      out->scale = red[1];
      if (which == 2) out->not_spd = red[2] != 0.0;  // distributed factorisation: the pivots live on their owners
    } else {
      out->max_diag = red[2];
    }
  }
}
__global__ void k_set_red_maxdiag(const Scalars *in, double *red) {
  if (threadIdx.x == 0 && blockIdx.x == 0) red[2] = in->max_diag;
}

// ---------------------------------------------------------------------------------------------
// Levenberg-Marquardt controller on the device (levenberg.cpp:61-164).  One thread each; `graph` != 0: the loop flags are
// also written into the conditional handles of the WHILE nodes that replay the captured iteration / trial bodies.
// ---------------------------------------------------------------------------------------------
__global__ void k_lm_begin(LmDev *lm, const LmIn *in, int graph, cudaGraphConditionalHandle h_iter) {
  if (threadIdx.x || blockIdx.x) return;
  lm->iters = in->iters, lm->max_trials = in->max_trials;
  lm->tau = in->tau, lm->good_upper = in->good_upper, lm->good_lower = in->good_lower, lm->chi_const = in->chi_const;
  lm->it = 0, lm->done = 0, lm->term = 0, lm->total_trials = 0, lm->stop_seen = 0, lm->accepted = 0, lm->qmax = 0;
  lm->lambda = -1.0, lm->ni = 2.0, lm->nBad = 0;
  lm->currentChi = lm->iniChi = lm->tempChi = lm->rho = lm->chi2_initial = 0.0;
  lm->iter_continue = in->iters > 0;
  lm->trial_continue = 0;
  if (graph) cudaGraphSetConditional(h_iter, lm->iter_continue);
}
// start of an outer iteration, after the linearisation: currentChi = activeRobustChi2() (:76), computeLambdaInit at iteration 0 (:93-97)
__global__ void k_lm_iter_begin(LmDev *lm, const Scalars *scal, int graph, cudaGraphConditionalHandle h_trial) {
  if (threadIdx.x || blockIdx.x) return;
  lm->currentChi = scal->chi2;
  lm->iniChi = lm->currentChi;
  if (lm->it == 0) {
    lm->chi2_initial = lm->currentChi;
    lm->lambda = lm->tau * scal->max_diag;
    lm->ni = 2.0;
    lm->nBad = 0;
  }
  lm->qmax = 0;
  lm->rho = 0.0;
  lm->accepted = 0;
  lm->trial_continue = 1;
  if (graph) cudaGraphSetConditional(h_trial, 1);
}
// end of a damped trial (:126-152) and, when the trial loop ends, of the outer iteration (:154-161; SparseOptimizer::optimize loop head)
__global__ void k_lm_decide(LmDev *lm, const Scalars *scal, const volatile int *stop, int graph, cudaGraphConditionalHandle h_trial,
                            cudaGraphConditionalHandle h_iter) {
  if (threadIdx.x || blockIdx.x) return;
  const bool stopped = stop && *stop;
  double tempChi = scal->chi2;
  if (scal->not_spd) tempChi = 1.7976931348623157e308;  // solve failed: std::numeric_limits<double>::max() (:126-127)
  double rho = lm->currentChi - tempChi;
  const double scale = scal->scale + 1e-3;  // (:129-131)
  rho /= scale;
  lm->tempChi = tempChi;
  if (rho > 0 && isfinite(tempChi)) {  // (:133-142)
    double alpha = 1. - pow((2 * rho - 1), 3);
    alpha = fmin(alpha, lm->good_upper);
    const double scaleFactor = fmax(lm->good_lower, alpha);
    lm->lambda *= scaleFactor;
    lm->ni = 2;
    lm->currentChi = tempChi;
    lm->accepted = 1;  // discardTop: the trial becomes the estimate (k_accept)
  } else {             // (:143-147)
    lm->lambda *= lm->ni;
    lm->ni *= 2;
    lm->accepted = 0;  // pop
  }
  lm->rho = rho;
  lm->qmax++;
  const bool again = rho < 0 && lm->qmax < lm->max_trials && !stopped;  // (:149)
  lm->trial_continue = again;
  if (graph) cudaGraphSetConditional(h_trial, again);
  if (again) return;
  // ---- the outer iteration is over ----
  const int it = lm->it;
  lm->done++;
  lm->total_trials += lm->qmax;
  if (it < PPO_TRACE_MAX) {
    ppo_ba_iter &r = lm->trace[it];
    r.chi2_before = lm->iniChi;
    r.chi2_after = lm->currentChi;
    r.lambda = lm->lambda;
    r.rho = rho;
    r.trials = lm->qmax;
    r.accepted = lm->accepted;
  }
  bool ok = true;
  if (lm->qmax == lm->max_trials || rho == 0) {  // Terminate (:151-152)
    ok = false;
    lm->term = 1;
  } else {  // Raul's rule (:155-161)
    if ((lm->iniChi - lm->currentChi) * 1e3 < lm->iniChi) lm->nBad++;
    else lm->nBad = 0;
    if (lm->nBad >= 3) {
      ok = false;
      lm->term = 1;
    }
  }
  lm->it = it + 1;
  if (stopped) lm->stop_seen = 1;
  lm->iter_continue = ok && lm->it < lm->iters && !stopped;
  if (graph) cudaGraphSetConditional(h_iter, lm->iter_continue);
}
// discardTop / pop of the estimate stack: an accepted trial is copied over the current estimates
__global__ void k_accept(DevGraph g, DevState cur, DevState tr, const LmDev *lm) {
  if (!lm->accepted) return;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = t0; i < 7 * (size_t)g.n_kf; i += stride) cur.kf_pose[i] = tr.kf_pose[i];
  for (size_t i = t0; i < 12 * (size_t)g.n_kf; i += stride) cur.kf_Rt[i] = tr.kf_Rt[i];
  for (size_t i = t0; i < 3 * (size_t)g.n_pt; i += stride) cur.pt[i] = tr.pt[i];
  for (size_t i = t0; i < 4 * (size_t)g.n_pl; i += stride) cur.pl[i] = tr.pl[i];
  for (size_t i = t0; i < 10 * (size_t)g.n_cu; i += stride) cur.cu[i] = tr.cu[i];
}

// ---------------------------------------------------------------------------------------------
// outlier pass (Optimizer.cc:2736-2833) and depth test
// ---------------------------------------------------------------------------------------------
__global__ void k_outlier_pass(DevGraph g, DevState s, double chi2_mono, double chi2_stereo, double chi2_plane, double chi2_vp,
                               double norm_bbox, double norm_corner, double norm_se3, int *n_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.n_pe) {
    const PointEdgeRec rec = g.pe_rec[i];
    const int pt = g.pe_pt[i];
    const double z = s.kf_Rt[12 * rec.kf + 6] * s.pt[3 * pt] + s.kf_Rt[12 * rec.kf + 7] * s.pt[3 * pt + 1] +
                     s.kf_Rt[12 * rec.kf + 8] * s.pt[3 * pt + 2] + s.kf_Rt[12 * rec.kf + 11];
    unsigned fl = g.pe_flags[i];
    if (g.pe_chi2[i] > (rec.ur < 0.f ? chi2_mono : chi2_stereo) || !(z > 0.0)) {
      if (!(fl & PPO_EF_LEVEL1_)) atomicAdd(&n_out[0], 1);
      fl |= PPO_EF_LEVEL1_;
    }
    fl &= ~PPO_EF_ROBUST_;
    g.pe_flags[i] = (uint8_t)fl;
  }
  if (i < g.n_cbe) {
    unsigned fl = g.cbe_flags[i];
    if (g.cbe_norm[i] > (g.cbe_kind[i] == 0 ? norm_bbox : (g.cbe_kind[i] == 2 ? norm_se3 : norm_corner))) {  // SE3: Optimizer.cc:1875-1882
      if (!(fl & PPO_EF_LEVEL1_)) atomicAdd(&n_out[2], 1);
      fl |= PPO_EF_LEVEL1_;
    }
    g.cbe_flags[i] = (uint8_t)fl;
  }
  if (i < g.n_ple) {
    unsigned fl = g.ple_flags[i];
    if (g.ple_chi2[i] > (g.ple_kind[i] == 0 ? chi2_plane : chi2_vp)) {
      if (!(fl & PPO_EF_LEVEL1_)) atomicAdd(&n_out[1], 1);
      fl |= PPO_EF_LEVEL1_;
    }
    fl &= ~PPO_EF_ROBUST_;
    g.ple_flags[i] = (uint8_t)fl;
  }
}
// Point edges the caller erases after the BA (Optimizer.cc:2840-2852): e->chi2() above the threshold of its kind or !e->isDepthPositive();
// compacted edge ids in arbitrary order (the host sorts the few it gets)
__global__ void k_point_edge_outliers(DevGraph g, DevState s, double th_mono, double th_stereo, int *idx, int *count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool out = false;
  if (i < g.n_pe) {
    const PointEdgeRec rec = g.pe_rec[i];
    const int pt = g.pe_pt[i];
    double p[3], X[3] = {s.pt[3 * pt], s.pt[3 * pt + 1], s.pt[3 * pt + 2]};
    double q[4] = {s.kf_pose[7 * rec.kf], s.kf_pose[7 * rec.kf + 1], s.kf_pose[7 * rec.kf + 2], s.kf_pose[7 * rec.kf + 3]};
    quat_rot(q, X, p);
    const bool depth_positive = (p[2] + s.kf_pose[7 * rec.kf + 6]) > 0.0;
    out = g.pe_chi2[i] > (rec.ur < 0 ? th_mono : th_stereo) || !depth_positive;
  }
  const unsigned mask = __ballot_sync(0xffffffffu, out);
  if (mask) {
    const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (out) idx[base + __popc(mask & ((1u << lane) - 1))] = i;
  }
}
__global__ void k_depth_flags(DevGraph g, DevState s, int kind, unsigned char *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (kind == 0 && i < g.n_pe) {
    const int kf = g.pe_rec[i].kf, pt = g.pe_pt[i];
    double p[3], X[3] = {s.pt[3 * pt], s.pt[3 * pt + 1], s.pt[3 * pt + 2]};
    double q[4] = {s.kf_pose[7 * kf], s.kf_pose[7 * kf + 1], s.kf_pose[7 * kf + 2], s.kf_pose[7 * kf + 3]};
    quat_rot(q, X, p);
    out[i] = (p[2] + s.kf_pose[7 * kf + 6]) > 0.0;
  }
  if (kind == 1 && i < g.n_ple) {  // EdgePlane::isDepthPositive, G2O_Plane3D.h:199-209
    double l[4], pc[4], Rt[12];
    for (int k = 0; k < 4; k++) pc[k] = s.pl[4 * g.ple_plane[i] + k];
    for (int k = 0; k < 12; k++) Rt[k] = s.kf_Rt[12 * g.ple_kf[i] + k];
    plane_transform(Rt, pc, l);
    out[i] = (-l[3]) > 0;
  }
}

}  // namespace ppo
