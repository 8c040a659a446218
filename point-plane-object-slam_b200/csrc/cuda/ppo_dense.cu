// Dense solve of the Schur-reduced pose system: the GPU counterpart of
// LinearSolverDense::solve (Thirdparty/g2o/g2o/solvers/linear_solver_dense.h:65-113, Eigen::LDLT +
// isPositive()).  Right-looking tiled Cholesky in double; a non-positive pivot raises `not_spd`,
// which the LM loop treats exactly like g2o's failed solve (step rejected).
//
// Storage (ppo_dense.h): the lower triangle of Hschur cut into 64 x 64 tiles, every tile contiguous in memory with its
// columns padded to 68 doubles, so that ONE TMA bulk copy (cp.async.bulk, 34816 bytes) lands a tile in shared memory in
// a bank-conflict-free layout for the FP64 tensor-core fragments.  The reduced gradient is an extra tile row below the
// matrix: carrying it through the panel solves and trailing updates turns it into y = L^-1 b (forward substitution for
// free); a second kernel does the backward substitution.
//
// ONE launch factorises the whole system (k_chol_dataflow).  Every tile (i, j) goes through the operations
//      U_0 .. U_{j-1}   C_ij -= P_ik P_jk^T                         (trailing updates, any CTA)
//      F_j   (i == j)   W_j = chol(C_jj)^-1                         (critical-path CTA)
//      T_j   (i >  j)   P_ij = C_ij W_j^T                           (any CTA; the tile right below the diagonal: critical-path CTA)
// and carries a version counter ver[i][j] = number of operations applied (tagged with the launch epoch, so the counters
// are never cleared).  The first CTA to start takes the critical path  F_k -> T_k(k+1) -> U_k(k+1,k+1) -> F_{k+1}  and keeps
// the diagonal tile in shared memory between steps; all other CTAs pull the remaining operations from a global queue that
// is ordered level by level (T_k first, then U_k by column), i.e. topologically: an operation only waits for operations
// that were claimed before it, by CTAs that are therefore running -- the kernel cannot deadlock whatever number of CTAs
// is resident, and it needs no cooperative launch.  Workers prefetch the operands of their next operation while the
// current one runs on the tensor cores (DMMA).  Every wait is bounded: on a time-out the kernel raises ctrl->err, all
// CTAs leave, and the solve counts as failed (the back-substitution kernel clears the flag for the next solve).
//
// What was measured on the way (tools/ubench/chol_test[_t]; n = 1644 is latency-bound by the 26-panel chain, n = 7794 throughput-bound):
//   * a 9th warp per CTA that does all the talking (claims, counter polls, TMA issue, publication)      0.454 -> 0.445 ms, 11.9 -> 10.1 ms
//   * warp 0 updates + factorises the first 16 x 16 sub-block of the next diagonal tile while warps 1..7
//     finish the tile's last trailing update                                                             0.445 -> 0.415 ms
//   * workers at n = 7794 spend 2700 of 8500 cycles per operation waiting for operands and 5750 computing + storing (4096 = the DMMA
//     time of a 64^3 product).  Tried and dropped: claiming runs of entries that share their column operand (a third less operand
//     traffic: 10.0 -> 9.9 ms, 1.70 -> 1.93 at n = 4000), a communication thread that runs ahead of the buffer ring (9.7 ms without the
//     device-scope fence before the release, which racecheck does not accept; 10.4 with it), three buffer sets with the target tile read
//     straight into registers (waiting 1900, computing 7500: 11.0 ms).
//   * per panel at n = 1644 (instrumented build, cycles): F 17.9 k (four 16 x 16 factorisations of 2.7 k each, i.e. ~300 per 2 x 2 pivot step,
//     plus the strips / updates between them), T 4.3 k, U 6.3 k (with the look-ahead factorisation inside), publish + operands 1.1 k.
//     Tried and dropped: (1) warp 0 carrying the whole serial chain of F alone (its own strips of L, bar.arrive hand-over, no block
//     barrier between (b) and (c)): 0.416 -> 0.418 ms, (b) was already one parallel round.  (2) T and U re-arranged around the chain:
//     all warps form rows 0..15 of P first, then warp 0 updates + factorises the first sub-block of the next tile while the other warps
//     form the remaining 48 rows and update the other 33 blocks (W double-buffered): T + U 10.5 k -> 8.2 k, but the look-ahead
//     factorisation next to 7 warps of DMMA + fragment loads takes 4.5 k instead of 2.7 k (shared-memory round trips of the pivot steps
//     queue behind the fragment traffic; keeping the warp on the same scheduler idle does not help): 0.423 / 0.427 ms.  The tensor work of
//     T + U on one SM is 1152 DMMAs = 4.6 k cycles at 16 cycles per DMMA and sub-partition, so the chain cannot hide behind it.
//     (3) inside the 16 x 16 factorisation: the pivot block by shuffle from its owners instead of through shared memory (kept: 2744 -> 2572
//     cycles per sub-block, 0.416 -> 0.414 ms).  Every lane forming the NEXT pivot block itself, so that only one FMA, the determinant
//     and the reciprocal are on the chain (dropped: 3400 cycles per sub-block, 0.457 ms -- the warp issues in order at 2 cycles per
//     FP64 instruction, and the extra off-chain arithmetic delays the chain more than the shorter dependency path gains).
//     (4) the three look-ahead blocks of warp 0 in U as three interleaved DMMA chains instead of one after the other: no change
//     (0.416 ms), U is not bound by that warp's tensor work.
//
// The second half of the file spreads the same factorisation over the GPUs of a node (k_chol_dist and friends).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "ppo_dense.h"

namespace ppo {

constexpr int NB = DENSE_NB;        // tile size
constexpr int CLD = DENSE_CLD;      // padded column length of a tile (doubles): 544-byte columns, 16-byte aligned
constexpr int TILE = DENSE_TILE;    // doubles per tile
constexpr unsigned TILE_BYTES = TILE * 8;
constexpr int SB = 16;              // sub-block of the diagonal-tile factorisation
#ifndef PPO_CHOL_TIMEOUT_SHIFT
#define PPO_CHOL_TIMEOUT_SHIFT 32
#endif
constexpr long long CHOL_TIMEOUT = 1ll << PPO_CHOL_TIMEOUT_SHIFT;  // cycles (~2 s)

#ifdef PPO_CHOL_TIMING
__device__ long long g_chol_t[16];
#define CH_STAMP(var) const long long var = clock64()
#define CH_ACC(slot, a, b) if (threadIdx.x == 0) g_chol_t[slot] += (b) - (a)
#else
#define CH_STAMP(var)
#define CH_ACC(slot, a, b)
#endif

struct CholCtrl {
  int ticket;  // role tickets of the running launch (first CTA to arrive = critical-path CTA)
  int qhead;   // next operation of the worker queue
  int done;    // CTAs that have left the kernel; the last one resets the block for the next launch
  int epoch;   // launch counter: flag values are epoch * 256 + level
  int err;     // 1: a wait timed out
  int bticket, bdone, bepoch;  // the same for the back-substitution kernel
};

// branch-free reciprocal / reciprocal square root of a normal positive double (hardware seed + polynomial / Newton steps);
// anything else yields garbage, which the caller has already flagged as "not positive definite".
__device__ __forceinline__ double pf_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d, y, 1.0);  // y (1 + e + e^2 + e^3): one quartic step from the >= 17-bit seed
  return fma(y, fma(fma(e, e, e), e, e), y);
}
__device__ __forceinline__ double pf_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);
  y = fma(y * e, fma(0.375, e, 0.5), y);
  return fma(y, fma(-0.5 * x * y, y, 0.5), y);
}

// mma.sync.aligned.m8n8k4.row.col.f64: A fragment a[row = lane/4][k = lane%4], B fragment b[k = lane%4][col = lane/4],
// C fragment c0,c1 = C[row = lane/4][col = 2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!ok);
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ bool flag_ready(const int *flag, int want) { return ld_acquire(flag) >= want; }
// bounded wait until *flag >= want; false (and ctrl->err raised) on time-out or when another CTA has raised it
__device__ __forceinline__ bool flag_wait(const int *flag, int want, CholCtrl *ctrl) {
  if (ld_acquire(flag) >= want) return true;
  const long long t0 = clock64();
  for (int it = 1;; it++) {
    if (ld_acquire(flag) >= want) return true;
    if ((it & 63) == 0) {
      if (*(volatile int *)&ctrl->err) return false;
      if (clock64() - t0 > CHOL_TIMEOUT) {
        atomicExch(&ctrl->err, 1);
        return false;
      }
    }
    __nanosleep(20);
  }
}

// system-scope variants: counters written by another GPU over NVLink (distributed factorisation below)
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ bool flag_wait_sys(const int *flag, int want, CholCtrl *ctrl) {
  if (ld_acquire_sys(flag) >= want) return true;
  const long long t0 = clock64();
  for (int it = 1;; it++) {
    if (ld_acquire_sys(flag) >= want) return true;
    if ((it & 63) == 0) {
      if (*(volatile int *)&ctrl->err) return false;
      if (clock64() - t0 > CHOL_TIMEOUT) {
        atomicExch(&ctrl->err, 1);
        return false;
      }
    }
    __nanosleep(20);
  }
}

typedef double (*TilePtr)[CLD];  // tile[col][row]

// Row block of a warp in the tile products.  Warps w and w + 4 issue on the same scheduler; in the lower-triangular products row
// block rb costs rb + 1 column blocks, so the pairs get (rb, 7 - rb): every scheduler carries 9 of the 36 blocks.
__device__ __forceinline__ int warp_row_block() {
  const int warp = threadIdx.x >> 5;
  return warp < 4 ? warp : 11 - warp;
}
// acc(i, j) += sum_m sA[m][i] sB[m][j] on a 64 x 64 x 64 tile; a warp owns the 8 rows of its row block.
// MODE 0: full; 1: lower triangle only (column blocks <= row block); 2: sB = W^T of a lower-triangular W (m <= j).
// Fully unrolled and software-pipelined by hand: the fragments of k-step m0 + 4 are loaded from shared memory while the DMMAs
// of step m0 issue (a DMMA that waits for its own operand load stalls the FP64 pipe for the shared-memory latency).
template <int MODE>
__device__ __forceinline__ void tile_mma64(const double (*sA)[CLD], const double (*sB)[CLD], double acc[8][2]) {
  const int lane = threadIdx.x & 31, rb = warp_row_block();
  const int row = rb * 8 + (lane >> 2), kk = lane & 3, lc = lane >> 2;
  const int nmax = MODE == 1 ? rb : 7;  // last column block of this warp (warp-uniform)
  double a_n = sA[kk][row], b_n[8];
#pragma unroll
  for (int nbk = 0; nbk < 8; nbk++) b_n[nbk] = (nbk <= nmax) ? sB[kk][nbk * 8 + lc] : 0.0;
#pragma unroll
  for (int m0 = 0; m0 < NB; m0 += 4) {
    const double a = a_n;
    double b[8];
#pragma unroll
    for (int nbk = 0; nbk < 8; nbk++) b[nbk] = b_n[nbk];
    if (m0 + 4 < NB) {
      a_n = sA[m0 + 4 + kk][row];
#pragma unroll
      for (int nbk = 0; nbk < 8; nbk++) {
        if (MODE == 2 && m0 + 4 > 8 * nbk + 7) continue;  // (compile time)
        if (nbk <= nmax) b_n[nbk] = sB[m0 + 4 + kk][nbk * 8 + lc];
      }
    }
#pragma unroll
    for (int nbk = 0; nbk < 8; nbk++) {
      if (MODE == 2 && m0 > 8 * nbk + 7) continue;  // (compile time)
      if (nbk <= nmax) dmma(acc[nbk][0], acc[nbk][1], a, b[nbk]);
    }
  }
}

// ---- diagonal tile: W = chol(A)^-1 by block Gauss-Jordan elimination of [A | I] with 16 x 16 sub-blocks ------------------------
// D[c][r] holds the symmetric tile (both triangles; identity beyond the valid size), X[c][r] must be zero on entry and holds
// W (lower triangular) on exit.  Per sub-block s:
//   (a) warp 0: V = chol(A_ss)^-1, warp-synchronous: lane i < 16 owns row i of A_ss, lane 16 + i row i of the identity part;
//       per pivot the column of A and the pivot row of X go through shared memory (one round trip), the rank-1 update is
//       16 FMAs per lane; no block barrier inside the 16 pivots.
//   (b) L_is = A_is V^T (rows below), X_s,: = V X_s,: (columns left)              -- DMMA, all warps
//   (c) A_ij -= L_is L_js^T, X_i,: -= L_is X_s,:                                  -- DMMA, all warps
struct __align__(16) FacSmem {
  // buf[half][slot][vector][index]: half 0 = the two pivot columns of A (read by the row lanes), half 1 = the two multiplier
  // vectors (read by the column lanes one step later).  The multipliers of step j go to slot (j/2 + 1) & 1, so that in every step
  // BOTH halves read slot (j/2) & 1 of their own half: one per-lane base address, no selects in the loop.
  double buf[2][2][2][SB];
  double rs[SB];  // 1 / sqrt(d_i)
};
__device__ __forceinline__ void sub_factor16(TilePtr D, TilePtr X, int s, FacSmem &fs, int *not_spd) {
  // Two pivots per step (2 x 2 block elimination): the serial chain per step is  shared-memory round trip -> determinant ->
  // reciprocal -> two multipliers -> update.  Lane i < 16 owns ROW i of A_ss; lane 16 + c owns COLUMN c of the identity part
  // X, so its pivot entries are its own registers and all it needs per step are the two multiplier vectors, which the row
  // lanes publish; the column lanes run one step behind, off the serial chain.  Both halves execute the same instruction
  // r[k] += vec0[k] * s0 + vec1[k] * s1   (rows: vec = pivot columns, s = own multipliers; columns: vec = multipliers, s = own pivots),
  // and only for the indices a step can still change (k >= j - 2).
  const int lane = threadIdx.x & 31, li = lane & 15;
  const bool lo = lane < 16;
  const int o = SB * s;
  double r[SB];
#pragma unroll
  for (int c = 0; c < SB; c++) r[c] = lo ? D[o + c][o + li] : (c == li ? 1.0 : 0.0);
  (&fs.buf[1][0][0][0])[lane] = 0.0, (&fs.buf[1][0][0][0])[lane + 32] = 0.0;  // the column lanes read their half before it is first written (times zero)
  double my_d = 1.0;
  const double *mine = &fs.buf[lo ? 0 : 1][0][0][0];
#pragma unroll
  for (int j = 0; j <= SB; j += 2) {  // (the last round only flushes the lagging identity half)
    const int bi = (j >> 1) & 1;
    // The 2 x 2 pivot block comes straight out of the registers of its owners (lanes j, j+1) by shuffle, and a lane's own entries of the
    // two pivot columns are its own r[j], r[j+1] (symmetry): the serial chain does not wait for the shared-memory round trip below, which
    // only feeds the rank-2 update at the end of the step.
    double b00 = 1.0, b10 = 0.0, b11 = 1.0;
    if (j < SB) {
      b00 = __shfl_sync(0xffffffffu, r[j], j);
      b10 = __shfl_sync(0xffffffffu, r[j], j + 1);
      b11 = __shfl_sync(0xffffffffu, r[j + 1], j + 1);
    }
    if (j < SB && lo) fs.buf[0][bi][0][li] = r[j], fs.buf[0][bi][1][li] = r[j + 1];  // columns j, j+1 of the current A (= rows, by symmetry)
    __syncwarp();
    double s0 = 0.0, s1 = 0.0;
    if (j < SB) {
      const double c0 = r[j], c1 = r[j + 1];
      double det = fma(b00, b11, -b10 * b10);
      const bool bad = !(b00 > 0.0) || !(det > 0.0);
      if (bad && lane == 0) *not_spd = 1;  // everything computed from here on is garbage and will be discarded
      b00 = bad ? 1.0 : b00, b10 = bad ? 0.0 : b10, b11 = bad ? 1.0 : b11, det = bad ? 1.0 : det;
      const double idet = pf_rcp(det), i00 = pf_rcp(b00);
      if (li == j) my_d = b00;
      if (li == j + 1) my_d = det * i00;  // d_{j+1} = b11 - b10^2 / b00
      // rows below the pivot pair: [t0 t1] = -[c0 c1] B^-1; row j+1 itself is eliminated by pivot j alone
      double t0 = -(c0 * b11 - c1 * b10) * idet, t1 = -(c1 * b00 - c0 * b10) * idet;
      t0 = li == j + 1 ? -b10 * i00 : t0, t1 = li == j + 1 ? 0.0 : t1;
      t0 = li <= j ? 0.0 : t0, t1 = li <= j ? 0.0 : t1;
      if (lo) {
        fs.buf[1][bi ^ 1][0][li] = t0, fs.buf[1][bi ^ 1][1][li] = t1;
        s0 = t0, s1 = t1;
      }
    }
    if (!lo && j >= 2) s0 = r[j >= 2 ? j - 2 : 0], s1 = r[j >= 2 ? j - 1 : 1];  // X(j-2, c), X(j-1, c) of my column c, before this step touches them
    const double *vec0 = mine + bi * 2 * SB, *vec1 = vec0 + SB;
    auto upd = [&](int c) {
      const double2 p = *reinterpret_cast<const double2 *>(&vec0[c]);
      const double2 q = *reinterpret_cast<const double2 *>(&vec1[c]);
      r[c] = fma(q.x, s1, fma(p.x, s0, r[c]));
      r[c + 1] = fma(q.y, s1, fma(p.y, s0, r[c + 1]));
    };
    if (j + 2 < SB) upd(j + 2);  // the next pivot pair first: the next step starts from it
#pragma unroll
    for (int c = (j >= 2 ? j - 2 : 2); c < SB; c += 2)
      if (c != j + 2) upd(c);
  }
  // V(i, c) = X(i, c) / sqrt(d_i): column lanes hold X(:, c)
  if (lo) fs.rs[li] = pf_rsqrt(my_d);
  __syncwarp();
  if (!lo) {
#pragma unroll
    for (int i = 0; i < SB; i++) X[o + li][o + i] = r[i] * fs.rs[i];
  }
  __syncwarp();
}
// one 8 x 8 output block of a 16-deep product on the tensor cores:  acc += sum_m Aop(r0 + ., m) Bop(m, c0 + .)
template <typename FA, typename FB>
__device__ __forceinline__ void blk_mma16(FA aop, FB bop, double &c0, double &c1) {
  const int lane = threadIdx.x & 31, rr = lane >> 2, kk = lane & 3;
#pragma unroll
  for (int ks = 0; ks < SB; ks += 4) dmma(c0, c1, aop(rr, ks + kk), bop(ks + kk, rr));
}
// barrier of the 8 compute warps (the 9th warp of the CTA only talks to the rest of the device)
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// FIRST_DONE: warp 0 has already factorised sub-block 0 (overlapped with the last trailing update of the tile, see k_chol_dataflow) and a
// barrier of the compute warps has passed since
template <bool FIRST_DONE = false>
__device__ __forceinline__ void diag_factor(TilePtr D, TilePtr X, FacSmem &fs, int *not_spd) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rr = lane >> 2, cc = 2 * (lane & 3);
  // trailing update of one 8 x 8 block (rb, cb <= rb) of A after sub-block s, mirrored
  auto upd_a = [&](int o, int R0, int rb, int cb) {
    double a0 = 0.0, a1 = 0.0;
    const int r0 = R0 + 8 * rb, c0 = R0 + 8 * cb;
    blk_mma16([&](int r, int m) { return D[o + m][r0 + r]; }, [&](int m, int c) { return D[o + m][c0 + c]; }, a0, a1);
    const double v0 = D[c0 + cc][r0 + rr] - a0, v1 = D[c0 + cc + 1][r0 + rr] - a1;
    D[c0 + cc][r0 + rr] = v0;
    D[c0 + cc + 1][r0 + rr] = v1;
    if (rb != cb) {  // mirror (the diagonal sub-blocks must stay symmetric for step (a))
      D[r0 + rr][c0 + cc] = v0;
      D[r0 + rr][c0 + cc + 1] = v1;
    }
  };
  if (!FIRST_DONE) {
    CH_STAMP(ta0);
    if (warp == 0) sub_factor16(D, X, 0, fs, not_spd);
    CH_STAMP(ta1);
    CH_ACC(4, ta0, ta1);
    cbar();
  }
  for (int s = 0; s < NB / SB; s++) {
    const int o = SB * s, R0 = o + SB;
    CH_STAMP(tb0);
    // (b) units: strips of 8 rows below the block (L_is = A_is V^T, in place) and 8-column blocks left of it (X_s = V X_s, in place)
    {
      const int n_strip = (NB - R0) / 8, n_cb = o / 8;
      for (int u = warp; u < n_strip + n_cb; u += 8) {
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        if (u < n_strip) {
          const int r0 = R0 + 8 * u;
#pragma unroll
          for (int h = 0; h < 2; h++)
            blk_mma16([&](int r, int m) { return D[o + m][r0 + r]; }, [&](int m, int c) { return X[o + m][o + 8 * h + c]; }, acc[h][0], acc[h][1]);
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; h++) {
            D[o + 8 * h + cc][r0 + rr] = acc[h][0];
            D[o + 8 * h + cc + 1][r0 + rr] = acc[h][1];
          }
        } else {
          const int c0 = 8 * (u - n_strip);
#pragma unroll
          for (int h = 0; h < 2; h++)
            blk_mma16([&](int r, int m) { return X[o + m][o + 8 * h + r]; }, [&](int m, int c) { return X[c0 + c][o + m]; }, acc[h][0], acc[h][1]);
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; h++) {
            X[c0 + cc][o + 8 * h + rr] = acc[h][0];
            X[c0 + cc + 1][o + 8 * h + rr] = acc[h][1];
          }
        }
      }
    }
    cbar();
    CH_STAMP(tb1);
    CH_ACC(5, tb0, tb1);
    // (c) warp 0: the three blocks of the next diagonal sub-block, then straight on to factorise it (look-ahead);
    //     warps 1..7: the other lower blocks of the trailing part of A and the rows below of X for all columns up to block s
    {
      const int q = (NB - R0) / 8;  // block rows below
      const int n1 = q * (q + 1) / 2, ncx = R0 / 8, n2 = q * ncx;
      if (warp == 0) {
        if (q > 0) {  // the 16 x 16 block (R0.., R0..): three 8 x 8 blocks, their DMMA chains interleaved
          double a00[2] = {0.0, 0.0}, a10[2] = {0.0, 0.0}, a11[2] = {0.0, 0.0};
          const int kk = lane & 3;
#pragma unroll
          for (int ks = 0; ks < SB; ks += 4) {
            const double p0 = D[o + ks + kk][R0 + rr], p1 = D[o + ks + kk][R0 + 8 + rr];  // A fragment of row blocks 0, 1 == B fragment of column blocks 0, 1
            dmma(a00[0], a00[1], p0, p0);
            dmma(a10[0], a10[1], p1, p0);
            dmma(a11[0], a11[1], p1, p1);
          }
          {
            const double v0 = D[R0 + cc][R0 + rr] - a00[0], v1 = D[R0 + cc + 1][R0 + rr] - a00[1];
            const double w0 = D[R0 + cc][R0 + 8 + rr] - a10[0], w1 = D[R0 + cc + 1][R0 + 8 + rr] - a10[1];
            const double z0 = D[R0 + 8 + cc][R0 + 8 + rr] - a11[0], z1 = D[R0 + 8 + cc + 1][R0 + 8 + rr] - a11[1];
            __syncwarp();
            D[R0 + cc][R0 + rr] = v0, D[R0 + cc + 1][R0 + rr] = v1;
            D[R0 + cc][R0 + 8 + rr] = w0, D[R0 + cc + 1][R0 + 8 + rr] = w1;
            D[R0 + 8 + rr][R0 + cc] = w0, D[R0 + 8 + rr][R0 + cc + 1] = w1;  // mirror
            D[R0 + 8 + cc][R0 + 8 + rr] = z0, D[R0 + 8 + cc + 1][R0 + 8 + rr] = z1;
          }
          __syncwarp();
          CH_STAMP(tc1);
          CH_ACC(6, tb1, tc1);
          sub_factor16(D, X, s + 1, fs, not_spd);
          CH_STAMP(tc2);
          CH_ACC(4, tc1, tc2);
        }
      } else {
        for (int u = 3 + (warp - 1); u < n1 + n2; u += 7) {
          if (u < n1) {
            int rb = 0, rem = u;
            while (rem > rb) rem -= ++rb;  // u -> (rb, cb = rem), cb <= rb
            upd_a(o, R0, rb, rem);
          } else {
            double a0 = 0.0, a1 = 0.0;
            const int v = u - n1, rb = v / ncx, cb = v - rb * ncx;
            const int r0 = R0 + 8 * rb, c0 = 8 * cb;
            blk_mma16([&](int r, int m) { return D[o + m][r0 + r]; }, [&](int m, int c) { return X[c0 + c][o + m]; }, a0, a1);
            X[c0 + cc][r0 + rr] -= a0;
            X[c0 + cc + 1][r0 + rr] -= a1;
          }
        }
      }
    }
    cbar();
    CH_STAMP(tc3);
    CH_ACC(7, tb1, tc3);
  }
}

// ---- persistent factorisation ----------------------------------------------------------------------------------------------
// Every CTA has 8 compute warps and one COMMUNICATION warp (one active thread).  The compute warps only ever touch shared memory,
// the tensor cores and plain global stores; everything that waits on the rest of the device -- claiming queue entries, polling
// version counters, issuing the TMA loads, and the device-scope fence + release that publishes a finished tile (~3000 cycles) --
// is done by the communication thread.  The two sides talk through mbarriers: full[s] (operands of buffer set s have landed;
// completed by the TMA engine), done[s] (all compute warps have stored their part of the result and no longer read set s).
struct ChSmem {
  double T[6][NB][CLD];  // critical path: 0 diagonal tile, 1 W, 2 tile below the diagonal, 3 next diagonal tile; workers: two sets {A, B, C}
  unsigned long long full[2];  // workers: per buffer set; critical path: [0] first diagonal tile, [1] operands of T / U
  unsigned long long done[3];  // workers: per buffer set; critical path: [0] W stored, [1] P stored, [2] A / C free again
  FacSmem fs;
  int op[2][4];  // decoded queue entries {type, k, i, j}
  int abort;     // a wait timed out somewhere on the device
};
struct CholArgs {
  double *S;
  int Tm;     // tile rows of the allocation (ppo_dense.h)
  int n, Tc;  // current system: Tc column tiles; row tiles 0 .. Tc, the last one holds the carried gradient
  double *Winv;
  int *ver;
  CholCtrl *ctrl;
  int *not_spd;
};
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// communication thread: bounded wait for an mbarrier phase
__device__ __forceinline__ bool mbar_wait_bounded(unsigned long long *bar, unsigned parity, CholCtrl *ctrl) {
  if (mbar_test(bar, parity)) return true;
  const long long t0 = clock64();
  for (int it = 1;; it++) {
    if (mbar_test(bar, parity)) return true;
    if ((it & 1023) == 0) {
      if (*(volatile int *)&ctrl->err) return false;
      if (clock64() - t0 > CHOL_TIMEOUT) {
        atomicExch(&ctrl->err, 1);
        return false;
      }
    }
  }
}
constexpr int CH_THREADS = 288;  // 8 compute warps + the communication warp

__global__ void __launch_bounds__(CH_THREADS, 1) k_chol_dataflow(CholArgs a) {
  extern __shared__ __align__(16) unsigned char dsm[];
  ChSmem &sm = *reinterpret_cast<ChSmem *>(dsm);
  __shared__ int s_role, s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double *S = a.S;
  const int Tm = a.Tm, Tc = a.Tc, vs = a.Tm;
  CholCtrl *ctrl = a.ctrl;
  if (tid == 0) {
    s_role = atomicAdd(&ctrl->ticket, 1);
    s_base = ld_acquire(&ctrl->epoch) * 256;
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    mbar_init(&sm.done[0], 8);
    mbar_init(&sm.done[1], 8);
    mbar_init(&sm.done[2], 8);
    sm.abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int role = s_role, base = s_base;
  const bool comm = warp == 8;
  const int il = warp_row_block() * 8 + (lane >> 2);  // (compute warps) my row of a tile in the DMMA C-fragment layout; my columns: 8 q + 2 (lane & 3) + h
  const int jc = 2 * (lane & 3);
  auto tile = [&](int i, int j) { return S + dense_tile_index(Tm, i, j) * (size_t)TILE; };
  // one elected lane per compute warp tells the communication thread that the warp's stores are issued / its reads are over
  auto warp_arrive = [&](unsigned long long *bar) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
  };
  auto publish = [&](int *flag, int value) {  // communication thread: the tile behind `flag` is complete in global memory
    __threadfence();
    st_release(flag, value);
  };
  if (role == 0) {
    // =============================== critical path ===============================================================
    TilePtr D = sm.T[0], W = sm.T[1], A = sm.T[2], C = sm.T[3];
    if (comm) {
      if (lane == 0) {
        mbar_expect_tx(&sm.full[0], TILE_BYTES);
        bulk_g2s(D, tile(0, 0), TILE_BYTES, &sm.full[0]);
        unsigned pd[3] = {0, 0, 0};
        for (int k = 0; k < Tc; k++) {
          const bool more = k + 1 < Tc;
          // operands of T_k(k+1) and U_k(k+1,k+1): as soon as their producers are done and the buffers are free again
          bool ok = k == 0 || mbar_wait_bounded(&sm.done[2], pd[2], ctrl);
          if (k > 0) pd[2] ^= 1;
          if (ok && k > 0) ok = flag_wait(&a.ver[(k + 1) * vs + k], base + k, ctrl) && (!more || flag_wait(&a.ver[(k + 1) * vs + k + 1], base + k, ctrl));
          if (ok) {
            fence_proxy_async();
            mbar_expect_tx(&sm.full[1], (more ? 2 : 1) * TILE_BYTES);
            bulk_g2s(A, tile(k + 1, k), TILE_BYTES, &sm.full[1]);
            if (more) bulk_g2s(C, tile(k + 1, k + 1), TILE_BYTES, &sm.full[1]);
          }
          // W_k, then P_{k+1,k}: published the moment the compute warps have stored them
          ok = ok && mbar_wait_bounded(&sm.done[0], pd[0], ctrl);
          pd[0] ^= 1;
          if (ok) publish(&a.ver[k * vs + k], base + k + 1);
          ok = ok && mbar_wait_bounded(&sm.done[1], pd[1], ctrl);
          pd[1] ^= 1;
          if (ok) publish(&a.ver[(k + 1) * vs + k], base + k + 1);
          if (!ok) {  // timed out (or another CTA did): release the compute warps, which may sit in an mbarrier wait
            sm.abort = 1;
            mbar_arrive(&sm.full[1]);
            break;
          }
        }
      }
    } else {
      unsigned pf1 = 0;
      mbar_wait(&sm.full[0], 0);
      {  // symmetric fill + identity padding of the first diagonal tile
        const int nb = min(NB, a.n);
        cbar();
        for (int e = tid; e < NB * NB; e += 256) {
          const int c = e >> 6, r = e & 63;
          if (r >= c) {
            const double v = (r < nb && c < nb) ? D[c][r] : (r == c ? 1.0 : 0.0);
            D[c][r] = v;
            D[r][c] = v;
          }
        }
        for (int e = tid; e < TILE; e += 256) (&W[0][0])[e] = 0.0;
        cbar();
      }
      for (int k = 0; k < Tc; k++) {
        CH_STAMP(t0);
        if (k == 0) diag_factor<false>(D, W, sm.fs, a.not_spd);  // ends with a barrier of the compute warps
        else diag_factor<true>(D, W, sm.fs, a.not_spd);         // (sub-block 0 was factorised next to the update that produced the tile)
        CH_STAMP(t1);
        {  // W_k to global memory (tile layout, so that workers fetch it with one bulk copy)
          double *Wg = a.Winv + (size_t)k * TILE;
          for (int e = tid; e < TILE; e += 256) Wg[e] = (&W[0][0])[e];
        }
        warp_arrive(&sm.done[0]);
        mbar_wait(&sm.full[1], pf1);
        pf1 ^= 1;
        if (*(volatile int *)&sm.abort) break;
        CH_STAMP(t2);
        // T_k(k+1): the tile right below the diagonal
        double acc[8][2];
#pragma unroll
        for (int q = 0; q < 8; q++) acc[q][0] = acc[q][1] = 0.0;
        const bool grad_tile = (k + 1 == Tc);  // the gradient tile has one valid row
        if (!grad_tile || warp == 0) tile_mma64<2>(A, W, acc);
        cbar();  // everybody has read A and W
        {
          double *dst = tile(k + 1, k);
#pragma unroll
          for (int q = 0; q < 8; q++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int jl = q * 8 + jc + h;
              dst[jl * CLD + il] = acc[q][h];
              A[jl][il] = acc[q][h];
            }
          for (int e = tid; e < TILE; e += 256) (&W[0][0])[e] = 0.0;  // cleared for the next panel
        }
        warp_arrive(&sm.done[1]);
        CH_STAMP(t3);
        if (k + 1 >= Tc) {
          CH_ACC(0, t0, t1);
          CH_ACC(1, t1, t2);
          CH_ACC(2, t2, t3);
          break;
        }
        cbar();  // P is complete in shared memory
        // U_k(k+1,k+1) on the next diagonal tile, which then stays in shared memory for F_{k+1}
        {
          // Look-ahead: warp 0 updates the first 16 x 16 sub-block alone (3 of the 36 lower 8 x 8 blocks) and goes straight on to
          // factorise it -- the serial part of F_{k+1} -- while warps 1..7 update the other 33 blocks.
          const int nb2 = min(NB, a.n - NB * (k + 1));
          const int rr = lane >> 2, kk = lane & 3;
          auto upd_block = [&](int rb, int cb) {  // D(rb, cb) = C - sum_m A[m][rb..] A[m][cb..]  (+ mirror), identity beyond the valid size
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int m0 = 0; m0 < NB; m0 += 4) dmma(c0, c1, A[m0 + kk][8 * rb + rr], A[m0 + kk][8 * cb + rr]);
            const int r = 8 * rb + rr;
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int c = 8 * cb + jc + h;
              if (r >= c) {
                const double v = (r < nb2 && c < nb2) ? C[c][r] - (h ? c1 : c0) : (r == c ? 1.0 : 0.0);
                D[c][r] = v;
                D[r][c] = v;
              }
            }
          };
          if (warp == 0) {
            upd_block(0, 0);
            upd_block(1, 0);
            upd_block(1, 1);
            __syncwarp();
            sub_factor16(D, W, 0, sm.fs, a.not_spd);
          } else {
            for (int u = warp - 1; u < 33; u += 7) {  // row blocks 2 .. 7, column blocks 0 .. rb
              int rb = 2, rem = u;
              while (rem > rb) rem -= ++rb;
              upd_block(rb, rem);
            }
          }
          (void)acc;
        }
        warp_arrive(&sm.done[2]);  // A and C may be refilled
        cbar();
        CH_STAMP(t4);
        CH_ACC(0, t0, t1);
        CH_ACC(1, t1, t2);
        CH_ACC(2, t2, t3);
        CH_ACC(3, t3, t4);
      }
    }
  } else if (comm) {
    // =============================== workers, communication thread ===============================================
    if (lane == 0) {
      int lvl = 0, lvl_start = 0;  // level of the last decoded entry and index of its first entry
      auto claim = [&](int slot) -> bool {  // next queue entry -> sm.op[slot]; false: the queue is exhausted
        const int idx = atomicAdd(&ctrl->qhead, 1);
        int type = -1, k = 0, i = 0, j = 0;
        for (; lvl < Tc; lvl++) {
          const int m = Tc - 1 - lvl;
          const int size = m + (m >= 1 ? (m + 1) * (m + 2) / 2 - 2 : 0);
          if (idx < lvl_start + size) break;
          lvl_start += size;
        }
        if (lvl < Tc) {
          k = lvl;
          const int m = Tc - 1 - k;
          int u = idx - lvl_start;
          if (u < m) {  // T_k(i), i = k+2 .. Tc
            type = 0, i = k + 2 + u, j = k;
          } else {      // U_k(i, j): column k+1 rows k+2..Tc, then columns j >= k+2 rows j..Tc
            type = 1;
            u -= m;
            if (u < m) {
              i = k + 2 + u, j = k + 1;
            } else {
              u -= m;
              for (j = k + 2;; j++) {
                const int cnt = Tc - j + 1;
                if (u < cnt) break;
                u -= cnt;
              }
              i = j + u;
            }
          }
        }
        sm.op[slot][0] = type, sm.op[slot][1] = k, sm.op[slot][2] = i, sm.op[slot][3] = j;
        return type >= 0;
      };
      auto deps_ready = [&](int slot) -> bool {  // are the producers of the entry done?
        const int type = sm.op[slot][0], k = sm.op[slot][1], i = sm.op[slot][2], j = sm.op[slot][3];
        if (type == 0) return flag_ready(&a.ver[k * vs + k], base + k + 1) && (k == 0 || flag_ready(&a.ver[i * vs + k], base + k));
        return flag_ready(&a.ver[i * vs + k], base + k + 1) && (j == i || flag_ready(&a.ver[j * vs + k], base + k + 1)) &&
               (k == 0 || flag_ready(&a.ver[i * vs + j], base + k));
      };
      auto issue = [&](int slot) {  // TMA loads of the entry's operands into buffer set `slot`
        const int type = sm.op[slot][0], k = sm.op[slot][1], i = sm.op[slot][2], j = sm.op[slot][3];
        fence_proxy_async();
        if (type == 0) {
          mbar_expect_tx(&sm.full[slot], 2 * TILE_BYTES);
          bulk_g2s(sm.T[3 * slot], tile(i, k), TILE_BYTES, &sm.full[slot]);
          bulk_g2s(sm.T[3 * slot + 1], a.Winv + (size_t)k * TILE, TILE_BYTES, &sm.full[slot]);
        } else {
          const bool diag = i == j;
          mbar_expect_tx(&sm.full[slot], (diag ? 2 : 3) * TILE_BYTES);
          bulk_g2s(sm.T[3 * slot], tile(i, k), TILE_BYTES, &sm.full[slot]);
          if (!diag) bulk_g2s(sm.T[3 * slot + 1], tile(j, k), TILE_BYTES, &sm.full[slot]);
          bulk_g2s(sm.T[3 * slot + 2], tile(i, j), TILE_BYTES, &sm.full[slot]);
        }
      };
      // ring of two buffer sets: `head` is filled next, `tail` retired next; the descriptor of an entry in flight stays in sm.op[slot]
      int head = 0, tail = 0, inflight = 0;
      bool claimed = false, exhausted = false;
      unsigned pd[2] = {0, 0};
      int rel_i[2] = {0, 0}, rel_j[2] = {0, 0}, rel_v[2] = {0, 0};
      long long t_idle = clock64();
      for (;;) {
        bool progress = false;
        if (inflight > 0 && mbar_test(&sm.done[tail], pd[tail])) {  // the oldest entry is computed and stored: publish its tile
          pd[tail] ^= 1;
          publish(&a.ver[rel_i[tail] * vs + rel_j[tail]], rel_v[tail]);
          tail ^= 1;
          inflight--;
          progress = true;
        }
        if (!claimed && !exhausted && inflight < 2) {
          if (claim(head)) claimed = true;
          else exhausted = true;
          progress = true;
        }
        if (claimed && deps_ready(head)) {
          rel_i[head] = sm.op[head][2], rel_j[head] = sm.op[head][3], rel_v[head] = base + sm.op[head][1] + 1;
          issue(head);
          head ^= 1;
          inflight++;
          claimed = false;
          progress = true;
        }
        if (exhausted && inflight == 0) {  // (sm.op[head][0] is -1: written by the failed claim) wake the compute warps for the last time
          mbar_arrive(&sm.full[head]);
          break;
        }
        if (progress) {
          t_idle = clock64();
        } else {
          __nanosleep(20);
          if (*(volatile int *)&ctrl->err || clock64() - t_idle > CHOL_TIMEOUT) {
            atomicExch(&ctrl->err, 1);
            sm.abort = 1;
            sm.op[head][0] = -1, sm.op[head ^ 1][0] = -1;
            mbar_arrive(&sm.full[0]);
            mbar_arrive(&sm.full[1]);
            break;
          }
        }
      }
    }
  } else {
    // =============================== workers, compute warps ======================================================
    unsigned pf[2] = {0, 0};
    for (int cur = 0;; cur ^= 1) {
      mbar_wait(&sm.full[cur], pf[cur]);
      pf[cur] ^= 1;
      const int type = sm.op[cur][0], k = sm.op[cur][1], i = sm.op[cur][2], j = sm.op[cur][3];
      if (type < 0 || *(volatile int *)&sm.abort) break;
      (void)k;
      TilePtr A = sm.T[3 * cur], B = sm.T[3 * cur + 1], C = sm.T[3 * cur + 2];
      const bool one_row = (i == Tc);  // gradient tile: only row 0 carries data (the rest is zero padding)
      if (!one_row || warp == 0) {
        double acc[8][2];
#pragma unroll
        for (int q = 0; q < 8; q++) acc[q][0] = acc[q][1] = 0.0;
        double *dst = tile(i, j);
        if (type == 0) {
          tile_mma64<2>(A, B, acc);
#pragma unroll
          for (int q = 0; q < 8; q++)
#pragma unroll
            for (int h = 0; h < 2; h++) dst[(q * 8 + jc + h) * CLD + il] = acc[q][h];
        } else {
          const bool diag = (i == j);
          if (diag) tile_mma64<1>(A, A, acc);
          else tile_mma64<0>(A, B, acc);
#pragma unroll
          for (int q = 0; q < 8; q++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int jl = q * 8 + jc + h;
              if (!diag || il >= jl) dst[jl * CLD + il] = C[jl][il] - acc[q][h];
            }
        }
      }
      warp_arrive(&sm.done[cur]);  // my stores are issued, my reads of this buffer set are over
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&ctrl->done, 1) == (int)gridDim.x - 1) {  // last CTA out: arm the control block for the next launch
      if (*(volatile int *)&ctrl->err) *a.not_spd = 1;      // a timed-out factorisation counts as a failed solve
      ctrl->ticket = 0;
      ctrl->qhead = 0;
      ctrl->done = 0;
      ctrl->epoch = ctrl->epoch + 1;
      __threadfence();
    }
  }
}

// ---- backward substitution L^T x = y in ONE launch --------------------------------------------------------------------
// CTA (by start ticket r) owns block column b = Tc-1-r: it folds x_c of the later blocks into its right-hand side as they
// are published ( acc -= L_cb^T x_c, tiles prefetched by TMA one ahead ) and finishes with x_b = W_b^T acc.  A CTA only waits
// for CTAs that started before it.  x_b travels between CTAs as 8-byte words {32 data bits, 32-bit launch tag} (the "LL"
// protocol of collective libraries): data and flag arrive in one atomic store, so the consumer needs no fence and no second
// round trip.  Chain per block: tagged words -> 64 x 64 mat-vec -> warp reduction -> 64 x 64 mat-vec -> tagged words.
struct BsSmem {
  double L[2][NB][CLD];
  double W[NB][CLD + 1];  // odd stride: conflict-free column reads
  double xc[NB];
  double accv[NB];
  double part[4][NB];
  unsigned long long mbar[2];
  int ok;
};
// DIST: the factor was assembled from tiles pushed by the other ranks (k_chol_dist); every tile is read only after its version
// counter `ver` shows the final value base + column + 1.
template <bool DIST>
__global__ void __launch_bounds__(256) k_backsolve_chain(const double *S, int Tm, int n, int Tc, const double *Winv, double *x, unsigned long long *xll, CholCtrl *ctrl,
                                                         const int *ver, int base, int *not_spd) {
  extern __shared__ __align__(16) unsigned char dsm[];
  BsSmem &sm = *reinterpret_cast<BsSmem *>(dsm);
  __shared__ int s_b, s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    s_b = Tc - 1 - atomicAdd(&ctrl->bticket, 1);
    s_base = ld_acquire(&ctrl->bepoch);
    mbar_init(&sm.mbar[0], 1);
    mbar_init(&sm.mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int b = s_b;
  const unsigned want = (unsigned)s_base;  // launch tag (never 0: the word buffer starts zeroed, the launch counter at 1)
  auto tile = [&](int i, int j) { return S + dense_tile_index(Tm, i, j) * (size_t)TILE; };
  if (b >= 0) {
    unsigned ph[2] = {0, 0};
    int buf = 0;
    auto tile_final = [&](int i, int j) {  // (thread 0) tile (i, j) of the factor has arrived
      if (DIST) {
        flag_wait_sys(&ver[i * Tm + j], base + j + 1, ctrl);
        fence_proxy_async();
      }
    };
    if (Tc - 1 > b && tid == 0) {  // first tile of the sweep: (Tc-1, b)
      tile_final(Tc - 1, b);
      mbar_expect_tx(&sm.mbar[0], TILE_BYTES);
      bulk_g2s(sm.L[0], tile(Tc - 1, b), TILE_BYTES, &sm.mbar[0]);
    }
    if (DIST) {  // W_b and the carried gradient tile
      if (tid == 0) {
        flag_wait_sys(&ver[b * Tm + b], base + b + 1, ctrl);
        flag_wait_sys(&ver[Tc * Tm + b], base + b + 1, ctrl);
      }
      __syncthreads();
    }
    {
      const double *W = Winv + (size_t)b * TILE;
      for (int e = tid; e < NB * NB; e += 256) sm.W[e >> 6][e & 63] = __ldcg(W + (e >> 6) * CLD + (e & 63));
    }
    // my slice of y (the carried gradient row, final since the factorisation ended): requested now, consumed after the sweep
    double yv = 0.0;
    if (lane < 8) {
      const int cl = 8 * warp + lane;
      yv = NB * b + cl < n ? __ldcg(tile(Tc, b) + cl * CLD) : 0.0;
    }
#ifdef PPO_CHOL_TIMING
    long long t_detect = clock64();
#endif
    double s[8];  // lane-partial sums of  sum_r L(r, 8 warp + jj) x_c[r]  over all tiles so far
#pragma unroll
    for (int jj = 0; jj < 8; jj++) s[jj] = 0.0;
    bool ok = true;
    if (tid == 0) sm.ok = 1;
    for (int c = Tc - 1; c > b; c--) {
      if (tid == 0 && c - 1 > b) {  // prefetch the next tile into the other buffer (its last readers are behind the barrier that ended the previous round)
        tile_final(c - 1, b);
        mbar_expect_tx(&sm.mbar[buf ^ 1], TILE_BYTES);
        bulk_g2s(sm.L[buf ^ 1], tile(c - 1, b), TILE_BYTES, &sm.mbar[buf ^ 1]);
      }
      if (tid < 2 * NB) {  // 128 threads poll the 128 tagged words of x_c
        const volatile unsigned long long *p = xll + (size_t)c * 2 * NB + tid;
        unsigned long long w = *p;
        if ((unsigned)(w >> 32) != want) {
          const long long t0 = clock64();
          for (int it = 1; (unsigned)(w >> 32) != want; it++) {
            w = *p;
            if ((it & 255) == 0 && (*(volatile int *)&ctrl->err || clock64() - t0 > CHOL_TIMEOUT)) {
              atomicExch(&ctrl->err, 1);
              sm.ok = 0;
              break;
            }
          }
        }
        const unsigned hi = __shfl_down_sync(0xffffffffu, (unsigned)w, 1);
        if (!(tid & 1)) sm.xc[tid >> 1] = __hiloint2double((int)hi, (int)(unsigned)w);
      }
#ifdef PPO_CHOL_TIMING
      if (c == b + 1) t_detect = clock64();
#endif
      __syncthreads();
      if (!sm.ok) {
        ok = false;
        break;
      }
      mbar_wait(&sm.mbar[buf], ph[buf]);
      ph[buf] ^= 1;
      const double x0 = sm.xc[lane], x1 = sm.xc[lane + 32];
#pragma unroll
      for (int jj = 0; jj < 8; jj++) s[jj] = fma(sm.L[buf][8 * warp + jj][lane], x0, fma(sm.L[buf][8 * warp + jj][lane + 32], x1, s[jj]));
      buf ^= 1;
      __syncthreads();  // the buffer may be refilled by the next prefetch
    }
    if (ok) {
#pragma unroll
      for (int jj = 0; jj < 8; jj++) {
        double v = s[jj];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const double yj = __shfl_sync(0xffffffffu, yv, jj);
        if (lane == 0) sm.accv[8 * warp + jj] = yj - v;
      }
      __syncthreads();
      {  // x_b = W^T acc  (W lower triangular: rows r >= i)
        const int i = tid & 63, p = tid >> 6;
        double t = 0.0;
        for (int r = i + p; r < NB; r += 4) t = fma(sm.W[i][r], sm.accv[r], t);
        sm.part[p][i] = t;
      }
      __syncthreads();
      if (tid < NB) {
        double v = sm.part[0][tid] + sm.part[1][tid] + sm.part[2][tid] + sm.part[3][tid];
        v = (NB * b + tid < n) ? v : 0.0;
        volatile unsigned long long *p = xll + (size_t)b * 2 * NB + 2 * tid;
        p[0] = ((unsigned long long)want << 32) | (unsigned)__double2loint(v);
        p[1] = ((unsigned long long)want << 32) | (unsigned)__double2hiint(v);
        x[NB * b + tid] = v;
      }
#ifdef PPO_CHOL_TIMING
      if (tid == 0 && b < Tc - 1) {
        atomicAdd((unsigned long long *)&g_chol_t[8], (unsigned long long)(clock64() - t_detect));
        atomicAdd((unsigned long long *)&g_chol_t[9], 1ull);
      }
#endif
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&ctrl->bdone, 1) == (int)gridDim.x - 1) {
      if (*(volatile int *)&ctrl->err) {  // a wait of this solve timed out: the solve counts as failed; the next one starts clean
        *not_spd = 1;
        ctrl->err = 0;
      }
      ctrl->bticket = 0;
      ctrl->bdone = 0;
      ctrl->bepoch = ctrl->bepoch + 1;
      __threadfence();
    }
  }
}

// ---- one window on several GPUs: distributed factorisation over NVLink peer memory ----------------------------------------
// Tile column j of the reduced system belongs to rank j mod world (1-D block-cyclic).  Every rank holds a full-size copy of S in
// peer-mapped memory (cudaIpc); only the owner's copy of a column is ever updated, the other copies receive the FINAL tiles.
//   k_dist_signal / k_dist_reduce   the partial Schur complements of the ranks (each rank summed its own landmarks into its own
//                                   S) are pulled over NVLink and added up by the owner of each column, in rank order
//                                   (a reduce-scatter whose chunks are the block-cyclic columns; no staging, no NCCL);
//   k_chol_dist                     the same dataflow factorisation as k_chol_dataflow, every rank executing the operations that
//                                   write ITS columns: F_k, T_k(i) on owner(k); U_k(i, j) on owner(j).  The finished panel tiles
//                                   P_ik = T_k(i) and W_k are staged in shared memory and PUSHED to all ranks by TMA bulk stores
//                                   (cp.async.bulk shared -> peer global), followed by a system-scope release of the tile's
//                                   version counter on every rank: consumers poll their LOCAL counter and fetch the tile from
//                                   their LOCAL copy.  The critical path F_k -> T_k(k+1) -> [NVLink] -> U_k(k+1,k+1) -> F_k+1
//                                   hops from rank to rank; the next owner is served first.
//   k_backsolve_chain<true>         every rank ends up with the whole factor and runs the backward substitution redundantly
//                                   (0.2 ms at n = 7794); it waits for the version counter of every tile it reads, which also
//                                   guarantees that no push into this rank's memory is still in flight when the solve returns.
__device__ __forceinline__ double2 ld_peer_v2(const double2 *p) {  // never from L1: the peer rewrites the location every solve
  double2 v;
  asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct CholDistArgs {
  CholArgs a;           // this rank's pointers
  DistPeers p;
  const unsigned *ops;  // this rank's worker queue: type << 24 | k << 16 | i << 8 | j, level-major (topological) order
  int n_ops;
  int base;             // version-counter tag of this solve (the same on every rank)
};

__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void st_relaxed_sys(int *p, int v) { asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// communication thread: the tile staged at `src` goes to offset `off` of the copies of ranks q with (mask >> q) & 1 (rank `first`, if any,
// is served first and sees its counter before the other copies are waited for), then counter `flag` of those ranks is released with
// `value`: completion of the bulk stores (wait_group) -> proxy fence -> ONE system-scope release fence -> relaxed counter stores.
__device__ __forceinline__ void dist_push_publish(const DistPeers &p, double *const *bases, size_t off, const void *src, int flag, int value, int first, unsigned mask) {
  if (first >= 0) {  // alone on the wire until its counter is out: a system-scope fence drains every store this thread has in flight
    bulk_s2g(bases[first] + off, src, TILE_BYTES);
    bulk_commit();
    bulk_wait<0>();
    fence_proxy_async();
    fence_acq_rel_sys();
    st_relaxed_sys(p.ver[first] + flag, value);
  }
  const int q0 = first >= 0 ? first + 1 : 0;  // the other copies in the order in which their ranks own the next columns
  bool any = false;
  for (int t = 0; t < p.world; t++) {
    const int q = (q0 + t) % p.world;
    if (q != first && ((mask >> q) & 1)) bulk_s2g(bases[q] + off, src, TILE_BYTES), any = true;
  }
  if (!any) return;
  bulk_commit();
  bulk_wait<0>();
  fence_proxy_async();
  fence_acq_rel_sys();
  for (int q = 0; q < p.world; q++)
    if (q != first && ((mask >> q) & 1)) st_relaxed_sys(p.ver[q] + flag, value);
}

__global__ void __launch_bounds__(CH_THREADS, 1) k_chol_dist(CholDistArgs d) {
  extern __shared__ __align__(16) unsigned char dsm[];
  ChSmem &sm = *reinterpret_cast<ChSmem *>(dsm);
  __shared__ int s_role;
  const CholArgs &a = d.a;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double *S = a.S;
  const int Tm = a.Tm, Tc = a.Tc, vs = a.Tm, base = d.base;
  const int R = d.p.rank, Wd = d.p.world;
  const unsigned all_ranks = (1u << Wd) - 1;
  CholCtrl *ctrl = a.ctrl;
  if (tid == 0) {
    s_role = atomicAdd(&ctrl->ticket, 1);
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    mbar_init(&sm.done[0], 8);
    mbar_init(&sm.done[1], 8);
    mbar_init(&sm.done[2], 8);
    sm.abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int role = s_role;
  const bool comm = warp == 8;
  const int il = warp_row_block() * 8 + (lane >> 2);
  const int jc = 2 * (lane & 3);
  auto tile_off = [&](int i, int j) { return dense_tile_index(Tm, i, j) * (size_t)TILE; };
  auto tile = [&](int i, int j) { return S + tile_off(i, j); };
  auto warp_arrive = [&](unsigned long long *bar) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
  };
  const int Bc = d.p.blk;  // columns per ownership block: column j belongs to rank (j / Bc) mod world
  auto own = [&](int j) { return (j / Bc) % Wd; };
  if (role == 0) {
    // =============================== critical path of my columns ================================================================
    // Inside a block of Bc consecutive columns the critical path stays on this GPU (the panel tile just computed is still in shared
    // memory, like in k_chol_dataflow); it crosses NVLink once per block.
    TilePtr D = sm.T[0], A = sm.T[2], C = sm.T[3], Pp = sm.T[4];
    if (comm) {
      if (lane == 0) {
        unsigned pd0 = 0, pd1 = 0, pd2 = 0;
        bool ok = true, early = false;
        int it = 0;
        for (int k = 0; k < Tc; k++) {
          if (own(k) != R) continue;
          TilePtr W = (it++ & 1) ? sm.T[5] : sm.T[1];
          const bool local_prev = k > 0 && own(k - 1) == R;
          // operands of U_{k-1}(k,k): P_{k,k-1} (still in A if column k-1 is mine, else pushed by its owner) and my tile (k,k) after U_0 .. U_{k-2}
          if (!early) {
            if (!local_prev && k > 0) ok = flag_wait_sys(&a.ver[k * vs + k - 1], base + k, ctrl);
            if (ok && k >= 2) ok = flag_wait_sys(&a.ver[k * vs + k], base + k - 1, ctrl);
            if (ok) {
              const bool need_p = !local_prev && k > 0;
              fence_proxy_async();
              mbar_expect_tx(&sm.full[0], (need_p ? 2 : 1) * TILE_BYTES);
              if (need_p) bulk_g2s(Pp, tile(k, k - 1), TILE_BYTES, &sm.full[0]);
              bulk_g2s(C, tile(k, k), TILE_BYTES, &sm.full[0]);
            }
          }
          early = false;
          // operand of T_k(k+1): my tile (k+1,k) after U_0 .. U_{k-1}; buffer A is free once the diagonal update has read it
          ok = ok && mbar_wait_bounded(&sm.done[2], pd2, ctrl);
          pd2 ^= 1;
          if (ok && k > 0) ok = flag_wait_sys(&a.ver[(k + 1) * vs + k], base + k, ctrl);
          if (ok) {
            fence_proxy_async();
            mbar_expect_tx(&sm.full[1], TILE_BYTES);
            bulk_g2s(A, tile(k + 1, k), TILE_BYTES, &sm.full[1]);
          }
          // W_k: to my own copy at once (my workers' panel operations wait for it); the other ranks only need it for the backward
          // substitution, so their copies go out last
          ok = ok && mbar_wait_bounded(&sm.done[0], pd0, ctrl);
          pd0 ^= 1;
          if (ok) dist_push_publish(d.p, d.p.Winv, (size_t)k * TILE, W, k * vs + k, base + k + 1, -1, 1u << R);
          ok = ok && mbar_wait_bounded(&sm.done[1], pd1, ctrl);
          pd1 ^= 1;
          if (ok) {
            const bool next_local = k + 1 < Tc && own(k + 1) == R;
            if (next_local) {  // the critical path continues here: get the next diagonal tile on its way before anything is pushed
              if (k + 1 >= 2) ok = flag_wait_sys(&a.ver[(k + 1) * vs + k + 1], base + k, ctrl);
              if (ok) {
                fence_proxy_async();
                mbar_expect_tx(&sm.full[0], TILE_BYTES);
                bulk_g2s(C, tile(k + 1, k + 1), TILE_BYTES, &sm.full[0]);
                early = true;
              }
            }
            if (ok) {
              // the panel tile P_{k+1,k}: the next owner's critical path waits for it (served first), everybody needs it later
              dist_push_publish(d.p, d.p.S, tile_off(k + 1, k), A, (k + 1) * vs + k, base + k + 1, (next_local || k + 1 >= Tc) ? -1 : own(k + 1), all_ranks);
              dist_push_publish(d.p, d.p.Winv, (size_t)k * TILE, W, k * vs + k, base + k + 1, -1, all_ranks & ~(1u << R));
            }
          }
          if (!ok) {  // timed out (here or elsewhere): release the compute warps from whichever wait they sit in
            atomicExch(&ctrl->err, 1);
            sm.abort = 1;
            mbar_arrive(&sm.full[0]);
            mbar_arrive(&sm.full[1]);
            break;
          }
        }
      }
    } else {
      unsigned pf0 = 0, pf1 = 0;
      int it = 0;
      for (int k = 0; k < Tc; k++) {
        if (own(k) != R) continue;
        TilePtr W = (it++ & 1) ? sm.T[5] : sm.T[1];
        const bool local_prev = k > 0 && own(k - 1) == R;
        TilePtr P = local_prev ? A : Pp;
        mbar_wait(&sm.full[0], pf0);
        pf0 ^= 1;
        if (*(volatile int *)&sm.abort) break;
        for (int e = tid; e < TILE; e += 256) (&W[0][0])[e] = 0.0;
        {  // D = C - P P^T (last update of the diagonal tile), symmetric, identity beyond the valid size; as in k_chol_dataflow warp 0
           // updates the first 16 x 16 sub-block alone and factorises it while the other warps update the rest
          const int nb = min(NB, a.n - NB * k);
          const int rr = lane >> 2, kk = lane & 3;
          auto upd_block = [&](int rb, int cb) {
            double c0 = 0.0, c1 = 0.0;
            if (k > 0) {
#pragma unroll
              for (int m0 = 0; m0 < NB; m0 += 4) dmma(c0, c1, P[m0 + kk][8 * rb + rr], P[m0 + kk][8 * cb + rr]);
            }
            const int r = 8 * rb + rr;
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int c = 8 * cb + jc + h;
              if (r >= c) {
                const double v = (r < nb && c < nb) ? C[c][r] - (h ? c1 : c0) : (r == c ? 1.0 : 0.0);
                D[c][r] = v;
                D[r][c] = v;
              }
            }
          };
          cbar();  // W is zero everywhere before warp 0 writes into it
          if (warp == 0) {
            upd_block(0, 0);
            upd_block(1, 0);
            upd_block(1, 1);
            __syncwarp();
            sub_factor16(D, W, 0, sm.fs, a.not_spd);
          } else {
            for (int u = warp - 1; u < 33; u += 7) {
              int rb = 2, rem = u;
              while (rem > rb) rem -= ++rb;
              upd_block(rb, rem);
            }
          }
        }
        warp_arrive(&sm.done[2]);  // A / Pp / C are read: A may be refilled
        cbar();
        diag_factor<true>(D, W, sm.fs, a.not_spd);  // ends with a barrier of the compute warps
        fence_proxy_async_smem();             // W is read by the TMA stores of the communication thread
        warp_arrive(&sm.done[0]);
        mbar_wait(&sm.full[1], pf1);
        pf1 ^= 1;
        if (*(volatile int *)&sm.abort) break;
        {  // T_k(k+1) in place: a warp reads and writes only its own 8 rows of A
          double acc[8][2];
#pragma unroll
          for (int q = 0; q < 8; q++) acc[q][0] = acc[q][1] = 0.0;
          const bool grad_tile = (k + 1 == Tc);
          if (!grad_tile || warp == 0) {
            tile_mma64<2>(A, W, acc);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; q++)
#pragma unroll
              for (int h = 0; h < 2; h++) A[q * 8 + jc + h][il] = acc[q][h];
          }
        }
        fence_proxy_async_smem();
        warp_arrive(&sm.done[1]);
      }
    }
  } else if (comm) {
    // =============================== workers, communication thread ===============================================
    if (lane == 0) {
      auto claim = [&](int slot) -> bool {
        const int idx = atomicAdd(&ctrl->qhead, 1);
        if (idx >= d.n_ops) {
          sm.op[slot][0] = -1;
          return false;
        }
        const unsigned o = d.ops[idx];
        sm.op[slot][0] = (int)(o >> 24), sm.op[slot][1] = (int)((o >> 16) & 255), sm.op[slot][2] = (int)((o >> 8) & 255), sm.op[slot][3] = (int)(o & 255);
        return true;
      };
      auto ready = [&](int i, int j, int want) { return ld_acquire_sys(&a.ver[i * vs + j]) >= want; };
      auto deps_ready = [&](int slot) -> bool {
        const int type = sm.op[slot][0], k = sm.op[slot][1], i = sm.op[slot][2], j = sm.op[slot][3];
        if (type == 0) return ready(k, k, base + k + 1) && (k == 0 || ready(i, k, base + k));
        if (type == 2) return ready(i, k, base + k + 1);  // the panel tile I forward has arrived from its producer
        return ready(i, k, base + k + 1) && (j == i || ready(j, k, base + k + 1)) && (k == 0 || ready(i, j, base + k));
      };
      auto issue = [&](int slot) {
        const int type = sm.op[slot][0], k = sm.op[slot][1], i = sm.op[slot][2], j = sm.op[slot][3];
        fence_proxy_async();
        if (type == 0) {
          mbar_expect_tx(&sm.full[slot], 2 * TILE_BYTES);
          bulk_g2s(sm.T[3 * slot], tile(i, k), TILE_BYTES, &sm.full[slot]);
          bulk_g2s(sm.T[3 * slot + 1], a.Winv + (size_t)k * TILE, TILE_BYTES, &sm.full[slot]);
        } else if (type == 2) {
          mbar_expect_tx(&sm.full[slot], TILE_BYTES);
          bulk_g2s(sm.T[3 * slot], tile(i, k), TILE_BYTES, &sm.full[slot]);
        } else {
          const bool diag = i == j;
          mbar_expect_tx(&sm.full[slot], (diag ? 2 : 3) * TILE_BYTES);
          bulk_g2s(sm.T[3 * slot], tile(i, k), TILE_BYTES, &sm.full[slot]);
          if (!diag) bulk_g2s(sm.T[3 * slot + 1], tile(j, k), TILE_BYTES, &sm.full[slot]);
          bulk_g2s(sm.T[3 * slot + 2], tile(i, j), TILE_BYTES, &sm.full[slot]);
        }
      };
      int head = 0, tail = 0, inflight = 0;
      bool claimed = false, exhausted = false;
      unsigned pd[2] = {0, 0};
      int rel_t[2] = {0, 0}, rel_i[2] = {0, 0}, rel_j[2] = {0, 0}, rel_v[2] = {0, 0};
      long long t_idle = clock64();
      for (;;) {
        bool progress = false;
        if (inflight > 0 && mbar_test(&sm.done[tail], pd[tail])) {
          pd[tail] ^= 1;
          if (rel_t[tail] == 0) {  // a panel tile: staged in the first buffer of the set -> every rank
            // (all panel tiles of column k become ready together, right after W_k: the owner of column k+1, whose critical path needs
            // the first of them next, gets its copies ahead of the rest of the burst)
            // With forwarding (d.p.fwd, PPO_DIST_FORWARD=1; off by default) a panel tile of tile row i leaves its producer ONCE, for the
            // owner of column i -- the rank that needs it first (its updates of column i all read it) --, which passes it on to the other
            // ranks (operation type 2 of its queue): every rank's links carry a share of the panel instead of the producer's links carrying
            // world - 1 copies of it.  Measured on 8 GPUs at n = 7794: 4.00 ms without, 4.61 ms with -- the broadcast is not what bounds the
            // solve, the second hop costs more than the spread egress gains.
            const int f = own(rel_i[tail]);
            if (d.p.fwd && f != R)
              dist_push_publish(d.p, d.p.S, tile_off(rel_i[tail], rel_j[tail]), sm.T[3 * tail], rel_i[tail] * vs + rel_j[tail], rel_v[tail], f, (1u << R) | (1u << f));
            else
              dist_push_publish(d.p, d.p.S, tile_off(rel_i[tail], rel_j[tail]), sm.T[3 * tail], rel_i[tail] * vs + rel_j[tail], rel_v[tail],
                                own(rel_v[tail] - base), all_ranks);
          } else if (rel_t[tail] == 2) {  // a panel tile of another rank's column that I pass on: to everybody but its producer and me
            dist_push_publish(d.p, d.p.S, tile_off(rel_i[tail], rel_j[tail]), sm.T[3 * tail], rel_i[tail] * vs + rel_j[tail], rel_v[tail], -1,
                              all_ranks & ~(1u << R) & ~(1u << own(rel_j[tail])));
          } else {                 // a trailing update of one of my tiles: stays here
            __threadfence();
            st_release(&a.ver[rel_i[tail] * vs + rel_j[tail]], rel_v[tail]);
          }
          tail ^= 1;
          inflight--;
          progress = true;
        }
        if (!claimed && !exhausted && inflight < 2) {
          if (claim(head)) claimed = true;
          else exhausted = true;
          progress = true;
        }
        if (claimed && deps_ready(head)) {
          rel_t[head] = sm.op[head][0], rel_i[head] = sm.op[head][2], rel_j[head] = sm.op[head][3], rel_v[head] = base + sm.op[head][1] + 1;
          issue(head);
          head ^= 1;
          inflight++;
          claimed = false;
          progress = true;
        }
        if (exhausted && inflight == 0) {
          mbar_arrive(&sm.full[head]);
          break;
        }
        if (progress) {
          t_idle = clock64();
        } else {
          __nanosleep(20);
          if (*(volatile int *)&ctrl->err || clock64() - t_idle > CHOL_TIMEOUT) {
            atomicExch(&ctrl->err, 1);
            sm.abort = 1;
            sm.op[head][0] = -1, sm.op[head ^ 1][0] = -1;
            mbar_arrive(&sm.full[0]);
            mbar_arrive(&sm.full[1]);
            break;
          }
        }
      }
    }
  } else {
    // =============================== workers, compute warps ======================================================
    unsigned pf[2] = {0, 0};
    for (int cur = 0;; cur ^= 1) {
      mbar_wait(&sm.full[cur], pf[cur]);
      pf[cur] ^= 1;
      const int type = sm.op[cur][0], i = sm.op[cur][2], j = sm.op[cur][3];
      if (type < 0 || *(volatile int *)&sm.abort) break;
      TilePtr A = sm.T[3 * cur], B = sm.T[3 * cur + 1], C = sm.T[3 * cur + 2];
      const bool one_row = (i == Tc);
      if (type != 2 && (!one_row || warp == 0)) {  // (type 2: the tile only passes through this CTA's shared memory)
        double acc[8][2];
#pragma unroll
        for (int q = 0; q < 8; q++) acc[q][0] = acc[q][1] = 0.0;
        if (type == 0) {  // result staged in place of A (a warp touches only its own rows), pushed by the communication thread
          tile_mma64<2>(A, B, acc);
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; q++)
#pragma unroll
            for (int h = 0; h < 2; h++) A[q * 8 + jc + h][il] = acc[q][h];
        } else {
          double *dst = tile(i, j);
          const bool diag = (i == j);
          if (diag) tile_mma64<1>(A, A, acc);
          else tile_mma64<0>(A, B, acc);
#pragma unroll
          for (int q = 0; q < 8; q++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int jl = q * 8 + jc + h;
              if (!diag || il >= jl) dst[jl * CLD + il] = C[jl][il] - acc[q][h];
            }
        }
      }
      if (type == 0) fence_proxy_async_smem();
      warp_arrive(&sm.done[cur]);
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&ctrl->done, 1) == (int)gridDim.x - 1) {
      if (*(volatile int *)&ctrl->err) *a.not_spd = 1;
      ctrl->ticket = 0;
      ctrl->qhead = 0;
      ctrl->done = 0;
      ctrl->epoch = ctrl->epoch + 1;
      __threadfence();
    }
  }
}

// "my partial Schur complement is complete": one system-scope release per peer (the kernel boundary before this launch ordered
// the accumulation kernels' writes)
__global__ void k_dist_signal(DistPeers p, int value) {
  if (threadIdx.x < p.world) {
    __threadfence_system();
    st_release_sys(p.sig[threadIdx.x] + p.rank, value);
  }
}
// owner-side sum of the partial systems: tile columns j = rank, rank + world, .. (tile rows j .. Tc), 16 bytes per lane and peer
__global__ void __launch_bounds__(256) k_dist_reduce(DistPeers p, int Tm, int Tc, int value, CholCtrl *ctrl) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < p.world) {
    if (!flag_wait_sys(p.sig[p.rank] + threadIdx.x, value, ctrl)) s_ok = 0;
  }
  __syncthreads();
  if (!s_ok) return;
  // flatten my columns into one index space of double2 elements, block-cyclic over the CTAs
  auto mine = [&](int j) { return (j / p.blk) % p.world == p.rank; };
  size_t total = 0;
  for (int j = 0; j < Tc; j++)
    if (mine(j)) total += (size_t)(Tc - j + 1) * (TILE / 2);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  int j = 0;
  while (j < Tc && !mine(j)) j++;
  size_t col_start = 0, col_len = j < Tc ? (size_t)(Tc - j + 1) * (TILE / 2) : 0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    while (e >= col_start + col_len) {
      col_start += col_len;
      j++;
      while (!mine(j)) j++;
      col_len = (size_t)(Tc - j + 1) * (TILE / 2);
    }
    const size_t off = dense_tile_index(Tm, j, j) * (size_t)(TILE / 2) + (e - col_start);
    double2 v[DIST_MAX];
#pragma unroll
    for (int q = 0; q < DIST_MAX; q++)
      if (q < p.world) v[q] = ld_peer_v2(reinterpret_cast<const double2 *>(p.S[q]) + off);
    double2 s = v[0];
#pragma unroll
    for (int q = 1; q < DIST_MAX; q++)
      if (q < p.world) s.x += v[q].x, s.y += v[q].y;
    reinterpret_cast<double2 *>(p.S[p.rank])[off] = s;
  }
}

// ---- host side --------------------------------------------------------------------------------------------------------
int dense_num_blocks(int n) { return (n + NB - 1) / NB; }
size_t dense_matrix_doubles(int max_n) {
  const int Tm = dense_num_blocks(max_n);
  return dense_tile_index(Tm, Tm, Tm) * (size_t)TILE + TILE;  // (column Tm does not exist: its offset is the total)
}
size_t dense_x_doubles(int max_n) { return (size_t)NB * dense_num_blocks(max_n) + NB; }
size_t dense_workspace_bytes(int max_n) {
  const size_t T = dense_num_blocks(max_n);
  return 256 + sizeof(int) * ((T + 1) * T + 2) + sizeof(unsigned long long) * 2 * NB * (T + 1);
}
void dense_workspace_init(void *ws, int max_n, cudaStream_t st) {
  cudaMemsetAsync(ws, 0, dense_workspace_bytes(max_n), st);
  const int one = 1;
  CholCtrl *c = reinterpret_cast<CholCtrl *>(ws);
  cudaMemcpyAsync(&c->epoch, &one, sizeof(int), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(&c->bepoch, &one, sizeof(int), cudaMemcpyHostToDevice, st);
}
static int sm_count(int dev) {
  static int cache[64] = {};
  if (dev >= 0 && dev < 64 && cache[dev]) return cache[dev];
  int v = 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
  if (v <= 0) v = 148;
  if (dev >= 0 && dev < 64) cache[dev] = v;
  return v;
}
void dense_setup_device(int dev) {  // per-device function attributes (> 48 KB of dynamic shared memory); called from ppo_ba_create
  (void)dev;
  cudaFuncSetAttribute(k_chol_dataflow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChSmem));
  cudaFuncSetAttribute(k_backsolve_chain<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BsSmem));
  cudaFuncSetAttribute(k_backsolve_chain<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BsSmem));
  cudaFuncSetAttribute(k_chol_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChSmem));
}

static cudaEvent_t g_mid_event = nullptr;  // test hook: recorded between the factorisation and the back-substitution
void dense_debug_set_mid_event(cudaEvent_t e) { g_mid_event = e; }

void dense_cholesky_solve(double *S, int n, int max_n, double *x, double *Winv, void *ws, int *not_spd, cudaStream_t st, long long *launches, int sm_cap) {
  if (n <= 0) return;
  const int Tm = dense_num_blocks(max_n), Tc = dense_num_blocks(n);
  int dev = 0;
  cudaGetDevice(&dev);
  CholCtrl *ctrl = reinterpret_cast<CholCtrl *>(ws);
  int *ver = reinterpret_cast<int *>(reinterpret_cast<char *>(ws) + 256);
  unsigned long long *xll = reinterpret_cast<unsigned long long *>(ver + (((size_t)(Tm + 1) * Tm + 1) & ~(size_t)1));  // tagged words of x, 8-byte aligned
  long long ops = 0;
  for (int k = 0; k < Tc; k++) {
    const long long m = Tc - 1 - k;
    ops += m + (m >= 1 ? (m + 1) * (m + 2) / 2 - 2 : 0);
  }
  // one persistent CTA per SM (208 KB of staging buffers each); sm_cap > 0: this window's share of the device when several windows are
  // solved side by side (their dependency chains overlap instead of queueing for all SMs one after the other)
  const long long cap = sm_cap > 0 ? std::min(sm_cap, sm_count(dev)) : sm_count(dev);
  const int grid = (int)(1 + (ops < cap - 1 ? ops : cap - 1));
  CholArgs a;
  a.S = S, a.Tm = Tm, a.n = n, a.Tc = Tc, a.Winv = Winv, a.ver = ver, a.ctrl = ctrl, a.not_spd = not_spd;
  k_chol_dataflow<<<grid, CH_THREADS, sizeof(ChSmem), st>>>(a);
  if (g_mid_event) cudaEventRecord(g_mid_event, st);
  k_backsolve_chain<false><<<Tc, 256, sizeof(BsSmem), st>>>(S, Tm, n, Tc, Winv, x, xll, ctrl, nullptr, 0, not_spd);
  (*launches) += 2;
}

// ---- LinearSolverEigen flavour: the solve of an INDEFINITE reduced system ---------------------------------------------------------
// LinearSolverEigen::solve (Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124) factorises with Eigen::SimplicialLDLT, an LDL^T
// WITHOUT pivoting that reports failure only on a pivot that is exactly zero: on an indefinite system it still returns the solution,
// where LinearSolverDense (Eigen::LDLT + isPositive()) and the tile Cholesky above report a failed solve.  Windows of the
// BlockSolver_6_3 / LinearSolverEigen stack (PPO_SOLVER_6_3) therefore keep a copy of the reduced system and, when the Cholesky has
// raised not_spd (or timed out), this kernel redoes the solve on the copy as a right-looking LDL^T in natural order (the fill-reducing
// ordering of the reference changes the rounding, not the solution).  The damped systems LM produces are positive definite, so this is
// the exceptional path: one CTA, one block barrier per column, no attempt at speed; it leaves at once when not_spd is clear.
// The gradient row rides along as row n, exactly as in the Cholesky: after column j it holds y = L^-1 b.
__global__ void __launch_bounds__(1024) k_ldlt_fallback(double *M, int Tm, int n, int Tc, double *x, int *not_spd) {
  if (*(volatile int *)not_spd == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grow = NB * Tc;
  for (int j = 0; j < n; j++) {
    const double d = M[dense_elem_index(Tm, j, j)];
    if (d == 0.0) return;  // SimplicialLDLT: "failure, D(k,k) is zero" -- not_spd stays set, the step is rejected
    const double inv = 1.0 / d;
    const double gj = M[dense_elem_index(Tm, grow, j)];
    for (int k = j + 1 + warp; k < n; k += 32) {
      const double lk = M[dense_elem_index(Tm, k, j)] * inv;
      for (int i = k + lane; i < n; i += 32) M[dense_elem_index(Tm, i, k)] -= M[dense_elem_index(Tm, i, j)] * lk;
      if (lane == 0) M[dense_elem_index(Tm, grow, k)] -= gj * lk;
    }
    __syncthreads();
  }
  // x <- D^-1 y, columns of L scaled, then the backward substitution L^T x = D^-1 y column by column
  for (int j = warp; j < n; j += 32) {
    const double inv = 1.0 / M[dense_elem_index(Tm, j, j)];
    for (int i = j + 1 + lane; i < n; i += 32) M[dense_elem_index(Tm, i, j)] *= inv;
    if (lane == 0) x[j] = M[dense_elem_index(Tm, grow, j)] * inv;
  }
  for (int i = n + tid; i < NB * Tc; i += 1024) x[i] = 0.0;
  __syncthreads();
  for (int k = n - 1; k > 0; k--) {
    const double xk = x[k];
    for (int i = tid; i < k; i += 1024) x[i] -= M[dense_elem_index(Tm, k, i)] * xk;
    __syncthreads();
  }
  if (tid == 0) *not_spd = 0;
}
size_t dense_used_doubles(int n, int max_n) { return dense_tile_index(dense_num_blocks(max_n), dense_num_blocks(n), dense_num_blocks(n)) * (size_t)TILE; }
void dense_ldlt_fallback(double *S_copy, int n, int max_n, double *x, int *not_spd, cudaStream_t st, long long *launches) {
  if (n <= 0) return;
  k_ldlt_fallback<<<1, 1024, 0, st>>>(S_copy, dense_num_blocks(max_n), n, dense_num_blocks(n), x, not_spd);
  (*launches)++;
}

// ---- distributed variant ------------------------------------------------------------------------------------------------------
// worker queue of one rank, level by level: T_k(i) for i = k+2 .. Tc when the rank owns column k, then U_k(i, j) for its columns
// j > k (the diagonal update U_k(k+1,k+1) and T_k(k+1) belong to the critical-path CTA of the owner)
void dense_dist_build_ops(int Tc, int rank, int world, std::vector<unsigned> *ops, int blk, int fwd) {
  ops->clear();
  auto pack = [](int type, int k, int i, int j) { return (unsigned)type << 24 | (unsigned)k << 16 | (unsigned)i << 8 | (unsigned)j; };
  auto own = [&](int j) { return (j / blk) % world; };
  for (int k = 0; k < Tc; k++) {
    if (own(k) == rank)
      for (int i = k + 2; i <= Tc; i++) ops->push_back(pack(0, k, i, k));
    else if (fwd && world > 2)  // panel tiles of my tile rows produced elsewhere: I pass them on (type 2), before my updates of the level
      for (int i = k + 2; i <= Tc; i++)
        if (own(i) == rank) ops->push_back(pack(2, k, i, k));
    for (int j = k + 1; j < Tc; j++) {
      if (own(j) != rank) continue;
      for (int i = (j == k + 1 ? j + 1 : j); i <= Tc; i++) ops->push_back(pack(1, k, i, j));
    }
  }
}
void dense_dist_reduce(const DistPeers &p, int n, int max_n, void *ws, int seq, cudaStream_t st, long long *launches, int sm_cap) {
  if (n <= 0) return;
  const int Tm = dense_num_blocks(max_n), Tc = dense_num_blocks(n);
  int dev = 0;
  cudaGetDevice(&dev);
  const int cap = sm_cap > 0 ? sm_cap : sm_count(dev);
  k_dist_signal<<<1, 32, 0, st>>>(p, seq);
  // (sm_cap > 0: several ranks share one device in the protocol test -- one CTA per SM at most, so that the other ranks' kernels
  // find idle SMs while this one waits for their signals)
  k_dist_reduce<<<sm_cap > 0 ? cap : cap * 4, 256, 0, st>>>(p, Tm, Tc, seq, reinterpret_cast<CholCtrl *>(ws));
  (*launches) += 2;
}
void dense_cholesky_solve_dist(const DistPeers &p, int n, int max_n, double *x, void *ws, int *not_spd, const unsigned *d_ops, int n_ops, int seq,
                               cudaStream_t st, long long *launches, int sm_cap) {
  if (n <= 0) return;
  const int Tm = dense_num_blocks(max_n), Tc = dense_num_blocks(n);
  int dev = 0;
  cudaGetDevice(&dev);
  CholCtrl *ctrl = reinterpret_cast<CholCtrl *>(ws);
  int *ver = reinterpret_cast<int *>(reinterpret_cast<char *>(ws) + 256);
  unsigned long long *xll = reinterpret_cast<unsigned long long *>(ver + (((size_t)(Tm + 1) * Tm + 1) & ~(size_t)1));
  const long long cap = sm_cap > 0 ? sm_cap : sm_count(dev);
  const int grid = (int)(1 + ((long long)n_ops < cap - 1 ? (long long)n_ops : cap - 1));
  CholDistArgs d;
  d.a.S = p.S[p.rank], d.a.Tm = Tm, d.a.n = n, d.a.Tc = Tc, d.a.Winv = p.Winv[p.rank], d.a.ver = ver, d.a.ctrl = ctrl, d.a.not_spd = not_spd;
  d.p = p, d.ops = d_ops, d.n_ops = n_ops, d.base = (seq & 0x3fffff) * 256;
  k_chol_dist<<<grid, CH_THREADS, sizeof(ChSmem), st>>>(d);
  if (g_mid_event) cudaEventRecord(g_mid_event, st);
  k_backsolve_chain<true><<<Tc, 256, sizeof(BsSmem), st>>>(p.S[p.rank], Tm, n, Tc, p.Winv[p.rank], x, xll, ctrl, ver, d.base, not_spd);
  (*launches) += 2;
}

#ifdef PPO_CHOL_TIMING
void dense_timing_fetch(long long out[16], bool reset) {
  cudaMemcpyFromSymbol(out, g_chol_t, sizeof(long long) * 16);
  if (reset) {
    long long z[16] = {};
    cudaMemcpyToSymbol(g_chol_t, z, sizeof z);
  }
}
#endif

}  // namespace ppo
