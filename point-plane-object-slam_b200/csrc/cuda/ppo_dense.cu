// Dense solve of the Schur-reduced pose system: the GPU counterpart of
// LinearSolverDense::solve (Thirdparty/g2o/g2o/solvers/linear_solver_dense.h:65-113, Eigen::LDLT +
// isPositive()).  Blocked right-looking Cholesky in double; a non-positive pivot raises `not_spd`,
// which the LM loop treats exactly like g2o's failed solve (step rejected).
//
// Storage: S is (n+1) x ld doubles.  Read row-major it holds the UPPER triangle of Hschur in rows
// 0..n-1 and the reduced gradient in column n; read column-major it is the LOWER triangle with the
// gradient as an extra row n.  Factoring the lower triangle while carrying row n through the panel
// solves and trailing updates turns row n into y = L^-1 b (forward substitution for free).
//
// Per 64-column panel:  pf_factor               L_kk = chol(A_kk) and W_kk = L_kk^-1  (one CTA: k_potrf_inv for panel 0,
//                                               afterwards CTA 0 of the previous panel's k_syrk_update)
//                       k_panel_gemm  (rows/16) A_ik <- A_ik W_kk^T           (TRSM as a GEMM, 16 panel rows per CTA)
//                       k_syrk_update (tiles)   A_ij -= A_ik A_jk^T           (FP64 tensor cores, DMMA)
// then per 64-block, last to first:  k_backsolve_step   x_B = W_BB^T y_B ; y_A -= L_BA^T x_B
#include <cuda_runtime.h>

#include <cstdlib>

#include "ppo_dense.h"

namespace ppo {

constexpr int NB = 64;  // panel width

#define A_(i, j) S[(size_t)(j) * ld + (i)]

// Programmatic dependent launch: every kernel of the panel chain is launched with the programmatic-stream-serialisation
// attribute and waits here for its predecessor; launch processing of kernel N+1 overlaps the tail of kernel N.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// --- diagonal block: Cholesky + explicit inverse of the 64 x 64 factor --------------------------------
// In-place Gaussian elimination of [A | I] without pivoting or scaling: when column j of A has been eliminated it is
// dead, and exactly then column j of the identity part starts to fill in, so a single 64 x 64 array holds the live
// window.  Step j:  u = column j below the pivot (saved: L(:,j) = u / sqrt(d_j)),  v = row j / d_j,  column j := e_j,
// then the rank-1 update  M -= u v^T.  At the end M = X = L_u^-1 (unit lower), and W = L^-1 = D^-1/2 X.
// 128 threads, thread (tr, tc) keeps the 8 x 4 block M(8 tr .. , 4 tc ..) in registers: per step 12 operands come
// through shared memory for 32 FMAs (a 1-D row or column distribution needs one operand per FMA and is bound by the
// 128 B/clk shared-memory return path).  The step loop is unrolled by 8 only, so that the pivot's position INSIDE a
// register block is a compile-time constant while the code stays instruction-cache resident.
constexpr int PF_THREADS = 128;
#ifdef PPO_POTRF_TIMING
__device__ long long g_potrf_t[8];
#define PF_STAMP(n) if (threadIdx.x == 0) g_potrf_t[n] = clock64()
#else
#define PF_STAMP(n)
#endif
// branch-free reciprocal / reciprocal square root of a normal positive double (hardware seed + one cubic step [+ one
// Newton step]); anything else yields garbage, which the caller has already flagged as "not positive definite".
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
// Shared-memory accesses of the step loop by explicit 32-bit address (computed once, outside the loop) and with
// predicated stores: the publishing threads differ per step and branches would serialise their warps.
__device__ __forceinline__ void sts_if(bool p, unsigned addr, double x) {
  asm volatile("{ .reg .pred q; setp.ne.b32 q, %0, 0; @q st.shared.f64 [%1], %2; }" ::"r"((int)p), "r"(addr), "d"(x) : "memory");
}
__device__ __forceinline__ void sts2_if(bool p, unsigned addr, double x, double y) {
  asm volatile("{ .reg .pred q; setp.ne.b32 q, %0, 0; @q st.shared.v2.f64 [%1], {%2, %3}; }" ::"r"((int)p), "r"(addr), "d"(x), "d"(y)
               : "memory");
}
__device__ __forceinline__ double lds1(unsigned addr) {
  double x;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(addr) : "memory");
  return x;
}
__device__ __forceinline__ void lds2(unsigned addr, double &x, double &y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr) : "memory");
}
constexpr int LDL = NB + 2;  // 16-byte aligned rows
struct PfSmem {
  double Lu[NB][LDL];           // first the staged symmetric tile; then Lu[j][i] = column j at step j
  double rowb[2][NB];           // pivot row (double buffered)
  double rinv[NB], dpiv[NB];    // 1 / d_j, d_j
  double rs[NB], sq[NB];        // 1/sqrt(d_j), sqrt(d_j)
};
// barrier of the PF_THREADS threads that run the factorisation (the fused caller has more threads in its CTA)
__device__ __forceinline__ void pf_bar() { asm volatile("bar.sync 1, %0;" ::"n"(PF_THREADS) : "memory"); }
// STAGED: the caller has already put the symmetric tile (identity padded beyond nb) into sm.Lu and synchronised.
template <bool STAGED>
__device__ __forceinline__ void pf_factor(PfSmem &sm, double *S, int ld, int k, int nb, double *Winv, int *not_spd) {
  PF_STAMP(0);
  double(*Lu)[LDL] = sm.Lu;
  double(*rowb)[NB] = sm.rowb;
  double *rinv = sm.rinv, *dpiv = sm.dpiv, *rs = sm.rs, *sq = sm.sq;
  const int t = threadIdx.x, tr = t >> 4, tc = t & 15;
  if (!STAGED) {
    const int i = t & 63, c0 = t >> 6;
    double v[NB / 2];
#pragma unroll
    for (int q = 0; q < NB / 2; q++) {
      const int c = c0 + 2 * q;
      v[q] = (i < nb && c < nb && i >= c) ? A_(k + i, k + c) : (i == c ? 1.0 : 0.0);
    }
#pragma unroll
    for (int q = 0; q < NB / 2; q++) {
      const int c = c0 + 2 * q;
      if (i >= c) Lu[c][i] = v[q], Lu[i][c] = v[q];
    }
    pf_bar();
  }
  double a[8][4];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const double2 lo = *reinterpret_cast<const double2 *>(&Lu[8 * tr + r][4 * tc]);
    const double2 hi = *reinterpret_cast<const double2 *>(&Lu[8 * tr + r][4 * tc + 2]);
    a[r][0] = lo.x, a[r][1] = lo.y, a[r][2] = hi.x, a[r][3] = hi.y;
  }
  pf_bar();
  // operands of step 0
  if (tc == 0) {
#pragma unroll
    for (int r = 0; r < 8; r++) Lu[0][8 * tr + r] = a[r][0];
  }
  if (tr == 0) {
#pragma unroll
    for (int c = 0; c < 4; c++) rowb[0][4 * tc + c] = a[0][c];
  }
  if (t == 0) {
    double d = a[0][0];
    if (!(d > 0.0)) {
      *not_spd = 1;
      d = 1.0;
    }
    dpiv[0] = d;
    rinv[0] = pf_rcp(d);
  }
  pf_bar();
  PF_STAMP(1);
  unsigned s_col = (unsigned)__cvta_generic_to_shared(&Lu[0][8 * tr]);    // + j * LDL * 8: column j, my 8 rows
  unsigned s_row = (unsigned)__cvta_generic_to_shared(&rowb[0][4 * tc]);  // + (j & 1) * NB * 8: pivot row, my 4 columns
  unsigned s_rinv = (unsigned)__cvta_generic_to_shared(&rinv[0]), s_dpiv = (unsigned)__cvta_generic_to_shared(&dpiv[0]);
  // opaque to the compiler: otherwise it re-derives the shared window base (a slow special-register read) in every step
  asm volatile("" : "+r"(s_col), "+r"(s_row), "+r"(s_rinv), "+r"(s_dpiv));
#pragma unroll 1
  for (int jb = 0; jb < NB / 8; jb++) {
    const unsigned c_col = s_col + jb * (8 * LDL * 8), c_rinv = s_rinv + jb * 64, c_dpiv = s_dpiv + jb * 64;
    const bool act = tr >= jb, prow = tr == jb;
#pragma unroll
    for (int js = 0; js < 8; js++) {
      const int pc = 2 * jb + (js >> 2), lc = js & 3;                 // thread column / register column of matrix column j = 8 jb + js
      const int rn = (js + 1) & 7, lcn = (js + 1) & 3;                // register row / column of the next pivot
      const int jbn = jb + (js == 7), pcn = 2 * jb + ((js + 1) >> 2);  // its thread row / column (none after the last step)
      if (act) {  // rows above the pivot block are final
        const double ri = lds1(c_rinv + 8 * js);
        double u[8], v[4];
#pragma unroll
        for (int c = 0; c < 4; c += 2) lds2(s_row + (js & 1) * (NB * 8) + 8 * c, v[c], v[c + 1]);
#pragma unroll
        for (int r = 0; r < 8; r += 2) lds2(c_col + js * (LDL * 8) + 8 * r, u[r], u[r + 1]);
#pragma unroll
        for (int c = 0; c < 4; c++) v[c] *= -ri;  // v = -(row j) / d_j
        const bool pcol = tc == pc;
#pragma unroll
        for (int r = 0; r <= js; r++) u[r] = prow ? 0.0 : u[r];  // rows up to the pivot do not move
        v[lc] = pcol ? -ri : v[lc];                              // column j restarts as e_j: X(i,j) = -u_i / d_j below the pivot
#pragma unroll
        for (int r = 0; r < 8; r++) a[r][lc] = pcol ? 0.0 : a[r][lc];
        // the next pivot row and column first: they are published while the rest of the update runs
#pragma unroll
        for (int c = 0; c < 4; c++) a[rn][c] = fma(u[rn], v[c], a[rn][c]);
        const bool piv = tr == jbn && tc == pcn;
        const double d = a[rn][lcn];
        const double rd = pf_rcp(d);  // every thread computes it on its own element (straight-line code), the owner publishes
        if (piv && !(d > 0.0)) *not_spd = 1;  // everything computed from here on is garbage and will be discarded
#pragma unroll
        for (int r = 0; r < 8; r++)
          if (r != rn) a[r][lcn] = fma(u[r], v[lcn], a[r][lcn]);
        // (js + 1) * 8 wraps into the next block of eight when js == 7; nobody matches the predicates after step 63
        sts_if(piv, c_rinv + 8 * (js + 1), rd);
        sts_if(piv, c_dpiv + 8 * (js + 1), d);
        sts2_if(tr == jbn, s_row + ((js + 1) & 1) * (NB * 8), a[rn][0], a[rn][1]);
        sts2_if(tr == jbn, s_row + ((js + 1) & 1) * (NB * 8) + 16, a[rn][2], a[rn][3]);
#pragma unroll
        for (int r = 0; r < 8; r++) sts_if(tc == pcn, c_col + (js + 1) * (LDL * 8) + 8 * r, a[r][lcn]);
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int c = 0; c < 4; c++)
            if (r != rn && c != lcn) a[r][c] = fma(u[r], v[c], a[r][c]);
      }
      pf_bar();
    }
  }
  PF_STAMP(2);
  if (t < NB) {
    const double d = dpiv[t], r = pf_rsqrt(d);
    rs[t] = r;
    sq[t] = d * r;
  }
  pf_bar();
  {  // L(i,c) = u_c(i) / sqrt(d_c), L(c,c) = sqrt(d_c): coalesced from shared memory
    const int i = t & 63;
#pragma unroll 8
    for (int c = t >> 6; c < nb; c += 2)
      if (i < nb && i >= c) A_(k + i, k + c) = i > c ? Lu[c][i] * rs[c] : sq[c];
  }
  // W(i,j) = X(i,j) / sqrt(d_i) below the diagonal, 1/sqrt(d_i) on it, exact zeros above
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const int j = 4 * tc + c;
#pragma unroll
    for (int r = 0; r < 8; r += 2) {
      const int i = 8 * tr + r;
      const double r0 = rs[i], r1 = rs[i + 1];
      const double w0 = i > j ? a[r][c] * r0 : (i == j ? r0 : 0.0);
      const double w1 = i + 1 > j ? a[r + 1][c] * r1 : (i + 1 == j ? r1 : 0.0);
      *reinterpret_cast<double2 *>(&Winv[(size_t)j * NB + i]) = make_double2(w0, w1);
    }
  }
  PF_STAMP(3);
}

__global__ void __launch_bounds__(PF_THREADS) k_potrf_inv(double *S, int ld, int k, int nb, double *Winv, int *not_spd) {
  __shared__ __align__(16) PfSmem sm;
  pdl_launch_dependents();
  pdl_wait();
  pf_factor<false>(sm, S, ld, k, nb, Winv, not_spd);
}

// --- 64 x 64 output tile of  C = sum_m A(i,m) B(j,m)  on the FP64 tensor cores ---------------------------
// mma.sync.aligned.m8n8k4.row.col.f64: A fragment a[row = lane/4][k = lane%4], B fragment b[k = lane%4][col = lane/4],
// C fragment c0,c1 = C[row = lane/4][col = 2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int TS = 64;
constexpr int KC = 64;       // the whole panel width is staged at once (dynamic shared memory, > 48 KB)
constexpr int SLD = TS + 4;  // padded leading dimension of the staged tiles (bank-conflict-free fragment loads)
// 8 warps; warp w owns rows 8w..8w+7 of the tile and all 64 columns (8 DMMA column blocks).
// sA[m][r] = A(i0 + r, m), sB[m][r] = B(j0 + r, m)
__device__ __forceinline__ void tile_mma(const double (*sA)[SLD], const double (*sB)[SLD], int mc, double acc[8][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = warp * 8 + (lane >> 2), kk = lane & 3;
  for (int m0 = 0; m0 < mc; m0 += 4) {
    const double a = sA[m0 + kk][row];
#pragma unroll
    for (int nbk = 0; nbk < 8; nbk++) {
      const double b = sB[m0 + kk][nbk * 8 + (lane >> 2)];
      dmma(acc[nbk][0], acc[nbk][1], a, b);
    }
  }
}

// --- panel: X = A(rows, k:k+nb) * W^T  (W = inverse of the diagonal factor) --------------------------------
// One CTA per 16 rows of the panel (in place: a CTA reads and writes only its own rows), so that the 64 x 64 x 64 product
// that sits on the critical path of every panel is spread over four SMs: warp w owns the 8 x 16 block
// (rows 8 (w & 1).., columns 16 (w >> 1)..) and issues 2 DMMAs per k-step instead of 8.
constexpr int GR = 16;        // rows per CTA
constexpr int GLD = GR + 4;   // padded leading dimension of the staged A rows
__global__ void __launch_bounds__(256) k_panel_gemm(double *S, int ld, int k, int nb, int n_rows_total, const double *Winv) {
  __shared__ double sA[NB][GLD];  // sA[m][r] = A(i0 + r, k + m)
  __shared__ double sB[NB][SLD];  // sB[m][j] = W(j, m)
  pdl_launch_dependents();
  pdl_wait();
  const int i0 = k + nb + blockIdx.x * GR;
  const int tid = threadIdx.x;
  {
    double ra[4], rb[16];
    const int r = tid % GR, ma = tid / GR;  // 16 rows x 16 column groups
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int m = ma + 16 * q;
      ra[q] = (m < nb && i0 + r < n_rows_total) ? A_(i0 + r, k + m) : 0.0;
    }
    const int j = tid % NB, mb = tid / NB;
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const int m = mb + 4 * q;
      rb[q] = (m < nb) ? Winv[(size_t)m * NB + j] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) sA[ma + 16 * q][r] = ra[q];
#pragma unroll
    for (int q = 0; q < 16; q++) sB[mb + 4 * q][j] = rb[q];
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const int rblk = warp & 1, cg = warp >> 1;
  const int row = rblk * 8 + (lane >> 2), kk = lane & 3;
  double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  const int mc = (nb + 3) & ~3;
  for (int m0 = 0; m0 < mc; m0 += 4) {
    const double a = sA[m0 + kk][row];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const double b = sB[m0 + kk][(2 * cg + q) * 8 + (lane >> 2)];
      dmma(acc[q][0], acc[q][1], a, b);
    }
  }
  const int i = i0 + row;
  if (i < n_rows_total) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int j = (2 * cg + q) * 8 + 2 * (lane & 3);
      if (j < nb) A_(i, k + j) = acc[q][0];
      if (j + 1 < nb) A_(i, k + j + 1) = acc[q][1];
    }
  }
}

// --- trailing update: C(i,j) -= sum_m P(i,m) P(j,m) over 64x64 tiles of the lower triangle ----------------
// CTA 0 owns the next diagonal tile: once it is updated the CTA goes straight on to factorise it (pf_factor), so the
// 64-step latency-bound factorisation of panel k+1 overlaps the rest of the trailing update of panel k.
struct TileSmem {
  double sA[KC][SLD];
  double sB[KC][SLD];
};
union SyrkSmem {
  TileSmem t;
  PfSmem pf;
};
__global__ void __launch_bounds__(256) k_syrk_update(double *S, int ld, int k, int nb, int n, int gr, double *Wnext, int *not_spd) {
  extern __shared__ __align__(16) unsigned char dsm[];
  SyrkSmem &sm = *reinterpret_cast<SyrkSmem *>(dsm);
  double(*sA)[SLD] = sm.t.sA;
  double(*sB)[SLD] = sm.t.sB;
  pdl_launch_dependents();
  // map linear block id -> (bi, bj), bj <= bi
  int bid = blockIdx.x, bi = 0;
  while (bid >= bi + 1) {
    bid -= bi + 1;
    bi++;
  }
  const int bj = bid;
  const int k2 = k + nb;
  const int i0 = k2 + bi * TS, j0 = k2 + bj * TS;
  const int tid = threadIdx.x;
  pdl_wait();
  const int lane = tid & 31, warp = tid >> 5;
  const int il = warp * 8 + (lane >> 2), i = i0 + il;  // my row of the tile; my columns: q * 8 + 2 (lane % 4) + h
  double acc[8][2], c[8][2];
#pragma unroll
  for (int q = 0; q < 8; q++) acc[q][0] = acc[q][1] = 0.0;
  // the C tile is read up front, together with the panel operands (one global-memory latency instead of two)
#pragma unroll
  for (int q = 0; q < 8; q++)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int j = j0 + q * 8 + 2 * (lane & 3) + h;
      c[q][h] = (i <= gr && j < n && i >= j) ? A_(i, j) : 0.0;  // columns >= n do not exist (row gr is the carried gradient)
    }
  for (int m0 = 0; m0 < nb; m0 += KC) {
    const int mc = min(KC, nb - m0);
    __syncthreads();
    {
      const int r = tid % TS, mb = tid / TS;
      double ra[KC / 4], rb[KC / 4];
#pragma unroll
      for (int q = 0; q < KC / 4; q++) {
        const int m = mb + 4 * q;
        ra[q] = (m < mc && i0 + r <= gr) ? A_(i0 + r, k + m0 + m) : 0.0;
        rb[q] = (m < mc && j0 + r <= gr) ? A_(j0 + r, k + m0 + m) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < KC / 4; q++) sA[mb + 4 * q][r] = ra[q], sB[mb + 4 * q][r] = rb[q];
    }
    __syncthreads();
    tile_mma(sA, sB, (mc + 3) & ~3, acc);
  }
  const bool next_diag = blockIdx.x == 0 && k2 < n;  // this tile is the diagonal block of the next panel
  const int nb2 = min(NB, n - k2);
  if (i <= gr) {
#pragma unroll
    for (int q = 0; q < 8; q++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int j = j0 + q * 8 + 2 * (lane & 3) + h;
        // (rows of the next diagonal block stay on chip: pf_factor overwrites them with the factor anyway)
        if (j < n && i >= j && !(next_diag && il < nb2)) A_(i, j) = c[q][h] - acc[q][h];
      }
  }
  if (next_diag) {
    __syncthreads();  // everybody is done with the staging buffers, which the factorisation's tile aliases
#pragma unroll
    for (int q = 0; q < 8; q++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int jl = q * 8 + 2 * (lane & 3) + h;
        if (il >= jl) {
          const double v = (il < nb2 && jl < nb2) ? c[q][h] - acc[q][h] : (il == jl ? 1.0 : 0.0);
          sm.pf.Lu[jl][il] = v;
          sm.pf.Lu[il][jl] = v;
        }
      }
    __syncthreads();
    if (tid < PF_THREADS) pf_factor<true>(sm.pf, S, ld, k2, nb2, Wnext, not_spd);
  }
}

// --- back substitution L^T x = y (y = row n), one launch per 64-block, last block first ------------------------
// every CTA recomputes x_B = W_BB^T y_B (64 x 64 mat-vec), CTA c then folds x_B into the 64 columns it owns:
// y_i -= sum_r L(b0 + r, i) x_B[r].
__global__ void __launch_bounds__(256) k_backsolve_step(double *S, int ld, int gr, int b0, int nb, const double *Winv, double *x) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double sW[NB][NB + 1];
  __shared__ double xb[NB];
  __shared__ double part[4][NB];
  __shared__ double yb[NB];
  const int tid = threadIdx.x;
  const int i = tid & 63, p = tid >> 6;
  // all global operands are requested up front: W_BB, y_B and (CTAs 1..) this CTA's 64 columns of the block row L(B, :)
  double w[16], l[16];
#pragma unroll
  for (int q = 0; q < 16; q++) w[q] = Winv[(size_t)(p + 4 * q) * NB + i];
  const int col = ((int)blockIdx.x - 1) * NB + i;
  const bool fold = blockIdx.x > 0 && col < b0;
#pragma unroll
  for (int q = 0; q < 16; q++) l[q] = (fold && p + 4 * q < nb) ? A_(b0 + p + 4 * q, col) : 0.0;
  if (tid < NB) yb[tid] = tid < nb ? A_(gr, b0 + tid) : 0.0;
#pragma unroll
  for (int q = 0; q < 16; q++) sW[i][p + 4 * q] = w[q];
  __syncthreads();
  {
    const int j = i;
    double s = 0.0;
    for (int r = j + p; r < nb; r += 4) s += sW[r][j] * yb[r];
    part[p][j] = s;
  }
  __syncthreads();
  if (tid < NB) {
    const double v = part[0][tid] + part[1][tid] + part[2][tid] + part[3][tid];
    xb[tid] = v;
    if (blockIdx.x == 0 && tid < nb) x[b0 + tid] = v;
  }
  __syncthreads();
  if (blockIdx.x == 0) return;  // CTA 0 only publishes x_B; CTAs 1.. own the columns [64 (c-1), 64 c)
  {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 16; q++) s += l[q] * xb[p + 4 * q];
    part[p][i] = s;
  }
  __syncthreads();
  if (tid < NB) {
    const int c = ((int)blockIdx.x - 1) * NB + tid;
    if (c < b0) A_(gr, c) -= part[0][tid] + part[1][tid] + part[2][tid] + part[3][tid];
  }
}

// =====================================================================================================================
// Persistent dataflow factorisation (default path).  ONE launch factorises the whole reduced system; a second launch
// does the backward substitution.  The matrix is cut into 64 x 64 tiles; every tile (i, j) goes through the operations
//      U_0 .. U_{j-1}   C_ij -= P_ik P_jk^T                         (trailing updates, any CTA)
//      F_j   (i == j)   L_jj = chol(C_jj), W_j = L_jj^-1            (critical-path CTA)
//      T_j   (i >  j)   P_ij = C_ij W_j^T                           (any CTA; the tile right below the diagonal: critical-path CTA)
// and carries a version counter ver[i][j] = number of operations applied (tagged with the launch epoch, so the counters
// are never cleared).  The first CTA to start takes the critical path  F_k -> T_k(k+1) -> U_k(k+1,k+1) -> F_{k+1}  and keeps
// the diagonal tile in shared memory between steps; all other CTAs pull the remaining operations from a global queue that
// is ordered level by level (T_k first, then U_k by column), i.e. topologically: an operation only waits for operations
// that were claimed before it, by CTAs that are therefore running -- the kernel cannot deadlock whatever number of CTAs
// is resident, and it needs no cooperative launch.  Operand tiles are staged by the TMA engine (cp.async.bulk, one 512-byte
// column per copy into a padded, bank-conflict-free layout, completion on an mbarrier); the products run on the FP64
// tensor cores (DMMA).  Every wait is bounded: on a time-out the kernel raises ctrl->err and all CTAs leave.
// =====================================================================================================================
struct CholCtrl {
  int ticket;  // role tickets of the running launch (first CTA to arrive = critical-path CTA)
  int qhead;   // next operation of the worker queue
  int done;    // CTAs that have left the kernel; the last one resets the block for the next launch
  int epoch;   // launch counter: flag values are epoch * 256 + level
  int err;     // 1: a wait timed out
  int bticket, bdone, bepoch;  // the same for the back-substitution kernel
};
constexpr int CLD = NB + 4;  // padded leading dimension of a staged tile (doubles): 544-byte columns, 16-byte aligned
constexpr long long CHOL_TIMEOUT = 1ll << 32;  // cycles (~2 s)

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
// one warp stages the 64 x 64 tile whose top-left element is `src` (column stride ld_src) into dst[col][row]
__device__ __forceinline__ void tile_load(double (*dst)[CLD], const double *src, size_t ld_src, unsigned long long *bar) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int c = lane; c < NB; c += 32) bulk_g2s(&dst[c][0], src + (size_t)c * ld_src, NB * 8, bar);
}

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

struct ChSmem {
  double A[NB][CLD];  // operand tile / result tile
  double B[NB][CLD];  // second operand / W
  PfSmem pf;          // diagonal tile of the critical-path CTA
  unsigned long long mbar;
  int op[4];          // broadcast of the decoded queue entry {type, k, i, j}
  int ok;
};
struct CholArgs {
  double *S;
  int ld, n, Tc;  // Tc column tiles; row tiles 0 .. Tc, the last one holds the carried gradient (row 64 Tc)
  double *Winv;
  int *ver;
  int vs;  // row stride of ver
  CholCtrl *ctrl;
  int *not_spd;
};
// acc(i, j) += sum_m sA[m][i] sB[m][j] on a 64 x 64 x 64 tile; warp w owns rows 8w..8w+7.
// MODE 0: full; 1: lower triangle only (column blocks <= row block); 2: sB = W^T of a lower-triangular W (m <= j)
template <int MODE>
__device__ __forceinline__ void tile_mma64(const double (*sA)[CLD], const double (*sB)[CLD], double acc[8][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = warp * 8 + (lane >> 2), kk = lane & 3;
#pragma unroll 4
  for (int m0 = 0; m0 < NB; m0 += 4) {
    const double a = sA[m0 + kk][row];
#pragma unroll
    for (int nbk = 0; nbk < 8; nbk++) {
      if (MODE == 1 && nbk > warp) continue;
      if (MODE == 2 && m0 > 8 * nbk + 7) continue;
      const double b = sB[m0 + kk][nbk * 8 + (lane >> 2)];
      dmma(acc[nbk][0], acc[nbk][1], a, b);
    }
  }
}
__device__ __forceinline__ size_t tile_off(int ld, int i, int j) { return (size_t)(NB * j) * ld + (size_t)NB * i; }

__global__ void __launch_bounds__(256) k_chol_dataflow(CholArgs a) {
  extern __shared__ __align__(16) unsigned char dsm[];
  ChSmem &sm = *reinterpret_cast<ChSmem *>(dsm);
  __shared__ int s_role, s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double *S = a.S;
  const int ld = a.ld, Tc = a.Tc, vs = a.vs;
  CholCtrl *ctrl = a.ctrl;
  if (tid == 0) {
    s_role = atomicAdd(&ctrl->ticket, 1);
    s_base = ld_acquire(&ctrl->epoch) * 256;
    mbar_init(&sm.mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int role = s_role, base = s_base;
  unsigned phase = 0;
  const int il = warp * 8 + (lane >> 2);  // my row of a tile in the DMMA C-fragment layout; my columns: 8 q + 2 (lane & 3) + h
  if (role == 0) {
    // ------------------------------- critical path -------------------------------------------------------------
    for (int k = 0; k < Tc; k++) {
      const int nb = min(NB, a.n - NB * k);
      if (tid < PF_THREADS) {
        if (k == 0) pf_factor<false>(sm.pf, S, ld, 0, nb, a.Winv, a.not_spd);
        else pf_factor<true>(sm.pf, S, ld, NB * k, nb, a.Winv + (size_t)k * NB * NB, a.not_spd);
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        st_release(&a.ver[k * vs + k], base + k + 1);  // W_k is published
        sm.ok = (k == 0) ? 1 : (int)flag_wait(&a.ver[(k + 1) * vs + k], base + k, ctrl);
      }
      // W_k (written by this CTA a moment ago) -> sm.B, generic loads
      {
        const double *W = a.Winv + (size_t)k * NB * NB;
        const int r = tid & 63, c0 = tid >> 6;
#pragma unroll
        for (int q = 0; q < 16; q++) sm.B[c0 + 4 * q][r] = W[(size_t)(c0 + 4 * q) * NB + r];
      }
      __syncthreads();
      if (!sm.ok) break;
      // T_k(k+1): the tile right below the diagonal
      if (warp == 0) {
        if (lane == 0) {
          fence_proxy_async();
          mbar_expect_tx(&sm.mbar, NB * NB * 8);
        }
        __syncwarp();
        tile_load(sm.A, S + tile_off(ld, k + 1, k), ld, &sm.mbar);
      }
      mbar_wait(&sm.mbar, phase);
      phase ^= 1;
      double acc[8][2];
#pragma unroll
      for (int q = 0; q < 8; q++) acc[q][0] = acc[q][1] = 0.0;
      const bool grad_tile = (k + 1 == Tc);  // the gradient tile has one valid row
      if (!grad_tile || warp == 0) tile_mma64<2>(sm.A, sm.B, acc);
      __syncthreads();  // everybody has read sm.A
      {
        double *dst = S + tile_off(ld, k + 1, k);
#pragma unroll
        for (int q = 0; q < 8; q++)
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int jl = q * 8 + 2 * (lane & 3) + h;
            dst[(size_t)jl * ld + il] = acc[q][h];
            sm.A[jl][il] = acc[q][h];
          }
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        st_release(&a.ver[(k + 1) * vs + k], base + k + 1);  // P_{k+1,k} is published
        if (k + 1 < Tc) sm.ok = (k == 0) ? 1 : (int)flag_wait(&a.ver[(k + 1) * vs + k + 1], base + k, ctrl);
      }
      if (k + 1 >= Tc) break;
      __syncthreads();
      if (!sm.ok) break;
      // U_k(k+1,k+1) on the next diagonal tile, which then stays in shared memory for F_{k+1}
      {
        const int nb2 = min(NB, a.n - NB * (k + 1));
        const double *src = S + tile_off(ld, k + 1, k + 1);
        double c[8][2];
#pragma unroll
        for (int q = 0; q < 8; q++)
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int jl = q * 8 + 2 * (lane & 3) + h;
            c[q][h] = (il >= jl) ? __ldcg(src + (size_t)jl * ld + il) : 0.0;
            acc[q][h] = 0.0;
          }
        tile_mma64<1>(sm.A, sm.A, acc);
#pragma unroll
        for (int q = 0; q < 8; q++)
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int jl = q * 8 + 2 * (lane & 3) + h;
            if (il >= jl) {
              const double v = (il < nb2 && jl < nb2) ? c[q][h] - acc[q][h] : (il == jl ? 1.0 : 0.0);
              sm.pf.Lu[jl][il] = v;
              sm.pf.Lu[il][jl] = v;
            }
          }
      }
      __syncthreads();
    }
  } else {
    // ------------------------------- workers: operations from the queue ---------------------------------------------
    int lvl = 0, lvl_start = 0;  // (thread 0) level of the last decoded entry and index of its first entry
    for (;;) {
      if (tid == 0) {
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
        bool ok = true;
        if (type == 0) {
          ok = flag_wait(&a.ver[k * vs + k], base + k + 1, ctrl);
          if (ok && k > 0) ok = flag_wait(&a.ver[i * vs + k], base + k, ctrl);
        } else if (type == 1) {
          ok = flag_wait(&a.ver[i * vs + k], base + k + 1, ctrl);
          if (ok && j != i) ok = flag_wait(&a.ver[j * vs + k], base + k + 1, ctrl);
          if (ok && k > 0) ok = flag_wait(&a.ver[i * vs + j], base + k, ctrl);
        }
        sm.op[0] = ok ? type : -1, sm.op[1] = k, sm.op[2] = i, sm.op[3] = j;
      }
      __syncthreads();
      const int type = sm.op[0], k = sm.op[1], i = sm.op[2], j = sm.op[3];
      if (type < 0) break;
      const bool one_row = (i == Tc);  // gradient tile: only row 0 carries data (the rest is zero padding)
      double acc[8][2];
#pragma unroll
      for (int q = 0; q < 8; q++) acc[q][0] = acc[q][1] = 0.0;
      if (type == 0) {
        if (warp == 0) {
          if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(&sm.mbar, 2 * NB * NB * 8);
          }
          __syncwarp();
          tile_load(sm.A, S + tile_off(ld, i, k), ld, &sm.mbar);
          tile_load(sm.B, a.Winv + (size_t)k * NB * NB, NB, &sm.mbar);
        }
        mbar_wait(&sm.mbar, phase);
        phase ^= 1;
        if (!one_row || warp == 0) {
          tile_mma64<2>(sm.A, sm.B, acc);
          double *dst = S + tile_off(ld, i, k);
#pragma unroll
          for (int q = 0; q < 8; q++)
#pragma unroll
            for (int h = 0; h < 2; h++) dst[(size_t)(q * 8 + 2 * (lane & 3) + h) * ld + il] = acc[q][h];
        }
      } else {
        const bool diag = (i == j);
        if (warp == 0) {
          if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(&sm.mbar, (diag ? 1 : 2) * NB * NB * 8);
          }
          __syncwarp();
          tile_load(sm.A, S + tile_off(ld, i, k), ld, &sm.mbar);
          if (!diag) tile_load(sm.B, S + tile_off(ld, j, k), ld, &sm.mbar);
        }
        double *ct = S + tile_off(ld, i, j);
        double c[8][2];
        if (!one_row || warp == 0) {
#pragma unroll
          for (int q = 0; q < 8; q++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int jl = q * 8 + 2 * (lane & 3) + h;
              c[q][h] = (!diag || il >= jl) ? __ldcg(ct + (size_t)jl * ld + il) : 0.0;
            }
        }
        mbar_wait(&sm.mbar, phase);
        phase ^= 1;
        if (!one_row || warp == 0) {
          if (diag) tile_mma64<1>(sm.A, sm.A, acc);
          else tile_mma64<0>(sm.A, sm.B, acc);
#pragma unroll
          for (int q = 0; q < 8; q++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int jl = q * 8 + 2 * (lane & 3) + h;
              if (!diag || il >= jl) ct[(size_t)jl * ld + il] = c[q][h] - acc[q][h];
            }
        }
      }
      __syncthreads();  // all stores of the tile issued; the staging buffers are free again
      if (tid == 0) {
        __threadfence();
        st_release(&a.ver[i * vs + j], base + k + 1);
      }
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
// for CTAs that started before it.  Chain per block: flag -> 64 x 64 mat-vec -> warp reduction -> 64 x 64 mat-vec -> flag.
struct BsSmem {
  double L[2][NB][CLD];
  double W[NB][NB + 1];
  double xc[NB];
  double accv[NB];
  double part[4][NB];
  unsigned long long mbar[2];
  int ok;
};
__global__ void __launch_bounds__(256) k_backsolve_chain(const double *S, int ld, int n, int Tc, const double *Winv, double *x, int *xrdy, CholCtrl *ctrl) {
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
  const int b = s_b, want = s_base + 1;
  const int grow = NB * Tc;
  if (b >= 0) {
    unsigned ph[2] = {0, 0};
    int buf = 0;
    if (Tc - 1 > b && warp == 0) {  // first tile of the sweep: (Tc-1, b)
      if (lane == 0) mbar_expect_tx(&sm.mbar[0], NB * NB * 8);
      __syncwarp();
      tile_load(sm.L[0], S + tile_off(ld, Tc - 1, b), ld, &sm.mbar[0]);
    }
    {  // W_b, column-major with an odd stride (conflict-free column reads)
      const double *W = Winv + (size_t)b * NB * NB;
      const int r = tid & 63, c0 = tid >> 6;
#pragma unroll
      for (int q = 0; q < 16; q++) sm.W[c0 + 4 * q][r] = __ldcg(W + (size_t)(c0 + 4 * q) * NB + r);
    }
    double s[8];  // lane-partial sums of  sum_r L(r, 8 warp + jj) x_c[r]  over all tiles so far
#pragma unroll
    for (int jj = 0; jj < 8; jj++) s[jj] = 0.0;
    bool ok = true;
    for (int c = Tc - 1; c > b; c--) {
      if (tid == 0) sm.ok = (int)flag_wait(&xrdy[c], want, ctrl);
      if (c - 1 > b && warp == 0) {  // prefetch the next tile into the other buffer (its last readers are two barriers back)
        if (lane == 0) mbar_expect_tx(&sm.mbar[buf ^ 1], NB * NB * 8);
        __syncwarp();
        tile_load(sm.L[buf ^ 1], S + tile_off(ld, c - 1, b), ld, &sm.mbar[buf ^ 1]);
      }
      __syncthreads();
      if (!sm.ok) {
        ok = false;
        break;
      }
      if (tid < NB) sm.xc[tid] = __ldcg(x + NB * c + tid);
      mbar_wait(&sm.mbar[buf], ph[buf]);
      ph[buf] ^= 1;
      __syncthreads();
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
        if (lane == 0) {
          const int col = NB * b + 8 * warp + jj;
          sm.accv[8 * warp + jj] = (col < n ? __ldcg(S + (size_t)col * ld + grow) : 0.0) - v;
        }
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
        const double v = sm.part[0][tid] + sm.part[1][tid] + sm.part[2][tid] + sm.part[3][tid];
        x[NB * b + tid] = (NB * b + tid < n) ? v : 0.0;
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        st_release(&xrdy[b], want);
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&ctrl->bdone, 1) == (int)gridDim.x - 1) {
      ctrl->bticket = 0;
      ctrl->bdone = 0;
      ctrl->bepoch = ctrl->bepoch + 1;
      __threadfence();
    }
  }
}

// ---- host side --------------------------------------------------------------------------------------------------------
int dense_num_blocks(int n) { return (n + NB - 1) / NB; }
int dense_ld(int max_n) { return NB * (dense_num_blocks(max_n) + 1); }
size_t dense_matrix_doubles(int max_n) { return (size_t)NB * dense_num_blocks(max_n) * dense_ld(max_n); }
size_t dense_x_doubles(int max_n) { return (size_t)NB * dense_num_blocks(max_n); }
size_t dense_workspace_bytes(int max_n) {
  const size_t T = dense_num_blocks(max_n);
  return 256 + sizeof(int) * ((T + 1) * T + T);
}
void dense_workspace_init(void *ws, int max_n, cudaStream_t st) {
  cudaMemsetAsync(ws, 0, dense_workspace_bytes(max_n), st);
  const int one = 1;
  CholCtrl *c = reinterpret_cast<CholCtrl *>(ws);
  cudaMemcpyAsync(&c->epoch, &one, sizeof(int), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(&c->bepoch, &one, sizeof(int), cudaMemcpyHostToDevice, st);
}

static bool legacy_path() {
  static const bool v = std::getenv("PPO_DENSE_LEGACY") != nullptr;
  return v;
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
  cudaFuncSetAttribute(k_syrk_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SyrkSmem));
  cudaFuncSetAttribute(k_chol_dataflow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChSmem));
  cudaFuncSetAttribute(k_backsolve_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BsSmem));
}

void dense_cholesky_solve(double *S, int n, int max_n, double *x, double *Winv, void *ws, int *not_spd, cudaStream_t st, long long *launches) {
  if (n <= 0) return;
  const int ld = dense_ld(max_n);
  const int Tc = dense_num_blocks(n), gr = NB * Tc;
  if (legacy_path()) {
    const int rows_total = gr + 1;  // rows 0..gr (row gr carries the gradient; rows n..gr-1 are zero padding)
    launch_pdl(k_potrf_inv, dim3(1), dim3(PF_THREADS), 0, st, S, ld, 0, n < NB ? n : NB, Winv, not_spd);
    (*launches)++;
    for (int k = 0, blk = 0; k < n; k += NB, blk++) {  // the diagonal block of panel k is already factorised
      const int nb = (n - k < NB) ? (n - k) : NB;
      double *W = Winv + (size_t)blk * NB * NB;
      const int T = (rows_total - (k + nb) + TS - 1) / TS;  // >= 1: row gr
      launch_pdl(k_panel_gemm, dim3((rows_total - (k + nb) + GR - 1) / GR), dim3(256), 0, st, S, ld, k, nb, rows_total, (const double *)W);
      launch_pdl(k_syrk_update, dim3(T * (T + 1) / 2), dim3(256), sizeof(SyrkSmem), st, S, ld, k, nb, n, gr, W + NB * NB, not_spd);
      (*launches) += 2;
    }
    for (int b = Tc - 1; b >= 0; b--) {
      const int b0 = b * NB, nb = (n - b0 < NB) ? (n - b0) : NB;
      launch_pdl(k_backsolve_step, dim3(1 + b), dim3(256), 0, st, S, ld, gr, b0, nb, (const double *)(Winv + (size_t)b * NB * NB), x);
      (*launches)++;
    }
    return;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  CholCtrl *ctrl = reinterpret_cast<CholCtrl *>(ws);
  int *ver = reinterpret_cast<int *>(reinterpret_cast<char *>(ws) + 256);
  const int Tm = dense_num_blocks(max_n);
  int *xrdy = ver + (size_t)(Tm + 1) * Tm;
  long long ops = 0;
  for (int k = 0; k < Tc; k++) {
    const long long m = Tc - 1 - k;
    ops += m + (m >= 1 ? (m + 1) * (m + 2) / 2 - 2 : 0);
  }
  static int occ[64] = {};  // resident CTAs per SM of the persistent kernel (registers / shared memory)
  if (dev >= 0 && dev < 64 && !occ[dev]) {
    int v = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_chol_dataflow, 256, sizeof(ChSmem));
    occ[dev] = v > 0 ? v : 1;
  }
  const long long cap = (long long)(dev >= 0 && dev < 64 ? occ[dev] : 1) * sm_count(dev);
  const int grid = (int)(1 + (ops < cap - 1 ? ops : cap - 1));
  CholArgs a;
  a.S = S, a.ld = ld, a.n = n, a.Tc = Tc, a.Winv = Winv, a.ver = ver, a.vs = Tm, a.ctrl = ctrl, a.not_spd = not_spd;
  k_chol_dataflow<<<grid, 256, sizeof(ChSmem), st>>>(a);
  k_backsolve_chain<<<Tc, 256, sizeof(BsSmem), st>>>(S, ld, n, Tc, Winv, x, xrdy, ctrl);
  (*launches) += 2;
}

}  // namespace ppo
