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
// Per 64-column panel:  k_potrf_inv   (1 CTA)   L_kk = chol(A_kk) and W_kk = L_kk^-1
//                       k_panel_gemm  (rows/64) A_ik <- A_ik W_kk^T           (TRSM as a GEMM)
//                       k_syrk_update (tiles)   A_ij -= A_ik A_jk^T           (FP64 tensor cores, DMMA)
// then per 64-block, last to first:  k_backsolve_step   x_B = W_BB^T y_B ; y_A -= L_BA^T x_B
#include <cuda_runtime.h>

#include "ppo_dense.h"

namespace ppo {

constexpr int NB = 64;  // panel width

#define A_(i, j) S[(size_t)(j) * ld + (i)]

// Programmatic dependent launch: every kernel of the panel chain is launched with the programmatic-stream-serialisation
// attribute and waits here for its predecessor; launch processing of kernel N+1 overlaps the tail of kernel N.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// --- diagonal block: Cholesky + explicit inverse of the 64 x 64 factor --------------------------------
// Gaussian elimination of [A | I] without scaling: A ends as L_u D (unit-lower factor times pivots), the
// identity part as X = L_u^-1; then L = L_u D^1/2 and W = L^-1 = D^-1/2 X.
// 128 threads: thread (i, h) keeps A(i, c) and X(i, c), c = 32 h + q, in REGISTERS.  Per column j two 64-vectors
// travel through (double-buffered) shared memory: va[c] = A(c, j) (pivot column, symmetric) and vx[c] = X(j, c).
// The 64 elimination steps are FULLY UNROLLED (per column half h), so every register index, every "is this column
// still live" test and the shared-memory addresses are compile-time constants: one FMA per (row, column) and step,
// one barrier and one reciprocal per column; square roots and the write-out happen after the loop.
constexpr int PF_THREADS = 128;
#ifdef PPO_POTRF_TIMING
__device__ long long g_potrf_t[8];
#define PF_STAMP(n) if (threadIdx.x == 0) g_potrf_t[n] = clock64()
#else
#define PF_STAMP(n)
#endif
template <int H>
__device__ __forceinline__ void pf_eliminate(double (&V)[32], double (&X)[32], double (*va)[NB], double (*vx)[NB], double *rinv,
                                             double (*Lu)[NB + 1], const int i, int *not_spd) {
#pragma unroll
  for (int j = 0; j < NB; j++) {
    __syncthreads();
    if ((i | 31) >= j) {  // otherwise all rows of this warp are final (warp-uniform)
      const int cur = j & 1, nxt = cur ^ 1;
      const double aij = va[cur][i];
      const double f = i > j ? aij * rinv[j] : 0.0;  // L_u(i,j); finished rows (and row j itself) do not move
      if (H == 0 && i >= j) Lu[i][j] = aij;
#pragma unroll
      for (int q = 0; q < 32; q++) {
        const int c = 32 * H + q;
        if (c > j) V[q] = fma(-f, va[cur][c], V[q]);  // live column of A
        else X[q] = fma(-f, vx[cur][c], X[q]);        // X(i,c) -= f X(j,c), X(j,j) = 1
      }
      const int jn = j + 1;
      if (jn < NB) {
        if (jn >= 32 * H && jn < 32 * H + 32 && i >= jn) {  // operand column of the next step: A(i, j+1)
          const double v = V[(jn - 32 * H) & 31];
          va[nxt][i] = v;
          if (i == jn) {
            double d = v;
            if (!(d > 0.0)) {
              *not_spd = 1;
              d = 1.0;
            }
            rinv[jn] = __drcp_rn(d);
          }
        }
        if (i == jn) {  // row j+1 of X is final (its unit diagonal is already in the register)
#pragma unroll
          for (int q = 0; q < 32; q++)
            if (32 * H + q <= jn) vx[nxt][32 * H + q] = X[q];
        }
      }
    }
  }
}
__global__ void __launch_bounds__(PF_THREADS) k_potrf_inv(double *S, int ld, int k, int nb, double *Winv, int *not_spd) {
  pdl_launch_dependents();
  pdl_wait();
  PF_STAMP(0);
  __shared__ __align__(16) double va[2][NB];
  __shared__ __align__(16) double vx[2][NB];
  __shared__ double rinv[NB];        // 1 / d_j
  __shared__ double Lu[NB][NB + 1];  // finished columns A(i,j) = L_u(i,j) d_j
  __shared__ double rs[NB], sq[NB];  // 1/sqrt(d_c), sqrt(d_c)
  const int tid = threadIdx.x;
  const int i = tid & 63, h = tid >> 6;
  double V[32], X[32];
#pragma unroll
  for (int q = 0; q < 32; q++) {
    const int c = 32 * h + q;
    V[q] = (i < nb && c < nb && i >= c) ? A_(k + i, k + c) : (i == c ? 1.0 : 0.0);
    X[q] = i == c ? 1.0 : 0.0;
  }
  if (h == 0) {
    va[0][i] = V[0];
    vx[0][i] = i == 0 ? 1.0 : 0.0;
    if (i == 0) {
      double d = V[0];
      if (!(d > 0.0)) {
        *not_spd = 1;
        d = 1.0;
      }
      rinv[0] = __drcp_rn(d);
    }
  }
  PF_STAMP(1);
  if (h == 0) pf_eliminate<0>(V, X, va, vx, rinv, Lu, i, not_spd);
  else pf_eliminate<1>(V, X, va, vx, rinv, Lu, i, not_spd);
  __syncthreads();
  PF_STAMP(2);
  if (tid < NB) {
    const double r = sqrt(rinv[tid]);
    rs[tid] = r;
    sq[tid] = 1.0 / r;
  }
  __syncthreads();
  const double rsi = rs[i];
#pragma unroll
  for (int q = 0; q < 32; q++) {
    const int c = 32 * h + q;
    Winv[(size_t)c * NB + i] = X[q] * rsi;  // W(i,c) = X(i,c) / sqrt(d_i)  (exact zeros above the diagonal)
    if (i < nb && c < nb && i >= c)        // L(i,c) = A(i,c) / sqrt(d_c) ; L(c,c) = sqrt(d_c)
      A_(k + i, k + c) = i > c ? Lu[i][c] * rs[c] : sq[c];
  }
  PF_STAMP(3);
}

// --- 64 x 64 output tile of  C = sum_m A(i,m) B(j,m)  on the FP64 tensor cores ---------------------------
// mma.sync.aligned.m8n8k4.row.col.f64: A fragment a[row = lane/4][k = lane%4], B fragment b[k = lane%4][col = lane/4],
// C fragment c0,c1 = C[row = lane/4][col = 2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int TS = 64;
constexpr int KC = 32;       // K staged per pass
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
__global__ void __launch_bounds__(256) k_panel_gemm(double *S, int ld, int k, int nb, int n_rows_total, const double *Winv) {
  __shared__ double sA[KC][SLD];
  __shared__ double sB[KC][SLD];
  pdl_launch_dependents();
  pdl_wait();
  const int i0 = k + nb + blockIdx.x * TS;
  const int tid = threadIdx.x;
  double acc[8][2];
#pragma unroll
  for (int q = 0; q < 8; q++) acc[q][0] = acc[q][1] = 0.0;
  for (int m0 = 0; m0 < nb; m0 += KC) {
    const int mc = min(KC, nb - m0);
    __syncthreads();
    {
      const int r = tid % TS, mb = tid / TS;  // 8 independent loads per operand in flight
      double ra[KC / 4], rb[KC / 4];
#pragma unroll
      for (int q = 0; q < KC / 4; q++) {
        const int m = mb + 4 * q;
        ra[q] = (m < mc && i0 + r < n_rows_total) ? A_(i0 + r, k + m0 + m) : 0.0;
        rb[q] = (m < mc) ? Winv[(size_t)(m0 + m) * NB + r] : 0.0;  // W(j = r, m)
      }
#pragma unroll
      for (int q = 0; q < KC / 4; q++) sA[mb + 4 * q][r] = ra[q], sB[mb + 4 * q][r] = rb[q];
    }
    __syncthreads();
    tile_mma(sA, sB, (mc + 3) & ~3, acc);
  }
  const int lane = tid & 31, warp = tid >> 5;
  const int i = i0 + warp * 8 + (lane >> 2);
  if (i < n_rows_total) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int j = q * 8 + 2 * (lane & 3);
      if (j < nb) A_(i, k + j) = acc[q][0];
      if (j + 1 < nb) A_(i, k + j + 1) = acc[q][1];
    }
  }
}

// --- trailing update: C(i,j) -= sum_m P(i,m) P(j,m) over 64x64 tiles of the lower triangle ----------------
__global__ void __launch_bounds__(256) k_syrk_update(double *S, int ld, int k, int nb, int n) {
  __shared__ double sA[KC][SLD];
  __shared__ double sB[KC][SLD];
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
  double acc[8][2];
#pragma unroll
  for (int q = 0; q < 8; q++) acc[q][0] = acc[q][1] = 0.0;
  for (int m0 = 0; m0 < nb; m0 += KC) {
    const int mc = min(KC, nb - m0);
    __syncthreads();
    {
      const int r = tid % TS, mb = tid / TS;
      double ra[KC / 4], rb[KC / 4];
#pragma unroll
      for (int q = 0; q < KC / 4; q++) {
        const int m = mb + 4 * q;
        ra[q] = (m < mc && i0 + r <= n) ? A_(i0 + r, k + m0 + m) : 0.0;
        rb[q] = (m < mc && j0 + r <= n) ? A_(j0 + r, k + m0 + m) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < KC / 4; q++) sA[mb + 4 * q][r] = ra[q], sB[mb + 4 * q][r] = rb[q];
    }
    __syncthreads();
    tile_mma(sA, sB, (mc + 3) & ~3, acc);
  }
  const int lane = tid & 31, warp = tid >> 5;
  const int i = i0 + warp * 8 + (lane >> 2);
  if (i <= n) {
    double c[8][2];
#pragma unroll
    for (int q = 0; q < 8; q++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int j = j0 + q * 8 + 2 * (lane & 3) + h;
        c[q][h] = (j < n && i >= j) ? A_(i, j) : 0.0;  // column n does not exist (row n is the carried gradient)
      }
#pragma unroll
    for (int q = 0; q < 8; q++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int j = j0 + q * 8 + 2 * (lane & 3) + h;
        if (j < n && i >= j) A_(i, j) = c[q][h] - acc[q][h];
      }
  }
}

// --- back substitution L^T x = y (y = row n), one launch per 64-block, last block first ------------------------
// every CTA recomputes x_B = W_BB^T y_B (64 x 64 mat-vec), CTA c then folds x_B into the 64 columns it owns:
// y_i -= sum_r L(b0 + r, i) x_B[r].
__global__ void __launch_bounds__(256) k_backsolve_step(double *S, int ld, int n, int b0, int nb, const double *Winv, double *x) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double sW[NB][NB + 1];
  __shared__ double xb[NB];
  __shared__ double part[4][NB];
  const int tid = threadIdx.x;
  __shared__ double yb[NB];
  {
    double w[16];
    const int i = tid & 63, jb = tid >> 6;
#pragma unroll
    for (int q = 0; q < 16; q++) w[q] = Winv[(size_t)(jb + 4 * q) * NB + i];
    if (tid < NB) yb[tid] = tid < nb ? A_(n, b0 + tid) : 0.0;
#pragma unroll
    for (int q = 0; q < 16; q++) sW[i][jb + 4 * q] = w[q];
  }
  __syncthreads();
  {
    const int j = tid & 63, p = tid >> 6;
    double s = 0.0;
    for (int i = j + p; i < nb; i += 4) s += sW[i][j] * yb[i];
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
    const int i = (blockIdx.x - 1) * NB + (tid & 63), p = tid >> 6;
    double s = 0.0;
    if (i < b0) {
      double l[16];
#pragma unroll
      for (int q = 0; q < 16; q++) l[q] = (p + 4 * q < nb) ? A_(b0 + p + 4 * q, i) : 0.0;
#pragma unroll
      for (int q = 0; q < 16; q++) s += l[q] * xb[p + 4 * q];
    }
    part[p][tid & 63] = s;
  }
  __syncthreads();
  if (tid < NB) {
    const int i = (blockIdx.x - 1) * NB + tid;
    if (i < b0) A_(n, i) -= part[0][tid] + part[1][tid] + part[2][tid] + part[3][tid];
  }
}

void dense_cholesky_solve(double *S, int n, int ld, double *x, double *Winv, int *not_spd, cudaStream_t st, long long *launches) {
  if (n <= 0) return;
  const int rows_total = n + 1;  // rows 0..n (row n carries the gradient)
  for (int k = 0, blk = 0; k < n; k += NB, blk++) {
    const int nb = (n - k < NB) ? (n - k) : NB;
    double *W = Winv + (size_t)blk * NB * NB;
    launch_pdl(k_potrf_inv, dim3(1), dim3(PF_THREADS), st, S, ld, k, nb, W, not_spd);
    (*launches)++;
    const int below = rows_total - (k + nb);
    if (below > 0) {
      const int T = (below + TS - 1) / TS;
      launch_pdl(k_panel_gemm, dim3(T), dim3(256), st, S, ld, k, nb, rows_total, (const double *)W);
      launch_pdl(k_syrk_update, dim3(T * (T + 1) / 2), dim3(256), st, S, ld, k, nb, n);
      (*launches) += 2;
    }
  }
  const int nblk = (n + NB - 1) / NB;
  for (int b = nblk - 1; b >= 0; b--) {
    const int b0 = b * NB, nb = (n - b0 < NB) ? (n - b0) : NB;
    launch_pdl(k_backsolve_step, dim3(1 + b), dim3(256), st, S, ld, n, b0, nb, (const double *)(Winv + (size_t)b * NB * NB), x);
    (*launches)++;
  }
}

int dense_num_blocks(int n) { return (n + NB - 1) / NB; }

}  // namespace ppo
