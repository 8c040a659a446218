// Dense solve of the Schur-reduced pose system: the GPU counterpart of
// LinearSolverDense::solve (Thirdparty/g2o/g2o/solvers/linear_solver_dense.h:65-113, Eigen::LDLT +
// isPositive()).  Blocked right-looking Cholesky in double; a non-positive pivot raises `not_spd`,
// which the LM loop treats exactly like g2o's failed solve (step rejected).
//
// Storage: S is (n+1) x ld doubles.  Read row-major it holds the UPPER triangle of Hschur in rows
// 0..n-1 and the reduced gradient in column n; read column-major it is the LOWER triangle with the
// gradient as an extra row n.  Factoring the lower triangle while carrying row n through the panel
// solves and trailing updates turns row n into y = L^-1 b (forward substitution for free).
#include <cuda_runtime.h>

#include "ppo_dense.h"

namespace ppo {

constexpr int NB = 64;  // panel width

#define A_(i, j) S[(size_t)(j) * ld + (i)]

// --- diagonal block: unblocked Cholesky of an nb x nb block in shared memory -----------------------
__global__ void __launch_bounds__(256) k_potrf_diag(double *S, int ld, int k, int nb, int *not_spd) {
  __shared__ double a[NB][NB + 1];
  const int tid = threadIdx.x;
  for (int t = tid; t < nb * nb; t += 256) {
    const int i = t % nb, j = t / nb;
    a[i][j] = (i >= j) ? A_(k + i, k + j) : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < nb; j++) {
    double d = a[j][j];
    __syncthreads();
    if (!(d > 0.0)) {
      if (tid == 0) *not_spd = 1;
      d = 1.0;
    }
    const double s = sqrt(d);
    if (tid == 0) a[j][j] = s;
    for (int i = j + 1 + tid; i < nb; i += 256) a[i][j] /= s;
    __syncthreads();
    // trailing rank-1 update of the lower triangle
    const int m = nb - j - 1;
    for (int t = tid; t < m * m; t += 256) {
      const int i = j + 1 + t % m, c = j + 1 + t / m;
      if (i >= c) a[i][c] -= a[i][j] * a[c][j];
    }
    __syncthreads();
  }
  for (int t = tid; t < nb * nb; t += 256) {
    const int i = t % nb, j = t / nb;
    if (i >= j) A_(k + i, k + j) = a[i][j];
  }
}

// --- panel: X = A(rows, k:k+nb) * L^-T, one row per thread -------------------------------------------
constexpr int TRSM_ROWS = 32;
__global__ void __launch_bounds__(TRSM_ROWS) k_trsm_panel(double *S, int ld, int k, int nb, int n_rows_total) {
  __shared__ double L[NB][NB];  // reads are warp-wide broadcasts: no padding needed (keeps static smem at 48 KB)
  __shared__ double xs[NB][TRSM_ROWS];
  const int tid = threadIdx.x;
  for (int t = tid; t < nb * nb; t += TRSM_ROWS) {
    const int i = t % nb, j = t / nb;
    L[i][j] = (i >= j) ? A_(k + i, k + j) : 0.0;
  }
  __syncthreads();
  const int i = k + nb + blockIdx.x * TRSM_ROWS + tid;
  if (i >= n_rows_total) return;
  for (int j = 0; j < nb; j++) {
    double v = A_(i, k + j);
    for (int m = 0; m < j; m++) v -= xs[m][tid] * L[j][m];
    v /= L[j][j];
    xs[j][tid] = v;
    A_(i, k + j) = v;
  }
}

// --- trailing update: C(i,j) -= sum_m P(i,m) P(j,m) over 64x64 tiles of the lower triangle -----------
constexpr int TS = 64;
__global__ void __launch_bounds__(256) k_syrk_update(double *S, int ld, int k, int nb, int n, int n_tiles_side) {
  constexpr int KC = 32;  // panel columns staged per pass (2 x 16 KB of shared memory)
  __shared__ __align__(16) double Pi[KC][TS];
  __shared__ __align__(16) double Pj[KC][TS];
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
  const int tr = tid % 16, tc = tid / 16;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0;
  for (int m0 = 0; m0 < nb; m0 += KC) {
    const int mc = min(KC, nb - m0);
    __syncthreads();
    for (int t = tid; t < mc * TS; t += 256) {
      const int r = t % TS, m = t / TS;
      Pi[m][r] = (i0 + r <= n) ? A_(i0 + r, k + m0 + m) : 0.0;
      Pj[m][r] = (j0 + r <= n) ? A_(j0 + r, k + m0 + m) : 0.0;
    }
    __syncthreads();
    for (int m = 0; m < mc; m++) {
      const double2 r01 = *reinterpret_cast<const double2 *>(&Pi[m][tr * 4]);
      const double2 r23 = *reinterpret_cast<const double2 *>(&Pi[m][tr * 4 + 2]);
      const double2 c01 = *reinterpret_cast<const double2 *>(&Pj[m][tc * 4]);
      const double2 c23 = *reinterpret_cast<const double2 *>(&Pj[m][tc * 4 + 2]);
      const double rv[4] = {r01.x, r01.y, r23.x, r23.y}, cv[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] += rv[a] * cv[b];
    }
  }
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const int j = j0 + tc * 4 + b;
    if (j >= n) continue;  // column n does not exist (row n is the carried gradient)
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int i = i0 + tr * 4 + a;
      if (i <= n && i >= j) A_(i, j) -= acc[a][b];
    }
  }
  (void)n_tiles_side;
}

// --- back substitution L^T x = y (y = row n), single CTA ---------------------------------------------
__global__ void __launch_bounds__(1024) k_backsolve(const double *S, int ld, int n, double *x) {
  __shared__ double Lb[NB][NB + 1];
  __shared__ double rhs[NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (n + NB - 1) / NB;
  for (int b = nblk - 1; b >= 0; b--) {
    const int b0 = b * NB, nb = min(NB, n - b0), tail = b0 + nb;
    // rhs_i = y_i - sum_{m >= tail} L(m, i) x_m   : one warp per column i, coalesced along m
    for (int ii = warp; ii < nb; ii += 32) {
      const int i = b0 + ii;
      double s = 0;
      for (int m = tail + lane; m < n; m += 32) s += A_(m, i) * x[m];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) rhs[ii] = A_(n, i) - s;
    }
    for (int t = tid; t < nb * nb; t += 1024) {
      const int i = t % nb, j = t / nb;
      Lb[i][j] = (i >= j) ? A_(b0 + i, b0 + j) : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      for (int i = nb - 1; i >= 0; i--) {
        double s = 0;
        for (int m = i + 1 + lane; m < nb; m += 32) s += Lb[m][i] * rhs[m];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) rhs[i] = (rhs[i] - s) / Lb[i][i];
        __syncwarp();
      }
      for (int i = lane; i < nb; i += 32) x[b0 + i] = rhs[i];
    }
    __threadfence_block();
    __syncthreads();
  }
}

void dense_cholesky_solve(double *S, int n, int ld, double *x, int *not_spd, cudaStream_t st, long long *launches) {
  if (n <= 0) return;
  const int rows_total = n + 1;  // rows 0..n (row n carries the gradient)
  for (int k = 0; k < n; k += NB) {
    const int nb = (n - k < NB) ? (n - k) : NB;
    k_potrf_diag<<<1, 256, 0, st>>>(S, ld, k, nb, not_spd);
    (*launches)++;
    const int below = rows_total - (k + nb);
    if (below > 0) {
      k_trsm_panel<<<(below + TRSM_ROWS - 1) / TRSM_ROWS, TRSM_ROWS, 0, st>>>(S, ld, k, nb, rows_total);
      (*launches)++;
      const int T = (below + TS - 1) / TS;
      k_syrk_update<<<T * (T + 1) / 2, 256, 0, st>>>(S, ld, k, nb, n, T);
      (*launches)++;
    }
  }
  k_backsolve<<<1, 1024, 0, st>>>(S, ld, n, x);
  (*launches)++;
}

}  // namespace ppo
