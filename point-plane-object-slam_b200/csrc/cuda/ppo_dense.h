#pragma once
#include <cuda_runtime.h>
#include <cstddef>
namespace ppo {
// Dense solve of the reduced pose system (LinearSolverDense::solve, solvers/linear_solver_dense.h:65-113).
// Layout of S, for a handle whose pose block has at most max_n scalars: column-major lower triangle (= row-major upper
// triangle) with leading dimension ld = dense_ld(max_n) = 64 (T + 1), T = ceil(max_n / 64); 64 T columns are allocated
// (dense_matrix_doubles).  For the current system of n <= max_n scalars the right-hand side is row 64 ceil(n / 64) (tile
// aligned), rows / columns n .. 64 ceil(n / 64) - 1 must be zero.  The factorisation is in place; x receives
// 64 ceil(n / 64) doubles (zeros beyond n).  *not_spd is set on a non-positive pivot.  All work is enqueued on `st`.
// Winv: dense_num_blocks(max_n) * 64 * 64 doubles; ws: dense_workspace_bytes(max_n), initialised once by dense_workspace_init.
void dense_cholesky_solve(double *S, int n, int max_n, double *x, double *Winv, void *ws, int *not_spd, cudaStream_t st, long long *launches);
int dense_num_blocks(int n);
int dense_ld(int max_n);
size_t dense_matrix_doubles(int max_n);
size_t dense_x_doubles(int max_n);
size_t dense_workspace_bytes(int max_n);
void dense_workspace_init(void *ws, int max_n, cudaStream_t st);
void dense_setup_device(int dev);
}  // namespace ppo
