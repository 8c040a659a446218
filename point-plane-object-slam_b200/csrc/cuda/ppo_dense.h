#pragma once
#include <cuda_runtime.h>
namespace ppo {
// In-place Cholesky of the lower (column-major) triangle of the n x n system stored in S with leading
// dimension ld, row n = right-hand side; solution written to x[0..n).  *not_spd set to 1 on a
// non-positive pivot.  All work is enqueued on `st`.
void dense_cholesky_solve(double *S, int n, int ld, double *x, int *not_spd, cudaStream_t st, long long *launches);
}  // namespace ppo
