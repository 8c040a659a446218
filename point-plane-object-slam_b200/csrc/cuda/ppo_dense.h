#pragma once
#include <cuda_runtime.h>
namespace ppo {
// In-place Cholesky of the lower (column-major) triangle of the n x n system stored in S with leading
// dimension ld, row n = right-hand side; solution written to x[0..n).  *not_spd set to 1 on a
// non-positive pivot.  All work is enqueued on `st`.
// Winv: scratch of dense_num_blocks(n) * 64 * 64 doubles (inverses of the diagonal factors).
void dense_cholesky_solve(double *S, int n, int ld, double *x, double *Winv, int *not_spd, cudaStream_t st, long long *launches);
int dense_num_blocks(int n);
}  // namespace ppo
