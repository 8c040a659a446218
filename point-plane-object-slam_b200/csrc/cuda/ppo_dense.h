#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <vector>
namespace ppo {
// Dense solve of the reduced pose system (LinearSolverDense::solve, solvers/linear_solver_dense.h:65-113).
//
// Layout of S for a handle whose pose block has at most max_n scalars (Tm = ceil(max_n / 64) tile columns): the LOWER
// triangle of Hschur in 64 x 64 tiles, packed column by column (tile column j holds tile rows j .. Tm), every tile
// contiguous: 64 columns of DENSE_CLD = 68 doubles (64 rows + 4 padding).  For the current system of n <= max_n scalars
// (Tc = ceil(n / 64)) the reduced gradient is row 0 of tile row Tc; everything else in tile columns 0 .. Tc-1 beyond the
// n x n matrix must be zero.  The factorisation is in place; x receives 64 Tc doubles (zeros beyond n).  *not_spd is set
// on a non-positive pivot.  All work is enqueued on `st`.
constexpr int DENSE_NB = 64;
constexpr int DENSE_CLD = 68;
constexpr int DENSE_TILE = DENSE_NB * DENSE_CLD;
__host__ __device__ inline size_t dense_tile_index(int Tm, int i, int j) { return (size_t)j * (Tm + 1) - (size_t)j * (j - 1) / 2 + (size_t)(i - j); }
// element (r, c), r >= c, of the lower triangle (r may also address the gradient row 64 Tc)
__host__ __device__ inline size_t dense_elem_index(int Tm, int r, int c) {
  return dense_tile_index(Tm, r >> 6, c >> 6) * (size_t)DENSE_TILE + (size_t)(c & 63) * DENSE_CLD + (size_t)(r & 63);
}
// Winv: dense_num_blocks(max_n) * DENSE_TILE doubles; ws: dense_workspace_bytes(max_n), initialised once by dense_workspace_init.
// sm_cap > 0 limits the persistent factorisation to that many CTAs (windows solved side by side share the SMs).
void dense_cholesky_solve(double *S, int n, int max_n, double *x, double *Winv, void *ws, int *not_spd, cudaStream_t st, long long *launches, int sm_cap = 0);
// LinearSolverEigen flavour (solvers/linear_solver_eigen.h:94-124): S_copy = the first dense_used_doubles(n, max_n) doubles of S as they
// were BEFORE dense_cholesky_solve; if that solve raised *not_spd the system is solved again by an unpivoted LDL^T (fails only on a zero
// pivot, like Eigen::SimplicialLDLT), x is overwritten and *not_spd cleared.  No-op (one empty launch) otherwise.
void dense_ldlt_fallback(double *S_copy, int n, int max_n, double *x, int *not_spd, cudaStream_t st, long long *launches);
size_t dense_used_doubles(int n, int max_n);
int dense_num_blocks(int n);
size_t dense_matrix_doubles(int max_n);
size_t dense_x_doubles(int max_n);
size_t dense_workspace_bytes(int max_n);
void dense_workspace_init(void *ws, int max_n, cudaStream_t st);
void dense_setup_device(int dev);
// One window on several GPUs (ppo_dense.cu, "distributed factorisation"): tile column j belongs to rank (j / blk) mod world.  All pointers
// are valid on the calling device: S / Winv / ver / sig of rank q are its peer-mapped (cudaIpc) buffers; ver = ws + 256 bytes and
// sig = ws + 64 bytes of rank q's workspace.  `seq` numbers the solves of the window and must agree on all ranks.
constexpr int DIST_MAX = 8;
constexpr int DIST_BLOCK_DEFAULT = 1;
constexpr int DIST_FORWARD_DEFAULT = 0;
struct DistPeers {
  int rank, world;
  int blk;  // ownership block: tile column j belongs to rank (j / blk) mod world (the critical path crosses NVLink once per block)
  int fwd;  // 1: a panel tile of tile row i goes from its producer to the owner of column i only, which forwards it (world > 2)
  double *S[DIST_MAX];
  double *Winv[DIST_MAX];
  int *ver[DIST_MAX];
  int *sig[DIST_MAX];
};
inline void dense_dist_set_peer(DistPeers *p, int q, double *S, double *Winv, void *ws) {
  p->S[q] = S, p->Winv[q] = Winv;
  p->ver[q] = reinterpret_cast<int *>(reinterpret_cast<char *>(ws) + 256);
  p->sig[q] = reinterpret_cast<int *>(reinterpret_cast<char *>(ws) + 64);
}
void dense_dist_build_ops(int Tc, int rank, int world, std::vector<unsigned> *ops, int blk = 1, int fwd = 0);
// every rank has accumulated ITS partial reduced system into its own S: owner-side sum of all ranks' copies (pulled over NVLink)
void dense_dist_reduce(const DistPeers &p, int n, int max_n, void *ws, int seq, cudaStream_t st, long long *launches, int sm_cap = 0);
// factorisation + backward substitution; d_ops = dense_dist_build_ops(...) on the device; x is computed on every rank
void dense_cholesky_solve_dist(const DistPeers &p, int n, int max_n, double *x, void *ws, int *not_spd, const unsigned *d_ops, int n_ops, int seq, cudaStream_t st,
                               long long *launches, int sm_cap = 0);
void dense_debug_set_mid_event(cudaEvent_t e);  // test hook (tools/ubench/chol_test.cu): times the two kernels separately
#ifdef PPO_CHOL_TIMING
void dense_timing_fetch(long long out[16], bool reset);
#endif
}  // namespace ppo
