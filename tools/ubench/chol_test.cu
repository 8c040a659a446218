// Standalone check + timing of the dense solve of the reduced pose system (ppo_dense.cu) on random SPD systems.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DPPO_CHOL_TIMING] -o tools/ubench/chol_test tools/ubench/chol_test.cu
//   tools/ubench/chol_test [n ...]
// Check: relative residual ||A x - b|| / ||b|| of the returned x (O(n^2) on the host), and agreement of a second run.
#include "../../point-plane-object-slam_b200/csrc/cuda/ppo_dense.cu"

#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

using namespace ppo;

static int run(int n, int max_n, int reps) {
  const int Tm = dense_num_blocks(max_n), Tc = dense_num_blocks(n), grow = 64 * Tc;
  const size_t nS = dense_matrix_doubles(max_n);
  std::vector<double> A((size_t)n * n), b(n), S(nS, 0.0);
  std::mt19937_64 rng(1234 + n);
  std::normal_distribution<double> nd(0.0, 1.0);
  // banded-ish SPD matrix: A = G G^T + n I with a sparse random G would cost O(n^3); use diagonally dominant symmetric noise
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      const double v = nd(rng) * std::exp(-0.002 * (i - j));
      A[(size_t)i * n + j] = A[(size_t)j * n + i] = v;
    }
  for (int i = 0; i < n; i++) {
    double s = 0;
    for (int j = 0; j < n; j++) s += std::fabs(A[(size_t)i * n + j]);
    A[(size_t)i * n + i] = s + 1.0;
    b[i] = nd(rng);
  }
  for (int j = 0; j < n; j++) {  // tiled lower triangle + gradient row
    for (int i = j; i < n; i++) S[dense_elem_index(Tm, i, j)] = A[(size_t)i * n + j];
    S[dense_elem_index(Tm, grow, j)] = b[j];
  }
  double *dS0, *dS, *dx, *dW;
  void *ws;
  int *dns;
  cudaMalloc(&dS0, nS * 8);
  cudaMalloc(&dS, nS * 8);
  cudaMalloc(&dx, dense_x_doubles(max_n) * 8);
  cudaMalloc(&dW, (size_t)dense_num_blocks(max_n) * DENSE_TILE * 8);
  cudaMalloc(&ws, dense_workspace_bytes(max_n));
  cudaMalloc(&dns, 4);
  cudaMemset(dns, 0, 4);
  cudaMemcpy(dS0, S.data(), nS * 8, cudaMemcpyHostToDevice);
  cudaStream_t st;
  cudaStreamCreate(&st);
  dense_workspace_init(ws, max_n, st);
  cudaEvent_t e0, e1, em;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventCreate(&em);
  dense_debug_set_mid_event(em);
  double best_bs = 1e30;
  long long launches = 0;
  std::vector<double> x(n), x2(n);
  double best = 1e30, total = 0;
  for (int r = 0; r < reps + 2; r++) {
    cudaMemcpyAsync(dS, dS0, nS * 8, cudaMemcpyDeviceToDevice, st);
    cudaEventRecord(e0, st);
    dense_cholesky_solve(dS, n, max_n, dx, dW, ws, dns, st, &launches);
    cudaEventRecord(e1, st);
    cudaError_t err = cudaStreamSynchronize(st);
    if (err != cudaSuccess) {
      printf("n=%d CUDA error: %s\n", n, cudaGetErrorString(err));
      return 1;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r >= 2) best = std::min(best, (double)ms), total += ms;
    cudaEventElapsedTime(&ms, em, e1);
    if (r >= 2) best_bs = std::min(best_bs, (double)ms);
    if (r == 0) cudaMemcpy(x.data(), dx, n * 8, cudaMemcpyDeviceToHost);
    if (r == 1) cudaMemcpy(x2.data(), dx, n * 8, cudaMemcpyDeviceToHost);
  }
  int ns = 0;
  cudaMemcpy(&ns, dns, 4, cudaMemcpyDeviceToHost);
#ifdef PPO_CHOL_TIMING
  {
    long long t[16];
    dense_timing_fetch(t, true);
    const double per = 1.0 / ((reps + 2) * (double)Tc);
    printf("   critical path, cycles per panel: factor %.0f  publish+operands %.0f  T-op %.0f  U-op %.0f\n", t[0] * per, t[1] * per, t[2] * per, t[3] * per);
    printf("   back-substitution: cycles from 'x of the next block seen' to 'own x stored', mean over blocks: %.0f\n", t[9] ? (double)t[8] / (double)t[9] : 0.0);
    printf("   inside the factorisation, cycles per panel: (a) 4 x 16x16 %.0f  (b) %.0f  (c, warp 0 part) %.0f  (c, whole incl. next (a)) %.0f\n", t[4] * per, t[5] * per, t[6] * per,
           t[7] * per);
  }
#endif
  double rn = 0, bn = 0, dmax = 0;
  for (int i = 0; i < n; i++) {
    double s = -b[i];
    for (int j = 0; j < n; j++) s += A[(size_t)i * n + j] * x[j];
    rn += s * s, bn += b[i] * b[i];
    dmax = std::max(dmax, std::fabs(x[i] - x2[i]));
  }
  const double rel = std::sqrt(rn / bn), fl = (double)n * n * n / 3.0 + 2.0 * n * n;
  printf("n=%5d max_n=%5d  residual %.3e  rerun-diff %.3e  not_spd %d  mean %.3f ms  best %.3f ms (back-substitution %.3f)  %.2f TFLOP/s  launches/solve %lld  %s\n", n,
         max_n, rel, dmax, ns, total / reps, best, best_bs, fl / (best * 1e-3) / 1e12, launches / (reps + 2), (rel < 1e-10 && ns == 0 && dmax == 0.0) ? "OK" : "FAIL");
  cudaFree(dS0), cudaFree(dS), cudaFree(dx), cudaFree(dW), cudaFree(ws), cudaFree(dns);
  return !(rel < 1e-10 && ns == 0);
}

int main(int argc, char **argv) {
  dense_setup_device(0);
  std::vector<int> ns;
  for (int i = 1; i < argc; i++) ns.push_back(atoi(argv[i]));
  if (ns.empty()) ns = {9, 54, 64, 100, 128, 384, 1000, 1644, 4000, 7794};
  int bad = 0;
  for (int n : ns) {
    bad += run(n, n, n > 3000 ? 3 : 10);
    if (n == 384) bad += run(300, 384, 5);  // a smaller system inside a larger allocation (round 2 of a BA call)
  }
  // an indefinite system must raise not_spd and still terminate
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}
