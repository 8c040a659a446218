// stand-alone timing of the dense-solver kernels (warm L2, CUDA events + in-kernel clock64 stamps)
#define PPO_POTRF_TIMING 1
#include "../../point-plane-object-slam_b200/csrc/cuda/ppo_dense.cu"
#include <cstdio>
#include <vector>
#include <cstdlib>
#include <cmath>
using namespace ppo;
int main() {
  const int n = 1644, ld = n + 1;
  std::vector<double> A((size_t)(n + 1) * ld, 0.0);
  srand(1);
  for (int j = 0; j < n; j++) {
    for (int i = j; i < n; i++) A[(size_t)j * ld + i] = (i == j) ? n + 1.0 : (rand() / (double)RAND_MAX - 0.5);
    A[(size_t)j * ld + n] = 1.0;
  }
  double *S, *W, *x; int *flag;
  cudaMalloc(&S, A.size() * 8); cudaMalloc(&W, 64 * 64 * 8 * 32); cudaMalloc(&x, n * 8); cudaMalloc(&flag, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 3; rep++) {
    cudaMemcpy(S, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaEventRecord(e0);
    for (int r = 0; r < 100; r++) k_potrf_inv<<<1, PF_THREADS>>>(S, ld, 0, 64, W, flag);  // refactors garbage after the first, timing only
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    long long t[8]; cudaMemcpyFromSymbol(t, g_potrf_t, sizeof t);
    printf("k_potrf_inv: %.2f us/launch ; cycles prologue %lld loop %lld epilogue %lld\n", ms * 10, t[1] - t[0], t[2] - t[1], t[3] - t[2]);
  }
  {  // correctness of one diagonal block: L L^T = A, W L = I
    cudaMemcpy(S, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    k_potrf_inv<<<1, PF_THREADS>>>(S, ld, 0, 64, W, flag);
    std::vector<double> L((size_t)64 * ld), Wh(64 * 64);
    cudaMemcpy(L.data(), S, L.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(Wh.data(), W, Wh.size() * 8, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int i = 0; i < 64; i++)
      for (int j = 0; j <= i; j++) {
        double s = 0, w = 0;
        for (int m = 0; m <= j; m++) s += L[(size_t)m * ld + i] * L[(size_t)m * ld + j];
        for (int m = j; m <= i; m++) w += Wh[m * 64 + i] * L[(size_t)j * ld + m];  // W(i,m) L(m,j)
        e1 = fmax(e1, fabs(s - A[(size_t)j * ld + i]));
        e2 = fmax(e2, fabs(w - (i == j ? 1.0 : 0.0)));
      }
    for (int i = 0; i < 64; i++)
      for (int j = i + 1; j < 64; j++) e2 = fmax(e2, fabs(Wh[j * 64 + i]));  // strictly upper part of W must be 0
    printf("potrf block check: max|LL^T-A| = %.3e  max|WL-I| = %.3e\n", e1, e2);
  }
  long long launches = 0;
  for (int rep = 0; rep < 3; rep++) {
    cudaMemcpy(S, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaEventRecord(e0);
    dense_cholesky_solve(S, n, ld, x, W, flag, 0, &launches);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("full solve n=%d: %.3f ms\n", n, ms);
    if (rep == 2) {
      std::vector<double> xh(n);
      cudaMemcpy(xh.data(), x, n * 8, cudaMemcpyDeviceToHost);
      double r = 0;
      for (int i = 0; i < n; i++) {
        double s = 0;
        for (int j = 0; j < n; j++) s += (i >= j ? A[(size_t)j * ld + i] : A[(size_t)i * ld + j]) * xh[j];
        r = fmax(r, fabs(s - 1.0));
      }
      printf("full solve residual max|Ax-b| = %.3e\n", r);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
