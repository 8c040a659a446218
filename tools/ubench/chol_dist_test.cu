// Check + timing of the DISTRIBUTED dense solve (ppo_dense.cu: k_dist_reduce, k_chol_dist, k_backsolve_chain<true>) in one process.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench/chol_dist_test tools/ubench/chol_dist_test.cu
//   tools/ubench/chol_dist_test <world> <virtual 0|1> [n ...]
// virtual = 1: all `world` ranks are kernels on separate streams of device 0, each limited to SMs / world persistent CTAs (so that all
//              ranks are resident at once) -- exercises the protocol (ownership, queue, pushes, counters) without a second GPU;
// virtual = 0: rank q runs on device q, buffers reached through cudaDeviceEnablePeerAccess (same addresses a cudaIpc mapping gives).
// Every rank starts from a random PARTIAL system; the partial systems add up to A | b.  Check: relative residual of every rank's x,
// all ranks bit-identical, and a second solve reproducing the first.
#include "../../point-plane-object-slam_b200/csrc/cuda/ppo_dense.cu"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

using namespace ppo;

#define CKE(x)                                                                    \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                    \
    }                                                                             \
  } while (0)

static int g_blk = 1;
static int g_fwd = 0;  // PPO_DIST_FORWARD=1: panel tiles go to the owner of their tile row's column, which forwards them
static int run(int n, int world, bool virt, int reps) {
  const int max_n = n;
  const int Tm = dense_num_blocks(max_n), Tc = dense_num_blocks(n), grow = 64 * Tc;
  const size_t nS = dense_matrix_doubles(max_n);
  std::vector<double> A((size_t)n * n), b(n);
  std::mt19937_64 rng(4321 + n);
  std::normal_distribution<double> nd(0.0, 1.0);
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
  // partial systems: rank q > 0 holds noise Z_q (only at valid positions), rank 0 holds the rest
  std::vector<std::vector<double>> Sp(world, std::vector<double>(nS, 0.0));
  for (int j = 0; j < n; j++) {
    for (int i = j; i <= n; i++) {
      const size_t e = i < n ? dense_elem_index(Tm, i, j) : dense_elem_index(Tm, grow, j);
      double rest = i < n ? A[(size_t)i * n + j] : b[j];
      for (int q = 1; q < world; q++) {
        const double z = nd(rng);
        Sp[q][e] = z;
        rest -= z;
      }
      Sp[0][e] = rest;
    }
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int cap = virt ? sms / world : 0;
  std::vector<double *> dS0(world), dS(world), dx(world), dW(world);
  std::vector<void *> ws(world);
  std::vector<int *> dns(world);
  std::vector<unsigned *> dops(world);
  std::vector<int> nops(world);
  std::vector<cudaStream_t> st(world);
  std::vector<cudaEvent_t> e0(world), e1(world);
  for (int q = 0; q < world; q++) {
    CKE(cudaSetDevice(virt ? 0 : q));
    if (!virt) {
      dense_setup_device(q);
      for (int r = 0; r < world; r++)
        if (r != q) cudaDeviceEnablePeerAccess(r, 0);
      cudaGetLastError();
    }
    CKE(cudaMalloc(&dS0[q], nS * 8));
    CKE(cudaMalloc(&dS[q], nS * 8));
    CKE(cudaMalloc(&dx[q], dense_x_doubles(max_n) * 8));
    CKE(cudaMalloc(&dW[q], (size_t)Tm * DENSE_TILE * 8));
    CKE(cudaMalloc(&ws[q], dense_workspace_bytes(max_n)));
    CKE(cudaMalloc(&dns[q], 4));
    CKE(cudaMemset(dns[q], 0, 4));
    CKE(cudaMemcpy(dS0[q], Sp[q].data(), nS * 8, cudaMemcpyHostToDevice));
    CKE(cudaStreamCreate(&st[q]));
    dense_workspace_init(ws[q], max_n, st[q]);
    std::vector<unsigned> ops;
    dense_dist_build_ops(Tc, q, world, &ops, g_blk, g_fwd);
    nops[q] = (int)ops.size();
    CKE(cudaMalloc(&dops[q], std::max<size_t>(4, ops.size() * 4)));
    CKE(cudaMemcpy(dops[q], ops.data(), ops.size() * 4, cudaMemcpyHostToDevice));
    CKE(cudaEventCreate(&e0[q]));
    CKE(cudaEventCreate(&e1[q]));
    CKE(cudaStreamSynchronize(st[q]));
  }
  std::vector<DistPeers> peers(world);
  for (int q = 0; q < world; q++) {
    peers[q].rank = q, peers[q].world = world, peers[q].blk = g_blk, peers[q].fwd = g_fwd;
    for (int r = 0; r < world; r++) dense_dist_set_peer(&peers[q], r, dS[r], dW[r], ws[r]);
  }
  long long launches = 0;
  std::vector<std::vector<double>> x(world, std::vector<double>(n)), x2(world, std::vector<double>(n));
  double best = 1e30, best_red = 1e30;
  std::vector<cudaEvent_t> em(world);
  for (int q = 0; q < world; q++) {
    CKE(cudaSetDevice(virt ? 0 : q));
    CKE(cudaEventCreate(&em[q]));
  }
  for (int r = 0; r < reps + 2; r++) {
    const int seq = r + 1;
    for (int q = 0; q < world; q++) {
      CKE(cudaSetDevice(virt ? 0 : q));
      CKE(cudaMemcpyAsync(dS[q], dS0[q], nS * 8, cudaMemcpyDeviceToDevice, st[q]));
    }
    for (int q = 0; q < world; q++) {
      CKE(cudaSetDevice(virt ? 0 : q));
      CKE(cudaStreamSynchronize(st[q]));
    }
    // phase by phase over the ranks, so that on ONE device no rank's spinning kernel can keep another rank's kernel from starting
    for (int q = 0; q < world; q++) {
      CKE(cudaSetDevice(virt ? 0 : q));
      CKE(cudaEventRecord(e0[q], st[q]));
      dense_dist_reduce(peers[q], n, max_n, ws[q], seq, st[q], &launches, cap);
      CKE(cudaEventRecord(em[q], st[q]));
    }
    const auto tw0 = std::chrono::steady_clock::now();
    if (virt)
      for (int q = 0; q < world; q++) CKE(cudaStreamSynchronize(st[q]));
    if (r < 3) printf("   rep %d: reduce phase drained after %.3f ms (host clock)\n", r, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count());
    if (r == 0 && n <= 2000) {  // the owner-side sums
      for (int q = 0; q < world; q++) {
        CKE(cudaSetDevice(virt ? 0 : q));
        CKE(cudaStreamSynchronize(st[q]));
        std::vector<double> Sq(nS);
        CKE(cudaMemcpy(Sq.data(), dS[q], nS * 8, cudaMemcpyDeviceToHost));
        double worst = 0;
        for (int j = 0; j < n; j++) {
          if (((j / 64) / g_blk) % world != q) continue;
          for (int i = j; i <= n; i++) {
            const size_t e = i < n ? dense_elem_index(Tm, i, j) : dense_elem_index(Tm, grow, j);
            const double want = (i < n ? A[(size_t)i * n + j] : b[j]);
            static int shown = 0;
            if (std::fabs(Sq[e] - want) > 1e-9 && shown < 12) {
              shown++;
              printf("      (%d,%d) e=%zu got %.6f want %.6f  partials:", i, j, e, Sq[e], want);
              for (int w = 0; w < world; w++) printf(" %.6f", Sp[w][e]);
              printf("\n");
            }
            worst = std::max(worst, std::fabs(Sq[e] - want));
          }
        }
        printf("   reduce check rank %d: max abs error %.3e\n", q, worst);
      }
    }
    for (int q = 0; q < world; q++) {
      CKE(cudaSetDevice(virt ? 0 : q));
      dense_cholesky_solve_dist(peers[q], n, max_n, dx[q], ws[q], dns[q], dops[q], nops[q], seq, st[q], &launches, cap);
      CKE(cudaEventRecord(e1[q], st[q]));
    }
    double worst = 0, worst_red = 0;
    for (int q = 0; q < world; q++) {
      CKE(cudaSetDevice(virt ? 0 : q));
      cudaError_t err = cudaStreamSynchronize(st[q]);
      if (err != cudaSuccess) {
        printf("n=%d rank %d CUDA error: %s\n", n, q, cudaGetErrorString(err));
        return 1;
      }
      float ms;
      cudaEventElapsedTime(&ms, em[q], e1[q]);
      worst = std::max(worst, (double)ms);
      cudaEventElapsedTime(&ms, e0[q], em[q]);
      worst_red = std::max(worst_red, (double)ms);
      if (r == 0) CKE(cudaMemcpy(x[q].data(), dx[q], n * 8, cudaMemcpyDeviceToHost));
      if (r == 1) CKE(cudaMemcpy(x2[q].data(), dx[q], n * 8, cudaMemcpyDeviceToHost));
    }
    if (r >= 2) best = std::min(best, worst), best_red = std::min(best_red, worst_red);
  }
  int bad = 0;
  for (int q = 0; q < world; q++) {
    int ns = 0;
    CKE(cudaSetDevice(virt ? 0 : q));
    CKE(cudaMemcpy(&ns, dns[q], 4, cudaMemcpyDeviceToHost));
    double rn = 0, bn = 0, dmax = 0, dr = 0;
    for (int i = 0; i < n; i++) {
      double s = -b[i];
      for (int j = 0; j < n; j++) s += A[(size_t)i * n + j] * x[q][j];
      rn += s * s, bn += b[i] * b[i];
      dmax = std::max(dmax, std::fabs(x[q][i] - x2[q][i]));
      dr = std::max(dr, std::fabs(x[q][i] - x[0][i]));
    }
    const double rel = std::sqrt(rn / bn);
    const bool ok = rel < 1e-10 && ns == 0 && dmax == 0.0 && dr == 0.0;
    printf("n=%5d world=%d rank %d  residual %.3e  rerun-diff %.3e  vs-rank0 %.3e  not_spd %d  ops %d  %s\n", n, world, q, rel, dmax, dr, ns, nops[q], ok ? "OK" : "FAIL");
    bad += !ok;
  }
  const double fl = (double)n * n * n / 3.0 + 2.0 * n * n;
  printf("n=%5d world=%d %s  reduce %.3f ms  factorise+back-substitute %.3f ms (max over ranks, best of %d)  %.2f TFLOP/s aggregate\n", n, world,
         virt ? "virtual ranks on one GPU" : "one GPU per rank", best_red, best, reps, fl / (best * 1e-3) / 1e12);
  for (int q = 0; q < world; q++) {
    CKE(cudaSetDevice(virt ? 0 : q));
    cudaFree(dS0[q]), cudaFree(dS[q]), cudaFree(dx[q]), cudaFree(dW[q]), cudaFree(ws[q]), cudaFree(dns[q]), cudaFree(dops[q]);
  }
  return bad;
}

int main(int argc, char **argv) {
  // the virtual ranks are kernels that wait for each other: every kernel must be loaded before the first one spins (lazy module
  // loading synchronises the context when it loads a kernel)
  setenv("CUDA_MODULE_LOADING", "EAGER", 1);
  {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    printf("device 0: %s, %d SMs, concurrentKernels %d, asyncEngineCount %d\n", pr.name, pr.multiProcessorCount, pr.concurrentKernels, pr.asyncEngineCount);
  }
  const int world = argc > 1 ? atoi(argv[1]) : 2;
  const bool virt = argc > 2 ? atoi(argv[2]) != 0 : true;
  dense_setup_device(0);
  if (getenv("PPO_DIST_BLOCK")) g_blk = std::max(1, atoi(getenv("PPO_DIST_BLOCK")));
  if (getenv("PPO_DIST_FORWARD")) g_fwd = atoi(getenv("PPO_DIST_FORWARD")) != 0;
  printf("ownership block: %d column(s), panel forwarding %s\n", g_blk, g_fwd ? "on" : "off");
  std::vector<int> ns;
  for (int i = 3; i < argc; i++) ns.push_back(atoi(argv[i]));
  if (ns.empty()) ns = {9, 64, 100, 384, 1000, 1644, 4000, 7794};
  int bad = 0;
  for (int n : ns) bad += run(n, world, virt, n > 3000 ? 3 : 5);
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}
