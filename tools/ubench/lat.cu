// Latency micro-benchmarks (cycles) for the FP64 building blocks of the dense solve: dependent DFMA / DMMA chains, reciprocal,
// shared-memory round trip, warp / block barrier.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/lat tools/ubench/lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double pf_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d, y, 1.0);
  return fma(y, fma(fma(e, e, e), e, e), y);
}
__global__ void k(long long *out, double *sink, double seed, int nwarps_active) {
  __shared__ double sm[1024];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double x = seed + lane, y = seed * 0.5, c0 = 0, c1 = 0;
  long long t0, t1;
  constexpr int N = 256;
  if (warp < nwarps_active) {
    // 1. dependent DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, y, 1.0);
    t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0);
    // 2. dependent DMMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) dmma(c0, c1, x, y);
    t1 = clock64();
    if (threadIdx.x == 0) out[1] = (t1 - t0);
    // 3. 8 independent DMMA chains (throughput per warp)
    double a[8][2] = {};
    t0 = clock64();
#pragma unroll 2
    for (int i = 0; i < N / 8; i++)
#pragma unroll
      for (int q = 0; q < 8; q++) dmma(a[q][0], a[q][1], x, y);
    t1 = clock64();
    if (threadIdx.x == 0) out[2] = (t1 - t0);
    for (int q = 0; q < 8; q++) c0 += a[q][0] + a[q][1];
    // 4. dependent reciprocal chain
    double r = x;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) r = pf_rcp(r + 1.5);
    t1 = clock64();
    if (threadIdx.x == 0) out[3] = (t1 - t0);
    // 5. shared-memory round trip: store, syncwarp, load from another lane
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) {
      sm[warp * 32 + lane] = r;
      __syncwarp();
      r = sm[warp * 32 + ((lane + 1) & 31)] + 1.0;
      __syncwarp();
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[4] = (t1 - t0);
    // 6. 8 independent DFMA chains (issue rate)
    double f[8];
    for (int q = 0; q < 8; q++) f[q] = r + q;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 8; i++)
#pragma unroll
      for (int q = 0; q < 8; q++) f[q] = fma(f[q], y, 1.0);
    t1 = clock64();
    if (threadIdx.x == 0) out[5] = (t1 - t0);
    for (int q = 0; q < 8; q++) c1 += f[q];
    x += r;
  }
  // 7. block barrier
  __syncthreads();
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; i++) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) out[6] = (t1 - t0);
  // 8. shuffle chain
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; i++) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
  t1 = clock64();
  if (threadIdx.x == 0) out[7] = (t1 - t0);
  sink[threadIdx.x] = x + c0 + c1;
}
int main() {
  long long *d, h[8];
  double *s;
  cudaMalloc(&d, 64);
  cudaMalloc(&s, 8 * 1024);
  const char *names[8] = {"dependent DFMA", "dependent DMMA m8n8k4", "DMMA, 8 independent chains (per instr)", "pf_rcp chain", "smem store+syncwarp+load+syncwarp",
                          "DFMA, 8 independent chains (per instr)", "__syncthreads", "shfl+dadd chain"};
  for (int threads : {32, 256}) {
    for (int active : {1, 8}) {
      if (active > threads / 32) continue;
      k<<<1, threads>>>(d, s, 1.0000001, active);
      k<<<1, threads>>>(d, s, 1.0000001, active);
      cudaDeviceSynchronize();
      cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
      printf("threads %d, warps doing the chains %d\n", threads, active);
      for (int i = 0; i < 8; i++) printf("  %-44s %.1f cycles\n", names[i], h[i] / 256.0);
    }
  }
  return 0;
}
