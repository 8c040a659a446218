// micro-benchmarks: FP64 FMA latency / throughput, DMMA rate, barrier + smem round trip on one SM
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double *out, int n, long long *cyc) {
  double a = out[0], b = out[1];
  long long t0 = clock64();
  for (int i = 0; i < n; i++) a = fma(a, b, 1e-9);
  long long t1 = clock64();
  out[2] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CH>
__global__ void k_thr(double *out, int n, long long *cyc) {
  double a[CH], b = out[1];
  for (int c = 0; c < CH; c++) a[c] = out[0] + c;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++)
#pragma unroll
    for (int c = 0; c < CH; c++) a[c] = fma(a[c], b, 1e-9);
  long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < CH; c++) s += a[c];
  out[2 + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dmma(double *out, int n, long long *cyc) {
  double c0[8], c1[8], a = out[0], b = out[1];
  for (int c = 0; c < 8; c++) c0[c] = c1[c] = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++)
#pragma unroll
    for (int c = 0; c < 8; c++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[c]), "+d"(c1[c]) : "d"(a), "d"(b));
  long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < 8; c++) s += c0[c] + c1[c];
  out[2 + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar(double *out, int n, long long *cyc) {
  __shared__ double sm[2][64];
  sm[0][threadIdx.x & 63] = 1.0;
  __syncthreads();
  double v = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    __syncthreads();
    v += sm[i & 1][(threadIdx.x + i) & 63];
    if ((threadIdx.x & 63) == (i & 63)) sm[(i & 1) ^ 1][threadIdx.x & 63] = v;
  }
  long long t1 = clock64();
  out[2 + threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rcp(double *out, int n, long long *cyc) {
  double a = out[0] + 1.5;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) a = __drcp_rn(a) + 1.25;
  long long t1 = clock64();
  out[2] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double *d; long long *c, h;
  cudaMalloc(&d, 8 * 4096); cudaMalloc(&c, 8);
  cudaMemset(d, 0, 8 * 4096);
  const int n = 4096;
  auto rep = [&](const char *name, double per) { cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-40s %10lld cycles  %8.2f cyc/iter\n", name, h, (double)h / per); };
  for (int w = 0; w < 2; w++) {
    k_lat<<<1, 32>>>(d, n, c); rep("dfma dependent chain, 1 warp", n);
    k_thr<8><<<1, 32>>>(d, n, c); rep("dfma 8 chains, 1 warp (per 8 dfma)", n);
    k_thr<8><<<1, 128>>>(d, n, c); rep("dfma 8 chains, 4 warps (per 8/warp)", n);
    k_thr<8><<<1, 512>>>(d, n, c); rep("dfma 8 chains, 16 warps (per 8/warp)", n);
    k_dmma<<<1, 32>>>(d, n, c); rep("dmma 8 chains, 1 warp (per 8 dmma)", n);
    k_dmma<<<1, 128>>>(d, n, c); rep("dmma 8 chains, 4 warps", n);
    k_dmma<<<1, 256>>>(d, n, c); rep("dmma 8 chains, 8 warps", n);
    k_bar<<<1, 128>>>(d, n, c); rep("barrier+lds+sts step, 4 warps", n);
    k_bar<<<1, 1024>>>(d, n, c); rep("barrier+lds+sts step, 32 warps", n);
    k_rcp<<<1, 32>>>(d, n, c); rep("drcp_rn + dadd dependent", n);
  }
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
