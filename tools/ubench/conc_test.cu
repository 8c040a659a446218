// Do two kernels on two streams of one device run concurrently on this box?  A spins on a flag that B sets.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void spin(volatile int *flag, int want, int *out) {
  const long long t0 = clock64();
  int seen = 0;
  while (clock64() - t0 < (1ll << 27)) {
    if (*flag >= want) { seen = 1; break; }
    __nanosleep(50);
  }
  if (threadIdx.x == 0) atomicAdd(&out[seen], 1);
}
__global__ void setf(int *flag, int v) { *flag = v; }
int main() {
  printf("CUDA_MODULE_LOADING=%s CUDA_LAUNCH_BLOCKING=%s CUDA_DEVICE_MAX_CONNECTIONS=%s\n", getenv("CUDA_MODULE_LOADING"), getenv("CUDA_LAUNCH_BLOCKING"), getenv("CUDA_DEVICE_MAX_CONNECTIONS"));
  int *flag, *out;
  cudaMalloc(&flag, 4); cudaMalloc(&out, 8);
  cudaStream_t a, b;
  for (int mode = 0; mode < 4; mode++) {
    if (mode & 1) { cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking); }
    else { cudaStreamCreate(&a); cudaStreamCreate(&b); }
    const int grid = (mode & 2) ? 296 : 1;
    for (int rep = 0; rep < 3; rep++) {
      cudaMemset(flag, 0, 4); cudaMemset(out, 0, 8);
      cudaDeviceSynchronize();
      spin<<<grid, 256, 0, a>>>(flag, 1, out);
      setf<<<1, 32, 0, b>>>(flag, 1);
      cudaDeviceSynchronize();
      int h[2];
      cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
      printf("mode %d (nonblocking %d, grid %d) rep %d: CTAs timed out %d, saw the flag %d\n", mode, mode & 1, grid, rep, h[0], h[1]);
    }
  }
  return 0;
}
