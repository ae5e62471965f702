// FP64 issue throughput per warp / per SM sub-partition on B200 (design input for the bookkeeper warp:
// six independent DADD chains per trial move).  threads = 32: one warp alone; 128: one warp per
// sub-partition; 512: four warps per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int CH>
__global__ void k(double *out, long long *cyc, double a, double b) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) x[c] = a + c;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = __dadd_rn(x[c], b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH> void run(double *out, long long *cyc) {
  for (int threads : {32, 128, 256, 512}) {
    long long c;
    for (int rep = 0; rep < 2; rep++) { k<CH><<<1, threads>>>(out, cyc, 1.0, 1e-9); cudaDeviceSynchronize(); }
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("chains=%d threads=%3d: %.2f cycles per DADD per warp (%.2f per group of %d)\n", CH, threads,
           (double)c / N / CH, (double)c / N, CH);
  }
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  run<1>(out, cyc); run<2>(out, cyc); run<6>(out, cyc); run<12>(out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
