// Dependent-issue latency microbenchmarks on one warp (B200 design input).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void k(double *out, long long *cyc, double a, double b, int *idx) {
  __shared__ double sm[1024];
  __shared__ int si[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) { sm[i] = a + i; si[i] = idx[i]; }
  __syncthreads();
  double x = a, y = b;
  int p = threadIdx.x & 31;
  float fx = (float)a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) {
    if (OP == 0) x = __dadd_rn(x, y);
    if (OP == 1) x = __dmul_rn(x, y);
    if (OP == 2) x = __fma_rn(x, y, y);
    if (OP == 3) p = si[p];                       // LDS chain
    if (OP == 4) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31);   // 2 SHFL chain
    if (OP == 5) fx = __fmaf_rn(fx, 1.0001f, 0.5f);
    if (OP == 6) p = p * 3 + 1;                   // IMAD chain
    if (OP == 7) { __syncthreads(); }
    if (OP == 8) x = __dadd_rn(x, sm[(i * 7) & 1023]);   // DADD with independent LDS operand
    if (OP == 9) p = __ldg(idx + p);              // LDG (L1 hit) chain
    if (OP == 10) x = __ddiv_rn(x, y);
    if (OP == 11) x = exp(x * 1e-3 - 1.0);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + y + p + fx;
}
int main() {
  double *out; long long *cyc; int *idx;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024); cudaMalloc(&idx, 4096);
  int h[1024]; for (int i = 0; i < 1024; i++) h[i] = (i * 37 + 11) & 1023;
  cudaMemcpy(idx, h, 4096, cudaMemcpyHostToDevice);
  const char *names[] = {"DADD", "DMUL", "DFMA", "LDS chain", "SHFL.f64", "FFMA", "IMAD", "BAR(1 warp)",
                         "DADD+LDS", "LDG L1 chain", "DDIV", "exp"};
  for (int threads : {32, 128}) {
    for (int op = 0; op < 12; op++) {
      long long c;
      for (int rep = 0; rep < 2; rep++) {
        switch (op) {
          case 0: k<0><<<1, threads>>>(out, cyc, 1.0, 1e-9, idx); break;
          case 1: k<1><<<1, threads>>>(out, cyc, 1.0, 1.0000001, idx); break;
          case 2: k<2><<<1, threads>>>(out, cyc, 1.0, 1e-9, idx); break;
          case 3: k<3><<<1, threads>>>(out, cyc, 1.0, 1e-9, idx); break;
          case 4: k<4><<<1, threads>>>(out, cyc, 1.0, 1e-9, idx); break;
          case 5: k<5><<<1, threads>>>(out, cyc, 1.0, 1e-9, idx); break;
          case 6: k<6><<<1, threads>>>(out, cyc, 1.0, 1e-9, idx); break;
          case 7: k<7><<<1, threads>>>(out, cyc, 1.0, 1e-9, idx); break;
          case 8: k<8><<<1, threads>>>(out, cyc, 1.0, 1e-9, idx); break;
          case 9: k<9><<<1, threads>>>(out, cyc, 1.0, 1e-9, idx); break;
          case 10: k<10><<<1, threads>>>(out, cyc, 1.0, 1.0000001, idx); break;
          case 11: k<11><<<1, threads>>>(out, cyc, 1.0, 1e-9, idx); break;
        }
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("threads=%3d %-14s %.1f cycles/op\n", threads, names[op], (double)c / N);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
