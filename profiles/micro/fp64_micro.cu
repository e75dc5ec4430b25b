// Micro-benchmarks used to calibrate the design (run on the B200 box):
//  1. DFMA throughput per SM with 8 / 16 / 32 warps and ILP 1 / 8
//  2. dependent shared-memory load chain latency
//  3. __syncthreads cost with 8 warps
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* out, int iters, int ilp8) {
  double a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  if (ilp8) { for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c); } }
  else { for (int it = 0; it < iters * 8; ++it) a[0] = fma(a[0], b, c); }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("dfma warps=%d ilp8=%d: %.2f cycles per warp-DFMA (per SM: %.2f warp-DFMA/cycle)\n",
      blockDim.x / 32, ilp8, (double)(t1 - t0) / (iters * 8.0), (blockDim.x / 32) * iters * 8.0 / (double)(t1 - t0));
}
__global__ void ffma(float* out, int iters) {
  float a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  float b = 1.0000001f, c = 1e-9f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c); }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("ffma warps=%d: per SM %.2f warp-FFMA/cycle\n", blockDim.x / 32, (blockDim.x / 32) * iters * 8.0 / (double)(t1 - t0));
}
__global__ void ldschain(int* out, int iters) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 37 + 11) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) p = idx[p];
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) printf("dependent LDS chain: %.1f cycles per load\n", (double)(t1 - t0) / iters);
  t0 = clock64();
  for (int it = 0; it < iters; ++it) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) printf("__syncthreads with %d warps: %.1f cycles\n", blockDim.x / 32, (double)(t1 - t0) / iters);
}
int main() {
  double* d; float* f; int* ii; cudaMalloc(&d, 1 << 20); cudaMalloc(&f, 1 << 20); cudaMalloc(&ii, 1 << 16);
  for (int w : {8, 16, 32}) for (int ilp : {0, 1}) { dfma<<<1, w * 32>>>(d, 20000, ilp); cudaDeviceSynchronize(); }
  dfma<<<148, 256>>>(d, 20000, 1); cudaDeviceSynchronize();
  ffma<<<1, 256>>>(f, 20000); cudaDeviceSynchronize();
  ldschain<<<1, 256>>>(ii, 20000); cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
