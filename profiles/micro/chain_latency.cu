// Micro-benchmark (round 2): what bounds the column-sweep substitution chain of one warp?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_latency chain_latency.cu
// Prints cycles per operation for: dependent DFMA, SHFL(double)+DFMA, the pair chain, the quad chain.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double shfl_d(double x, int src) { return __shfl_sync(0xffffffffu, x, src); }
__global__ void k(double* out, long long* cyc, int reps) {
  __shared__ double T[32 * 33];
  const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) T[i] = 1e-3 * ((i * 7) % 13);
  __syncthreads();
  if (w != 0) return;
  double x = 1.0 + l, a = 1e-6 * l;
  long long t0, t1;
  // 1. dependent DFMA
  t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x = fma(x, a, 1e-9);
  }
  t1 = clock64(); if (l == 0) cyc[0] = t1 - t0;
  // 2. shfl + DFMA dependent
  t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int j = 0; j < 32; ++j) { const double xc = shfl_d(x, j); x -= a * xc; }
  }
  t1 = clock64(); if (l == 0) cyc[1] = t1 - t0;
  // 3. pair chain with operands from shared memory (as warp_trisolve_lower<1>)
  t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    const double* col = T + l;
#pragma unroll 1
    for (int c0 = 0; c0 < 32; c0 += 8) {
      double av[8], l10[4];
#pragma unroll
      for (int j = 0; j < 8; ++j) av[j] = (l > c0 + j) ? col[j * 33] : 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) l10[q] = T[(c0 + 2 * q) * 33 + c0 + 2 * q + 1];
      col += 8 * 33;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double xc = shfl_d(x, c0 + 2 * q), x1p = shfl_d(x, c0 + 2 * q + 1);
        const double x1f = x1p - l10[q] * xc;
        x -= av[2 * q] * xc; x -= av[2 * q + 1] * x1f;
      }
    }
  }
  t1 = clock64(); if (l == 0) cyc[2] = t1 - t0;
  // 4. quad chain: four unknowns per shuffle round
  t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    const double* col = T + l;
#pragma unroll 1
    for (int c0 = 0; c0 < 32; c0 += 8) {
      double av[8], m[2][6];
#pragma unroll
      for (int j = 0; j < 8; ++j) av[j] = (l > c0 + j) ? col[j * 33] : 0.0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + 4 * h;
        m[h][0] = T[c * 33 + c + 1]; m[h][1] = T[c * 33 + c + 2]; m[h][2] = T[(c + 1) * 33 + c + 2];
        m[h][3] = T[c * 33 + c + 3]; m[h][4] = T[(c + 1) * 33 + c + 3]; m[h][5] = T[(c + 2) * 33 + c + 3];
      }
      col += 8 * 33;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + 4 * h;
        const double x0 = shfl_d(x, c), x1 = shfl_d(x, c + 1), x2 = shfl_d(x, c + 2), x3 = shfl_d(x, c + 3);
        const double x1f = x1 - m[h][0] * x0;
        double x2f = x2 - m[h][1] * x0; x2f -= m[h][2] * x1f;
        double x3f = x3 - m[h][3] * x0; x3f -= m[h][4] * x1f; x3f -= m[h][5] * x2f;
        x -= av[4 * h] * x0; x -= av[4 * h + 1] * x1f; x -= av[4 * h + 2] * x2f; x -= av[4 * h + 3] * x3f;
      }
    }
  }
  t1 = clock64(); if (l == 0) cyc[3] = t1 - t0;
  // 5. LDS dependent chain (pointer chase in shared memory)
  {
    __shared__ int nxt[32];
    nxt[l] = (l * 5 + 1) & 31; __syncwarp();
    int p = l;
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int j = 0; j < 32; ++j) p = nxt[p];
    }
    t1 = clock64(); if (l == 0) cyc[4] = t1 - t0;
    x += p;
  }
  out[l] = x;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 256); cudaMalloc(&cyc, 64);
  const int reps = 200;
  for (int nthr : {32, 256}) {
    k<<<1, nthr>>>(out, cyc, reps); cudaDeviceSynchronize();
    k<<<1, nthr>>>(out, cyc, reps); cudaDeviceSynchronize();
    long long h[5]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("block %d threads: DFMA dependent %.1f cyc | SHFL(double)+DFMA %.1f cyc/col | pair chain %.1f cyc/col | quad chain %.1f cyc/col | LDS chase %.1f cyc\n",
           nthr, h[0] / (32.0 * reps), h[1] / (32.0 * reps), h[2] / (32.0 * reps), h[3] / (32.0 * reps), h[4] / (32.0 * reps));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
