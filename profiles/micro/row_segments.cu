// Micro-benchmark: what the HBM system delivers for K3's access pattern without any arithmetic --
// every CTA owns a tile of TC consecutive cells of [row][ncell] arrays, reads the tile's segment
// of NR rows and writes its segment of NW rows (segments of TC*8 bytes at a stride of ncell*8).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o row_segments row_segments.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CPL>
__global__ void __launch_bounds__(512, 2) seg_kernel(int ncell, int nr, int nw, const double* __restrict__ in, double* __restrict__ out) {
  constexpr int TC = 32 * CPL;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int tile = blockIdx.x; tile * TC < ncell; tile += gridDim.x) {
    const int cell = tile * TC + CPL * l;
    double2 acc = make_double2(0.0, 0.0);
    for (int r = w; r < nr; r += 16) {
#pragma unroll
      for (int q = 0; q < CPL / 2; ++q) { const double2 v = __ldg((const double2*)(in + (size_t)r * ncell + cell) + q); acc.x += v.x; acc.y += v.y; }
    }
    for (int r = w; r < nw; r += 16) {
#pragma unroll
      for (int q = 0; q < CPL / 2; ++q) __stcs((double2*)(out + (size_t)r * ncell + cell) + q, acc);
    }
  }
}
// interleaved: 1 read row per 3 write rows, like the Jacobian kernel's phases
template <int CPL>
__global__ void __launch_bounds__(512, 2) seg_kernel_mix(int ncell, int nr, int nw, const double* __restrict__ in, double* __restrict__ out) {
  constexpr int TC = 32 * CPL;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int tile = blockIdx.x; tile * TC < ncell; tile += gridDim.x) {
    const int cell = tile * TC + CPL * l;
    int rw = w;
    for (int r = w; r < nr; r += 16) {
      double2 v[CPL / 2];
#pragma unroll
      for (int q = 0; q < CPL / 2; ++q) v[q] = __ldg((const double2*)(in + (size_t)r * ncell + cell) + q);
      for (int k = 0; k < 3 && rw < nw; ++k, rw += 16) {
#pragma unroll
        for (int q = 0; q < CPL / 2; ++q) __stcs((double2*)(out + (size_t)rw * ncell + cell) + q, v[q]);
      }
    }
  }
}
int main() {
  const int ncell = 75776, nr = 4801 + 468, nw = 13469;
  double *in, *out;
  cudaMalloc(&in, (size_t)nr * ncell * 8); cudaMalloc(&out, (size_t)nw * ncell * 8);
  cudaMemset(in, 0, (size_t)nr * ncell * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double gb = (double)(nr + nw) * ncell * 8 / 1e9;
  for (int var = 0; var < 6; ++var) {
    float best = 1e9f;
    for (int it = 0; it < 5; ++it) {
      cudaEventRecord(e0);
      const int g2 = 296;
      if (var == 0) seg_kernel<2><<<g2, 512>>>(ncell, nr, nw, in, out);
      if (var == 1) seg_kernel<4><<<g2, 512>>>(ncell, nr, nw, in, out);
      if (var == 2) seg_kernel<8><<<g2, 512>>>(ncell, nr, nw, in, out);
      if (var == 3) seg_kernel_mix<2><<<g2, 512>>>(ncell, nr, nw, in, out);
      if (var == 4) seg_kernel_mix<4><<<g2, 512>>>(ncell, nr, nw, in, out);
      if (var == 5) seg_kernel_mix<8><<<g2, 512>>>(ncell, nr, nw, in, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
    }
    printf("%s segments of %4d B: %.3f ms  %.0f GB/s\n", var < 3 ? "read-all-then-write-all" : "1 read : 3 writes mixed ", (var % 3 == 0 ? 512 : var % 3 == 1 ? 1024 : 2048), best, gb / (best * 1e-3));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
