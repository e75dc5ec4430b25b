// Stand-alone timing of the register-tiled dense tail LU (same code as racg_integrate.cu tail_lu<TL>)
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NT = 256;
// helpers with a compile-time tile index so that register arrays stay in registers
template <int TL, int B> struct PubCol { template <class T> static __device__ __forceinline__ void go(const T& t, double* pc, int ti) {
#pragma unroll
  for (int a = 0; a < TL; ++a) pc[ti + 16 * a] = t[a][B]; } };
template <int TL, int A> struct PubRow { template <class T> static __device__ __forceinline__ void go(const T& t, double* pr, int tj) {
#pragma unroll
  for (int b = 0; b < TL; ++b) pr[tj + 16 * b] = t[A][b]; } };

// one elimination step for the tile block row/col KA (uniform across the CTA)
template <int TL, int KA>
__device__ __forceinline__ void tail_step(double (&t)[TL][TL], int k, int kr, int ti, int tj, double* pc, double* pr, int* flag) {
  if (tj == kr) PubCol<TL, KA>::go(t, pc, ti);
  if (ti == kr) PubRow<TL, KA>::go(t, pr, tj);
  __syncthreads();
  const double piv = pc[k];
  if (piv == 0.0 || isnan(piv)) { if (threadIdx.x == 0) *flag = 1; }
  const double inv = 1.0 / piv;
  // rows a > KA and columns b > KA are fully active, a == KA / b == KA are active for ti > kr / tj > kr
  double lr[TL], uc[TL];
#pragma unroll
  for (int a = KA; a < TL; ++a) lr[a] = pc[ti + 16 * a] * inv;
#pragma unroll
  for (int b = KA; b < TL; ++b) uc[b] = pr[tj + 16 * b];
  const bool rowKA = ti > kr, colKA = tj > kr;
  if (!rowKA) lr[KA] = 0.0;
  if (!colKA) uc[KA] = 0.0;
#pragma unroll
  for (int a = KA; a < TL; ++a)
#pragma unroll
    for (int b = KA; b < TL; ++b) t[a][b] -= lr[a] * uc[b];
  // column k keeps the multipliers (its uc was zero, so t was untouched above)
  if (tj == kr) {
#pragma unroll
    for (int a = KA + 1; a < TL; ++a) t[a][KA] = lr[a];
    if (rowKA) t[KA][KA] = lr[KA];
  }
}

template <int TL>
__device__ __noinline__ void tail_lu(double* Dt, int nt, int ldt, double* pub, int* flag) {
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  double t[TL][TL];
#pragma unroll
  for (int a = 0; a < TL; ++a)
#pragma unroll
    for (int b = 0; b < TL; ++b) t[a][b] = Dt[(tj + 16 * b) * ldt + ti + 16 * a];
  for (int k = 0; k < nt; ++k) {
    double* pc = pub + (k & 1) * (2 * 128 + 8);
    double* pr = pc + 128;
    const int ka = k >> 4, kr = k & 15;
    switch (ka) {   // uniform
      case 0: tail_step<TL, 0>(t, k, kr, ti, tj, pc, pr, flag); break;
      case 1: if (TL > 1) tail_step<TL, (TL > 1 ? 1 : 0)>(t, k, kr, ti, tj, pc, pr, flag); break;
      case 2: if (TL > 2) tail_step<TL, (TL > 2 ? 2 : 0)>(t, k, kr, ti, tj, pc, pr, flag); break;
      case 3: if (TL > 3) tail_step<TL, (TL > 3 ? 3 : 0)>(t, k, kr, ti, tj, pc, pr, flag); break;
      case 4: if (TL > 4) tail_step<TL, (TL > 4 ? 4 : 0)>(t, k, kr, ti, tj, pc, pr, flag); break;
      case 5: if (TL > 5) tail_step<TL, (TL > 5 ? 5 : 0)>(t, k, kr, ti, tj, pc, pr, flag); break;
      case 6: if (TL > 6) tail_step<TL, (TL > 6 ? 6 : 0)>(t, k, kr, ti, tj, pc, pr, flag); break;
      default: if (TL > 7) tail_step<TL, (TL > 7 ? 7 : 0)>(t, k, kr, ti, tj, pc, pr, flag); break;
    }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < TL; ++a)
#pragma unroll
    for (int b = 0; b < TL; ++b) Dt[(tj + 16 * b) * ldt + ti + 16 * a] = t[a][b];
  __syncthreads();
}
__global__ void __launch_bounds__(NT, 1) k(int reps) {
  extern __shared__ double sm[];
  __shared__ int flag;
  const int nt = 112, ldt = 113;
  double* Dt = sm; double* pub = sm + ldt * nt;
  long long tot = 0;
  for (int r = 0; r < reps; ++r) {
    for (int e = threadIdx.x; e < ldt * nt; e += NT) Dt[e] = ((e % (ldt + 1)) == 0 ? 10.0 : 0.01 * ((e * 7) % 13));
    __syncthreads();
    long long t0 = clock64();
    tail_lu<7>(Dt, nt, ldt, pub, &flag);
    tot += clock64() - t0;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("tail_lu<7>: %.0f cycles per factorisation (%.1f per step), D[5]=%g\n", (double)tot / reps, (double)tot / reps / nt, Dt[5]);
}
int main() {
  size_t smem = (113 * 112 + 600) * 8;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<148, NT, smem>>>(50); cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
