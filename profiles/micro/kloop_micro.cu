// Stand-alone timing of the sparse row-elimination k-loop of racg_integrate.cu (factor():
// "tail rows against the head pivots"), synthetic structure of the rate06 size.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
constexpr int NT = 256, NW = 8;
struct Tab { const int4* pivmeta; const uint16_t* lc_col; const int* lc_ptr; int nh, nt, n, n_hh; const uint16_t* hh_col_g; };
__global__ void __launch_bounds__(NT, 1) k(Tab tb, int reps, int variant) {
  extern __shared__ double sm[];
  const int n = tb.n, nh = tb.nh, nt = tb.nt;
  double* hh = sm; double* dinv = hh + tb.n_hh; double* X = dinv + nh;
  uint16_t* hhcol = (uint16_t*)(X + NW * n);
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  for (int q = tid; q < tb.n_hh; q += NT) { hh[q] = 1e-3 * ((q * 7) % 11); hhcol[q] = tb.hh_col_g[q]; }
  for (int q = tid; q < nh; q += NT) dinv[q] = 0.5;
  __syncthreads();
  double* wrow = X + w * n;
  long long tot = 0; long ksteps = 0;
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    long long t0 = clock64();
    for (int a = w; a < nt; a += NW) {
      const int base = tb.lc_ptr[a], nl = tb.lc_ptr[a + 1] - base;
      for (int q = l; q < n; q += 32) wrow[q] = 1.0 + q * 1e-3;
      __syncwarp();
      if (variant == 0) {
        for (int q0 = 0; q0 < nl; q0 += 32) {
          const int myk = (q0 + l < nl) ? (int)__ldg(tb.lc_col + base + q0 + l) : 0;
          const int4 mm = __ldg(tb.pivmeta + myk);
          const int cnt = (nl - q0) < 32 ? (nl - q0) : 32;
          for (int j = 0; j < cnt; ++j) {
            const int kk = __shfl_sync(0xffffffffu, myk, j);
            const int kb = __shfl_sync(0xffffffffu, mm.x, j), klen = __shfl_sync(0xffffffffu, mm.y, j);
            const double lv = wrow[kk] * dinv[kk];
            __syncwarp();
            for (int t = l; t < klen; t += 32) wrow[hhcol[kb + t]] -= lv * hh[kb + t];
            __syncwarp();
          }
        }
      } else {
        // variant 1: the U-row entries of the NEXT pivot are fetched (index + value) before the
        // current one is applied: the dependent chain per k-step shrinks to LDS w[k] -> DMUL -> (LDS w[c] -> DFMA -> STS)
        for (int q0 = 0; q0 < nl; q0 += 32) {
          const int myk = (q0 + l < nl) ? (int)__ldg(tb.lc_col + base + q0 + l) : 0;
          const int4 mm = __ldg(tb.pivmeta + myk);
          const int cnt = (nl - q0) < 32 ? (nl - q0) : 32;
          int kk = __shfl_sync(0xffffffffu, myk, 0), kb = __shfl_sync(0xffffffffu, mm.x, 0), klen = __shfl_sync(0xffffffffu, mm.y, 0);
          int c0 = (l < klen) ? hhcol[kb + l] : 0; double v0 = (l < klen) ? hh[kb + l] : 0.0;
          double di = dinv[kk];
          for (int j = 0; j < cnt; ++j) {
            const int jn = (j + 1 < cnt) ? j + 1 : j;
            const int kkn = __shfl_sync(0xffffffffu, myk, jn), kbn = __shfl_sync(0xffffffffu, mm.x, jn), klenn = __shfl_sync(0xffffffffu, mm.y, jn);
            const int c1 = (l < klenn) ? hhcol[kbn + l] : 0; const double v1 = (l < klenn) ? hh[kbn + l] : 0.0;
            const double din = dinv[kkn];
            const double lv = wrow[kk] * di;
            __syncwarp();
            if (l < klen) wrow[c0] -= lv * v0;
            for (int t = l + 32; t < klen; t += 32) wrow[hhcol[kb + t]] -= lv * hh[kb + t];
            __syncwarp();
            kk = kkn; kb = kbn; klen = klenn; c0 = c1; v0 = v1; di = din;
          }
        }
      }
      if (w == 0 && l == 0) ksteps += nl;
    }
    __syncthreads();
    tot += clock64() - t0;
  }
  if (tid == 0 && blockIdx.x == 0) printf("variant %d: %.0f cycles per pass, %.1f cycles per k-step of one warp (k-steps warp0 per pass %ld), w=%g\n",
      variant, (double)tot / reps, (double)tot / ksteps, ksteps / reps, wrow[3]);
}
int main() {
  // synthetic structure: nh = 355 head rows with U rows of 8..40 entries, nt = 112 tail rows with L_C rows of 54 pivots on average
  const int nh = 355, nt = 112, n = nh + nt;
  std::vector<int4> piv(nh); std::vector<uint16_t> hcol; std::vector<int> lptr(nt + 1, 0); std::vector<uint16_t> lcol;
  unsigned s = 12345; auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
  for (int k2 = 0; k2 < nh; ++k2) {
    int len = 8 + rnd() % 33; int start = (int)hcol.size() + 1; hcol.push_back(k2);
    for (int e = 0; e < len; ++e) hcol.push_back((uint16_t)(k2 + 1 + rnd() % (n - k2 - 1)));
    piv[k2] = make_int4(start, len, 0, 0);
  }
  for (int a = 0; a < nt; ++a) { int nl = 10 + rnd() % 90; for (int e = 0; e < nl; ++e) lcol.push_back((uint16_t)(rnd() % nh)); lptr[a + 1] = (int)lcol.size(); }
  Tab tb; tb.nh = nh; tb.nt = nt; tb.n = n; tb.n_hh = (int)hcol.size();
  int4* dp; uint16_t *dl, *dh; int* dptr;
  cudaMalloc(&dp, piv.size() * 16); cudaMemcpy(dp, piv.data(), piv.size() * 16, cudaMemcpyHostToDevice);
  cudaMalloc(&dl, lcol.size() * 2); cudaMemcpy(dl, lcol.data(), lcol.size() * 2, cudaMemcpyHostToDevice);
  cudaMalloc(&dh, hcol.size() * 2); cudaMemcpy(dh, hcol.data(), hcol.size() * 2, cudaMemcpyHostToDevice);
  cudaMalloc(&dptr, lptr.size() * 4); cudaMemcpy(dptr, lptr.data(), lptr.size() * 4, cudaMemcpyHostToDevice);
  tb.pivmeta = dp; tb.lc_col = dl; tb.lc_ptr = dptr; tb.hh_col_g = dh;
  size_t smem = (tb.n_hh + nh + NW * n) * 8 + tb.n_hh * 2 + 64;
  printf("n_hh %d, L_C entries %zu, smem %zu\n", tb.n_hh, lcol.size(), smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int v = 0; v < 2; ++v) { k<<<148, NT, smem>>>(tb, 20, v); cudaDeviceSynchronize(); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
