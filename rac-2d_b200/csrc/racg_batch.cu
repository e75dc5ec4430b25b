// racg_batch.cu -- batched stand-alone kernels of libracg (sm_100a):
//   K1 rates_kernel     chem_cal_rates for every cell      (src/chemistry.f90:591-966)
//   K2 rhs_kernel       chem_ode_f for every cell          (src/disk.f90:4569-4659)
//   K3 jac_kernel       whole chem_ode_jac for every cell  (src/disk.f90:4746-4903)
// Layout: every array is [item][cell] (Fortran a(ncell, item)), so that a warp
// working on 32 consecutive cells of one item reads/writes 256 contiguous bytes.
#include <cuda_runtime.h>
#include "racg_dev.cuh"
#include "racg_rates.cuh"

namespace racg {

// ---------------------------------------------------------------------------
// K1: thread = cell (coalesced along cells), loop over reactions.  FP64/SFU bound.
__global__ void __launch_bounds__(128)
rates_kernel(const DevNet net, int ncell, const double* __restrict__ cellpar, double* __restrict__ rates) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  CellCommon cc;
  cell_common(net.cfg, [&](int k) { return cellpar[(size_t)k * ncell + cell]; }, cc);
  for (int r = 0; r < net.R; ++r) rates[(size_t)r * ncell + cell] = rate_coeff(net, cc, r);
  // duplicate-set resolution (zeroing commutes, see racg_rates.cuh)
  for (int d = 0; d < net.ndup; ++d)
    resolve_dupli(net, cc.Tgas, d, [&](int z) { rates[(size_t)z * ncell + cell] = 0.0; });
}

// ---------------------------------------------------------------------------
// K2: RHS for a tile of TC cells per CTA.
//   phase 1: flux[r][c] for all reactions of the tile -> shared memory
//            (rates streamed once from HBM, 8*TC contiguous bytes per reaction)
//   phase 2: segmented-ELL gather over species (net stoichiometry), writes ydot
template <int TC>
__global__ void __launch_bounds__(256)
rhs_kernel(const DevNet net, int ncell, const double* __restrict__ cellpar,
           const double* __restrict__ y, const double* __restrict__ rates, double* __restrict__ ydot) {
  extern __shared__ __align__(16) double sm[];
  const int n = net.n, R = net.R, NEQ = net.NEQ;
  double* ys = sm;                       // [n][TC]
  double* fx = ys + (size_t)n * TC;      // [R][TC]
  double* px = fx + (size_t)R * TC;      // [npartial][TC]
  double* ds = px + (size_t)net.rhs.npartial * TC;  // [TC]
  const int tid = threadIdx.x, NTH = blockDim.x;
  const int c = tid % TC, q = tid / TC, NQ = NTH / TC;
  for (int tile = blockIdx.x; tile * TC < ncell; tile += gridDim.x) {
    const int cell = tile * TC + c;
    const bool ok = cell < ncell;
    __syncthreads();
    for (int i = q; i < n; i += NQ) ys[(size_t)i * TC + c] = ok ? y[(size_t)i * ncell + cell] : 0.0;
    if (q == 0) ds[c] = ok ? cellpar[(size_t)RACG_P_ratioDust2HnucNum * ncell + cell] *
                              cellpar[(size_t)RACG_P_SitesPerGrain * ncell + cell] : 1.0;
    __syncthreads();
    const double DS = ds[c];
    for (int r = q; r < R; r += NQ) {
      const uint32_t w = __ldg(net.fw + r);
      const double k = ok ? __ldcs(rates + (size_t)r * ncell + cell) : 0.0;
      // flux_of with the tile-strided y
      const int kind = (w >> 20) & 3;
      const double y1 = ys[(size_t)(w & 1023) * TC + c];
      double f;
      if (kind == FK_ONE) f = k * y1;
      else if (kind == FK_TWO) {
        const double y2 = ys[(size_t)((w >> 10) & 1023) * TC + c];
        f = k * y1 * y2;
        if (y1 < 0.0 && y2 < 0.0) f = -f;
      } else if (kind == FK_SAT) {
        const double tmp1 = DS * net.sat_c[w >> 22];
        if (tmp1 <= 0.0) f = k;
        else { const double tmp = y1 / tmp1; f = (tmp <= 1e-4) ? k * tmp : k * (1.0 - exp(-tmp)); }
      } else f = 0.0;
      fx[(size_t)r * TC + c] = f;
    }
    __syncthreads();
    // gather: work item = (sub-row, cell)
    const GatherDev& g = net.rhs;
    const int nsub = g.nblk * 32;
    for (int s = q; s < nsub; s += NQ) {
      const int t = g.sub_target[s];
      if (t == -1) continue;
      const int b = s >> 5, l = s & 31;
      const uint32_t* e = g.ent + g.blk_off[b] + l;
      // true length of this sub-row is unknown here; padded entries carry coef 0
      const int width = g.blk_width[b];
      double acc = 0.0;
      for (int j = 0; j < width; ++j) {
        const uint32_t v = __ldg(e + j * 32);
        const int cf = (int)(v >> 24) - 4;
        if (cf != 0) acc += (double)cf * fx[(size_t)(v & 0xffffffu) * TC + c];
      }
      if (t >= 0) { if (ok) __stcs(ydot + (size_t)t * ncell + cell, acc); }
      else px[(size_t)(-2 - t) * TC + c] = acc;
    }
    __syncthreads();
    for (int m = q; m < g.ncombine; m += NQ) {
      double sacc = 0.0;
      for (int p = g.comb_ptr[m]; p < g.comb_ptr[m + 1]; ++p) sacc += px[(size_t)p * TC + c];
      if (ok) __stcs(ydot + (size_t)g.comb_row[m] * ncell + cell, sacc);
    }
    if (ok && q == 0) ydot[(size_t)(NEQ - 1) * ncell + cell] = 0.0;   // T slot, evolT = .false.
  }
}

// ---------------------------------------------------------------------------
// K3: whole Jacobian in the user's CSC slot order, pd[slot][cell].
// A warp owns 32 consecutive cells; the CTA walks the columns of the Jacobian in
// groups.  For a column j only reactions that consume j contribute, so per group
//   phase 1: partial derivatives d[e][c] of the group's (reaction, reactant) pairs
//   phase 2: every CSC slot of the group gathers sum coef*d and is written as one
//            256-byte coalesced row.
// Tables: jc (column-group gather schedule) built on the host.

__global__ void __launch_bounds__(256)
jac_kernel(const DevNet net, const JacColTables jc, int ncell, const double* __restrict__ cellpar,
           const double* __restrict__ y, const double* __restrict__ rates, double* __restrict__ pd) {
  extern __shared__ __align__(16) double sm[];
  const int NWARP = blockDim.x >> 5;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  double* dbuf = sm;   // [max_pairs][32]
  for (int tile = blockIdx.x; tile * 32 < ncell; tile += gridDim.x) {
    const int cell = tile * 32 + l;
    const bool ok = cell < ncell;
    const double DS = ok ? cellpar[(size_t)RACG_P_ratioDust2HnucNum * ncell + cell] *
                           cellpar[(size_t)RACG_P_SitesPerGrain * ncell + cell] : 1.0;
    for (int z = w; z < jc.nzero; z += NWARP) if (ok) __stcs(pd + (size_t)jc.zero_slots[z] * ncell + cell, 0.0);
    for (int g = 0; g < jc.ngroups; ++g) {
      __syncthreads();
      const int pb = jc.grp_pair_ptr[g], pe = jc.grp_pair_ptr[g + 1];
      for (int p = pb + w; p < pe; p += NWARP) {
        const uint32_t pr = __ldg(jc.pair + p);
        const int r = pr & 0xffff, which = pr >> 16;
        const uint32_t fwv = __ldg(net.fw + r);
        const int kind = (fwv >> 20) & 3;
        const int r1 = fwv & 1023, r2 = (fwv >> 10) & 1023;
        const double k = ok ? __ldg(rates + (size_t)r * ncell + cell) : 0.0;
        double d = 0.0;
        if (kind == FK_ONE) d = k;
        else if (kind == FK_TWO) {
          const double y1 = ok ? __ldg(y + (size_t)r1 * ncell + cell) : 0.0;
          const double y2 = (r2 == r1) ? y1 : (ok ? __ldg(y + (size_t)r2 * ncell + cell) : 0.0);
          if (r1 != r2) d = (which == 0) ? k * y2 : k * y1;
          else d = 2.0 * k * y2;
          if (y1 < 0.0 && y2 < 0.0) d = -d;
        } else if (kind == FK_SAT) {
          const double tmp2 = DS * net.sat_c[fwv >> 22];
          if (tmp2 > 0.0) {
            const double tmp1 = 1.0 / tmp2;
            const double y1 = ok ? __ldg(y + (size_t)r1 * ncell + cell) : 0.0;
            const double tmp = y1 * tmp1;
            d = (tmp <= 1e-4) ? k * tmp1 : k * tmp1 * exp(-tmp);
          }
        }
        dbuf[(size_t)(p - pb) * 32 + l] = d;
      }
      __syncthreads();
      const int sb = jc.grp_slot_ptr[g], se = jc.grp_slot_ptr[g + 1];
      const bool accum = jc.grp_accum[g] != 0;
      for (int s = sb + w; s < se; s += NWARP) {
        double acc = (accum && ok) ? pd[(size_t)jc.slot_id[s] * ncell + cell] : 0.0;
        for (int e = jc.slot_ent_ptr[s]; e < jc.slot_ent_ptr[s + 1]; ++e) {
          const uint32_t v = __ldg(jc.ent + e);
          acc += (double)((int)(v >> 24) - 4) * dbuf[(size_t)(v & 0xffffffu) * 32 + l];
        }
        if (ok) __stcs(pd + (size_t)jc.slot_id[s] * ncell + cell, acc);
      }
    }
  }
}

cudaError_t launch_rates(const DevNet& net, int ncell, const double* cellpar, double* rates, cudaStream_t st) {
  rates_kernel<<<(ncell + 127) / 128, 128, 0, st>>>(net, ncell, cellpar, rates);
  return cudaGetLastError();
}

constexpr int RHS_TC = 4;
size_t rhs_smem_bytes(const DevNet& net) {
  return ((size_t)net.n + net.R + net.rhs.npartial + 1) * RHS_TC * sizeof(double);
}
cudaError_t launch_rhs(const DevNet& net, int ncell, const double* cellpar, const double* y, const double* rates,
                       double* ydot, int nsm, cudaStream_t st) {
  const size_t smem = rhs_smem_bytes(net);
  cudaError_t e = cudaFuncSetAttribute(rhs_kernel<RHS_TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int ntiles = (ncell + RHS_TC - 1) / RHS_TC;
  int grid = ntiles < nsm ? ntiles : nsm;
  rhs_kernel<RHS_TC><<<grid, 256, smem, st>>>(net, ncell, cellpar, y, rates, ydot);
  return cudaGetLastError();
}

cudaError_t launch_jac(const DevNet& net, const JacColTables& jc, int ncell, const double* cellpar, const double* y,
                       const double* rates, double* pd, int nsm, cudaStream_t st) {
  const size_t smem = (size_t)jc.max_pairs * 32 * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(jac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int ntiles = (ncell + 31) / 32;
  int per_sm = (int)(200 * 1024 / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int grid = ntiles < nsm * per_sm ? ntiles : nsm * per_sm;
  jac_kernel<<<grid, 256, smem, st>>>(net, jc, ncell, cellpar, y, rates, pd);
  return cudaGetLastError();
}

}  // namespace racg
