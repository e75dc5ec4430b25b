// racg_batch.cu -- batched stand-alone kernels of libracg (sm_100a):
//   K1 rates_kernel     chem_cal_rates for every cell      (src/chemistry.f90:591-966)
//   K2 rhs_kernel       chem_ode_f for every cell          (src/disk.f90:4569-4659)
//   K3 jac_kernel       whole chem_ode_jac for every cell  (src/disk.f90:4746-4903)
// Layout: every array is [item][cell] (Fortran a(ncell, item)), so that a warp
// working on 32 consecutive cells of one item reads/writes 256 contiguous bytes.
#include <cuda.h>            // CUtensorMap and its enums only; the encoder is fetched at run time
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
__global__ void __launch_bounds__(1024)
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
#pragma unroll 8
    for (int i = q; i < n; i += NQ) ys[(size_t)i * TC + c] = ok ? __ldcs(y + (size_t)i * ncell + cell) : 0.0;
    if (q == 0) ds[c] = ok ? cellpar[(size_t)RACG_P_ratioDust2HnucNum * ncell + cell] *
                              cellpar[(size_t)RACG_P_SitesPerGrain * ncell + cell] : 1.0;
    __syncthreads();
    const double DS = ds[c];
    for (int r0 = q; r0 < R; r0 += 8 * NQ) {
      uint32_t wv[8]; double kv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = r0 + u * NQ;
        wv[u] = (r < R) ? __ldg(net.fw + r) : (3u << 20);
        kv[u] = (r < R && ok) ? __ldcs(rates + (size_t)r * ncell + cell) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = r0 + u * NQ;
        if (r >= R) continue;
        const uint32_t w = wv[u];
        const double k = kv[u];
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
      for (int j0 = 0; j0 < width; j0 += 8) {
        uint32_t ev[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ev[j] = (j0 + j < width) ? __ldg(e + (j0 + j) * 32) : (4u << 24);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int cf = (int)(ev[j] >> 24) - 4;
          if (cf != 0) acc += (double)cf * fx[(size_t)(ev[j] & 0xffffffu) * TC + c];
        }
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
// K2, streaming variant (the one launch_rhs uses whenever the network fits): HBM-bound design.
//   * a CTA of 32 warps owns a tile of 16 consecutive cells: half-warp lane = cell, so every global
//     access is a 128-byte row segment of the [item][cell] arrays and every shared-memory access is
//     conflict-free; the two halves of a warp work on different reactions / species;
//   * the tile's abundances y[n][16] stay in shared memory for the whole tile;
//   * the rate coefficients -- 84 % of the bytes -- stream through a ring of NSTAGE shared-memory
//     stages of RC reactions x 16 cells, each filled by TMA tile loads (cp.async.bulk.tensor.2d over
//     a tensor map of rates[R][ncell], box 16 x 128, out-of-range rows/cells zero-filled,
//     completion counted on an mbarrier) issued NSTAGE-1 chunks ahead of the arithmetic, across
//     tile boundaries;
//   * per chunk the fluxes are formed in place in the stage, then every half-warp adds them to the
//     register accumulators of the <= SPW species it owns (HostNet::RhsChunks run lists: 16-bit
//     indices of the chunk's rows, consumed terms then produced terms, each in reaction order);
//     chunks are large (RC = 640) so that a run has several terms and its overhead is amortised;
//   * ydot leaves as one coalesced row segment per species.
// Persistent grid: one CTA per SM, tiles dealt round-robin.
// (Tried first: 32-cell tiles with 128-reaction chunks -- 12 k runs of 1.5 terms per tile, 27 k
// warp-instructions per cell, 5.2 ms for 75 776 cells; and one cp.async.bulk per row segment --
// thousands of small copies per tile saturate the copy engine, 7.1 ms.)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

constexpr int K2_TC = 16;          // cells per tile
template <int SPW>
__global__ void __launch_bounds__(1024, 1)
rhs_stream_kernel(const DevNet net, const RhsChunkDev rc, const __grid_constant__ CUtensorMap rates_map, int ncell,
                  int nstage, const double* __restrict__ cellpar, const double* __restrict__ y,
                  double* __restrict__ ydot) {
  extern __shared__ __align__(128) unsigned char smb[];
  const int n = net.n, R = net.R, NEQ = net.NEQ, RC = rc.RC, nchunk = rc.nchunk;
  const int STG = (RC + 1) * (K2_TC * 8);                // bytes per stage: RC rows + one zero row
  double* const ys = (double*)smb;                       // [n][16]
  unsigned char* const kb0 = smb + (size_t)n * (K2_TC * 8);
  uint64_t* const bars = (uint64_t*)(kb0 + (size_t)nstage * STG);
  const int tid = threadIdx.x, cl = tid & (K2_TC - 1), hw = tid >> 4;     // hw: half-warp 0..63
  const int ntile = (ncell + K2_TC - 1) / K2_TC;
  const int my_tiles = (ntile > (int)blockIdx.x) ? (ntile - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int G = my_tiles * nchunk;                       // chunks this CTA will consume
  if (tid == 0) {
    for (int s2 = 0; s2 < nstage; ++s2) mbar_init(bars + s2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < K2_TC) for (int s2 = 0; s2 < nstage; ++s2) ((double*)(kb0 + (size_t)s2 * STG))[RC * K2_TC + tid] = 0.0;
  __syncthreads();
  // producer (thread 0): chunk g of this CTA's stream -> stage g % nstage, RC/128 TMA boxes of 128 x 16
  auto issue = [&](int g) {
    const int ti = g / nchunk, c = g - ti * nchunk;
    const int cell0 = ((int)blockIdx.x + ti * (int)gridDim.x) * K2_TC;
    const int s2 = g % nstage;
    mbar_expect_tx(bars + s2, (uint32_t)(RC * K2_TC * 8));    // whole boxes land, zero-filled where out of range
    for (int b = 0; b < RC / 128; ++b)
      tma_load_2d(kb0 + (size_t)s2 * STG + (size_t)b * 128 * (K2_TC * 8), &rates_map, cell0, c * RC + b * 128, bars + s2);
  };
  if (tid == 0) for (int g = 0; g < nstage - 1 && g < G; ++g) issue(g);
  for (int ti = 0; ti < my_tiles; ++ti) {
    const int cell0 = ((int)blockIdx.x + ti * (int)gridDim.x) * K2_TC;
    const int cell = cell0 + cl;
    const bool ok = cell < ncell;
    for (int i = hw; i < n; i += 64) ys[i * K2_TC + cl] = ok ? __ldcs(y + (size_t)i * ncell + cell) : 0.0;
    const double DS = ok ? cellpar[(size_t)RACG_P_ratioDust2HnucNum * ncell + cell] *
                           cellpar[(size_t)RACG_P_SitesPerGrain * ncell + cell] : 1.0;
    double acc[SPW];
#pragma unroll
    for (int k = 0; k < SPW; ++k) acc[k] = 0.0;
    __syncthreads();
    for (int c = 0; c < nchunk; ++c) {
      const int g = ti * nchunk + c, s2 = g % nstage;
      if (tid == 0 && g + nstage - 1 < G) {
        // the stage about to be refilled was last written through the generic proxy (fluxes)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(g + nstage - 1);
      }
      mbar_wait(bars + s2, (uint32_t)((g / nstage) & 1));
      double* const kb = (double*)(kb0 + (size_t)s2 * STG) + cl;
      // ---- fluxes in place (branches of chem_ode_f, src/disk.f90:4583-4643), one reaction per half-warp
      for (int rl = hw; rl < RC; rl += 64) {
        const int r = c * RC + rl;
        double f = 0.0;
        if (r < R) {
          const uint32_t fwv = __ldg(net.fw + r);
          const double k = kb[rl * K2_TC];
          const int kind = (fwv >> 20) & 3;
          const double y1 = ys[(fwv & 1023) * K2_TC + cl];
          if (kind == FK_ONE) f = k * y1;
          else if (kind == FK_TWO) {
            const double y2 = ys[((fwv >> 10) & 1023) * K2_TC + cl];
            f = k * y1 * y2;
            if (y1 < 0.0 && y2 < 0.0) f = -f;
          } else if (kind == FK_SAT) {
            const double tmp1 = DS * net.sat_c[fwv >> 22];
            if (tmp1 <= 0.0) f = k;
            else { const double tmp = y1 / tmp1; f = (tmp <= 1e-4) ? k * tmp : k * (1.0 - exp(-tmp)); }
          }
        }
        kb[rl * K2_TC] = f;
      }
      __syncthreads();
      // ---- this half-warp's species: run lists of the chunk
      {
        const uint32_t* p = rc.stream + __ldg(rc.off + hw * nchunk + c);
        const int nr = __ldg(rc.nrun + hw * nchunk + c);
        uint32_t nxt = __ldg(p);     // the stream is read one word ahead
        int q = 0;
        // runs are listed by ascending slot: a static walk over the slots keeps acc[] in registers
#pragma unroll
        for (int kk = 0; kk < SPW; ++kk) {
          if (q < nr && (int)(nxt & 31u) == kk) {
            const int nm = (nxt >> 5) & 0x1fff, np = nxt >> 18;
            nxt = __ldg(++p);
            double a = acc[kk];
            for (int i = 0; i < nm; ++i) {
              const uint32_t e = nxt; nxt = __ldg(++p);
              a -= kb[(e & 0xffffu) * K2_TC]; a -= kb[(e >> 16) * K2_TC];
            }
            for (int i = 0; i < np; ++i) {
              const uint32_t e = nxt; nxt = __ldg(++p);
              a += kb[(e & 0xffffu) * K2_TC]; a += kb[(e >> 16) * K2_TC];
            }
            acc[kk] = a;
            ++q;
          }
        }
      }
      __syncthreads();     // every warp is done with the stage: it may be refilled
    }
    if (ok) {
#pragma unroll
      for (int k = 0; k < SPW; ++k) {
        const int sp = (k < rc.spw) ? __ldg(rc.slot_species + hw * rc.spw + k) : -1;
        if (sp >= 0) __stcs(ydot + (size_t)sp * ncell + cell, acc[k]);
      }
      if (hw == 0) ydot[(size_t)(NEQ - 1) * ncell + cell] = 0.0;   // T slot, evolT = .false.
    }
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
      for (int p0 = pb + w; p0 < pe; p0 += 4 * NWARP) {
        uint32_t fwv4[4]; int which4[4]; double k4[4], ya4[4], yb4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int p = p0 + u * NWARP;
          const bool on = p < pe;
          const uint32_t pr = on ? __ldg(jc.pair + p) : 0u;
          const int r = pr & 0xffff;
          which4[u] = pr >> 16;
          fwv4[u] = on ? __ldg(net.fw + r) : (3u << 20);
          k4[u] = (on && ok) ? __ldg(rates + (size_t)r * ncell + cell) : 0.0;
          const int kind = (fwv4[u] >> 20) & 3;
          const int r1 = fwv4[u] & 1023, r2 = (fwv4[u] >> 10) & 1023;
          ya4[u] = (on && ok && kind != FK_ONE) ? __ldg(y + (size_t)r1 * ncell + cell) : 0.0;
          yb4[u] = (on && ok && kind == FK_TWO && r2 != r1) ? __ldg(y + (size_t)r2 * ncell + cell) : ya4[u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int p = p0 + u * NWARP;
          if (p >= pe) continue;
          const uint32_t fwv = fwv4[u];
          const int kind = (fwv >> 20) & 3, which = which4[u];
          const int r1 = fwv & 1023, r2 = (fwv >> 10) & 1023;
          const double k = k4[u];
          double d = 0.0;
          if (kind == FK_ONE) d = k;
          else if (kind == FK_TWO) {
            const double y1 = ya4[u], y2 = yb4[u];
            if (r1 != r2) d = (which == 0) ? k * y2 : k * y1;
            else d = 2.0 * k * y2;
            if (y1 < 0.0 && y2 < 0.0) d = -d;
          } else if (kind == FK_SAT) {
            const double tmp2 = DS * net.sat_c[fwv >> 22];
            if (tmp2 > 0.0) {
              const double tmp1 = 1.0 / tmp2;
              const double tmp = ya4[u] * tmp1;
              d = (tmp <= 1e-4) ? k * tmp1 : k * tmp1 * exp(-tmp);
            }
          }
          dbuf[(size_t)(p - pb) * 32 + l] = d;
        }
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

// K3, wide variant (ncell % 4 == 0): a lane owns 4 consecutive cells (32-byte vector loads and
// stores, 1 KB contiguous per warp access), a CTA of 32 warps a tile of 128 cells, so that every
// Jacobian slot row is written in 1 KB pieces and each instruction moves 4x the bytes.
struct D4 { double v[4]; };
__device__ __forceinline__ D4 ld4(const double* p, bool ok) {
  D4 r;
  if (ok) { const double2 a = __ldg((const double2*)p), b = __ldg((const double2*)p + 1); r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; }
  else { r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.0; }
  return r;
}
__global__ void __launch_bounds__(1024)
jac_kernel4(const DevNet net, const JacColTables jc, int ncell, const double* __restrict__ cellpar,
            const double* __restrict__ y, const double* __restrict__ rates, double* __restrict__ pd) {
  extern __shared__ __align__(16) double sm[];
  const int NWARP = blockDim.x >> 5;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  double* dbuf = sm;   // [max_pairs][128]
  for (int tile = blockIdx.x; tile * 128 < ncell; tile += gridDim.x) {
    const int cell = tile * 128 + 4 * l;
    const bool ok = cell < ncell;          // ncell % 4 == 0: all four cells or none
    const D4 a1 = ld4(cellpar + (size_t)RACG_P_ratioDust2HnucNum * ncell + cell, ok);
    const D4 a2 = ld4(cellpar + (size_t)RACG_P_SitesPerGrain * ncell + cell, ok);
    double DS[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) DS[q] = ok ? a1.v[q] * a2.v[q] : 1.0;
    if (ok) {
      const double2 z2 = make_double2(0.0, 0.0);
      for (int z = w; z < jc.nzero; z += NWARP) {
        double2* o = (double2*)(pd + (size_t)jc.zero_slots[z] * ncell + cell);
        __stcs(o, z2); __stcs(o + 1, z2);
      }
    }
    for (int g = 0; g < jc.ngroups; ++g) {
      __syncthreads();
      const int pb = jc.grp_pair_ptr[g], pe = jc.grp_pair_ptr[g + 1];
      for (int p0 = pb + w; p0 < pe; p0 += 2 * NWARP) {
        uint32_t fwv2[2]; int which2[2]; D4 k2[2], ya2[2], yb2[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int p = p0 + u * NWARP;
          const bool on = p < pe;
          const uint32_t pr = on ? __ldg(jc.pair + p) : 0u;
          const int r = pr & 0xffff;
          which2[u] = pr >> 16;
          fwv2[u] = on ? __ldg(net.fw + r) : (3u << 20);
          const int kind = (fwv2[u] >> 20) & 3;
          const int r1 = fwv2[u] & 1023, r2 = (fwv2[u] >> 10) & 1023;
          k2[u] = ld4(rates + (size_t)r * ncell + cell, on && ok);
          ya2[u] = ld4(y + (size_t)r1 * ncell + cell, on && ok && kind != FK_ONE);
          yb2[u] = (kind == FK_TWO && r2 != r1) ? ld4(y + (size_t)r2 * ncell + cell, on && ok) : ya2[u];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int p = p0 + u * NWARP;
          if (p >= pe) continue;
          const uint32_t fwv = fwv2[u];
          const int kind = (fwv >> 20) & 3, which = which2[u];
          const int r1 = fwv & 1023, r2 = (fwv >> 10) & 1023;
          double d[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double k = k2[u].v[q];
            double dd = 0.0;
            if (kind == FK_ONE) dd = k;
            else if (kind == FK_TWO) {
              const double y1 = ya2[u].v[q], y2 = yb2[u].v[q];
              if (r1 != r2) dd = (which == 0) ? k * y2 : k * y1;
              else dd = 2.0 * k * y2;
              if (y1 < 0.0 && y2 < 0.0) dd = -dd;
            } else if (kind == FK_SAT) {
              const double tmp2 = DS[q] * net.sat_c[fwv >> 22];
              if (tmp2 > 0.0) {
                const double tmp1 = 1.0 / tmp2;
                const double tmp = ya2[u].v[q] * tmp1;
                dd = (tmp <= 1e-4) ? k * tmp1 : k * tmp1 * exp(-tmp);
              }
            }
            d[q] = dd;
          }
          double2* o = (double2*)(dbuf + (size_t)(p - pb) * 128 + 4 * l);
          o[0] = make_double2(d[0], d[1]); o[1] = make_double2(d[2], d[3]);
        }
      }
      __syncthreads();
      const int sb = jc.grp_slot_ptr[g], se = jc.grp_slot_ptr[g + 1];
      const bool accum = jc.grp_accum[g] != 0;
      for (int s = sb + w; s < se; s += NWARP) {
        const int slot = __ldg(jc.slot_id + s), e0 = __ldg(jc.slot_ent_ptr + s), e1 = __ldg(jc.slot_ent_ptr + s + 1);
        double* op = pd + (size_t)slot * ncell + cell;
        // coherent read (not the read-only path): an earlier group of this kernel stored the value
        D4 acc;
        if (accum && ok) { const double2 a = __ldcg((const double2*)op), b2 = __ldcg((const double2*)op + 1); acc.v[0] = a.x; acc.v[1] = a.y; acc.v[2] = b2.x; acc.v[3] = b2.y; }
        else { acc.v[0] = acc.v[1] = acc.v[2] = acc.v[3] = 0.0; }
        for (int e = e0; e < e1; ++e) {
          const uint32_t v = __ldg(jc.ent + e);
          const double cf = (double)((int)(v >> 24) - 4);
          const double2* dp = (const double2*)(dbuf + (size_t)(v & 0xffffffu) * 128 + 4 * l);
          const double2 da = dp[0], db = dp[1];
          acc.v[0] += cf * da.x; acc.v[1] += cf * da.y; acc.v[2] += cf * db.x; acc.v[3] += cf * db.y;
        }
        if (ok) { __stcs((double2*)op, make_double2(acc.v[0], acc.v[1])); __stcs((double2*)op + 1, make_double2(acc.v[2], acc.v[3])); }
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
// tensor map of rates[R][ncell] (f64, row pitch ncell*8) with a box of 16 cells x 128 reactions
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_rates_map(CUtensorMap* map, const double* rates, int R, int ncell) {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    else cudaGetLastError();
  }
  if (!fn) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)ncell, (cuuint64_t)R};
  const cuuint64_t gstr[1] = {(cuuint64_t)ncell * 8};
  const cuuint32_t box[2] = {(cuuint32_t)K2_TC, 128u};
  const cuuint32_t estr[2] = {1u, 1u};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)rates, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launch_rhs(const DevNet& net, const RhsChunkDev& rc, int ncell, const double* cellpar, const double* y,
                       const double* rates, double* ydot, int nsm, cudaStream_t st) {
  // streaming variant: the tensor map needs a 16-byte row pitch and base (even ncell); <= 16 species per half-warp
  if (rc.spw <= 16 && rc.RC % 128 == 0 && rc.RC < 65535 && (ncell & 1) == 0 && ((size_t)rates & 15) == 0) {
    const size_t stg = (size_t)(rc.RC + 1) * (K2_TC * 8), ybytes = (size_t)net.n * (K2_TC * 8);
    int nstage = (int)((227 * 1024 - 256 - ybytes) / stg);
    if (nstage > 3) nstage = 3;
    CUtensorMap map;
    if (nstage >= 2 && make_rates_map(&map, rates, net.R, ncell)) {
      const size_t smem2 = ybytes + nstage * stg + 64;
      const int ntile = (ncell + K2_TC - 1) / K2_TC;
      const int grid = ntile < nsm ? ntile : nsm;
      cudaError_t e2;
      if (rc.spw <= 8) {
        e2 = cudaFuncSetAttribute(rhs_stream_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e2 != cudaSuccess) return e2;
        rhs_stream_kernel<8><<<grid, 1024, smem2, st>>>(net, rc, map, ncell, nstage, cellpar, y, ydot);
      } else {
        e2 = cudaFuncSetAttribute(rhs_stream_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e2 != cudaSuccess) return e2;
        rhs_stream_kernel<16><<<grid, 1024, smem2, st>>>(net, rc, map, ncell, nstage, cellpar, y, ydot);
      }
      return cudaGetLastError();
    }
  }
  // fallback (odd ncell, very large networks): 4-cell tiles with the whole flux tile in shared memory
  if (rhs_smem_bytes(net) > 227 * 1024) return cudaErrorInvalidConfiguration;
  const size_t smem = rhs_smem_bytes(net);
  cudaError_t e = cudaFuncSetAttribute(rhs_kernel<RHS_TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int ntiles = (ncell + RHS_TC - 1) / RHS_TC;
  int grid = ntiles < nsm ? ntiles : nsm;
  rhs_kernel<RHS_TC><<<grid, 1024, smem, st>>>(net, ncell, cellpar, y, rates, ydot);
  return cudaGetLastError();
}

cudaError_t launch_jac(const DevNet& net, const JacColTables& jc, int ncell, const double* cellpar, const double* y,
                       const double* rates, double* pd, int nsm, cudaStream_t st) {
  const size_t smem4 = (size_t)jc.max_pairs * 128 * sizeof(double);
  if (ncell % 4 == 0 && smem4 <= 226 * 1024 && ((size_t)pd % 32) == 0 && ((size_t)y % 32) == 0 &&
      ((size_t)rates % 32) == 0 && ((size_t)cellpar % 32) == 0) {
    cudaError_t e4 = cudaFuncSetAttribute(jac_kernel4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
    if (e4 != cudaSuccess) return e4;
    const int ntiles4 = (ncell + 127) / 128;
    const int per_sm4 = (smem4 + 1024) * 2 <= 227 * 1024 ? 2 : 1;
    jac_kernel4<<<ntiles4 < nsm * per_sm4 ? ntiles4 : nsm * per_sm4, 1024, smem4, st>>>(net, jc, ncell, cellpar, y, rates, pd);
    return cudaGetLastError();
  }
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
