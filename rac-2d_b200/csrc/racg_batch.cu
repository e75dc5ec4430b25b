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
//   * a CTA of 32 warps owns a tile of 32 consecutive cells: lane = cell, so every global access is
//     a 256-byte row segment of the [item][cell] arrays and every shared-memory access is
//     conflict-free;
//   * the rate coefficients -- 84 % of the bytes -- stream through a ring of 2 shared-memory stages
//     of RC = 384 reactions x 32 cells, each filled by TMA tile loads (cp.async.bulk.tensor.2d over a
//     tensor map of rates[R][ncell], box 32 x 128, out-of-range rows/cells zero-filled, completion
//     counted on an mbarrier) issued one chunk ahead of the arithmetic, across tile boundaries;
//   * per chunk the fluxes are formed in place in the stage from host-built lists sorted by flux
//     kind (no branches in the loops; the abundances come straight from global memory: the tile's
//     y rows are L1/L2 hits);
//   * then every warp adds the fluxes to the register accumulators of the <= SPW species it owns
//     (HostNet::RhsChunks run lists: byte offsets of the chunk's rows, 16-byte groups, consumed
//     terms then produced terms, each in reaction order);
//   * ydot leaves as one coalesced row segment per species.
// Persistent grid: one CTA per SM, tiles dealt round-robin.
// Measured on the way (75 776 cells): one cp.async.bulk per 256-byte row segment 7.1 ms (thousands
// of small copies per tile saturate the copy engine); 32-cell tiles with 128-reaction chunks and y
// in shared memory 5.2 ms (12 k runs of 1.5 terms per tile: 27 k warp-instructions per cell);
// 16-cell tiles with 640-reaction chunks 3.25 ms (half-warps diverge, 21 k per cell).
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

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int K2_TC = 32;          // cells per tile
// Round-2 additions (ncu: the flux phase and the run lists were stalled on dependent global loads):
//   * the warp's flux words of a chunk are held one per lane (reaction f0 + w + 32 t in lane t),
//     requested a chunk ahead and broadcast by shuffle, so the abundance loads of a batch issue at once;
//   * the warp's run list of the chunk (headers + entry groups, contiguous in rc.stream) is copied
//     into a private shared-memory slice with cp.async before the flux phase and read from there
//     (the first `capl` 16-byte groups; the rest of a very long hub list still comes from L2);
//   * the per-(warp, chunk) offsets / counts and the chunk's list bounds sit in shared memory.
template <int SPW>
__global__ void __launch_bounds__(1024, 1)
rhs_stream_kernel(const DevNet net, const RhsChunkDev rc, const __grid_constant__ CUtensorMap rates_map, int ncell,
                  int nstage, int capl, const double* __restrict__ cellpar, const double* __restrict__ y,
                  double* __restrict__ ydot) {
  extern __shared__ __align__(128) unsigned char smb[];
  const int NEQ = net.NEQ, RC = rc.RC, nchunk = rc.nchunk;
  const int STG = (RC + 1) * 256;                        // bytes per stage: RC rows + one zero row
  unsigned char* const kb0 = smb;
  uint64_t* const bars = (uint64_t*)(kb0 + (size_t)nstage * STG);
  int* const meta = (int*)(bars + 8);                    // [32 warps][nchunk][2]: list offset (words), nrun | len4 << 8
  int* const flo = meta + 64 * nchunk;                   // [nchunk][4]: bounds of the chunk's ONE / TWO / SAT flux lists
  uint4* const lists = (uint4*)(((uintptr_t)(flo + 4 * nchunk) + 15) & ~(uintptr_t)15);   // [32][capl]
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int ntile = (ncell + 31) >> 5;
  const int my_tiles = (ntile > (int)blockIdx.x) ? (ntile - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int G = my_tiles * nchunk;                       // chunks this CTA will consume
  if (tid == 0) {
    for (int s2 = 0; s2 < nstage; ++s2) mbar_init(bars + s2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) for (int s2 = 0; s2 < nstage; ++s2) ((double*)(kb0 + (size_t)s2 * STG))[RC * 32 + tid] = 0.0;
  for (int i = tid; i < 32 * nchunk; i += 1024) {
    meta[2 * i] = (int)__ldg(rc.off + i);
    meta[2 * i + 1] = __ldg(rc.nrun + i) | (__ldg(rc.len4 + i) << 8);
  }
  for (int i = tid; i < 4 * nchunk; i += 1024) flo[i] = __ldg(rc.fl_off + i);
  __syncthreads();
  const int* const mymeta = meta + 2 * w * nchunk;
  uint4* const mylist = lists + (size_t)w * capl;
  const uint4* const stream4 = (const uint4*)rc.stream;
  // producer (thread 0): chunk g of this CTA's stream -> stage g % nstage, RC/128 TMA boxes of 128 x 32
  auto issue = [&](int g) {
    const int ti = g / nchunk, c = g - ti * nchunk;
    const int cell0 = ((int)blockIdx.x + ti * (int)gridDim.x) * 32;
    const int s2 = g % nstage;
    mbar_expect_tx(bars + s2, (uint32_t)(RC * 256));       // whole boxes land, zero-filled where out of range
    for (int b = 0; b < RC / 128; ++b)
      tma_load_2d(kb0 + (size_t)s2 * STG + (size_t)b * 128 * 256, &rates_map, cell0, c * RC + b * 128, bars + s2);
  };
  if (tid == 0) for (int g = 0; g < nstage - 1 && g < G; ++g) issue(g);
  // flux word of this lane for chunk c: reaction f0 + w + 32 * lane of the chunk's list
  auto flux_word = [&](int c) -> uint32_t {
    const int i = flo[4 * c] + w + 32 * lane;
    return (i < flo[4 * c + 3]) ? __ldg(rc.flux + i) : 0u;
  };
  uint32_t fw_next = (my_tiles > 0) ? flux_word(0) : 0u;
  for (int ti = 0; ti < my_tiles; ++ti) {
    const int cell0 = ((int)blockIdx.x + ti * (int)gridDim.x) * 32;
    const int cell = cell0 + lane;
    const bool ok = cell < ncell;
    const double* const yc = y + (ok ? cell : 0);          // this lane's column of y
    const double DS = ok ? cellpar[(size_t)RACG_P_ratioDust2HnucNum * ncell + cell] *
                           cellpar[(size_t)RACG_P_SitesPerGrain * ncell + cell] : 1.0;
    const unsigned ncu = (unsigned)ncell;
    double acc[SPW];
#pragma unroll
    for (int k = 0; k < SPW; ++k) acc[k] = 0.0;
    for (int c = 0; c < nchunk; ++c) {
      const int g = ti * nchunk + c, s2 = g % nstage;
      if (tid == 0 && g + nstage - 1 < G) {
        // the stage about to be refilled was last written through the generic proxy (fluxes)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(g + nstage - 1);
      }
      // this chunk's flux words are here; the next chunk's are requested now
      const uint32_t fw_cur = fw_next;
      fw_next = flux_word(c + 1 < nchunk ? c + 1 : 0);
      // the warp's run list of the chunk -> its shared-memory slice
      const int loff4 = mymeta[2 * c] >> 2, nr = mymeta[2 * c + 1] & 255, len4 = mymeta[2 * c + 1] >> 8;
      {
        const int ncopy = len4 < capl ? len4 : capl;
        for (int i = lane; i < ncopy; i += 32) cp_async16(mylist + i, stream4 + loff4 + i);
      }
      mbar_wait(bars + s2, (uint32_t)((g / nstage) & 1));
      unsigned char* const kbb = kb0 + (size_t)s2 * STG + lane * 8;     // this lane's column of the stage
      // ---- fluxes in place (branches of chem_ode_f, src/disk.f90:4583-4643), one reaction per warp
      {
        const int f0 = flo[4 * c], f1 = flo[4 * c + 1], f2 = flo[4 * c + 2], f3 = flo[4 * c + 3];
        const int base = f0 + w;
        auto first_t = [&](int a) { const int d = a - base; return d > 0 ? (d + 31) >> 5 : 0; };
#pragma unroll 4
        for (int t = 0; base + 32 * t < f1; ++t) {           // k * y1
          const uint32_t v = __shfl_sync(0xffffffffu, fw_cur, t);
          double* kp = (double*)(kbb + (v & 511u) * 256);
          *kp = *kp * __ldg(yc + (size_t)((unsigned long long)((v >> 9) & 1023u) * ncu));
        }
#pragma unroll 4
        for (int t = first_t(f1); base + 32 * t < f2; ++t) { // k * y1 * y2, sign as the reference
          const uint32_t v = __shfl_sync(0xffffffffu, fw_cur, t);
          double* kp = (double*)(kbb + (v & 511u) * 256);
          const double y1 = __ldg(yc + (size_t)((unsigned long long)((v >> 9) & 1023u) * ncu));
          const double y2 = __ldg(yc + (size_t)((unsigned long long)(v >> 19) * ncu));
          double f = *kp * y1 * y2;
          if (y1 < 0.0 && y2 < 0.0) f = -f;
          *kp = f;
        }
        for (int t = first_t(f2); base + 32 * t < f3; ++t) { // saturating desorption
          const uint32_t v = __shfl_sync(0xffffffffu, fw_cur, t);
          double* kp = (double*)(kbb + (v & 511u) * 256);
          const double k = *kp, y1 = __ldg(yc + (size_t)((unsigned long long)((v >> 9) & 1023u) * ncu));
          const double tmp1 = DS * net.sat_c[v >> 19];
          double f = k;
          if (tmp1 > 0.0) { const double tmp = y1 / tmp1; f = (tmp <= 1e-4) ? k * tmp : k * (1.0 - exp(-tmp)); }
          *kp = f;
        }
      }
      cp_async_wait_all();
      __syncthreads();
      // ---- this warp's species: run lists of the chunk (headers, then the entry groups)
      {
        const uint32_t* const hp = (const uint32_t*)mylist;
        const uint4* const gl = stream4 + loff4;
        int j = (nr + 3) >> 2;                               // entry groups follow the padded headers
        auto group = [&](int jj) -> uint4 { return jj < capl ? mylist[jj] : __ldg(gl + jj); };
        for (int q = 0; q < nr; ++q) {
          const uint32_t h = hp[q];
          const int k = h & 31, nm = (h >> 5) & 0x1fff, np = h >> 18;
          double a = 0.0;
          for (int i = 0; i < nm; ++i) {
            const uint4 e = group(j++);
            a -= *(const double*)(kbb + e.x); a -= *(const double*)(kbb + e.y);
            a -= *(const double*)(kbb + e.z); a -= *(const double*)(kbb + e.w);
          }
          for (int i = 0; i < np; ++i) {
            const uint4 e = group(j++);
            a += *(const double*)(kbb + e.x); a += *(const double*)(kbb + e.y);
            a += *(const double*)(kbb + e.z); a += *(const double*)(kbb + e.w);
          }
          switch (k) {     // uniform across the warp; keeps acc[] in registers
#define K2_CASE(n) case n: if (n < SPW) acc[n < SPW ? n : 0] += a; break;
            K2_CASE(0) K2_CASE(1) K2_CASE(2) K2_CASE(3) K2_CASE(4) K2_CASE(5) K2_CASE(6) K2_CASE(7)
            K2_CASE(8) K2_CASE(9) K2_CASE(10) K2_CASE(11) K2_CASE(12) K2_CASE(13) K2_CASE(14) K2_CASE(15)
            K2_CASE(16) K2_CASE(17) K2_CASE(18) K2_CASE(19) K2_CASE(20) K2_CASE(21) K2_CASE(22) K2_CASE(23)
#undef K2_CASE
            default: break;
          }
        }
      }
      __syncthreads();     // every warp is done with the stage: it may be refilled
    }
    if (ok) {
#pragma unroll
      for (int k = 0; k < SPW; ++k) {
        const int sp = (k < rc.spw) ? __ldg(rc.slot_species + w * rc.spw + k) : -1;
        if (sp >= 0) __stcs(ydot + (size_t)sp * ncell + cell, acc[k]);
      }
      if (w == 0) ydot[(size_t)(NEQ - 1) * ncell + cell] = 0.0;   // T slot, evolT = .false.
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

// K3, wide variant (ncell % CPL == 0): a lane owns CPL consecutive cells (16- or 32-byte vector
// loads and stores, 32*CPL*8 contiguous bytes per warp access); a CTA of NTH threads owns a tile of
// 32*CPL cells.  <4, 1024>: one CTA per SM, every Jacobian slot row written in 1 KB pieces;
// <2, 512>: two CTAs per SM, so that one CTA's derivative phase (global loads) overlaps the
// other's gather/store phase across the group barriers.
template <int CPL> struct DV { double v[CPL]; };
template <int CPL>
__device__ __forceinline__ DV<CPL> ldv(const double* p, bool ok) {
  DV<CPL> r;
  if (ok) {
#pragma unroll
    for (int q = 0; q < CPL; q += 2) { const double2 a = __ldg((const double2*)p + (q >> 1)); r.v[q] = a.x; r.v[q + 1] = a.y; }
  } else {
#pragma unroll
    for (int q = 0; q < CPL; ++q) r.v[q] = 0.0;
  }
  return r;
}
template <int CPL, int NTH>
__global__ void __launch_bounds__(NTH, 2048 / NTH > 2 ? 2 : 2048 / NTH / (CPL == 4 ? 2 : 1))
jac_kernel_wide(const DevNet net, const JacColTables jc, int ncell, const double* __restrict__ cellpar,
                const double* __restrict__ y, const double* __restrict__ rates, double* __restrict__ pd) {
  extern __shared__ __align__(16) double sm[];
  constexpr int TC = 32 * CPL;
  const int NWARP = NTH >> 5;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  double* dbuf = sm;   // [max_pairs][TC]
  for (int tile = blockIdx.x; tile * TC < ncell; tile += gridDim.x) {
    const int cell = tile * TC + CPL * l;
    const bool ok = cell < ncell;          // ncell % CPL == 0: all cells of a lane or none
    const DV<CPL> a1 = ldv<CPL>(cellpar + (size_t)RACG_P_ratioDust2HnucNum * ncell + cell, ok);
    const DV<CPL> a2 = ldv<CPL>(cellpar + (size_t)RACG_P_SitesPerGrain * ncell + cell, ok);
    double DS[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) DS[q] = ok ? a1.v[q] * a2.v[q] : 1.0;
    if (ok) {
      const double2 z2 = make_double2(0.0, 0.0);
      for (int z = w; z < jc.nzero; z += NWARP) {
        double2* o = (double2*)(pd + (size_t)jc.zero_slots[z] * ncell + cell);
#pragma unroll
        for (int q = 0; q < CPL / 2; ++q) __stcs(o + q, z2);
      }
    }
    for (int g = 0; g < jc.ngroups; ++g) {
      __syncthreads();
      const int pb = jc.grp_pair_ptr[g], pe = jc.grp_pair_ptr[g + 1];
      for (int p0 = pb + w; p0 < pe; p0 += 2 * NWARP) {
        uint32_t fwv2[2]; int which2[2]; DV<CPL> k2[2], ya2[2], yb2[2];
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
          k2[u] = ldv<CPL>(rates + (size_t)r * ncell + cell, on && ok);
          ya2[u] = ldv<CPL>(y + (size_t)r1 * ncell + cell, on && ok && kind != FK_ONE);
          yb2[u] = (kind == FK_TWO && r2 != r1) ? ldv<CPL>(y + (size_t)r2 * ncell + cell, on && ok) : ya2[u];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int p = p0 + u * NWARP;
          if (p >= pe) continue;
          const uint32_t fwv = fwv2[u];
          const int kind = (fwv >> 20) & 3, which = which2[u];
          const int r1 = fwv & 1023, r2 = (fwv >> 10) & 1023;
          double d[CPL];
          if (kind == FK_ONE) {
#pragma unroll
            for (int q = 0; q < CPL; ++q) d[q] = k2[u].v[q];
          } else if (kind == FK_TWO) {
            // d(k y1 y2)/dy: the other factor (both factors when r1 == r2), sign as chem_ode_jac
            const bool same = r1 == r2;
            const double fac = same ? 2.0 : 1.0;
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
              const double y1 = ya2[u].v[q], y2 = yb2[u].v[q];
              double dd = fac * k2[u].v[q] * ((same || which == 0) ? y2 : y1);
              if (y1 < 0.0 && y2 < 0.0) dd = -dd;
              d[q] = dd;
            }
          } else if (kind == FK_SAT) {
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
              double dd = 0.0;
              const double tmp2 = DS[q] * net.sat_c[fwv >> 22];
              if (tmp2 > 0.0) {
                const double tmp1 = 1.0 / tmp2;
                const double tmp = ya2[u].v[q] * tmp1;
                dd = (tmp <= 1e-4) ? k2[u].v[q] * tmp1 : k2[u].v[q] * tmp1 * exp(-tmp);
              }
              d[q] = dd;
            }
          } else {
#pragma unroll
            for (int q = 0; q < CPL; ++q) d[q] = 0.0;
          }
          double2* o = (double2*)(dbuf + (size_t)(p - pb) * TC + CPL * l);
#pragma unroll
          for (int q = 0; q < CPL; q += 2) o[q >> 1] = make_double2(d[q], d[q + 1]);
        }
      }
      __syncthreads();
      const int sb = jc.grp_slot_ptr[g], se = jc.grp_slot_ptr[g + 1];
      const bool accum = jc.grp_accum[g] != 0;
      for (int s = sb + w; s < se; s += NWARP) {
        const int slot = __ldg(jc.slot_id + s), e0 = __ldg(jc.slot_ent_ptr + s), e1 = __ldg(jc.slot_ent_ptr + s + 1);
        double* op = pd + (size_t)slot * ncell + cell;
        // coherent read (not the read-only path): an earlier group of this kernel stored the value
        double acc[CPL];
#pragma unroll
        for (int q = 0; q < CPL; q += 2) {
          if (accum && ok) { const double2 a = __ldcg((const double2*)op + (q >> 1)); acc[q] = a.x; acc[q + 1] = a.y; }
          else { acc[q] = 0.0; acc[q + 1] = 0.0; }
        }
        for (int e = e0; e < e1; ++e) {
          const uint32_t v = __ldg(jc.ent + e);
          const double cf = (double)((int)(v >> 24) - 4);
          const double2* dp = (const double2*)(dbuf + (size_t)(v & 0xffffffu) * TC + CPL * l);
#pragma unroll
          for (int q = 0; q < CPL; q += 2) { const double2 da = dp[q >> 1]; acc[q] += cf * da.x; acc[q + 1] += cf * da.y; }
        }
        if (ok) {
#pragma unroll
          for (int q = 0; q < CPL; q += 2) __stcs((double2*)op + (q >> 1), make_double2(acc[q], acc[q + 1]));
        }
      }
    }
  }
}

// K3, pipelined variant (the default when ncell is a multiple of 4, or even): same column groups
// as jac_kernel_wide, a lane owns CPL consecutive cells and a CTA of NTH threads a tile of 32*CPL
// cells, two CTAs per SM; but no dependent index chains and no register-staged rate loads:
//   * phase 1a: the rate rows of ALL pairs of the group are requested at once with cp.async
//     (16-byte pieces straight into the pair's row of the derivative buffer), so a whole group
//     (<= 96 rows x 1 KB) is in flight per CTA instead of two rows per warp;
//   * phase 1b: the pair words (flux word + reaction, one 8-byte load: HostNet::JacCols::pairw) are
//     read again (L1 hits), the abundances fetched PB pairs at a time, and the derivative
//     overwrites the rate in place;
//   * phase 2: listed slots heaviest first, their words (slotw: CSC slot, entry offset, count) two
//     slots ahead, the first entry group and -- for the chunks of a hub column -- the value an
//     earlier chunk stored one slot ahead; entries in groups of four (uint4), padded with +1 x a
//     zero row; an entry carries the high 16 bits of its coefficient as a double (no int->double
//     conversion in the loop).
// CPL = 4 halves the per-pair / per-slot / per-entry instruction overhead per cell against CPL = 2
// (measured with ncu: the CPL = 2 form issues 3.0 G warp-instructions for 75 776 cells and is
// issue-bound at 36 % of the HBM peak).  The group pointers sit in shared memory behind the buffer.
template <int CPL>
__device__ __forceinline__ DV<CPL> ldv_g(const double* p) {
  DV<CPL> r;
#pragma unroll
  for (int q = 0; q < CPL; q += 2) { const double2 a = __ldg((const double2*)p + (q >> 1)); r.v[q] = a.x; r.v[q + 1] = a.y; }
  return r;
}
template <int CPL>
__device__ __forceinline__ DV<CPL> ldv_s(const double* p) {
  DV<CPL> r;
#pragma unroll
  for (int q = 0; q < CPL; q += 2) { const double2 a = *((const double2*)p + (q >> 1)); r.v[q] = a.x; r.v[q + 1] = a.y; }
  return r;
}

// Phase 2 of jac_kernel_pipe for one warp: slots s, s + NWARP, ... < se of the group.  sbase = the
// shared-space address of this lane's columns of the derivative buffer; an entry = byte offset of
// the pair's row (17 bits) | the top 15 bits of its coefficient as a double.  ACCUM: a later chunk of
// a hub column adds to what an earlier group stored (coherent read, one slot ahead).
template <int CPL>
__device__ __forceinline__ void k3_fma_entry(double (&acc)[CPL], uint32_t sbase, uint32_t ev) {
  const double cf = __hiloint2double((int)(ev & 0xfffe0000u), 0);
  const uint32_t addr = sbase + (ev & 0x1ffffu);
#pragma unroll
  for (int q = 0; q < CPL; q += 2) {
    double dx, dy;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(dx), "=d"(dy) : "r"(addr + 8 * q));
    acc[q] += cf * dx; acc[q + 1] += cf * dy;
  }
}
template <int CPL, int NWARP, bool ACCUM>
__device__ __forceinline__ void k3_gather(uint32_t sbase, const uint2* __restrict__ slotw, const uint4* __restrict__ ent4,
                                          int s, int se, double* pdc, unsigned ncu, bool ok, uint2 c0, uint2 c1, uint4 e0) {
  // c0, c1 = the words of slots s and s + NWARP, e0 = the first entry group of slot s: the caller
  // requested them before the derivative phase
  const uint2 none = make_uint2(0u, 0u);
  const double2 z2 = make_double2(0.0, 0.0);
  double a0[CPL];
#pragma unroll
  for (int q = 0; q < CPL; q += 2) {
    const double2 t = (ACCUM && s < se) ? __ldcg((const double2*)(pdc + (size_t)((unsigned long long)c0.x * ncu)) + (q >> 1)) : z2;
    a0[q] = t.x; a0[q + 1] = t.y;
  }
#pragma unroll 2
  for (; s < se; s += NWARP) {
    const uint2 c2 = (s + 2 * NWARP < se) ? __ldg(slotw + s + 2 * NWARP) : none;
    const uint4 e1 = __ldg(ent4 + (c1.y & 0xffffffu));
    double a1[CPL];
#pragma unroll
    for (int q = 0; q < CPL; q += 2) {
      const double2 t = (ACCUM && s + NWARP < se) ? __ldcg((const double2*)(pdc + (size_t)((unsigned long long)c1.x * ncu)) + (q >> 1)) : z2;
      a1[q] = t.x; a1[q + 1] = t.y;
    }
    const int n4 = (int)(c0.y >> 24);
    const uint4* const ep = ent4 + (c0.y & 0xffffffu);
    uint4 e = e0;
    for (int i4 = 0;;) {
      k3_fma_entry<CPL>(a0, sbase, e.x); k3_fma_entry<CPL>(a0, sbase, e.y);
      k3_fma_entry<CPL>(a0, sbase, e.z); k3_fma_entry<CPL>(a0, sbase, e.w);
      if (++i4 >= n4) break;
      e = __ldg(ep + i4);
    }
    if (ok) {
      double2* const o = (double2*)(pdc + (size_t)((unsigned long long)c0.x * ncu));
#pragma unroll
      for (int q = 0; q < CPL; q += 2) __stcs(o + (q >> 1), make_double2(a0[q], a0[q + 1]));
    }
    c0 = c1; c1 = c2; e0 = e1;
#pragma unroll
    for (int q = 0; q < CPL; ++q) a0[q] = a1[q];
  }
}

template <int CPL, int NTH, int PB, int NBUF>
__global__ void __launch_bounds__(NTH, 1024 / NTH)
jac_kernel_pipe(const DevNet net, const JacColTables jc, int ncell, const double* __restrict__ cellpar,
                const double* __restrict__ y, const double* __restrict__ rates, double* __restrict__ pd) {
  extern __shared__ __align__(16) double sm[];
  constexpr int TC = 32 * CPL, NWARP = NTH >> 5;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  double* const dbuf = sm;                                   // NBUF x [max_pairs + 1][TC], the last row of each = 0
  const int BUFD = (jc.max_pairs + 1) * TC;                  // doubles per buffer
  int* const gp = (int*)(dbuf + NBUF * (size_t)BUFD);        // pair_ptr[ng+1] | slot_ptr[ng+1] | accum[ng] | ...
  const int ng = jc.ngroups;
  int* const gs = gp + ng + 1;
  int* const ga = gs + ng + 1;
  int* const g2 = ga + ng;            // first two-body pair
  int* const g3 = g2 + ng;            // first saturating pair
  for (int i = threadIdx.x; i <= ng; i += NTH) { gp[i] = __ldg(jc.grp_pair_ptr + i); gs[i] = __ldg(jc.grp_slot_ptr + i); }
  for (int i = threadIdx.x; i < ng; i += NTH) { ga[i] = __ldg(jc.grp_accum + i); g2[i] = __ldg(jc.grp_two_ptr + i); g3[i] = __ldg(jc.grp_sat_ptr + i); }
  const unsigned ncu = (unsigned)ncell;
  for (int i = threadIdx.x; i < NBUF * TC; i += NTH) dbuf[(size_t)(i / TC) * BUFD + (size_t)jc.max_pairs * TC + (i % TC)] = 0.0;
  const uint2* const pairw = (const uint2*)jc.pairw;
  const uint2* const slotw = (const uint2*)jc.slotw;
  const uint4* const ent4 = (const uint4*)jc.ent4;
  const double2 z2 = make_double2(0.0, 0.0);
  for (int tile = blockIdx.x; tile * TC < ncell; tile += gridDim.x) {
    const int cell = tile * TC + CPL * l;
    const bool ok = cell < ncell;                            // ncell % CPL == 0: all cells of a lane or none
    const double* const yc = y + (ok ? cell : 0);
    double* const pdc = pd + (ok ? cell : 0);
    double DS[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) DS[q] = 1.0;
    if (ok) {
      const DV<CPL> a1 = ldv_g<CPL>(cellpar + (size_t)RACG_P_ratioDust2HnucNum * ncell + cell);
      const DV<CPL> a2 = ldv_g<CPL>(cellpar + (size_t)RACG_P_SitesPerGrain * ncell + cell);
#pragma unroll
      for (int q = 0; q < CPL; ++q) DS[q] = a1.v[q] * a2.v[q];
      for (int z = w; z < jc.nzero; z += NWARP) {
        double2* const o = (double2*)(pd + (size_t)__ldg(jc.zero_slots + z) * ncell + cell);
#pragma unroll
        for (int q = 0; q < CPL / 2; ++q) __stcs(o + q, z2);
      }
    }
    // NBUF = 1: two barriers per group (the buffer is refilled after the gather).  NBUF = 2: group g
    // lives in buffer g & 1, its rate rows are requested (1a) right after the barrier that ends the
    // derivative phase of group g - 1, so they travel while that group is gathered, and there is ONE
    // barrier per group: a warp that finishes its slots early goes on to the next group's
    // derivatives in the other buffer (last read by the gather of group g - 2, which every warp
    // finished before that barrier).
    __syncthreads();                                         // the previous tile is done with the buffers (and gp[] is set)
    uint2 pwl = make_uint2(0u, 0u);                          // lane i: the word of this warp's i-th pair of the group
    auto request_rows = [&](int g, uint2 pw) {
      const int pb = gp[g], pe = gp[g + 1];
      const double* const rc = rates + (ok ? cell : 0);
      double* const dcol = dbuf + (NBUF == 2 ? (g & 1) * BUFD : 0) + CPL * l;
#pragma unroll 4
      for (int p = pb + w, i = 0; p < pe; p += NWARP, ++i) {
        const double* const src = rc + (size_t)((unsigned long long)__shfl_sync(0xffffffffu, pw.y, i) * ncu);
        double* const dst = dcol + (p - pb) * TC;
        if (ok) {
#pragma unroll
          for (int q = 0; q < CPL; q += 2) cp_async16(dst + q, src + q);
        }
      }
    };
    if (ng > 0) {
      const int q = gp[0] + w + NWARP * l;
      if (q < gp[1]) pwl = __ldg(pairw + q);
      request_rows(0, pwl);
    }
    for (int g = 0; g < ng; ++g) {
      const int pb = gp[g], pe = gp[g + 1];
      double* const myd = dbuf + (NBUF == 2 ? (g & 1) * BUFD : 0) + CPL * l;   // this lane's columns of the group's buffer
      const uint32_t sbase = smem_u32(myd);
      // the pair words of the next group: needed after the barrier
      uint2 pwn = make_uint2(0u, 0u);
      if (g + 1 < ng) {
        const int q = gp[g + 1] + w + NWARP * l;
        if (q < gp[g + 2]) pwn = __ldg(pairw + q);
      }
      // ---- 1b: derivatives in place (branches of chem_ode_jac, src/disk.f90:4765-4866).  One-body
      // pairs: the derivative is the rate, nothing to do.  Two-body pairs, PB at a time:
      // the gather phase's first slot words and entry group are requested here, behind the rate rows
      const int s0 = gs[g] + w, se = gs[g + 1];
      const uint2 c0 = (s0 < se) ? __ldg(slotw + s0) : make_uint2(0u, 0u);
      const uint2 c1 = (s0 + NWARP < se) ? __ldg(slotw + s0 + NWARP) : make_uint2(0u, 0u);
      bool landed = false;
      const int p2 = g2[g], p3 = g3[g];
      // (a pair is finished by the warp that requested its rate row: p = pb + w modulo NWARP, so that
      // the warp's own cp.async wait covers it)
      for (int p0 = p2 + ((w - (p2 - pb)) & (NWARP - 1)); p0 < p3; p0 += PB * NWARP) {
        uint32_t xw[PB]; DV<CPL> yo[PB], ys[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
          const int p = p0 + u * NWARP;
          const uint32_t xs = __shfl_sync(0xffffffffu, pwl.x, ((p - pb) / NWARP) & 31);
          xw[u] = (p < p3) ? xs : 0u;
          yo[u] = ldv_g<CPL>(yc + (size_t)((unsigned long long)(xw[u] & 1023u) * ncu));
          ys[u] = ldv_g<CPL>(yc + (size_t)((unsigned long long)((xw[u] >> 10) & 1023u) * ncu));
        }
        if (!landed) { cp_async_wait_all(); landed = true; }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
          const int p = p0 + u * NWARP;
          if (p < p3) {
            double* const kp = myd + (p - pb) * TC;
            const DV<CPL> k = ldv_s<CPL>(kp);
            const double fac = (xw[u] >> 20) ? 2.0 : 1.0;
            double d[CPL];
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
              const double dd = fac * k.v[q] * yo[u].v[q];
              d[q] = (yo[u].v[q] < 0.0 && ys[u].v[q] < 0.0) ? -dd : dd;
            }
#pragma unroll
            for (int q = 0; q < CPL; q += 2) *((double2*)kp + (q >> 1)) = make_double2(d[q], d[q + 1]);
          }
        }
      }
      // saturating desorption (a handful per network)
      for (int p = p3 + ((w - (p3 - pb)) & (NWARP - 1)); p < pe; p += NWARP) {
        const uint32_t x = __shfl_sync(0xffffffffu, pwl.x, ((p - pb) / NWARP) & 31);
        const DV<CPL> y1 = ldv_g<CPL>(yc + (size_t)((unsigned long long)(x & 1023u) * ncu));
        if (!landed) { cp_async_wait_all(); landed = true; }
        double* const kp = myd + (p - pb) * TC;
        const DV<CPL> k = ldv_s<CPL>(kp);
        const double sc = net.sat_c[x >> 10];
        double d[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          double dd = 0.0;
          const double tmp2 = DS[q] * sc;
          if (tmp2 > 0.0) {
            const double tmp1 = 1.0 / tmp2, tmp = y1.v[q] * tmp1;
            dd = (tmp <= 1e-4) ? k.v[q] * tmp1 : k.v[q] * tmp1 * exp(-tmp);
          }
          d[q] = dd;
        }
#pragma unroll
        for (int q = 0; q < CPL; q += 2) *((double2*)kp + (q >> 1)) = make_double2(d[q], d[q + 1]);
      }
      const uint4 e0 = __ldg(ent4 + (c0.y & 0xffffffu));
      if (!landed) cp_async_wait_all();
      __syncthreads();
      // ---- 2: gather the listed slots
      if (NBUF == 2 && g + 1 < ng) request_rows(g + 1, pwn);
      if (ga[g] != 0 && ok) k3_gather<CPL, NWARP, true>(sbase, slotw, ent4, s0, se, pdc, ncu, ok, c0, c1, e0);
      else k3_gather<CPL, NWARP, false>(sbase, slotw, ent4, s0, se, pdc, ncu, ok, c0, c1, e0);
      pwl = pwn;
      if (NBUF == 1 && g + 1 < ng) { __syncthreads(); request_rows(g + 1, pwn); }
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
  // streaming variant: the tensor map needs a 16-byte row pitch and base (even ncell); <= 24 species per warp
  if (rc.spw <= 24 && rc.RC % 128 == 0 && rc.RC <= 512 && (ncell & 1) == 0 && ((size_t)rates & 15) == 0) {
    const size_t stg = (size_t)(rc.RC + 1) * 256;
    const int nstage = 2;
    CUtensorMap map;
    // behind the stages: 8 mbarrier slots, the per-(warp, chunk) list words, the chunk bounds, the list slices
    const size_t fixed = nstage * stg + 64 + (size_t)(64 + 4) * rc.nchunk * sizeof(int) + 16;
    if (fixed + 32 * 16 * 8 <= 227 * 1024 && make_rates_map(&map, rates, net.R, ncell)) {
      int capl = (int)((227 * 1024 - fixed) / (32 * 16));
      if (capl > rc.max_len4) capl = rc.max_len4 > 0 ? rc.max_len4 : 1;
      const size_t smem2 = fixed + (size_t)capl * 32 * 16;
      const int ntile = (ncell + K2_TC - 1) / K2_TC;
      const int grid = ntile < nsm ? ntile : nsm;
      cudaError_t e2;
      if (rc.spw <= 16) {
        e2 = cudaFuncSetAttribute(rhs_stream_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e2 != cudaSuccess) return e2;
        rhs_stream_kernel<16><<<grid, 1024, smem2, st>>>(net, rc, map, ncell, nstage, capl, cellpar, y, ydot);
      } else {
        e2 = cudaFuncSetAttribute(rhs_stream_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e2 != cudaSuccess) return e2;
        rhs_stream_kernel<24><<<grid, 1024, smem2, st>>>(net, rc, map, ncell, nstage, capl, cellpar, y, ydot);
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

int g_k3_variant = 3;   // 3: jac_kernel_pipe (default); 2: jac_kernel_wide<2,512>; 4: jac_kernel_wide<4,1024> (A/B measurements)
cudaError_t launch_jac(const DevNet& net, const JacColTables& jc, int ncell, const double* cellpar, const double* y,
                       const double* rates, double* pd, int nsm, cudaStream_t st) {
  const bool aligned = ((size_t)pd % 32) == 0 && ((size_t)y % 32) == 0 && ((size_t)rates % 32) == 0 && ((size_t)cellpar % 32) == 0;
  // two 64-cell CTAs per SM (lane = 2 cells) when their derivative buffers fit side by side
  const size_t smem2 = (size_t)jc.max_pairs * 64 * sizeof(double);
  // pipelined kernel, 64-cell tiles (lane = 2 cells): two 512-thread CTAs per SM with one derivative buffer each
  // (variant 3, default), or one 1024-thread CTA per SM with two buffers and one barrier per group (variant 5)
  const size_t smemp1 = ((size_t)jc.max_pairs + 1) * 64 * sizeof(double) + (5 * (size_t)jc.ngroups + 2) * sizeof(int);
  const size_t smemp2 = smemp1 + ((size_t)jc.max_pairs + 1) * 64 * sizeof(double);
  if (g_k3_variant == 5 && ncell % 2 == 0 && aligned && smemp2 + 1024 <= 227 * 1024) {
    auto kern = jac_kernel_pipe<2, 1024, 4, 2>;
    cudaError_t e3 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemp2);
    if (e3 != cudaSuccess) return e3;
    const int ntiles = (ncell + 63) / 64;
    kern<<<ntiles < nsm ? ntiles : nsm, 1024, smemp2, st>>>(net, jc, ncell, cellpar, y, rates, pd);
    return cudaGetLastError();
  }
  if ((g_k3_variant == 3 || g_k3_variant == 5) && ncell % 2 == 0 && aligned && (smemp1 + 1024) * 2 <= 227 * 1024) {
    auto kern = jac_kernel_pipe<2, 512, 4, 1>;
    cudaError_t e3 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemp1);
    if (e3 != cudaSuccess) return e3;
    const int ntiles = (ncell + 63) / 64;
    kern<<<ntiles < 2 * nsm ? ntiles : 2 * nsm, 512, smemp1, st>>>(net, jc, ncell, cellpar, y, rates, pd);
    return cudaGetLastError();
  }
  if (g_k3_variant != 4 && ncell % 2 == 0 && aligned && (smem2 + 1024) * 2 <= 227 * 1024) {
    auto kern = jac_kernel_wide<2, 512>;
    cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e2 != cudaSuccess) return e2;
    const int ntiles = (ncell + 63) / 64;
    kern<<<ntiles < 2 * nsm ? ntiles : 2 * nsm, 512, smem2, st>>>(net, jc, ncell, cellpar, y, rates, pd);
    return cudaGetLastError();
  }
  const size_t smem4 = (size_t)jc.max_pairs * 128 * sizeof(double);
  if (ncell % 4 == 0 && smem4 <= 226 * 1024 && aligned) {
    auto kern = jac_kernel_wide<4, 1024>;
    cudaError_t e4 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
    if (e4 != cudaSuccess) return e4;
    const int ntiles4 = (ncell + 127) / 128;
    kern<<<ntiles4 < nsm ? ntiles4 : nsm, 1024, smem4, st>>>(net, jc, ncell, cellpar, y, rates, pd);
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
