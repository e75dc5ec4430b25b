// racg_integrate.cu -- the persistent per-cell stiff integrator of libracg (sm_100a).
//
// One CTA (256 threads) integrates one grid cell at a time (north_star (c)); a
// persistent grid of one CTA per SM pulls cells from an atomic work queue (K7) so that
// cells of very different stiffness balance.  What it restates, per cell, is the
// reference's
//   chem_evol_solve loop            src/chemistry.f90:391-588
//   DLSODES driver (MF=21, ITASK=4) src/opkdmain.f:3069-3588
//   DSTODE / DPRJS / DSOLSS         src/opkda1.f:629-1126, 1664-1942
//   DEWSET / DVNORM / DINTDY        src/opkda1.f:1127-1209, 174-281
//   chem_ode_f / chem_ode_jac       src/disk.f90:4569-4659, 4746-4903 (evolT=.false.)
//   chem_cal_rates                  src/chemistry.f90:591-966
// with YSMP's sparse LU replaced by a fixed-pattern LU on the host-computed ordering.
//
// Data placement (per cell), level-parallel mode (integrate_kernel<EPT, true>; networks whose
// factor does not fit in 227 KB run the generic mode <EPT, false> with head rows in L2):
//   registers      Nordsieck array YH (6 columns), ACOR, EWT, RTOL/ATOL: element i of every
//                  vector lives in thread i mod 256 (all controller updates are elementwise)
//   shared memory  y / savf / solve vector / 1/pivots, the schedules' descriptors, and the
//                  whole LU factor V in storage order [head x head | U_B | L_C | dense tail];
//                  the U_B/L_C part is the scratch X between factorisations (fluxes, gather
//                  partials, inverses of the S diagonal blocks, tables of the staged solves)
//   L2-resident    per-CTA workspace: Jacobian in storage order, ELL copies of the two
//                  coupling blocks U_B / L_C for the solves, the cell's rate coefficients
// Factorisation: factor_glu (levels of independent pivots, gather schedules, rank-1 levels,
// register-tiled dense tail, explicit block inverses); solves: solve_glu (staged sweeps).
#include <cuda_runtime.h>
#include <cstdio>
#include <mutex>
#include "racg_dev.cuh"
#include "racg_rates.cuh"

namespace racg {

#ifndef RACG_NT
#define RACG_NT 256
#endif
constexpr int NT = RACG_NT;       // threads per CTA = per cell (256 or 512)
constexpr int NW = NT / 32;       // warps per CTA
constexpr int TJ = NT / 16;       // dense tail LU: 16 x TJ thread grid
constexpr int LPB = NT / 32;      // lanes per row in the 32-row block stages of the solves
constexpr int LGLPB = (LPB == 8) ? 3 : 4;
constexpr int MAXTL = 8;          // tail rows per thread: nt <= 16 * MAXTL = 128
static_assert(NT == 256 || NT == 512, "RACG_NT must be 256 or 512");

// per-phase cycle counters (racg_phase_cycles): compiled in only for the profiling build
// (make PROF=1 ../libracg_prof.so); the production build keeps only the total and the cell count
#ifndef RACG_PHASE_TIMERS
#define RACG_PHASE_TIMERS 0
#endif
#if RACG_PHASE_TIMERS
__device__ __forceinline__ long long pclock() { return clock64(); }
#else
__device__ __forceinline__ long long pclock() { return 0; }
#endif

enum Phase { PH_RATES = 0, PH_F, PH_JAC, PH_FACT_HEAD, PH_FACT_SCHUR, PH_FACT_TAIL, PH_SOLVE,
             PH_VEC, PH_G_LOOP, PH_TOTAL, PH_NCELL, PH_PBUILD, PH_TAILINV, PH_S_FWD, PH_S_TAIL, PH_S_BWD,
             PH_F_FLUX, PH_F_GATHER, PH_G_PIVMUL, PH_G_FLAT, PH_G_NARROW, PH_G_WIDE, PH_S_SPMV, PH_G_COPY,
             PH_T_L0, PH_T_LALL, PH_T_W0MID, PH_T_U0, PH_T_UTOP, PH_B_DIAG, PH_B_PANEL, PH_B_UPD, PH_COUNT };

struct Smem {
  double* y;      // [n]  argument of f / right-hand side and result of the linear solve
  double* savf;   // [n]  f(y)
  double* xb;     // solve vector in elimination order: aliases savf (dead while P x = b is solved)
  double* dinv;   // [n] 1 / pivot: head rows, then the dense tail
  double* par;    // [32]
  double* red;    // [2*NW]
  double* Dt;     // [ldt*nt] dense tail, column-major; diagonal 32-blocks hold their inverses
  double* hh;     // [n_hh] head x head block (shared memory, or workspace when it does not fit)
  double* X;      // scratch
  // read-only index tables staged once per CTA (they sit on dependent-load chains)
  const uint16_t* hhcol;   // [n_hh]
  const int4* fthin;       // forward thin-level rows {row, start, len, 0}
  const int4* bthin;       // backward thin-level rows
  const int* flptr;        // [nflev+1] forward level pointers
  const int* suptr;        // [nsu+1] backward level pointers
};

struct Ws {          // per-CTA global workspace (L2 resident)
  double* J;         // [nstore] Jacobian in storage order
  double* ubE;       // [ubE.nval] U_B values, ELL order
  double* lcE;       // [lcE.nval] L_C values, ELL order
  double* ksave;     // [R] rate coefficients of the cell
  double* hhG;       // [n_hh] fallback home of hh
  double* ubG;       // [n_ub] fallback home of the factorisation's U_B
  double* good;      // [n] last record whose T and X(H2) are not NaN (F15 harvest, src/disk.f90:1716-1733)
};

struct Layout { int hh_smem, ub_smem, nwr, glu; size_t xdoubles, total, voff; };

// the one dynamic shared-memory block; device functions address it through this symbol so
// that the compiler emits shared-space (LDS/STS) instead of generic accesses
extern __shared__ __align__(16) double smem_raw[];

// Network descriptor in constant memory: every device function reads it as a direct constant-
// bank operand (passing the ~1 KB struct by reference would spill it to local memory, and an
// indexed array of descriptors costs a separate LDC per use: +17 % run time was measured).
// There is ONE descriptor per device, so launch_integrate() serialises the integrator launches
// of all handles of a device through an event: a launch first waits for the previous
// integrator launch on that device (whatever handle or stream issued it), then uploads its own
// descriptor in stream order -- skipped when the device already holds it.  The persistent
// kernel fills every SM, so two integrator launches could not overlap anyway.
__constant__ DevNet c_net;

// Shared-memory plan.  Preferred: everything (head x head block, the factorisation's copy
// of U_B) on chip AND the total under 196 KB, so that the SM keeps >= 60 KB of L1 for the
// read-only index tables the sparse phases chase (they sit on dependent-load chains: at L2
// latency they dominated the first version of this kernel).  nwr = number of warps that
// own a dense work row during the row phases of the factorisation.
__host__ __device__ inline Layout make_layout(const DevNet& net) {
  Layout L;
  const size_t n = net.n, R = net.R;
  const size_t nthin_f = (size_t)(net.nh - net.flev_nfat_rows), nthin_b = (size_t)(net.nh - net.su_nfat_rows);
  const size_t tables = ((size_t)net.n_hh * 2 + (nthin_f + nthin_b) * 16 +
                         (size_t)(net.nflev + net.nsu + 2) * 4 + 64 + 7) / 8;
  const size_t fixed = 2 * n + n + 32 + 2 * NW + (size_t)net.ldt * net.nt + tables;   // y, savf, 1/pivots (head + tail)
  // scratch X: fluxes + gather partials | dflux + partials | SpMV partials | tolerance staging |
  // tail_lu publication buffers; during the factorisation: nwr work rows + U_B values + U_B columns
  size_t x_f = R + (size_t)net.rhs.npartial + 32;
  const size_t xj0 = R + (size_t)net.jac[0].npartial + 32, xj1 = R + (size_t)net.jac[1].npartial + 32;
  if (xj0 > x_f) x_f = xj0;
  if (xj1 > x_f) x_f = xj1;
  const size_t xs = (size_t)net.ubE.npartial + (size_t)net.lcE.npartial + 32;
  if (xs > x_f) x_f = xs;
  if (2 * n > x_f) x_f = 2 * n;
  if (600 > x_f) x_f = 600;
  const size_t ubx = (size_t)net.n_ub + ((size_t)net.n_ub * 2 + 7) / 8 + 2;   // values + u16 columns
  const size_t big = (227 * 1024 - 1024) / 8;
  L.hh_smem = 1; L.ub_smem = 1; L.nwr = NW; L.glu = 0; L.voff = 0;
  if (net.glu.on) {   // planned by integrate_smem_bytes()
    L.glu = 1; L.voff = (size_t)net.glu.voff; L.xdoubles = (size_t)(net.o_tl - net.o_ub);
    L.total = L.voff + net.nstore + 2;
    return L;
  }
  bool ok = false;
  for (int nwr = NW; nwr >= 4 && !ok; --nwr) {
    size_t x = (size_t)nwr * n + ubx;
    if (x < x_f) x = x_f;
    if (fixed + net.n_hh + x <= big) { L.nwr = nwr; ok = true; }
  }
  if (!ok) { L.nwr = NW > 8 ? 8 : NW; L.hh_smem = 0; L.ub_smem = 0; }   // large networks: L2 workspace
  size_t x = (size_t)L.nwr * n + (L.ub_smem ? ubx : 0);
  if (x < x_f) x = x_f;
  L.xdoubles = x;
  L.total = fixed + (L.hh_smem ? net.n_hh : 0) + x;
  return L;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deterministic block-wide sum; result returned to every thread
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < NW; ++k) s += red[k];
  return s;
}

__device__ __forceinline__ double block_max(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = red[0];
#pragma unroll
  for (int k = 1; k < NW; ++k) s = fmax(s, red[k]);
  return s;
}

// ---------------------------------------------------------------------------
// segmented-ELL gather: out[target] (+)= sum coef * src[idx]; all index loads of a
// block are issued before the first use (one L2 round trip per block)
template <bool GLOBAL_OUT>
__device__ __forceinline__ void run_gather(const GatherDev& g, const double* src, double* out,
                                           double* partial) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int b = w; b < g.nblk; b += NW) {
    const int width = __ldg(g.blk_width + b);
    const uint32_t* e = g.ent + __ldg(g.blk_off + b) + l;
    uint32_t ev[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) ev[j] = (j < width) ? __ldcg(e + j * 32) : (4u << 24);
    // ONE accumulator, terms in reaction order as the reference's loop adds them: forward and
    // reverse reactions sit next to each other in the network file and cancel term by term.
    // Interleaved partial sums lose that cancellation and the extra round-off in f makes the
    // corrector of very stiff cells fail (3x more steps were measured).
    double a = 0.0;
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += 8) {
      if (j0 < width) {
        double sv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sv[j] = src[ev[j0 + j] & 0xffffffu];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = (int)(ev[j0 + j] >> 24) - 4;
          a += (c != 0) ? (double)c * sv[j] : 0.0;
        }
      }
    }
    const int t = __ldg(g.sub_target + b * 32 + l);
    if (t >= 0) {
      if (GLOBAL_OUT && g.sub_add && g.sub_add[b * 32 + l]) out[t] += a; else out[t] = a;
    } else if (t <= -2) partial[-2 - t] = a;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < g.ncombine; q += NT) {
    double s = 0.0;
    const int p0 = __ldg(g.comb_ptr + q), p1 = __ldg(g.comb_ptr + q + 1);
    for (int p = p0; p < p1; ++p) s += partial[p];
    const int row = __ldg(g.comb_row + q);
    if (GLOBAL_OUT && g.comb_add && g.comb_add[q]) out[row] += s; else out[row] = s;
  }
  __syncthreads();
}

// flux of reaction r (branches of chem_ode_f, src/disk.f90:4583-4643)
__device__ __forceinline__ double flux_of(uint32_t w, double k, const double* y,
                                          double DS) {
  const int kind = (w >> 20) & 3;
  const double y1 = y[w & 1023];
  if (kind == FK_ONE) return k * y1;
  if (kind == FK_TWO) {
    const double y2 = y[(w >> 10) & 1023];
    double r = k * y1 * y2;
    if (y1 < 0.0 && y2 < 0.0) r = -r;
    return r;
  }
  if (kind == FK_SAT) {
    const double tmp1 = DS * c_net.sat_c[w >> 22];
    if (tmp1 <= 0.0) return k;
    const double tmp = y1 / tmp1;
    return (tmp <= 1e-4) ? k * tmp : k * (1.0 - exp(-tmp));
  }
  return 0.0;
}

// d flux / d y(r1) (which = 0) or d flux / d y(r2) (which = 1)
// (branches of chem_ode_jac, src/disk.f90:4765-4866)
__device__ __forceinline__ double dflux_of(uint32_t w, double k, const double* y,
                                           double DS, int which) {
  const int kind = (w >> 20) & 3;
  const int r1 = w & 1023, r2 = (w >> 10) & 1023;
  if (kind == FK_ONE) return which == 0 ? k : 0.0;
  if (kind == FK_TWO) {
    const double y1 = y[r1], y2 = y[r2];
    double d;
    if (r1 != r2) d = (which == 0) ? k * y2 : k * y1;
    else d = (which == 0) ? 2.0 * k * y2 : 0.0;
    if (y1 < 0.0 && y2 < 0.0) d = -d;
    return d;
  }
  if (kind == FK_SAT) {
    if (which != 0) return 0.0;
    const double tmp2 = DS * c_net.sat_c[w >> 22];
    if (tmp2 <= 0.0) return 0.0;
    const double tmp1 = 1.0 / tmp2;
    const double tmp = y[r1] * tmp1;
    return (tmp <= 1e-4) ? k * tmp1 : k * tmp1 * exp(-tmp);
  }
  return 0.0;
}

// chem_ode_f: savf = S * flux(k, y).  k streamed from the workspace (L2), fluxes in the
// scratch region X (x_off = its offset in the shared-memory block); y = smem_raw[0..n),
// savf = smem_raw[n..2n)
__device__ __forceinline__ void eval_f(const double* __restrict__ ks, int x_off, double DS, unsigned long long* ph) {
  const DevNet& net = c_net;
  const int R = net.R;
  double* const fx = smem_raw + x_off;
  double* const px = fx + R;
  const double* const yv = smem_raw;
  double* const out = smem_raw + net.n;
  __syncthreads();   // y was just written by its owner threads
  const long long tf0 = pclock();
  for (int r0 = threadIdx.x; r0 < R; r0 += 4 * NT) {
    double kk[4]; uint32_t ww[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * NT;
      kk[u] = (r < R) ? __ldcg(ks + r) : 0.0;
      ww[u] = (r < R) ? __ldg(net.fw + r) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * NT;
      if (r < R) fx[r] = flux_of(ww[u], kk[u], yv, DS);
    }
  }
  for (int i = threadIdx.x; i < net.n; i += NT) out[i] = 0.0;
  __syncthreads();
  const long long tg = pclock();
  run_gather<false>(net.rhs, fx, out, px);
  if (threadIdx.x == 0) { ph[PH_F_FLUX] += tg - tf0; ph[PH_F_GATHER] += pclock() - tg; }
}

// chem_ode_jac for all columns at once -> ws.J (two passes over the reactions)
__device__ __forceinline__ void eval_jac(const double* __restrict__ ks, int x_off, double* J, double DS) {
  const DevNet& net = c_net;
  double* const dfx = smem_raw + x_off;
  double* const px = dfx + net.R;
  const double* const yv = smem_raw;
  __syncthreads();
  for (int pass = 0; pass < 2; ++pass) {
    for (int r = threadIdx.x; r < net.R; r += NT) dfx[r] = dflux_of(__ldg(net.fw + r), __ldcg(ks + r), yv, DS, pass);
    __syncthreads();
    run_gather<true>(net.jac[pass], dfx, J, px);
  }
}

// ---------------------------------------------------------------------------
// dense no-pivot LU of the nt x nt tail block held in sm.Dt (column-major, ld = ldt),
// register tiled: thread (ti,tj) of a 16 x TJ grid owns rows ti+16a (a < TL) and columns
// tj+TJ*b (b < TC = ceil(16 TL / TJ); columns >= nt are padding and never touch memory).
// One barrier per elimination step: the owners of column k / row k publish the unscaled
// column, the row and the pivot; everybody scales and updates its tile.
// helpers with a compile-time tile index so that register arrays stay in registers
template <int TL, int B> struct PubCol { template <class T> static __device__ __forceinline__ void go(const T& t, double* pc, int ti) {
#pragma unroll
  for (int a = 0; a < TL; ++a) pc[ti + 16 * a] = t[a][B]; } };
template <int TC, int A> struct PubRow { template <class T> static __device__ __forceinline__ void go(const T& t, double* pr, int tj) {
#pragma unroll
  for (int b = 0; b < TC; ++b) pr[tj + TJ * b] = t[A][b]; } };

// one elimination step k = 16*KA + kr (KA uniform across the CTA): tile row index KA, tile
// column index KC = k / TJ
template <int TL, int TC, int KA>
__device__ __forceinline__ void tail_step(double (&t)[TL][TC], int k, int ti, int tj, double* pc, double* pr, int* flag, double* rdt) {
  constexpr int KC = (16 * KA) / TJ;
  const int kr = k & 15, kc = k & (TJ - 1);
  if (tj == kc) PubCol<TL, KC>::go(t, pc, ti);
  if (ti == kr) PubRow<TC, KA>::go(t, pr, tj);
  __syncthreads();
  const double piv = pc[k];
  if (piv == 0.0 || isnan(piv)) { if (threadIdx.x == 0) *flag = 1; }
  const double inv = 1.0 / piv;
  if (threadIdx.x == 0) rdt[k] = inv;     // 1 / U(k,k) for the back substitution
  // rows a > KA and columns b > KC are fully active, a == KA / b == KC are active for ti > kr / tj > kc
  double lr[TL], uc[TC];
#pragma unroll
  for (int a = KA; a < TL; ++a) lr[a] = pc[ti + 16 * a] * inv;
#pragma unroll
  for (int b = KC; b < TC; ++b) uc[b] = pr[tj + TJ * b];
  const bool rowKA = ti > kr, colKC = tj > kc;
  if (!rowKA) lr[KA] = 0.0;
  if (!colKC) uc[KC] = 0.0;
#pragma unroll
  for (int a = KA; a < TL; ++a)
#pragma unroll
    for (int b = KC; b < TC; ++b) t[a][b] -= lr[a] * uc[b];
  // column k keeps the multipliers (its uc was zero, so t was untouched above)
  if (tj == kc) {
#pragma unroll
    for (int a = KA + 1; a < TL; ++a) t[a][KC] = lr[a];
    if (rowKA) t[KA][KC] = lr[KA];
  }
}

// ---------------------------------------------------------------------------
// Blocked variant of the dense tail LU (256 threads, 16 x 16 grid): right-looking with 16-column
// panels.  The one-pivot-per-barrier scheme below spends ~950 cycles per pivot on two barriers,
// two shared-memory round trips and a division; here a pivot costs one shuffle round inside a
// single warp, and the other 112 x 112 x 16 multiply-adds of a panel run without barriers:
//   1. diag16_lu     warp 0 factors the 16 x 16 diagonal block, one row per lane, pivot row by shuffles
//   2. panel_solve   threads 0.. solve the rows of L21 = A21 U11^-1, threads 128.. the columns of
//                    U12 = L11^-1 A12 (16 unknowns each, operands broadcast from shared memory)
//   3. trailing update A22 -= L21 U12 on the register tiles (16 k-steps, no barrier), then only the
//      next panel's column and row slices go back to shared memory
// Every element receives the same multiply-adds in the same order (k ascending, l = a * (1/pivot))
// as in the unblocked scheme: the factor is bit-identical.
// (LD = leading dimension of the tail block = nt + 1, a compile-time constant of the TL variant:
// every shared-memory access below is base register + immediate)
template <int LD>
__device__ __noinline__ void diag16_lu(int blk_off, int rd_off, int row_off, int* flag) {
  constexpr int ld = LD;
  double* const B = smem_raw + blk_off;
  double* const rdt = smem_raw + rd_off;
  double* const rowbuf = smem_raw + row_off;      // 2 x 16 doubles, 16-byte aligned
  const int l = threadIdx.x & 31, i = l & 15;
  double r[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) r[j] = B[j * ld + i];
  // The pivot row goes through shared memory (one store by its owner, broadcast loads by all):
  // a shuffle costs ~8 issue cycles per 32-bit half on this part, 30 of them per pivot were the
  // whole cost of this routine.
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    double2* const rb = (double2*)(rowbuf + 16 * (k & 1));
    if (l == k) {
#pragma unroll
      for (int j = (k & ~1); j < 16; j += 2) rb[j >> 1] = make_double2(r[j], r[j + 1]);
    }
    __syncwarp();
    double u[16];
#pragma unroll
    for (int j = (k & ~1); j < 16; j += 2) { const double2 v = rb[j >> 1]; u[j] = v.x; u[j + 1] = v.y; }
    const double piv = u[k];
    if (piv == 0.0 || isnan(piv)) { if (l == 0) *flag = 1; }
    const double inv = 1.0 / piv;
    if (l == 0) rdt[k] = inv;     // 1 / U(k,k) for the back substitution
    const double m = r[k] * inv;
    if (i > k) {
      r[k] = m;
#pragma unroll
      for (int j = k + 1; j < 16; ++j) r[j] -= m * u[j];
    }
  }
  if (l < 16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) B[j * ld + i] = r[j];
  }
}
template <int LD>
__device__ __noinline__ void panel_solve(int dt_off, int c0, int rd_off) {
  constexpr int ld = LD, nt = LD - 1;
  double* const Dt = smem_raw + dt_off;
  const double* const rd = smem_raw + rd_off;     // rd[j] = 1 / U(c0+j, c0+j)
  const int tid = threadIdx.x, nrem = nt - c0 - 16;
  const double* const D = Dt + c0 * ld + c0;      // the factored diagonal block
  if (tid < nrem) {                               // one row of L21
    double* const row = Dt + c0 * ld + c0 + 16 + tid;
    double x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = row[j * ld];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      double s = x[j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= x[k] * D[j * ld + k];
      x[j] = s * rd[j];
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) row[j * ld] = x[j];
  } else if (tid >= 128 && tid < 128 + nrem) {    // one column of U12
    double* const col = Dt + (c0 + 16 + tid - 128) * ld + c0;
    double x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = col[k];
#pragma unroll
    for (int k = 1; k < 16; ++k) {
      double s = x[k];
#pragma unroll
      for (int m = 0; m < k; ++m) s -= D[m * ld + k] * x[m];
      x[k] = s;
    }
#pragma unroll
    for (int k = 1; k < 16; ++k) col[k] = x[k];
  }
}
template <int TL>
__device__ __forceinline__ void tail_lu_blocked(int dt_off, int pub_off, int* flag, unsigned long long* ph) {
  const DevNet& net = c_net;
  constexpr int ld = 16 * TL + 1;      // = net.ldt (checked at launch)
  const int rd_off = 2 * net.n + net.nh;
  double* const Dt = smem_raw + dt_off;
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  double t[TL][TL];      // rows ti + 16 a, columns tj + 16 b; the first block row / column never enters registers
#pragma unroll
  for (int a = 1; a < TL; ++a)
#pragma unroll
    for (int b = 1; b < TL; ++b) t[a][b] = Dt[(tj + 16 * b) * ld + ti + 16 * a];
#pragma unroll
  for (int p = 0; p < TL; ++p) {
    const int c0 = 16 * p;
    const long long tb0 = pclock();
    if (threadIdx.x < 32) diag16_lu<ld>(dt_off + c0 * ld + c0, rd_off + c0, (pub_off + 1) & ~1, flag);
    if (RACG_PHASE_TIMERS && threadIdx.x == 0) ph[PH_B_DIAG] += pclock() - tb0;
    if (p < TL - 1) {
      __syncthreads();
      const long long tb1 = pclock();
      panel_solve<ld>(dt_off, c0, rd_off + c0);
      __syncthreads();
      const long long tb2 = pclock();
      if (RACG_PHASE_TIMERS && threadIdx.x == 0) ph[PH_B_PANEL] += tb2 - tb1;
      const double* const Lb = Dt + c0 * ld + ti;      // L21(ti + 16 a, k) = Lb[k * ld + 16 a]
      const double* const Ub = Dt + tj * ld + c0;      // U12(k, tj + 16 b) = Ub[16 b * ld + k]
#pragma unroll 4
      for (int k = 0; k < 16; ++k) {
        double lr[TL], uc[TL];
#pragma unroll
        for (int a = p + 1; a < TL; ++a) lr[a] = Lb[k * ld + 16 * a];
#pragma unroll
        for (int b = p + 1; b < TL; ++b) uc[b] = Ub[16 * b * ld + k];
#pragma unroll
        for (int a = p + 1; a < TL; ++a)
#pragma unroll
          for (int b = p + 1; b < TL; ++b) t[a][b] -= lr[a] * uc[b];
      }
      // the next panel's column slice and row slice go back to shared memory; the rest stays in registers
#pragma unroll
      for (int a = p + 1; a < TL; ++a) Dt[(tj + 16 * (p + 1)) * ld + ti + 16 * a] = t[a][p + 1];
#pragma unroll
      for (int b = p + 2; b < TL; ++b) Dt[(tj + 16 * b) * ld + ti + 16 * (p + 1)] = t[p + 1][b];
      __syncthreads();
      if (RACG_PHASE_TIMERS && threadIdx.x == 0) ph[PH_B_UPD] += pclock() - tb2;
    }
  }
  __syncthreads();
}

template <int TL>
__device__ __noinline__ void tail_lu(int dt_off, int pub_off, int* flag, unsigned long long* ph) {   // reciprocal pivots -> dinv[nh + k]
  if constexpr (TJ == 16) { tail_lu_blocked<TL>(dt_off, pub_off, flag, ph); return; }
  const DevNet& net = c_net;
  constexpr int TC = (16 * TL + TJ - 1) / TJ;
  const int nt = net.nt, ldt = net.ldt;
  double* const Dt = smem_raw + dt_off;     // offsets into the shared-memory block: keeps
  double* const pub = smem_raw + pub_off;   // the accesses in the shared address space
  double* const rdt = smem_raw + 2 * net.n + net.nh;
  const int ti = threadIdx.x / TJ, tj = threadIdx.x & (TJ - 1);
  double t[TL][TC];
#pragma unroll
  for (int a = 0; a < TL; ++a)
#pragma unroll
    for (int b = 0; b < TC; ++b) t[a][b] = (tj + TJ * b < nt) ? Dt[(tj + TJ * b) * ldt + ti + 16 * a] : 0.0;
  for (int k = 0; k < nt; ++k) {
    double* pc = pub + (k & 1) * (2 * 128 + 8);
    double* pr = pc + 128;
    switch (k >> 4) {   // uniform across the CTA
      case 0: tail_step<TL, TC, 0>(t, k, ti, tj, pc, pr, flag, rdt); break;
      case 1: if (TL > 1) tail_step<TL, TC, (TL > 1 ? 1 : 0)>(t, k, ti, tj, pc, pr, flag, rdt); break;
      case 2: if (TL > 2) tail_step<TL, TC, (TL > 2 ? 2 : 0)>(t, k, ti, tj, pc, pr, flag, rdt); break;
      case 3: if (TL > 3) tail_step<TL, TC, (TL > 3 ? 3 : 0)>(t, k, ti, tj, pc, pr, flag, rdt); break;
      case 4: if (TL > 4) tail_step<TL, TC, (TL > 4 ? 4 : 0)>(t, k, ti, tj, pc, pr, flag, rdt); break;
      case 5: if (TL > 5) tail_step<TL, TC, (TL > 5 ? 5 : 0)>(t, k, ti, tj, pc, pr, flag, rdt); break;
      case 6: if (TL > 6) tail_step<TL, TC, (TL > 6 ? 6 : 0)>(t, k, ti, tj, pc, pr, flag, rdt); break;
      default: if (TL > 7) tail_step<TL, TC, (TL > 7 ? 7 : 0)>(t, k, ti, tj, pc, pr, flag, rdt); break;
    }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < TL; ++a)
#pragma unroll
    for (int b = 0; b < TC; ++b) if (tj + TJ * b < nt) Dt[(tj + TJ * b) * ldt + ti + 16 * a] = t[a][b];
  __syncthreads();
}

// ---------------------------------------------------------------------------
// Shared-memory view of the level-parallel mode, derived from the smem_raw symbol so that
// every access below is a shared-space access.
struct GSm { double *y, *xb, *dinv, *V, *X, *Dt; const int4 *lvl, *grp, *st, *r1; };
__device__ __forceinline__ GSm glu_smem() {
  const DevNet& net = c_net;
  GSm g;
  g.y = smem_raw; g.xb = smem_raw + net.n; g.dinv = smem_raw + 2 * net.n;
  g.V = smem_raw + net.glu.voff; g.X = g.V + net.o_ub; g.Dt = g.V + net.o_tl;
  g.lvl = (const int4*)(smem_raw + net.glu.doff); g.grp = g.lvl + net.glu.nlev + 1; g.st = g.grp + net.glu.ngrp; g.r1 = g.st + net.glu.nst;
  return g;
}

// Triangular solves with a dense factored block (<= 128 rows) by ONE warp: plain substitution, as
// the reference's nntc does (src/opkda1.f:3750-3801), organised as a column sweep with the
// right-hand side in registers -- lane l owns row l of the current 32-row block (x0) and of the
// three blocks that follow it in sweep direction (x1..x3) -- so that the dependent chain is one
// shuffle and one FMA per column (forward) / one multiply by the stored reciprocal pivot more
// (backward).  Operands are offsets into the shared-memory block: T column-major with leading
// dimension ld, unit-lower L below and U on/above the diagonal; rd[i] = 1 / U(i,i); xs = the
// vector, solved in place.  One compact loop body shared by every caller (not inlined): a single
// warp runs this while the others wait, so its instruction fetches are fully exposed.
// (Explicit inverses of the diagonal blocks were equally fast to apply but cost a factorisation
// pass and lost the corrector on hot, very stiff cells: pivots spanning many decades.)
// WIN = number of 32-row blocks in the register window (1: a single 32-row block, 4: nt <= 128)
template <int WIN>
__device__ __noinline__ void warp_trisolve_lower(int t_off, int ld, int nt, int xs_off) {
  const double* const T = smem_raw + t_off;
  double* const xs = smem_raw + xs_off;
  const int l = threadIdx.x & 31, nb = (nt + 31) >> 5;
  double x[WIN];
#pragma unroll
  for (int m = 0; m < WIN; ++m) x[m] = (32 * m + l < nt) ? xs[32 * m + l] : 0.0;
#pragma unroll 1
  for (int mb = 0; mb < nb; ++mb) {
    const int base = 32 * mb, cmax = (nt - base) < 32 ? (nt - base) : 32;
    bool on[WIN];
#pragma unroll
    for (int m = 0; m < WIN; ++m) on[m] = base + 32 * m + l < nt;
    const double* col = T + base * ld + base + l;
#pragma unroll 1
    for (int c0 = 0; c0 < cmax; c0 += 8) {
      double a[WIN][8], l10[4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[0][j] = (on[0] && l > c0 + j) ? col[j * ld] : 0.0;
#pragma unroll
        for (int m = 1; m < WIN; ++m) a[m][j] = on[m] ? col[j * ld + 32 * m] : 0.0;
      }
      // columns go in pairs: both right-hand-side values are shuffled at once and every lane forms
      // the second unknown itself, x_{c+1} = b_{c+1} - L(c+1,c) x_c (bit-identical to what its owner
      // computes), so the dependent chain per PAIR is one shuffle and two FMAs
#pragma unroll
      for (int q = 0; q < 4; ++q) l10[q] = T[(base + c0 + 2 * q) * ld + base + c0 + 2 * q + 1];
      col += 8 * ld;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double xc = __shfl_sync(0xffffffffu, x[0], c0 + 2 * q);
        const double x1p = __shfl_sync(0xffffffffu, x[0], c0 + 2 * q + 1);
        const double x1f = x1p - l10[q] * xc;
#pragma unroll
        for (int m = 0; m < WIN; ++m) { x[m] -= a[m][2 * q] * xc; x[m] -= a[m][2 * q + 1] * x1f; }
      }
    }
    if (on[0]) xs[base + l] = x[0];
#pragma unroll
    for (int m = 0; m + 1 < WIN; ++m) x[m] = x[m + 1];
    x[WIN - 1] = 0.0;
  }
}
template <int WIN>
__device__ __noinline__ void warp_trisolve_upper(int t_off, int ld, int nt, int rd_off, int xs_off) {
  const double* const T = smem_raw + t_off;
  const double* const rd = smem_raw + rd_off;
  double* const xs = smem_raw + xs_off;
  const int l = threadIdx.x & 31, nb = (nt + 31) >> 5;
  const int top = 32 * (nb - 1);
  double x[WIN];
#pragma unroll
  for (int m = 0; m < WIN; ++m) x[m] = (m < nb && top - 32 * m + l < nt) ? xs[top - 32 * m + l] : 0.0;
#pragma unroll 1
  for (int mb = nb - 1; mb >= 0; --mb) {
    const int base = 32 * mb, cmax = (nt - base) < 32 ? (nt - base) : 32;
    // the current block is carried scaled by its reciprocal pivots, s_r = x_r / U(r,r), and so are
    // its U entries: lane c then holds the finished x_c when column c is reached, and the chain is
    // shuffle + FMA as in the forward sweep
    const double rdl = (base + l < nt) ? rd[base + l] : 1.0;
    x[0] *= rdl;
    const double* col = T + (base + cmax - 8) * ld + base + l;
#pragma unroll 1
    for (int c0 = cmax - 8; c0 >= 0; c0 -= 8) {
      double a[WIN][8], u10[4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[0][j] = (base + l < nt) ? col[j * ld] : 0.0;
#pragma unroll
        for (int m = 1; m < WIN; ++m) a[m][j] = (mb >= m) ? col[j * ld - 32 * m] : 0.0;
      }
      // pairs of columns (c, c-1) as in the forward sweep: x_{c-1} = s_{c-1} - (U(c-1,c) rd_{c-1}) x_c
      double r10[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        u10[q] = T[(base + c0 + 2 * q + 1) * ld + base + c0 + 2 * q];
        r10[q] = rd[base + c0 + 2 * q];
      }
      col -= 8 * ld;
      // all loads first, then the scaling: a multiply issued right behind its load would stall
      // the only running warp for the full shared-memory latency, eight times per group
#pragma unroll
      for (int j = 0; j < 8; ++j) a[0][j] = (l < c0 + j) ? a[0][j] * rdl : 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) u10[q] *= r10[q];
#pragma unroll
      for (int q = 3; q >= 0; --q) {
        const double xc = __shfl_sync(0xffffffffu, x[0], c0 + 2 * q + 1);
        const double x1p = __shfl_sync(0xffffffffu, x[0], c0 + 2 * q);
        const double x1f = x1p - u10[q] * xc;
#pragma unroll
        for (int m = 0; m < WIN; ++m) { x[m] -= a[m][2 * q + 1] * xc; x[m] -= a[m][2 * q] * x1f; }
      }
    }
    if (base + l < nt) xs[base + l] = x[0];
#pragma unroll
    for (int m = 0; m + 1 < WIN; ++m) x[m] = x[m + 1];
    x[WIN - 1] = 0.0;
  }
}

// Dense-tail substitution (L x = b, then U x = y; nt <= 128) by up to four warps, tid < 128.
// Warp w owns the 32-row block w of the right-hand side, one row per lane, in a register.  A
// sweep visits the diagonal blocks in order; the owner of the current one runs the column-pair
// chain of warp_trisolve_* on it and publishes every finished group of 8 unknowns to xs[]
// followed by a progress counter; the owners of the blocks still to come wait on that counter
// and subtract the group's columns from their rows right away, operands fetched before the wait.
// When a chain ends the next owner is a few FMAs behind, so the critical path is the diagonal
// chains alone -- the off-diagonal updates, which used to share the chain warp's issue slots,
// run beside it.  Every row sees the same operations in the same order as in the single-warp
// sweep: results are bit-identical.  ctr: two ints, zeroed by the caller before a barrier
// (forward progress = columns done from the left, backward = from the right).
// Ordering of the hand-off: the unknowns and the counter are stored by the same warp to shared
// memory in program order with a __syncwarp() between (lanes c0..c0+7, then lane 0), and read by
// the consumer counter first, data (volatile) after the branch that depends on it;
// RACG_TAIL_FENCE=1 makes both sides explicit release / acquire operations (measured: no
// difference in time, identical results).
// (A variant that broadcast four unknowns at a time through shared memory instead of shuffling
// pairs was bit-identical and not faster: the chains run ~4 groups per call and their time is
// dominated by the cold start of each code path, see DESIGN.md.)
#ifndef RACG_TAIL_FENCE
#define RACG_TAIL_FENCE 0
#endif
__device__ __forceinline__ void st_release_smem(int* p, int v) {
#if RACG_TAIL_FENCE
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
#else
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ int ld_acquire_smem(const int* p) {
  int v;
#if RACG_TAIL_FENCE
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
#else
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
#endif
  return v;
}
__device__ __noinline__ void tail_trisolve_mw(int t_off, int ld, int nt, int rd_off, int xs_off, int ctr_off, unsigned long long* ph) {
  const double* const T = smem_raw + t_off;
  const double* const rd = smem_raw + rd_off;
  double* const xs = smem_raw + xs_off;
  int* const ctrL = (int*)(smem_raw + ctr_off);
  int* const ctrU = ctrL + 1;
  const int l = threadIdx.x & 31, w = threadIdx.x >> 5, nb = (nt + 31) >> 5;
  if (w >= nb) return;
  const int base = 32 * w, cmax = (nt - base) < 32 ? (nt - base) : 32;
  const bool on = l < cmax;
  double x = on ? xs[base + l] : 0.0;
  const long long tp0 = pclock();
  // ---------------- forward
#pragma unroll 1
  for (int cb = 0; cb < w; ++cb) {
    const double* col = T + (32 * cb) * ld + base + l;
    const volatile double* xv = xs + 32 * cb;
#pragma unroll 1
    for (int c0 = 0; c0 < 32; c0 += 8) {
      double a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = on ? col[j * ld] : 0.0;
      col += 8 * ld;
      const int need = 32 * cb + c0 + 8;
      while (ld_acquire_smem(ctrL) < need) {}
      double v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = xv[c0 + j];
#pragma unroll
      for (int j = 0; j < 8; ++j) x -= a[j] * v[j];
    }
  }
  {
    const double* col = T + base * ld + base + l;
#pragma unroll 1
    for (int c0 = 0; c0 < cmax; c0 += 8) {
      double a[8], l10[4];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = (on && l > c0 + j) ? col[j * ld] : 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) l10[q] = T[(base + c0 + 2 * q) * ld + base + c0 + 2 * q + 1];
      col += 8 * ld;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double xc = __shfl_sync(0xffffffffu, x, c0 + 2 * q);
        const double x1p = __shfl_sync(0xffffffffu, x, c0 + 2 * q + 1);
        const double x1f = x1p - l10[q] * xc;
        x -= a[2 * q] * xc;
        x -= a[2 * q + 1] * x1f;
      }
      if (l >= c0 && l < c0 + 8) xs[base + l] = x;
      __syncwarp();
      if (l == 0) st_release_smem(ctrL, base + c0 + 8);
    }
  }
  const long long tp1 = pclock();
  // ---------------- backward (x = this block's forward result)
#pragma unroll 1
  for (int cb = nb - 1; cb > w; --cb) {
    const int cbm = (nt - 32 * cb) < 32 ? (nt - 32 * cb) : 32;
    const double* col = T + (32 * cb + cbm - 8) * ld + base + l;
    const volatile double* xv = xs + 32 * cb;
#pragma unroll 1
    for (int c0 = cbm - 8; c0 >= 0; c0 -= 8) {
      double a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = col[j * ld];
      col -= 8 * ld;
      const int need = nt - (32 * cb + c0);
      while (ld_acquire_smem(ctrU) < need) {}
      double v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = xv[c0 + j];
#pragma unroll
      for (int j = 7; j >= 0; --j) x -= a[j] * v[j];
    }
  }
  const long long tp2 = pclock();
  {
    // the block is carried scaled by its reciprocal pivots (see warp_trisolve_upper)
    const double rdl = on ? rd[base + l] : 1.0;
    x *= rdl;
    const double* col = T + (base + cmax - 8) * ld + base + l;
#pragma unroll 1
    for (int c0 = cmax - 8; c0 >= 0; c0 -= 8) {
      double a[8], u10[4], r10[4];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = on ? col[j * ld] : 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        u10[q] = T[(base + c0 + 2 * q + 1) * ld + base + c0 + 2 * q];
        r10[q] = rd[base + c0 + 2 * q];
      }
      col -= 8 * ld;
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = (l < c0 + j) ? a[j] * rdl : 0.0;   // loads first, scaling after
#pragma unroll
      for (int q = 0; q < 4; ++q) u10[q] *= r10[q];
#pragma unroll
      for (int q = 3; q >= 0; --q) {
        const double xc = __shfl_sync(0xffffffffu, x, c0 + 2 * q + 1);
        const double x1p = __shfl_sync(0xffffffffu, x, c0 + 2 * q);
        const double x1f = x1p - u10[q] * xc;
        x -= a[2 * q + 1] * xc;
        x -= a[2 * q] * x1f;
      }
      if (l >= c0 && l < c0 + 8) xs[base + l] = x;
      __syncwarp();
      if (l == 0) st_release_smem(ctrU, nt - (base + c0));
    }
  }
  if (RACG_PHASE_TIMERS && l == 0 && ph != nullptr) {
    const long long tp3 = pclock();
    if (w == 0) { ph[PH_T_L0] += tp1 - tp0; ph[PH_T_W0MID] += tp2 - tp1; ph[PH_T_U0] += tp3 - tp2; }
    if (w == nb - 1 && w > 0) { ph[PH_T_LALL] += tp1 - tp0; ph[PH_T_UTOP] += tp3 - tp2; }
  }
}

// One ELL group of the level-parallel factorisation with a compile-time width W (entries per
// target): NB targets per lane and step, all index loads first, then all operand reads, then
// the arithmetic.  ep/tp point at this lane's first entry / target of the warp's block range.
template <int W, int NB>
__device__ __forceinline__ void glu_group(const uint32_t* __restrict__ ep, const uint16_t* __restrict__ tp,
                                          int nb, double* V, int zp, uint32_t ZZ) {
  for (int b0 = 0; b0 < nb; b0 += NB) {
    uint32_t e[NB][W], t[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const bool on = b0 + u < nb;
      t[u] = on ? (uint32_t)__ldcg(tp + (b0 + u) * 32) : (uint32_t)(zp + 1);
#pragma unroll
      for (int j = 0; j < W; ++j) e[u][j] = on ? __ldcg(ep + ((b0 + u) * W + j) * 32) : ZZ;
    }
    double lv[NB][W], uv[NB][W], tv[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      if (t[u] == 0xFFFFu) t[u] = (uint32_t)(zp + 1);     // idle lane: the trash slot
      tv[u] = V[t[u]];
#pragma unroll
      for (int j = 0; j < W; ++j) { lv[u][j] = V[e[u][j] & 0xffffu]; uv[u][j] = V[e[u][j] >> 16]; }
    }
#pragma unroll
    for (int j = 0; j < W; ++j)
#pragma unroll
      for (int u = 0; u < NB; ++u) tv[u] -= lv[u][j] * uv[u][j];
#pragma unroll
    for (int u = 0; u < NB; ++u) V[t[u]] = tv[u];
  }
}

// P = I - hl0*J and its LU, level-parallel ("gather") formulation: see HostNet::LevelLU.
// The whole factor V (storage order) is in shared memory.  Returns 0 ok / 1 zero pivot.
__device__ __noinline__ int factor_glu(Ws ws, double con, int* flag, unsigned long long* ph) {
  const GSm sm = glu_smem();
  const DevNet& net = c_net;
  const GluDev& g = net.glu;
  const int nh = net.nh, nt = net.nt, ldt = net.ldt;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, tid = threadIdx.x;
  double* V = sm.V;
  if (tid == 0) *flag = 0;
  const long long t0 = pclock();
  {
    const double2* J2 = (const double2*)ws.J;
    double2* V2 = (double2*)V;
    const int n2 = net.nstore >> 1;
#pragma unroll 8
    for (int q = tid; q < n2; q += NT) { double2 v = __ldcs(J2 + q); v.x *= con; v.y *= con; V2[q] = v; }
    if (tid == 0) {
      if (net.nstore & 1) V[net.nstore - 1] = ws.J[net.nstore - 1] * con;
      V[g.zpos] = 0.0;
    }
  }
  __syncthreads();
  for (int i = tid; i < nh; i += NT) V[__ldg(net.pivmeta + i).x - 1] += 1.0;
  for (int a = tid; a < nt; a += NT) sm.Dt[a * ldt + a] += 1.0;
  __syncthreads();
  const long long t0b = pclock();
  const uint32_t ZZ = (uint32_t)g.zpos | ((uint32_t)g.zpos << 16);
  const int zp = g.zpos;
  // index entries of the next level's pivot/multiplier phase are fetched one level ahead
  uint32_t pe, m0, m1;
  {
    const int4 L0 = sm.lvl[0], L1 = sm.lvl[1];
    pe = (L0.x + tid < L1.x) ? __ldg(g.piv + L0.x + tid) : ZZ;
    m0 = (L0.y + tid < L1.y) ? __ldg(g.mul + L0.y + tid) : ZZ;
    m1 = (L0.y + tid + NT < L1.y) ? __ldg(g.mul + L0.y + tid + NT) : ZZ;
  }
  for (int lev = 0; lev < g.nlev; ++lev) {
    const long long tlev = pclock();
    const int4 L0 = sm.lvl[lev], L1 = sm.lvl[lev + 1];
    // rank-1 level: the first target chunk pair and the first multiplier position of this warp
    // are requested now and land during the multiplier phase
    int4 A = make_int4(0, 0, 0, 0), B = make_int4(0, 0, 0, 0);
    const uint2* tg = nullptr;
    uint2 ca = make_uint2(0u, 0u), cb = ca;
    uint32_t lpc = ZZ;
    if (L0.w >= 0) {
      A = sm.r1[2 * L0.w]; B = sm.r1[2 * L0.w + 1];   // {ua0, nua4, ub0, nub4}, {tgt_off, nr, nj4, 0}
      tg = (const uint2*)g.r1tgt + B.x + l;
      if (w < B.z && B.y > 0) {
        ca = __ldcg(tg + w * 32);
        cb = (w + NW < B.z) ? __ldcg(tg + (w + NW) * 32) : ca;
        lpc = (l < B.y) ? __ldg(g.mul + L0.y + l) : ZZ;
      }
    }
    // multipliers l(i,k) = a(i,k) / a(k,k); the pivots' reciprocals are kept for the solves
    if (L0.x + tid < L1.x) {
      const double d = V[pe & 0xffffu];
      if (d == 0.0 || isnan(d)) *flag = 1;
      sm.dinv[pe >> 16] = 1.0 / d;
    }
    {
      const double a0 = V[m0 & 0xffffu], d0 = V[m0 >> 16], a1 = V[m1 & 0xffffu], d1 = V[m1 >> 16];
      if (L0.y + tid < L1.y) V[m0 & 0xffffu] = a0 / d0;
      if (L0.y + tid + NT < L1.y) V[m1 & 0xffffu] = a1 / d1;
    }
    for (int q = L0.x + NT + tid; q < L1.x; q += NT) {
      const uint32_t e = __ldg(g.piv + q);
      const double d = V[e & 0xffffu];
      if (d == 0.0 || isnan(d)) *flag = 1;
      sm.dinv[e >> 16] = 1.0 / d;
    }
    for (int q = L0.y + 2 * NT + tid; q < L1.y; q += NT) {
      const uint32_t e = __ldg(g.mul + q);
      V[e & 0xffffu] = V[e & 0xffffu] / V[e >> 16];
    }
    __syncthreads();
    long long tq = pclock();
    if (tid == 0) ph[PH_G_PIVMUL] += tq - tlev;
    if (lev + 1 < g.nlev) {
      const int4 L2 = sm.lvl[lev + 2];
      pe = (L1.x + tid < L2.x) ? __ldg(g.piv + L1.x + tid) : ZZ;
      m0 = (L1.y + tid < L2.y) ? __ldg(g.mul + L1.y + tid) : ZZ;
      m1 = (L1.y + tid + NT < L2.y) ? __ldg(g.mul + L1.y + tid + NT) : ZZ;
    }
    // updates a(i,j) -= sum_k l(i,k) u(k,j)
    if (L0.w >= 0) {
      // single pivot k: outer product rows(k) x cols(k).  Lane = row of a 32-row chunk (its
      // multiplier in a register), warp = every NW-th chunk of 4 columns, two chunks per step;
      // the target positions of the next step are in flight while this one is applied
      const int nr = B.y, nj4 = B.z;
      int rc = 0, j4 = w;
      bool have = (w < nj4) && nr > 0;
      double lv = 0.0;
      bool fresh = true;
      while (have) {
        if (fresh) { lv = (rc * 32 + l < nr) ? V[lpc & 0xffffu] : 0.0; fresh = false; }
        int nrc = rc, nj = j4 + 2 * NW;
        if (nj >= nj4) { nj = w; ++nrc; }
        const bool hnext = nrc * 32 < nr;
        uint2 na = ca, nb2 = cb;
        uint32_t lpn = lpc;
        if (hnext) {
          const uint2* tp = tg + (size_t)nrc * nj4 * 32;
          na = __ldcg(tp + nj * 32);
          nb2 = (nj + NW < nj4) ? __ldcg(tp + (nj + NW) * 32) : na;
          if (nrc != rc) lpn = (nrc * 32 + l < nr) ? __ldg(g.mul + L0.y + nrc * 32 + l) : ZZ;
        }
        const int j4b = j4 + NW;
        const bool two = j4b < nj4;
        const int ua = (j4 < A.y) ? A.x + 4 * j4 : A.z + 4 * (j4 - A.y);
        const int ub = (j4b < A.y) ? A.x + 4 * j4b : A.z + 4 * (j4b - A.y);
        const int t0 = ca.x & 0xffffu, t1 = ca.x >> 16, t2 = ca.y & 0xffffu, t3 = ca.y >> 16;
        double u[4], tv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) u[c] = V[ua + c];
        tv[0] = V[t0]; tv[1] = V[t1]; tv[2] = V[t2]; tv[3] = V[t3];
        if (two) {
          const int s0 = cb.x & 0xffffu, s1 = cb.x >> 16, s2 = cb.y & 0xffffu, s3 = cb.y >> 16;
          double u2[4], sv[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) u2[c] = V[ub + c];
          sv[0] = V[s0]; sv[1] = V[s1]; sv[2] = V[s2]; sv[3] = V[s3];
#pragma unroll
          for (int c = 0; c < 4; ++c) { tv[c] -= lv * u[c]; sv[c] -= lv * u2[c]; }
          V[t0] = tv[0]; V[t1] = tv[1]; V[t2] = tv[2]; V[t3] = tv[3];
          V[s0] = sv[0]; V[s1] = sv[1]; V[s2] = sv[2]; V[s3] = sv[3];
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) tv[c] -= lv * u[c];
          V[t0] = tv[0]; V[t1] = tv[1]; V[t2] = tv[2]; V[t3] = tv[3];
        }
        fresh = nrc != rc;
        ca = na; cb = nb2; lpc = lpn; rc = nrc; j4 = nj; have = hnext;
      }
      if (tid == 0) { const long long tn2 = pclock(); ph[PH_G_FLAT] += tn2 - tq; tq = tn2; }
    }
    int rot = 0;
    for (int gi = L0.z; gi < L1.z; ++gi) {
      const int4 G = sm.grp[gi];     // {width, nblk, ent_off, tgt_off}
      // contiguous block range of this warp; the warp that gets the remainder rotates from group
      // to group so that runs of one-block groups spread over all warps
      const int wr = (w + rot) & (NW - 1);
      rot += G.y;
      const int b_lo = (G.y * wr) / NW, nb = (G.y * (wr + 1)) / NW - b_lo;
      const uint32_t* ep = g.ent + G.z + (size_t)b_lo * G.x * 32 + l;
      const uint16_t* tp = g.tgt + G.w + b_lo * 32 + l;
      if (G.x <= 8) {
        switch (G.x) {     // uniform
          case 1: glu_group<1, 8>(ep, tp, nb, V, zp, ZZ); break;
          case 2: glu_group<2, 8>(ep, tp, nb, V, zp, ZZ); break;
          case 3: glu_group<3, 4>(ep, tp, nb, V, zp, ZZ); break;
          case 4: glu_group<4, 4>(ep, tp, nb, V, zp, ZZ); break;
          case 5: glu_group<5, 2>(ep, tp, nb, V, zp, ZZ); break;
          case 6: glu_group<6, 2>(ep, tp, nb, V, zp, ZZ); break;
          case 7: glu_group<7, 2>(ep, tp, nb, V, zp, ZZ); break;
          default: glu_group<8, 2>(ep, tp, nb, V, zp, ZZ); break;
        }
      } else {
        for (int b = 0; b < nb; ++b) {
          const uint32_t* eb = ep + (size_t)b * G.x * 32;
          const uint32_t t = (uint32_t)__ldcg(tp + b * 32);
          double acc[4] = {0.0, 0.0, 0.0, 0.0};
          uint32_t e[16], en[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) e[j] = (j < G.x) ? __ldcg(eb + j * 32) : ZZ;
          for (int j0 = 0; j0 < G.x; j0 += 16) {
            if (j0 + 16 < G.x) {
#pragma unroll
              for (int j = 0; j < 16; ++j) en[j] = (j0 + 16 + j < G.x) ? __ldcg(eb + (j0 + 16 + j) * 32) : ZZ;
            }
            double lv[16], uv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { lv[j] = V[e[j] & 0xffffu]; uv[j] = V[e[j] >> 16]; }
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j & 3] += lv[j] * uv[j];
#pragma unroll
            for (int j = 0; j < 16; ++j) e[j] = en[j];
          }
          if (t != 0xFFFFu) V[t] -= (acc[0] + acc[1]) + (acc[2] + acc[3]);
        }
      }
      if (tid == 0) { const long long tn2 = pclock(); ph[G.x == 1 ? PH_G_FLAT : (G.x <= 8 ? PH_G_NARROW : PH_G_WIDE)] += tn2 - tq; tq = tn2; }
    }
    __syncthreads();
    if (tid == 0) ph[PH_G_PIVMUL] += pclock() - tq;
  }
  const long long t1 = pclock();
  // ---- U_B / L_C in ELL order for the solves
  for (int q = tid; q < net.n_ub; q += NT) ws.ubE[__ldg(net.ub_ellpos + q)] = V[net.o_ub + q];
  for (int q = tid; q < net.n_lc; q += NT) ws.lcE[__ldg(net.lc_ellpos + q)] = V[net.o_lc + q];
  __syncthreads();   // X (= the U_B/L_C part of V) is scratch from here on
  if (tid == 0) { ph[PH_G_COPY] += pclock() - t1; ph[PH_G_LOOP] += t1 - t0b; }
  const long long t2 = pclock();
  // ---- tables of the staged solves and dense copies of the S diagonal blocks -> upper part of X
  {
    uint32_t* tab = (uint32_t*)(sm.X + net.ss.tab);
    for (int q = tid; q < net.ss.blob_words; q += NT) tab[q] = __ldg(net.ss.blob + q);
    double* tiles = sm.X + net.ss.sinv;
    const int ntile = net.ss.nblkS * (33 * 32);
    for (int q = tid; q < ntile; q += NT) { const int t = q % (33 * 32); tiles[q] = (t / 33 == t % 33) ? 1.0 : 0.0; }
    __syncthreads();
    for (int q = tid; q < net.ss.next; q += NT) {
      const uint32_t e = __ldg(net.ss.ext + q);
      tiles[e >> 16] = V[e & 0xffffu];
    }
  }
  // ---- dense tail
  const int dto = (int)(sm.Dt - smem_raw), pbo = (int)(sm.X - smem_raw);
  switch (nt >> 4) {
    case 1: tail_lu<1>(dto, pbo, flag, ph); break; case 2: tail_lu<2>(dto, pbo, flag, ph); break;
    case 3: tail_lu<3>(dto, pbo, flag, ph); break; case 4: tail_lu<4>(dto, pbo, flag, ph); break;
    case 5: tail_lu<5>(dto, pbo, flag, ph); break; case 6: tail_lu<6>(dto, pbo, flag, ph); break;
    case 7: tail_lu<7>(dto, pbo, flag, ph); break; default: tail_lu<8>(dto, pbo, flag, ph); break;
  }
  const long long t2b = pclock();
  const int res = *flag;
  __syncthreads();
  if (tid == 0) {
    const long long t3 = pclock();
    ph[PH_FACT_HEAD] += t1 - t0; ph[PH_FACT_SCHUR] += t2 - t1; ph[PH_FACT_TAIL] += t3 - t2;
    ph[PH_PBUILD] += t0b - t0; ph[PH_TAILINV] += t3 - t2b;
  }
  return res;
}

// P = I - hl0*J (WK = J*CON, +1 on the diagonal; src/opkda1.f:1763-1764) and its numeric LU.
// Returns (to all threads) 0 ok / 1 zero pivot.
template <bool ALLSMEM>
__device__ __forceinline__ int factor(Ws ws, Smem sm, Layout lay, double con, int* flag, unsigned long long* ph) {
  const DevNet& net = c_net;
  const int n = net.n, nh = net.nh, nt = net.nt, ldt = net.ldt;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, tid = threadIdx.x;
  const int NWR = lay.nwr;
  double* wrow = sm.X + w * n;                                  // per-warp dense work row
  // factorisation copy of U_B (CSR values + tail-local columns), in X behind the work rows
  double* ub = (ALLSMEM || lay.ub_smem) ? sm.X + NWR * n : ws.ubG;
  const uint16_t* ubcol = (ALLSMEM || lay.ub_smem) ? (const uint16_t*)(sm.X + NWR * n + net.n_ub) : net.ub_col;
  if (tid == 0) *flag = 0;
  long long t0 = pclock();
  // ---- build P
  for (int s = tid; s < net.n_hh; s += NT) sm.hh[s] = __ldcs(ws.J + s) * con;
  for (int s = tid; s < net.n_ub; s += NT) ub[s] = __ldcs(ws.J + net.o_ub + s) * con;
  if (ALLSMEM || lay.ub_smem) {
    uint16_t* uc = (uint16_t*)(sm.X + NWR * n + net.n_ub);
    for (int s = tid; s < net.n_ub; s += NT) uc[s] = __ldg(net.ub_col + s);
  }
  for (int e = tid; e < ldt * nt; e += NT) sm.Dt[e] = __ldcs(ws.J + net.o_tl + e) * con;
  __syncthreads();
  for (int i = tid; i < nh; i += NT) sm.hh[__ldg(net.pivmeta + i).x - 1] += 1.0;
  for (int a = tid; a < nt; a += NT) sm.Dt[a * ldt + a] += 1.0;
  __syncthreads();
  long long t0b = pclock();
  // ---- head rows, level by level, one warp per row (up-looking through a dense work row over
  // all n columns).  The pivots k of a row and their metadata are fetched 32 at a time (one
  // lane each) so that the sequential k-loop touches shared memory only.
  for (int lev = 0; lev < net.nflev; ++lev) {
    const int rb = sm.flptr[lev], re = sm.flptr[lev + 1];
    for (int ri = rb + w; ri < re && w < NWR; ri += NWR) {
      const int4 fm = __ldg(net.fmeta + ri);
      const int i = fm.x, base = fm.y, nl = fm.z;
      if (nl == 0) {   // nothing to eliminate: U(i,:) = P(i,:)
        if (l == 0) {
          const double d = sm.hh[base];
          if (d == 0.0 || isnan(d)) *flag = 1;
          sm.dinv[i] = 1.0 / d;
        }
        continue;
      }
      const int4 pm = __ldg(net.pivmeta + i);
      const int len = pm.x + pm.y - base, u0 = pm.z, ulen = pm.w;
      for (int q = l; q < len; q += 32) wrow[sm.hhcol[base + q]] = sm.hh[base + q];
      for (int q = l; q < ulen; q += 32) wrow[nh + ubcol[u0 + q]] = ub[u0 + q];
      __syncwarp();
      for (int q0 = 0; q0 < nl; q0 += 32) {
        const int myk = (q0 + l < nl) ? (int)sm.hhcol[base + q0 + l] : 0;
        const int4 mm = __ldg(net.pivmeta + myk);
        const int cnt = (nl - q0) < 32 ? (nl - q0) : 32;
        for (int j = 0; j < cnt; ++j) {
          const int k = __shfl_sync(0xffffffffu, myk, j);
          const int kb = __shfl_sync(0xffffffffu, mm.x, j), klen = __shfl_sync(0xffffffffu, mm.y, j);
          const int ku = __shfl_sync(0xffffffffu, mm.z, j), kul = __shfl_sync(0xffffffffu, mm.w, j);
          const double lv = wrow[k] * sm.dinv[k];
          __syncwarp();
          if (l == 0) sm.hh[base + q0 + j] = lv;
          for (int t = l; t < klen; t += 32) wrow[sm.hhcol[kb + t]] -= lv * sm.hh[kb + t];
          for (int t = l; t < kul; t += 32) wrow[nh + ubcol[ku + t]] -= lv * ub[ku + t];
          __syncwarp();
        }
      }
      if (l == 0) {
        const double d = wrow[i];
        if (d == 0.0 || isnan(d)) *flag = 1;
        sm.dinv[i] = 1.0 / d;
        sm.hh[base + nl] = d;
      }
      for (int q = nl + 1 + l; q < len; q += 32) sm.hh[base + q] = wrow[sm.hhcol[base + q]];
      for (int q = l; q < ulen; q += 32) ub[u0 + q] = wrow[nh + ubcol[u0 + q]];
      __syncwarp();
    }
    __syncthreads();
  }
  long long t1 = pclock();
  // ---- tail rows against the head pivots (rows are independent; longest first):
  // multipliers L_C(a,:) -> ELL copy for the solves, Schur row -> Dt
  for (int ai = w; ai < nt && w < NWR; ai += NWR) {
    const int a = net.tail_order[ai];
    const int base = net.lc_ptr[a], nl = net.lc_ptr[a + 1] - base;
    for (int q = l; q < nl; q += 32) wrow[__ldg(net.lc_col + base + q)] = __ldcs(ws.J + net.o_lc + base + q) * con;
    for (int b = l; b < nt; b += 32) wrow[nh + b] = sm.Dt[(size_t)b * ldt + a];
    __syncwarp();
    for (int q0 = 0; q0 < nl; q0 += 32) {
      const int myk = (q0 + l < nl) ? (int)__ldg(net.lc_col + base + q0 + l) : 0;
      const int myp = (q0 + l < nl) ? __ldg(net.lc_ellpos + base + q0 + l) : 0;
      const int4 mm = __ldg(net.pivmeta + myk);
      const int cnt = (nl - q0) < 32 ? (nl - q0) : 32;
      double mylv = 0.0;
      for (int j = 0; j < cnt; ++j) {
        const int k = __shfl_sync(0xffffffffu, myk, j);
        const int kb = __shfl_sync(0xffffffffu, mm.x, j), klen = __shfl_sync(0xffffffffu, mm.y, j);
        const int ku = __shfl_sync(0xffffffffu, mm.z, j), kul = __shfl_sync(0xffffffffu, mm.w, j);
        const double lv = wrow[k] * sm.dinv[k];
        __syncwarp();
        if (l == j) mylv = lv;
        for (int t = l; t < klen; t += 32) wrow[sm.hhcol[kb + t]] -= lv * sm.hh[kb + t];
        for (int t = l; t < kul; t += 32) wrow[nh + ubcol[ku + t]] -= lv * ub[ku + t];
        __syncwarp();
      }
      if (q0 + l < nl) ws.lcE[myp] = mylv;      // multipliers -> ELL copy, one coalesced-ish store per chunk
    }
    for (int b = l; b < nt; b += 32) sm.Dt[(size_t)b * ldt + a] = wrow[nh + b];
    __syncwarp();
  }
  __syncthreads();
  // U_B in ELL order for the solves
  for (int s = tid; s < net.n_ub; s += NT) ws.ubE[__ldg(net.ub_ellpos + s)] = ub[s];
  __syncthreads();   // X (work rows, U_B) is reused as scratch by the dense tail below
  if (!lay.hh_smem) __threadfence_block();
  long long t2 = pclock();
  // ---- dense tail
  const int dto = (int)(sm.Dt - smem_raw), pbo = (int)(sm.X - smem_raw);
  switch (nt >> 4) {
    case 1: tail_lu<1>(dto, pbo, flag, ph); break; case 2: tail_lu<2>(dto, pbo, flag, ph); break;
    case 3: tail_lu<3>(dto, pbo, flag, ph); break; case 4: tail_lu<4>(dto, pbo, flag, ph); break;
    case 5: tail_lu<5>(dto, pbo, flag, ph); break; case 6: tail_lu<6>(dto, pbo, flag, ph); break;
    case 7: tail_lu<7>(dto, pbo, flag, ph); break; default: tail_lu<8>(dto, pbo, flag, ph); break;
  }
  long long t2b = pclock();
  const int res = *flag;
  __syncthreads();
  if (tid == 0) {
    long long t3 = pclock();
    ph[PH_FACT_HEAD] += t1 - t0; ph[PH_FACT_SCHUR] += t2 - t1; ph[PH_FACT_TAIL] += t3 - t2;
    ph[PH_PBUILD] += t0b - t0; ph[PH_TAILINV] += t3 - t2b;
  }
  return res;
}

// SpMV with an ELL coupling block: out[row] -= sum val * x[col]
__device__ __forceinline__ void spmv_sub(const EllDev& e, const double* __restrict__ val, const double* x,
                                         double* out, double* partial) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int b = w; b < e.nblk; b += NW) {
    const int off = e.blk_off[b], width = e.blk_width[b];
    double acc = 0.0;
    for (int j0 = 0; j0 < width; j0 += 16) {
      double vv[16]; int cc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        vv[j] = (j0 + j < width) ? __ldcs(val + off + (j0 + j) * 32 + l) : 0.0;
        cc[j] = (j0 + j < width) ? (int)__ldg(e.col + off + (j0 + j) * 32 + l) : 0;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) if (j0 + j < width) acc += vv[j] * x[cc[j]];
    }
    const int t = e.sub_target[b * 32 + l];
    if (t >= 0) out[t] -= acc;
    else if (t <= -2) partial[-2 - t] = acc;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < e.ncombine; q += NT) {
    double s = 0.0;
    for (int p = e.comb_ptr[q]; p < e.comb_ptr[q + 1]; ++p) s += partial[p];
    out[e.comb_row[q]] -= s;
  }
  __syncthreads();
}

// DSOLSS for the level-parallel mode: staged head sweeps (HostNet::SolveSched, tables in
// shared memory), SpMVs with the coupling blocks whose values are fetched into registers
// one phase ahead; the diagonal blocks of the S rows and the dense tail are solved by plain
// substitution on warp 0 (warp_trisolve).
__device__ __forceinline__ double group_sum(double v, int lpr) {
  for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void head_stage(const int4 S, const GSm& sm, const uint32_t* ent, const uint16_t* rp,
                                           const uint16_t* rows, double* tmp, bool upper) {
  const DevNet& net = c_net;
  const int tid = threadIdx.x;
  const int kind = S.x & 255, lg = (S.x >> 8) & 255, lpr = 1 << lg, nrows = S.y & 0xffff, blk = S.y >> 16;
  const int r = tid >> lg, sub = tid & (lpr - 1);
  double a0 = 0.0, a1 = 0.0;
  int row = 0;
  if (r < nrows) {
    row = rows[S.z + r];
    int q = rp[S.w + r] + sub;
    const int q1 = rp[S.w + r + 1];
    for (; q + lpr < q1; q += 2 * lpr) {
      const uint32_t e0 = ent[q], e1 = ent[q + lpr];
      a0 += sm.V[e0 & 0xffffu] * sm.xb[e0 >> 16];
      a1 += sm.V[e1 & 0xffffu] * sm.xb[e1 >> 16];
    }
    if (q < q1) { const uint32_t e0 = ent[q]; a0 += sm.V[e0 & 0xffffu] * sm.xb[e0 >> 16]; }
  }
  const double acc = group_sum(a0 + a1, lpr);
  if (kind != 2) {
    if (r < nrows && sub == 0) {
      const double v = sm.xb[row] - acc;
      sm.xb[row] = (kind == 1) ? v * sm.dinv[row] : v;
    }
    __syncthreads();
    return;
  }
  // S block: the rows' coupling inside the 32-row diagonal block (dense copy at X + sinv, ld 33,
  // identity-padded) is resolved by substitution on warp 0
  if (sub == 0) tmp[r] = (r < nrows) ? sm.xb[row] - acc : 0.0;
  __syncthreads();
  if (tid < 32) {
    const int t_off = (int)(sm.X - smem_raw) + net.ss.sinv + blk * (33 * 32), tmp_off = (int)(tmp - smem_raw);
    const int myrow = (tid < nrows) ? (int)rows[S.z + tid] : 0;
    if (upper) {
      tmp[32 + tid] = (tid < nrows) ? sm.dinv[myrow] : 1.0;     // reciprocal pivots of the block's rows
      __syncwarp();
      warp_trisolve_upper<1>(t_off, 33, 32, tmp_off + 32, tmp_off);
    } else warp_trisolve_lower<1>(t_off, 33, 32, tmp_off);
    __syncwarp();
    if (tid < nrows) sm.xb[myrow] = tmp[tid];
  }
  __syncthreads();
}

// coupling-block SpMV, split in two so that the loads fly while other phases run:
// each warp owns blocks w (and w + NW when U = 2) of the ELL copy (<= 16 entries per sub-row)
template <int U> struct EllRegs { double v[U][16]; uint32_t c[U][8]; };
template <int U>
__device__ __forceinline__ void ell_fetch(const EllDev& e, const double* __restrict__ val, EllRegs<U>& R) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int b = w + u * NW;
    const bool on = b < e.nblk;
    const int off = on ? __ldg(e.blk_off + b) : 0, width = on ? __ldg(e.blk_width + b) : 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) R.v[u][j] = (j < width) ? __ldcg(val + off + j * 32 + l) : 0.0;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const uint32_t c0 = (j < width) ? (uint32_t)__ldg(e.col + off + j * 32 + l) : 0u;
      const uint32_t c1 = (j + 1 < width) ? (uint32_t)__ldg(e.col + off + (j + 1) * 32 + l) : 0u;
      R.c[u][j >> 1] = c0 | (c1 << 16);
    }
  }
}
template <int U>
__device__ __forceinline__ void ell_apply(const EllDev& e, const EllRegs<U>& R, const double* x, double* out, double* partial) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int b = w + u * NW;
    if (b < e.nblk) {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        a0 += R.v[u][j] * x[R.c[u][j >> 1] & 0xffffu];
        a1 += R.v[u][j + 1] * x[R.c[u][j >> 1] >> 16];
      }
      const int t = __ldg(e.sub_target + b * 32 + l);
      if (t >= 0) out[t] -= a0 + a1;
      else if (t <= -2) partial[-2 - t] = a0 + a1;
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < e.ncombine; q += NT) {
    double s = 0.0;
    for (int p = __ldg(e.comb_ptr + q); p < __ldg(e.comb_ptr + q + 1); ++p) s += partial[p];
    out[__ldg(e.comb_row + q)] -= s;
  }
  __syncthreads();
}

template <int U>
__device__ __noinline__ void solve_glu(Ws ws, unsigned long long* ph) {
  const DevNet& net = c_net;
  const GSm sm = glu_smem();
  const int n = net.n, nh = net.nh, nt = net.nt, ldt = net.ldt;
  const int tid = threadIdx.x;
  double* tmp = sm.X;            // [64] block right-hand side + reciprocal pivots; the SpMV partials start behind
  double* part = sm.X + 64;
  const uint32_t* ent = (const uint32_t*)(sm.X + net.ss.tab);
  const uint16_t* rp = (const uint16_t*)(ent + net.ss.nent);
  const uint16_t* rows = rp + ((net.ss.nrp + 1) & ~1);
  const long long t0 = pclock();
  EllRegs<U> R;
  ell_fetch(net.lcE, ws.lcE, R);           // lands while the head sweep runs
  for (int i = tid; i < n; i += NT) sm.xb[i] = sm.y[__ldg(net.perm + i)];
  if (tid == 0) *(long long*)(smem_raw + 3 * n + 32) = 0;     // progress counters of the tail sweeps
  __syncthreads();
  for (int st = 0; st < net.ss.nf; ++st) head_stage(sm.st[st], sm, ent, rp, rows, tmp, false);
  const long long t1 = pclock();
  // ---- tail right-hand side: x_T -= L_C x_H
  ell_apply(net.lcE, R, sm.xb, sm.xb + nh, part);
  ell_fetch(net.ubE, ws.ubE, R);           // lands while the tail is solved
  const long long t1b = pclock();
  // ---- dense tail: forward and backward substitution on warp 0
  if (tid < 128)
    tail_trisolve_mw((int)(sm.Dt - smem_raw), ldt, nt, (int)(sm.dinv - smem_raw) + nh, (int)(sm.xb - smem_raw) + nh, 3 * n + 32, ph);
  __syncthreads();
  const long long t1c = pclock();
  // ---- head right-hand side: x_H -= U_B x_T
  ell_apply(net.ubE, R, sm.xb + nh, sm.xb, part);
  const long long t2 = pclock();
  for (int st = net.ss.nf; st < net.ss.nf + net.ss.nb; ++st) head_stage(sm.st[st], sm, ent, rp, rows, tmp, true);
  for (int i = tid; i < n; i += NT) sm.y[__ldg(net.perm + i)] = sm.xb[i];
  __syncthreads();
  if (tid == 0) {
    const long long t3 = pclock();
    ph[PH_S_FWD] += t1 - t0; ph[PH_S_TAIL] += t2 - t1; ph[PH_S_BWD] += t3 - t2;
    ph[PH_S_SPMV] += (t1b - t1) + (t2 - t1c);
  }
}

// DSOLSS: sm.y <- P^{-1} sm.y (original species order in, original order out)
// (wiped: P is pw*I, see DPRJS below)
__device__ __forceinline__ void solve(Ws ws, Smem sm, bool wiped, double pw) {
  const DevNet& net = c_net;
  const int n = net.n, nh = net.nh, nt = net.nt, ldt = net.ldt;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, tid = threadIdx.x;
  if (wiped) {
    for (int i = tid; i < n; i += NT) sm.y[i] = sm.y[i] / pw;
    __syncthreads();
    return;
  }
  for (int i = tid; i < n; i += NT) sm.xb[i] = sm.y[net.perm[i]];
  if (tid == 0) *(long long*)(smem_raw + 3 * n + 32) = 0;     // progress counters of the tail sweeps
  __syncthreads();
  // ---- forward, head: fat levels by the CTA, thin levels (<= 32 rows) by warp 0
  for (int lev = 1; lev < net.nfat_f; ++lev) {
    const int rb = sm.flptr[lev], re = sm.flptr[lev + 1];
    for (int ri = rb + tid; ri < re; ri += NT) {
      const int4 fm = __ldg(net.fmeta + ri);
      double s = sm.xb[fm.x];
      for (int q = 0; q < fm.z; ++q) s -= sm.hh[fm.y + q] * sm.xb[sm.hhcol[fm.y + q]];
      sm.xb[fm.x] = s;
    }
    __syncthreads();
  }
  if (w == 0) {
    const int off = net.flev_nfat_rows;
    for (int lev = (net.nfat_f > 1 ? net.nfat_f : 1); lev < net.nflev; ++lev) {
      const int rb = sm.flptr[lev], re = sm.flptr[lev + 1];
      if (rb + l < re) {
        const int4 fm = sm.fthin[rb + l - off];
        double s = sm.xb[fm.x];
        for (int q = 0; q < fm.z; ++q) s -= sm.hh[fm.y + q] * sm.xb[sm.hhcol[fm.y + q]];
        sm.xb[fm.x] = s;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // ---- tail right-hand side: x_T -= L_C x_H
  spmv_sub(net.lcE, ws.lcE, sm.xb, sm.xb + nh, sm.X);
  // ---- dense tail: forward and backward substitution on warp 0
  if (tid < 128)
    tail_trisolve_mw((int)(sm.Dt - smem_raw), ldt, nt, (int)(sm.dinv - smem_raw) + nh, (int)(sm.xb - smem_raw) + nh, 3 * n + 32, nullptr);
  __syncthreads();
  // ---- head right-hand side: x_H -= U_B x_T
  spmv_sub(net.ubE, ws.ubE, sm.xb + nh, sm.xb, sm.X);
  // ---- backward, head
  for (int lev = 0; lev < net.nfat_b; ++lev) {
    const int rb = sm.suptr[lev], re = sm.suptr[lev + 1];
    for (int ri = rb + tid; ri < re; ri += NT) {
      const int4 bm = __ldg(net.bmeta + ri);
      double s = sm.xb[bm.x];
      for (int q = 0; q < bm.z; ++q) s -= sm.hh[bm.y + q] * sm.xb[sm.hhcol[bm.y + q]];
      sm.xb[bm.x] = s * sm.dinv[bm.x];
    }
    __syncthreads();
  }
  if (w == 0) {
    const int off = net.su_nfat_rows;
    for (int lev = net.nfat_b; lev < net.nsu; ++lev) {
      const int rb = sm.suptr[lev], re = sm.suptr[lev + 1];
      if (rb + l < re) {
        const int4 bm = sm.bthin[rb + l - off];
        double s = sm.xb[bm.x];
        for (int q = 0; q < bm.z; ++q) s -= sm.hh[bm.y + q] * sm.xb[sm.hhcol[bm.y + q]];
        sm.xb[bm.x] = s * sm.dinv[bm.x];
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += NT) sm.y[net.perm[i]] = sm.xb[i];
  __syncthreads();
}

// ---------------------------------------------------------------------------
// labels of the DSTODE state machine (src/opkda1.f:746-1124)
enum { L100, L150, L160, L170, L175, L200, L220, L250, L270, L410, L430, L450, L500, L520, L540,
       L610, L620, L630, L640, L660, L670, L680, L690, L700, L720 };

struct Lsodes {      // COMMON /DLS001/ + /DLSS01/ (uniform across the CTA, held in registers)
  double CONIT, CRATE, EL1, EL2, EL3, EL4, EL5, EL6, HOLD, RMAX, EL0, H, HMXI, HU, RC, TN;
  double CON0, CONMIN, TCRIT;
  int MXSTEP, NSLAST, IALTH, IPUP, NSLP, ICF, IERPJ, IERSL, JCUR, JSTART, KFLAG, L;
  int NQ, NST, NFE, NJE, NQU, NSLJ, NLU, IMXER, IPLOST, INIT;
  // constants of the method as the reference's driver sets them (src/opkdmain.f:3305-3322)
  static constexpr double UROUND = 2.220446049250313e-16, CCMAX = 0.3, CCMXJ = 0.2, HMIN = 0.0,
                          PSMALL = 1000.0 * 2.220446049250313e-16, RBIG = 0.01 / (1000.0 * 2.220446049250313e-16);
  static constexpr int MAXORD = 5, LMAX = 6, MAXCOR = 3, MSBP = 20, MXNCF = 10, MSBJ = 50;
  int wiped;         // saved P zeroed by ISTATE=1/3 preprocessing and not yet rebuilt from a fresh J
  double pw;         // while wiped: P = pw * I
  bool IHIT;
  long long n_solve, n_cfail, n_efail;
};

template <int EPT, bool GLU>
__global__ void __launch_bounds__(NT, 1)
integrate_kernel(const BatchArgs args) {
  const DevNet& net = c_net;
  __shared__ int s_cell, s_flag;
  __shared__ unsigned long long s_ph[PH_COUNT];   // per-phase cycle counters (thread 0 only)
  const int n = net.n, NEQ = net.NEQ, R = net.R;
  const int tid = threadIdx.x;
  const Layout lay = make_layout(net);
  Ws ws;
  {
    double* p = args.ws + (size_t)blockIdx.x * args.ws_stride;
    ws.J = p; p += net.nstore; ws.ubE = p; p += net.ubE.nval; ws.lcE = p; p += net.lcE.nval;
    ws.ksave = p; p += R; ws.hhG = p; p += net.n_hh; ws.ubG = p; p += net.n_ub;
    ws.good = p;
  }
  Smem sm;
  if (GLU) {
    double* p = smem_raw;
    sm.y = p; p += n; sm.savf = p; p += n; sm.xb = sm.savf; sm.dinv = p; p += n;
    sm.par = p; p += 32; sm.red = p;
    double* V = smem_raw + lay.voff;      // the factor in storage order: [hh | U_B | L_C | tail]
    sm.hh = V; sm.X = V + net.o_ub; sm.Dt = V + net.o_tl;
    sm.hhcol = nullptr; sm.fthin = nullptr; sm.bthin = nullptr; sm.flptr = nullptr; sm.suptr = nullptr;
    int4* dsc = (int4*)(smem_raw + net.glu.doff);
    for (int q = tid; q < net.glu.ndesc; q += NT) dsc[q] = net.glu.desc[q];
  } else {
    double* p = smem_raw;
    sm.y = p; p += n; sm.savf = p; p += n; sm.xb = sm.savf; sm.dinv = p; p += n;
    sm.par = p; p += 32; sm.red = p; p += 2 * NW; sm.Dt = p; p += (size_t)net.ldt * net.nt;
    if (lay.hh_smem) { sm.hh = p; p += net.n_hh; } else sm.hh = ws.hhG;
    // index tables (staged below, once per CTA)
    const int nthin_f = net.nh - net.flev_nfat_rows, nthin_b = net.nh - net.su_nfat_rows;
    if ((size_t)p & 15) p += 1;   // int4 tables need 16-byte alignment
    int4* ft = (int4*)p; p += 2 * nthin_f; int4* bt = (int4*)p; p += 2 * nthin_b;
    int* fl = (int*)p; int* su = fl + (net.nflev + 1);
    p += ((size_t)(net.nflev + net.nsu + 2) * 4 + 7) / 8;
    uint16_t* hc = (uint16_t*)p; p += ((size_t)net.n_hh * 2 + 7) / 8;
    for (int q = tid; q < nthin_f; q += NT) ft[q] = net.fmeta[net.flev_nfat_rows + q];
    for (int q = tid; q < nthin_b; q += NT) bt[q] = net.bmeta[net.su_nfat_rows + q];
    for (int q = tid; q <= net.nflev; q += NT) fl[q] = net.flev_ptr[q];
    for (int q = tid; q <= net.nsu; q += NT) su[q] = net.su_ptr[q];
    for (int q = tid; q < net.n_hh; q += NT) hc[q] = net.hh_col[q];
    sm.fthin = ft; sm.bthin = bt; sm.flptr = fl; sm.suptr = su; sm.hhcol = hc;
    sm.X = p;
  }
  const int x_off = (int)(sm.X - smem_raw);   // scratch region: fluxes [R] | gather partials
  const bool ell1 = net.ubE.nblk <= NW && net.lcE.nblk <= NW;   // one ELL block per warp in the solves' SpMVs
  const double* ks = ws.ksave;
  unsigned long long* const ph = s_ph;
  if (tid < PH_COUNT) s_ph[tid] = 0;
  __syncthreads();
  const long long tk0 = clock64();
  const int ncell = args.ncell;
  // element e of this thread is species i = tid + e*NT
  double yh[EPT][6], acor[EPT], ewt[EPT], rt[EPT], at[EPT];
#define FORE _Pragma("unroll") for (int e = 0, i = tid; e < EPT; ++e, i += NT) if (i < n)
  for (;;) {
    __syncthreads();
    if (tid == 0) s_cell = atomicAdd(args.queue, 1);
    __syncthreads();
    if (s_cell >= ncell) break;
    const int cell = args.order ? args.order[s_cell] : s_cell;
    long long tc = pclock();
    // ---- load the cell
    if (tid < RACG_NPAR) sm.par[tid] = args.cellpar[(size_t)tid * ncell + cell];
    FORE sm.y[i] = args.y0[(size_t)i * ncell + cell];
    const double Tslot = args.y0[(size_t)(NEQ - 1) * ncell + cell];
    __syncthreads();
    const double DS = sm.par[RACG_P_ratioDust2HnucNum] * sm.par[RACG_P_SitesPerGrain];
    // tolerances: given, or chem_set_solver_flags_alt(j) (src/chemistry.f90:205-268)
    double rtT, atT;
    if (args.rtol) {
      FORE { rt[e] = args.rtol[(size_t)i * ncell + cell]; at[e] = args.atol[(size_t)i * ncell + cell]; }
      rtT = args.rtol[(size_t)(NEQ - 1) * ncell + cell]; atT = args.atol[(size_t)(NEQ - 1) * ncell + cell];
    } else {
      const int j = args.sp.tol_policy_j;
      const double RT = args.sp.RTOL, AT = args.sp.ATOL, D = sm.par[RACG_P_ratioDust2HnucNum];
      double r, a;
      if (j == 1) { r = RT; a = AT; rtT = 1e-3; atT = 1e-1; }
      else if (j == 2) { r = fmin(RT * 1e1, 1e-4); a = fmin(AT * 1e5, 1e-25); rtT = 1e-2; atT = 1e-1; }
      else if (j == 3) { r = fmin(RT * 1e2, 1e-4); a = fmin(AT * 1e10, 1e-20); rtT = 1e-3; atT = 1.0; }
      else if (j == 4) { r = fmin(RT * 1e2, 1e-4); a = fmin(AT * 1e10, 1e-18); rtT = 1e-3; atT = 1.0; }
      else { r = fmin(RT * pow(2.0, (double)j), 1e-3); a = fmin(AT * pow(1e2, (double)j), 1e-15); rtT = 1e-2; atT = 1.0; }
      // stage through shared memory (xb is free here) to apply the per-species overrides in order
      double* trt = sm.X; double* tat = sm.X + n;
      FORE { trt[i] = r; tat[i] = a; }
      __syncthreads();
      if (tid < 10 && net.hc_idx[tid] >= 0) { trt[net.hc_idx[tid]] = fmax(RT, 1e-4); tat[net.hc_idx[tid]] = fmax(AT, 1e-30); }
      __syncthreads();
      if (tid == 0 && net.iGrain0 >= 0) {
        const int g3[3] = {net.iGrain0, net.iGrainM, net.iGrainP};
        for (int q = 0; q < 3; ++q) if (g3[q] >= 0) { trt[g3[q]] = 1e-4; tat[g3[q]] = fmax(D * 1e-6, 1e-30); }
      }
      __syncthreads();
      for (int q = tid; q < net.ngrain; q += NT) { trt[net.grain_idx[q]] = fmax(RT, 1e-3); tat[net.grain_idx[q]] = fmax(AT, D * 1e-8); }
      __syncthreads();
      FORE { rt[e] = trt[i]; at[e] = tat[i]; }
      __syncthreads();
    }
    // ---- K1: rate coefficients of this cell -> workspace (L2)
    {
      CellCommon cc;
      cell_common(net.cfg, [&](int k) { return sm.par[k]; }, cc);
      for (int r = tid; r < R; r += NT) ws.ksave[r] = rate_coeff(net, cc, r);
      __syncthreads();
      for (int d = tid; d < net.ndup; d += NT) resolve_dupli(net, cc.Tgas, d, [&](int z) { ws.ksave[z] = 0.0; });
      __syncthreads();
    }
    if (tid == 0) { long long t = pclock(); ph[PH_RATES] += t - tc; }
    if (args.dbg_J) {   // diagnostics: in-kernel f and J at y0
      eval_f(ks, x_off, DS, ph);
      for (int i = tid; i < n; i += NT) args.y_final[(size_t)i * ncell + cell] = sm.savf[i];
      for (int q = tid; q < net.nstore; q += NT) ws.J[q] = 0.0;
      __syncthreads();
      eval_jac(ks, x_off, ws.J, DS);
      for (int q = tid; q < net.nstore; q += NT) args.dbg_J[(size_t)q * ncell + cell] = ws.J[q];
      __syncthreads();
      if (args.dbg_con != 0.0) {
        const int fl = GLU ? factor_glu(ws, args.dbg_con, &s_flag, ph) : factor<false>(ws, sm, lay, args.dbg_con, &s_flag, ph);
        for (int i = tid; i < n; i += NT) sm.y[i] = sm.savf[i];
        __syncthreads();
        if (GLU) { if (ell1) solve_glu<1>(ws, ph); else solve_glu<2>(ws, ph); } else solve(ws, sm, false, 1.0);
        for (int i = tid; i < n; i += NT) args.y_final[(size_t)i * ncell + cell] = fl ? nan("") : sm.y[i];
        __syncthreads();
      }
      continue;
    }

    // ---- chem_evol_solve (src/chemistry.f90:391-588)
    const double t_max = args.tmax[cell], t_start = args.t0[cell];
    const double ratio = args.sp.ratio_tstep;
    // n_record of chem_evol_solve_prepare_run_once (src/chemistry.f90:1894-1899); nrec_max is only
    // the capacity of touts/record: records beyond it are not stored, the loop is not shortened
    const int n_record = (int)ceil(log((t_max - t_start) / args.dt_first[cell] * (ratio - 1.0) + 1.0) / log(ratio)) + 1;
    const int n_record_formula = n_record;
    // work budget: the reference's cpu_time clock (src/chemistry.f90:428-438, 480-494) replaced by
    // the deterministic model  seconds = c_f NFE + c_jac NJE + c_lu NLU + c_solve n_solve + c_step NST
    const bool budget = args.sp.max_runtime_allowed > 0.0;
    const double max_runtime = args.sp.max_runtime_allowed;
    const double max_time_per_step = 5.0 / (double)n_record * max_runtime;
    double time_laststep = 0.0, runtime_laststep = 1.7976931348623157e308;
    int premature = 0;
    double t = t_start, t_step = args.dt_first[cell], tout = t + t_step;
    int NERR = 0, nerr_c = 0, quality = 0, ISTATE = 1, n_record_real = 1;
    long long aNST = 0, aNFE = 0, aNJE = 0, aNLU = 0, nrestart = 0;
    Lsodes s;
    s.NST = s.NFE = s.NJE = s.NLU = s.NQU = 0; s.HU = 0.0; s.INIT = 0; s.IMXER = 0;
    s.n_solve = s.n_cfail = s.n_efail = 0; s.wiped = 0; s.pw = 0.0;
    s.NQ = 1; s.L = 2; s.H = 0.0; s.TN = t; s.IHIT = false;
    // the harvest rule of calc_this_cell (src/disk.f90:1716-1733): the caller takes the LAST record
    // whose T and X(H2) are not NaN -- tracked here so that the records need not be stored
    double t_good = t_start; int isav = 0;
    auto record_out = [&](int irec) {   // touts(i) = t; record(:,i) = y
      if (args.y_good) {
        const bool fine = !isnan(Tslot) && !(net.iH2 >= 0 && isnan(sm.y[net.iH2]));
        if (fine) {
          for (int i = tid; i < n; i += NT) ws.good[i] = sm.y[i];
          t_good = t; isav = irec;
        }
      }
      if (irec > args.sp.nrec_max) return;
      if (args.touts && tid == 0) args.touts[(size_t)(irec - 1) * ncell + cell] = t;
      if (args.record) {
        double* rp = args.record + (size_t)(irec - 1) * NEQ * ncell + cell;
        for (int i = tid; i < n; i += NT) rp[(size_t)i * ncell] = sm.y[i];
        if (tid == 0) rp[(size_t)(NEQ - 1) * ncell] = Tslot;
      }
    };
    // DVNORM of a register vector against EWT (N = NEQ as in the reference)
    auto wrms_reg = [&](auto getv) -> double {
      double sq = 0.0;
      FORE { const double a = getv(e, i) * ewt[e]; sq += a * a; }
      return sqrt(block_sum(sq, sm.red) / (double)NEQ);
    };
    record_out(1);
    int irec;
    for (irec = 2; irec <= n_record; ++irec) {
      if (tout >= t_max) tout = t_max;
      if (ISTATE == 1) { aNST += s.NST; aNFE += s.NFE; aNJE += s.NJE; aNLU += s.NLU; s.NST = s.NFE = s.NJE = s.NLU = 0; ++nrestart; }
      // =================== DLSODES(ITASK=4, TCRIT=HMAX=t_max, IOPT=1) ===================
      {
        const double TOUT = tout;
        int dl;   // driver label
        enum { D_BLOCKC, D200, D245, D250, D270, D_STEP, D_AFTER, D_INTERP, D345, D400, D420, D560, D580, D_RET };
        if (ISTATE == 1) s.INIT = 0;
        if (ISTATE == 1 && TOUT == t) { dl = D_RET; }
        else if (ISTATE != 1 && s.INIT == 0) { ISTATE = -3; dl = D_RET; }   // "ISTATE > 1 but DLSODES not initialized" (src/opkdmain.f:3087, error 603):
                                                                              // the first step after a restart failed, then chem_evol_solve asks for ISTATE = 3
        else if (ISTATE == 2) dl = D200;
        else {
          // Block B
          s.MXSTEP = args.sp.mxstep_per_interval; if (s.MXSTEP == 0) s.MXSTEP = 500;
          s.HMXI = (t_max > 0.0) ? 1.0 / t_max : 0.0;
          s.wiped = 1; s.pw = 0.0;   // DPREP zeroes the saved P (src/opkda1.f:1493-1494)
          if (ISTATE == 3) { s.JSTART = -1; dl = D200; }
          else dl = D_BLOCKC;
        }
        double H0 = 0.0;
        for (;;) {
          if (dl == D_RET) break;
          switch (dl) {
            case D_BLOCKC: {
              s.TN = t; s.NST = 0; s.H = 1.0;
              FORE { yh[e][0] = sm.y[i]; yh[e][2] = 0.0; yh[e][3] = 0.0; yh[e][4] = 0.0; yh[e][5] = 0.0; }
              { long long ta = pclock(); eval_f(ks, x_off, DS, ph); if (tid == 0) ph[PH_F] += pclock() - ta; }
              FORE yh[e][1] = sm.savf[i];
              s.NFE = 1;
              bool bad = false;
              FORE { const double ew = rt[e] * fabs(yh[e][0]) + at[e]; if (ew <= 0.0) bad = true; ewt[e] = 1.0 / ew; }
              if (rtT * fabs(Tslot) + atT <= 0.0) bad = true;   // EWT(NEQ): the temperature slot
              if (__syncthreads_or(bad)) { ISTATE = -3; dl = D_RET; break; }
              s.TCRIT = t_max;
              if ((s.TCRIT - TOUT) * (TOUT - t) < 0.0) { ISTATE = -3; dl = D_RET; break; }
              s.JSTART = 0; s.NSLJ = 0; s.NJE = 0; s.NLU = 0; s.NSLAST = 0; s.HU = 0.0; s.NQU = 0;
              s.IPLOST = 0; s.CON0 = 0.0; s.CONMIN = 0.0;
              const double TDIST = fabs(TOUT - t), W0 = fmax(fabs(t), fabs(TOUT));
              if (TDIST < 2.0 * s.UROUND * W0) { ISTATE = -3; dl = D_RET; break; }
              double TOL = rtT;
              FORE TOL = fmax(TOL, rt[e]);
              TOL = block_max(TOL, sm.red);
              if (TOL <= 0.0) {
                double tl = 0.0;
                FORE { const double ay = fabs(yh[e][0]); if (ay != 0.0) tl = fmax(tl, at[e] / ay); }
                tl = block_max(tl, sm.red);
                TOL = fmax(TOL, tl);
                if (Tslot != 0.0) TOL = fmax(TOL, atT / fabs(Tslot));
              }
              TOL = fmax(TOL, 100.0 * s.UROUND);
              TOL = fmin(TOL, 0.001);
              double SUM = wrms_reg([&](int e, int) { return yh[e][1]; });
              SUM = 1.0 / (TOL * W0 * W0) + TOL * SUM * SUM;
              H0 = 1.0 / sqrt(SUM);
              H0 = fmin(H0, TDIST);
              H0 = copysign(H0, TOUT - t);
              const double RH = fabs(H0) * s.HMXI;
              if (RH > 1.0) H0 = H0 / RH;
              s.H = H0;
              FORE yh[e][1] = H0 * yh[e][1];
              dl = D270;
              break;
            }
            case D200: {
              s.NSLAST = s.NST;
              s.TCRIT = t_max;
              if ((s.TN - s.TCRIT) * s.H > 0.0) { ISTATE = -3; dl = D_RET; break; }
              if ((s.TCRIT - TOUT) * s.H < 0.0) { ISTATE = -3; dl = D_RET; break; }
              if ((s.TN - TOUT) * s.H < 0.0) { dl = D245; break; }
              dl = D_INTERP;
              break;
            }
            case D245: {
              const double HMX = fabs(s.TN) + fabs(s.H);
              s.IHIT = fabs(s.TN - s.TCRIT) <= 100.0 * s.UROUND * HMX;
              if (s.IHIT) { dl = D400; break; }
              const double TNEXT = s.TN + s.H * (1.0 + 4.0 * s.UROUND);
              if ((TNEXT - s.TCRIT) * s.H <= 0.0) { dl = D250; break; }
              s.H = (s.TCRIT - s.TN) * (1.0 - 4.0 * s.UROUND);
              if (ISTATE == 2) s.JSTART = -2;
              dl = D250;
              break;
            }
            case D250: {
              if ((s.NST - s.NSLAST) >= s.MXSTEP) { ISTATE = -1; dl = D580; break; }
              bool bad = false;
              FORE { const double ew = rt[e] * fabs(yh[e][0]) + at[e]; if (ew <= 0.0) bad = true; ewt[e] = 1.0 / ew; }
              if (rtT * fabs(Tslot) + atT <= 0.0) bad = true;
              if (__syncthreads_or(bad)) { ISTATE = -6; dl = D580; break; }
              dl = D270;
              break;
            }
            case D270: {
              double sq = 0.0;
              FORE { const double a = yh[e][0] * ewt[e]; sq += a * a; }
              sq = block_sum(sq, sm.red);
              { const double eT = rtT * fabs(Tslot) + atT; const double a = Tslot / eT; sq += a * a; }
              double TOLSF = s.UROUND * sqrt(sq / (double)NEQ);
              if (TOLSF > 1.0) {
                if (s.NST == 0) { ISTATE = -3; dl = D_RET; break; }
                ISTATE = -2; dl = D580; break;
              }
              dl = D_STEP;
              break;
            }
            case D_STEP: {
              // ======================= DSTODE =======================
              long long tv = pclock();
              int pc;
              double DCON, DDN, DEL = 0.0, DELP = 0.0, DSM = 0.0, DUP, R_, RH = 0.0, RHDN, RHSM, RHUP = 0.0, TOLD;
              int IREDO = 0, IRET = 0, M = 0, NCF = 0, NEWQ = 0;
              s.KFLAG = 0; TOLD = s.TN; s.IERPJ = 0; s.IERSL = 0; s.JCUR = 0; s.ICF = 0;
              if (s.JSTART > 0) pc = L200;
              else if (s.JSTART == -1) pc = L100;
              else if (s.JSTART == -2) pc = L160;
              else {
                s.NQ = 1; s.L = 2; s.IALTH = 2; s.RMAX = 10000.0; s.RC = 0.0;
                s.EL0 = 1.0; s.CRATE = 0.7; s.HOLD = s.H; s.NSLP = 0; s.IPUP = 1; IRET = 3;
                pc = L150;
              }
              for (;;) {
                if (pc == L720) break;
                switch (pc) {
                  case L100:
                    s.IPUP = 1;
                    if (s.IALTH == 1) s.IALTH = 2;
                    pc = L160;
                    break;
                  case L150:
                    s.EL1 = net.el[s.NQ][1]; s.EL2 = net.el[s.NQ][2]; s.EL3 = net.el[s.NQ][3];
                    s.EL4 = net.el[s.NQ][4]; s.EL5 = net.el[s.NQ][5]; s.EL6 = net.el[s.NQ][6];
                    s.RC = s.RC * s.EL1 / s.EL0;
                    s.EL0 = s.EL1;
                    s.CONIT = 0.5 / (s.NQ + 2);
                    pc = (IRET == 1) ? L160 : (IRET == 2) ? L170 : L200;
                    break;
                  case L160:
                    if (s.H == s.HOLD) { pc = L200; break; }
                    RH = s.H / s.HOLD; s.H = s.HOLD; IREDO = 3;
                    pc = L175;
                    break;
                  case L170:
                    RH = fmax(RH, s.HMIN / fabs(s.H));
                  case L175: {
                    RH = fmin(RH, s.RMAX);
                    RH = RH / fmax(1.0, fabs(s.H) * s.HMXI * RH);
                    double Rj = 1.0;
#pragma unroll
                    for (int j = 1; j < 6; ++j) { if (j < s.L) { Rj = Rj * RH; FORE yh[e][j] = yh[e][j] * Rj; } }
                    s.H = s.H * RH; s.RC = s.RC * RH; s.IALTH = s.L;
                    pc = (IREDO == 0) ? L690 : L200;
                    break;
                  }
                  case L200: {
                    if (fabs(s.RC - 1.0) > s.CCMAX) s.IPUP = 1;
                    if (s.NST >= s.NSLP + s.MSBP) s.IPUP = 1;
                    s.TN = s.TN + s.H;
                    // Pascal-triangle prediction, same per-element operation order as the flat
                    // YH1 sweep of src/opkda1.f:868-874
#pragma unroll
                    for (int JB = 1; JB <= 5; ++JB) {
                      if (JB <= s.NQ) {
#pragma unroll
                        for (int j = 0; j < 5; ++j) { if (j >= s.NQ - JB && j < s.NQ) FORE yh[e][j] = yh[e][j] + yh[e][j + 1]; }
                      }
                    }
                    pc = L220;
                    break;
                  }
                  case L220: {
                    M = 0;
                    FORE sm.y[i] = yh[e][0];
                    if (tid == 0) ph[PH_VEC] += pclock() - tv;
                    { long long ta = pclock(); eval_f(ks, x_off, DS, ph); if (tid == 0) ph[PH_F] += pclock() - ta; }
                    tv = pclock();
                    s.NFE = s.NFE + 1;
                    if (s.IPUP <= 0) { pc = L250; break; }
                    // ---------------- DPRJS (src/opkda1.f:1735-1838) ----------------
                    {
                      const double HL0 = s.H * s.EL0, CON = -HL0;
                      int JOK = 1;
                      if (s.NST == 0 || s.NST >= s.NSLJ + s.MSBJ) JOK = 0;
                      if (s.ICF == 1 && fabs(s.RC - 1.0) < s.CCMXJ) JOK = 0;
                      if (s.ICF == 2) JOK = 0;
                      if (JOK == 1) {
                        // label 250: the reference rescales the saved P in place,
                        // P <- (P - I)*RCON + I; we rebuild it from the saved J instead.
                        s.JCUR = 0;
                        const double RCON = CON / s.CON0;
                        const double RCONT = fabs(CON) / s.CONMIN;
                        if (RCONT > s.RBIG && s.IPLOST == 1) JOK = 0;
                        else if (s.wiped) {
                          // saved P was zeroed by the ISTATE=1/3 preprocessing (DPREP,
                          // src/opkda1.f:1493-1494): P stays a multiple pw of the identity
                          if (fabs(s.pw - 1.0) < s.PSMALL) { s.IPLOST = 1; s.CONMIN = fmin(fabs(s.CON0), s.CONMIN); }
                          s.pw = (s.pw - 1.0) * RCON + 1.0;
                        } else {
                          // the T-slot diagonal holds P = 1 exactly, so |P-1| < PSMALL and the
                          // reference sets IPLOST on every reuse (src/opkda1.f:1813-1817)
                          s.IPLOST = 1; s.CONMIN = fmin(fabs(s.CON0), s.CONMIN);
                        }
                      }
                      if (JOK == 0) {
                        s.JCUR = 1; s.NJE = s.NJE + 1; s.NSLJ = s.NST; s.IPLOST = 0; s.CONMIN = fabs(CON);
                        if (tid == 0) ph[PH_VEC] += pclock() - tv;
                        long long ta = pclock();
                        eval_jac(ks, x_off, ws.J, DS);
                        if (tid == 0) ph[PH_JAC] += pclock() - ta;
                        tv = pclock();
                        s.wiped = 0;
                      }
                      s.NLU = s.NLU + 1;
                      int flag;
                      if (s.wiped) {
                        flag = (s.pw == 0.0 || isnan(s.pw)) ? 1 : 0;
                      } else {
                        if (tid == 0) ph[PH_VEC] += pclock() - tv;
                        flag = GLU ? factor_glu(ws, CON, &s_flag, ph) : factor<false>(ws, sm, lay, CON, &s_flag, ph);
                        tv = pclock();
                      }
                      s.CON0 = CON;
                      s.IERPJ = flag ? 1 : 0;
                    }
                    s.IPUP = 0; s.RC = 1.0; s.NSLP = s.NST; s.CRATE = 0.7;
                    if (s.IERPJ != 0) { pc = L430; break; }
                    pc = L250;
                    break;
                  }
                  case L250:
                    FORE acor[e] = 0.0;
                  case L270: {
                    __syncthreads();
                    FORE sm.y[i] = s.H * sm.savf[i] - (yh[e][1] + acor[e]);
                    __syncthreads();
                    if (tid == 0) ph[PH_VEC] += pclock() - tv;
                    { long long ta = pclock(); if (s.wiped) { for (int i = tid; i < n; i += NT) sm.y[i] = sm.y[i] / s.pw; __syncthreads(); }
                      else if (GLU) { if (ell1) solve_glu<1>(ws, ph); else solve_glu<2>(ws, ph); }
                      else solve(ws, sm, false, s.pw);
                      if (tid == 0) ph[PH_SOLVE] += pclock() - ta; }
                    tv = pclock();
                    s.n_solve++;
                    DEL = wrms_reg([&](int, int i) { return sm.y[i]; });
                    FORE { acor[e] = acor[e] + sm.y[i]; sm.y[i] = yh[e][0] + s.EL1 * acor[e]; }
                    if (M != 0) s.CRATE = fmax(0.2 * s.CRATE, DEL / DELP);
                    DCON = DEL * fmin(1.0, 1.5 * s.CRATE) / (net.tesco[s.NQ][2] * s.CONIT);
                    if (DCON <= 1.0) { pc = L450; break; }
                    M = M + 1;
                    if (M == s.MAXCOR) { pc = L410; break; }
                    if (M >= 2 && DEL > 2.0 * DELP) { pc = L410; break; }
                    DELP = DEL;
                    if (tid == 0) ph[PH_VEC] += pclock() - tv;
                    { long long ta = pclock(); __syncthreads(); eval_f(ks, x_off, DS, ph); if (tid == 0) ph[PH_F] += pclock() - ta; }
                    tv = pclock();
                    s.NFE = s.NFE + 1;
                    pc = L270;
                    break;
                  }
                  case L410:
                    if (s.JCUR == 1) { pc = L430; break; }
                    s.ICF = 1; s.IPUP = 1;
                    pc = L220;
                    break;
                  case L430: {
                    s.ICF = 2; NCF = NCF + 1; s.n_cfail++; s.RMAX = 2.0; s.TN = TOLD;
#pragma unroll
                    for (int JB = 1; JB <= 5; ++JB) {
                      if (JB <= s.NQ) {
#pragma unroll
                        for (int j = 0; j < 5; ++j) { if (j >= s.NQ - JB && j < s.NQ) FORE yh[e][j] = yh[e][j] - yh[e][j + 1]; }
                      }
                    }
                    if (s.IERPJ < 0 || s.IERSL < 0) { pc = L680; break; }
                    if (fabs(s.H) <= s.HMIN * 1.00001) { pc = L670; break; }
                    if (NCF == s.MXNCF) { pc = L670; break; }
                    RH = 0.25; s.IPUP = 1; IREDO = 1;
                    pc = L170;
                    break;
                  }
                  case L450: {
                    s.JCUR = 0;
                    if (M == 0) DSM = DEL / net.tesco[s.NQ][2];
                    if (M > 0) DSM = wrms_reg([&](int e, int) { return acor[e]; }) / net.tesco[s.NQ][2];
                    if (DSM > 1.0) { pc = L500; break; }
                    s.KFLAG = 0; IREDO = 0; s.NST = s.NST + 1; s.HU = s.H; s.NQU = s.NQ;
#pragma unroll
                    {
                      const double elv[6] = {s.EL1, s.EL2, s.EL3, s.EL4, s.EL5, s.EL6};
#pragma unroll
                      for (int j = 0; j < 6; ++j) { if (j < s.L) FORE yh[e][j] = yh[e][j] + elv[j] * acor[e]; }
                    }
                    s.IALTH = s.IALTH - 1;
                    if (s.IALTH == 0) { pc = L520; break; }
                    if (s.IALTH > 1) { pc = L700; break; }
                    if (s.L == s.LMAX) { pc = L700; break; }
                    FORE yh[e][5] = acor[e];     // YH(:,LMAX), LMAX = 6
                    pc = L700;
                    break;
                  }
                  case L500: {
                    s.KFLAG = s.KFLAG - 1; s.n_efail++; s.TN = TOLD;
#pragma unroll
                    for (int JB = 1; JB <= 5; ++JB) {
                      if (JB <= s.NQ) {
#pragma unroll
                        for (int j = 0; j < 5; ++j) { if (j >= s.NQ - JB && j < s.NQ) FORE yh[e][j] = yh[e][j] - yh[e][j + 1]; }
                      }
                    }
                    s.RMAX = 2.0;
                    if (fabs(s.H) <= s.HMIN * 1.00001) { pc = L660; break; }
                    if (s.KFLAG <= -3) { pc = L640; break; }
                    IREDO = 2; RHUP = 0.0;
                    pc = L540;
                    break;
                  }
                  case L520: {
                    RHUP = 0.0;
                    if (s.L == s.LMAX) { pc = L540; break; }
                    DUP = wrms_reg([&](int e, int) { return acor[e] - yh[e][5]; }) / net.tesco[s.NQ][3];
                    RHUP = 1.0 / (1.4 * pow(DUP, 1.0 / (s.L + 1)) + 0.0000014);
                  }
                  case L540: {
                    RHSM = 1.0 / (1.2 * pow(DSM, 1.0 / s.L) + 0.0000012);
                    RHDN = 0.0;
                    if (s.NQ != 1) {
                      const int Lc = s.L - 1;   // YH(:,L)
                      DDN = wrms_reg([&](int e, int) {
                        double v = 0.0;
#pragma unroll
                        for (int j = 0; j < 6; ++j) if (j == Lc) v = yh[e][j];
                        return v; }) / net.tesco[s.NQ][1];
                      RHDN = 1.0 / (1.3 * pow(DDN, 1.0 / s.NQ) + 0.0000013);
                    }
                    // labels 560-590
                    int sel;   // 0: same order, 1: order down, 2: order up
                    if (RHSM >= RHUP) sel = (RHSM < RHDN) ? 1 : 0;
                    else sel = (RHUP > RHDN) ? 2 : 1;
                    if (sel == 0) { NEWQ = s.NQ; RH = RHSM; pc = L620; }
                    else if (sel == 1) { NEWQ = s.NQ - 1; RH = RHDN; if (s.KFLAG < 0 && RH > 1.0) RH = 1.0; pc = L620; }
                    else {
                      NEWQ = s.L; RH = RHUP;
                      if (RH < 1.1) { pc = L610; break; }
                      R_ = (s.L == 2 ? s.EL2 : s.L == 3 ? s.EL3 : s.L == 4 ? s.EL4 : s.L == 5 ? s.EL5 : s.EL6) / s.L;
#pragma unroll
                      for (int j = 0; j < 6; ++j) { if (j == NEWQ) FORE yh[e][j] = acor[e] * R_; }   // YH(:,NEWQ+1)
                      pc = L630;
                    }
                    break;
                  }
                  case L610:
                    s.IALTH = 3;
                    pc = L700;
                    break;
                  case L620:
                    if (s.KFLAG == 0 && RH < 1.1) { pc = L610; break; }
                    if (s.KFLAG <= -2) RH = fmin(RH, 0.2);
                    if (NEWQ == s.NQ) { pc = L170; break; }
                  case L630:
                    s.NQ = NEWQ; s.L = s.NQ + 1; IRET = 2;
                    pc = L150;
                    break;
                  case L640: {
                    if (s.KFLAG == -10) { pc = L660; break; }
                    RH = 0.1;
                    RH = fmax(s.HMIN / fabs(s.H), RH);
                    s.H = s.H * RH;
                    __syncthreads();
                    FORE sm.y[i] = yh[e][0];
                    if (tid == 0) ph[PH_VEC] += pclock() - tv;
                    { long long ta = pclock(); __syncthreads(); eval_f(ks, x_off, DS, ph); if (tid == 0) ph[PH_F] += pclock() - ta; }
                    tv = pclock();
                    s.NFE = s.NFE + 1;
                    FORE yh[e][1] = s.H * sm.savf[i];
                    s.IPUP = 1; s.IALTH = 5;
                    if (s.NQ == 1) { pc = L200; break; }
                    s.NQ = 1; s.L = 2; IRET = 3;
                    pc = L150;
                    break;
                  }
                  case L660: s.KFLAG = -1; pc = L720; break;
                  case L670: s.KFLAG = -2; pc = L720; break;
                  case L680: s.KFLAG = -3; pc = L720; break;
                  case L690:
                    s.RMAX = 10.0;
                  case L700: {
                    R_ = 1.0 / net.tesco[s.NQU][2];
                    FORE acor[e] = acor[e] * R_;
                    pc = L720;
                    break;
                  }
                }
              }
              s.HOLD = s.H; s.JSTART = 1;
              if (tid == 0) ph[PH_VEC] += pclock() - tv;
              // ===================== end DSTODE =====================
              if (s.KFLAG == 0) { dl = D_AFTER; break; }
              if (s.KFLAG == -1) { ISTATE = -4; dl = D560; break; }
              if (s.KFLAG == -2) { ISTATE = -5; dl = D560; break; }
              ISTATE = -7; dl = D580;
              break;
            }
            case D_AFTER: {
              s.INIT = 1;
              if ((s.TN - TOUT) * s.H < 0.0) { dl = D345; break; }
              dl = D_INTERP;
              break;
            }
            case D_INTERP: {   // DINTDY(TOUT, 0, ...) ; T = TOUT ; goto 420
              const double S_ = (TOUT - s.TN) / s.H;
              __syncthreads();
              FORE {
                double v = 0.0;
#pragma unroll
                for (int j = 5; j >= 0; --j) { if (j == s.NQ) v = yh[e][j]; else if (j < s.NQ) v = yh[e][j] + S_ * v; }
                sm.y[i] = v;
              }
              __syncthreads();
              t = TOUT;
              dl = D420;
              break;
            }
            case D345: {
              const double HMX = fabs(s.TN) + fabs(s.H);
              s.IHIT = fabs(s.TN - s.TCRIT) <= 100.0 * s.UROUND * HMX;
              if (s.IHIT) { dl = D400; break; }
              const double TNEXT = s.TN + s.H * (1.0 + 4.0 * s.UROUND);
              if ((TNEXT - s.TCRIT) * s.H <= 0.0) { dl = D250; break; }
              s.H = (s.TCRIT - s.TN) * (1.0 - 4.0 * s.UROUND);
              s.JSTART = -2;
              dl = D250;
              break;
            }
            case D400: {
              __syncthreads();
              FORE sm.y[i] = yh[e][0];
              __syncthreads();
              t = s.TN;
              if (s.IHIT) t = s.TCRIT;
              dl = D420;
              break;
            }
            case D420: ISTATE = 2; dl = D_RET; break;
            case D560: {   // IMXER = first index of max |ACOR*EWT| (T slot contributes 0)
              double big = 0.0; int im = 0x7fffffff;
              FORE { const double sz = fabs(acor[e] * ewt[e]); if (sz > big) { big = sz; im = i; } }
              for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, big, o); const int oi = __shfl_xor_sync(0xffffffffu, im, o);
                if (ob > big || (ob == big && oi < im)) { big = ob; im = oi; }
              }
              __syncthreads();
              if ((tid & 31) == 0) { sm.red[tid >> 5] = big; sm.red[NW + (tid >> 5)] = (double)im; }
              __syncthreads();
              big = sm.red[0]; im = (int)sm.red[NW];
              for (int k = 1; k < NW; ++k) {
                const double ob = sm.red[k]; const int oi = (int)sm.red[NW + k];
                if (ob > big || (ob == big && oi < im)) { big = ob; im = oi; }
              }
              s.IMXER = (big > 0.0) ? im + 1 : 1;
              dl = D580;
              break;
            }
            case D580: {
              __syncthreads();
              FORE sm.y[i] = yh[e][0];
              __syncthreads();
              t = s.TN;
              dl = D_RET;
              break;
            }
          }
        }
      }
      // =================== back in chem_evol_solve ===================
      __syncthreads();
      record_out(irec);
      n_record_real = irec;
      if (budget) {
        const double time_thisstep = net.rt_coef[0] * (double)(aNFE + s.NFE) + net.rt_coef[1] * (double)(aNJE + s.NJE) +
                                     net.rt_coef[2] * (double)(aNLU + s.NLU) + net.rt_coef[3] * (double)s.n_solve +
                                     net.rt_coef[4] * (double)(aNST + s.NST);
        const double runtime_thisstep = time_thisstep - time_laststep;
        if (runtime_thisstep > fmax(10.0 * runtime_laststep, 0.5 * max_runtime) || time_thisstep > max_runtime) {
          premature = 1; break;   // 'Premature finish', src/chemistry.f90:482-488
        }
        if (runtime_thisstep > max_time_per_step) ISTATE = 1;
        time_laststep = time_thisstep; runtime_laststep = runtime_thisstep;
      }
      if (t >= t_max) break;
      if (ISTATE < 0) {
        NERR += 1; nerr_c += 1;
        if (ISTATE == -4 || ISTATE == -5) {   // ode_solver_error_handling, src/chemistry.f90:326-337
          const int idx = s.IMXER - 1;
          FORE { if (i == idx) { rt[e] = fmin(rt[e] * 10.0, 1e-3); at[e] = fmin(at[e] * 100.0, 1e-20); } }
        }
        if (ISTATE == -7) { quality += 1024; break; }
        if (ISTATE == -3) { quality += 256; break; }
        if (nerr_c < 3) ISTATE = 3; else { ISTATE = 1; nerr_c = 0; }
      }
      {
        bool bad = isnan(Tslot) || Tslot <= 0.0;
        if (net.igH2 >= 0 && fabs(sm.y[net.igH2]) > 1.0) bad = true;
        if (net.igH2O >= 0 && fabs(sm.y[net.igH2O]) > 1.0) bad = true;
        if (net.igH >= 0 && fabs(sm.y[net.igH]) > 1.0) bad = true;
        if (net.iH >= 0 && fabs(sm.y[net.iH]) > 2.0) bad = true;
        if (net.iE >= 0 && fabs(sm.y[net.iE]) > 1.0) bad = true;
        if (bad) { quality += 512; break; }
      }
      if (irec % args.sp.steps_reset_solver == 0) ISTATE = 1;
      t_step = t_step * ratio;
      tout = t + t_step;
    }
    aNST += s.NST; aNFE += s.NFE; aNJE += s.NJE; aNLU += s.NLU;
    // records after an early exit are filled with the last state (src/chemistry.f90:570-575)
    if (args.touts || args.record) for (int r2 = n_record_real + 1; r2 <= args.sp.nrec_max; ++r2) record_out(r2);
    if (NERR > (int)(0.1f * (float)n_record_formula)) quality += 1;
    if (t <= 0.5 * t_max) quality += 2;
    // ---- write results
    if (args.y_good) {
      __syncthreads();
      double gsum = 0.0;     // get_ice_coverage (src/chemistry.f90:989-1003): molecules per grain
      for (int i = tid; i < n; i += NT) args.y_good[(size_t)i * ncell + cell] = ws.good[i];
      for (int q = tid; q < net.ngrain; q += NT) gsum += ws.good[net.grain_idx[q]];
      gsum = block_sum(gsum, sm.red);
      if (tid == 0) {
        args.y_good[(size_t)(NEQ - 1) * ncell + cell] = Tslot;
        args.t_good[cell] = t_good; args.isav[cell] = isav;
        if (args.side) {
          CellCommon cc;
          cell_common(net.cfg, [&](int k) { return sm.par[k]; }, cc);
          args.side[cell] = (net.h2form_reac >= 0) ? rate_coeff_raw(net, cc, net.h2form_reac) : 0.0;
          args.side[ncell + cell] = gsum / sm.par[RACG_P_ratioDust2HnucNum];
        }
      }
    }
    for (int i = tid; i < n; i += NT) args.y_final[(size_t)i * ncell + cell] = sm.y[i];
    if (tid == 0) {
      args.y_final[(size_t)(NEQ - 1) * ncell + cell] = Tslot;
      args.t_final[cell] = t; args.nrec_real[cell] = n_record_real; args.istate[cell] = ISTATE;
      args.quality[cell] = quality;
      double* st = args.stats + cell;
      const double sv[RACG_NSTAT] = {(double)aNST, (double)aNFE, (double)aNJE, (double)aNLU, (double)s.NQU,
                                     (double)s.n_solve, (double)NERR, (double)nrestart, (double)s.n_cfail,
                                     (double)s.n_efail, (double)n_record_real, (double)ISTATE, s.HU,
                                     net.rt_coef[0] * (double)aNFE + net.rt_coef[1] * (double)aNJE + net.rt_coef[2] * (double)aNLU +
                                         net.rt_coef[3] * (double)s.n_solve + net.rt_coef[4] * (double)aNST,
                                     (double)premature, 0};
      for (int k = 0; k < RACG_NSTAT; ++k) st[(size_t)k * ncell] = sv[k];
      ph[PH_NCELL] += 1;
    }
  }
  if (tid == 0 && args.phase) {
    ph[PH_TOTAL] = clock64() - tk0;
#pragma unroll
    for (int k = 0; k < PH_COUNT; ++k) atomicAdd(&args.phase[k], ph[k]);
  }
#undef FORE
}

// Decides whether the level-parallel mode fits (whole factor V = [hh | U_B | L_C | tail] in
// shared memory, scratch X = the U_B/L_C part of V, free between factorisations) and plans X:
//   [0, xlow)            scratch of f / Jacobian / SpMV partials / tail_lu buffers / inverse scratch
//   [sinv, +nblkS*33*32) inverses of the S diagonal blocks (column-major, ld 33)
//   [tab, +blob)         tables of the staged head solves
// cost of every cell of the batch just integrated (in units of one triangular solve), kept by
// the handle: when the next batch has the same number of cells -- the disk code re-integrates
// the same grid every structure iteration -- its work queue is served heaviest first.  The
// order is built on the device by a counting sort over 4096 logarithmic cost buckets (no host
// synchronisation; the order inside a bucket is arbitrary, results do not depend on it).
__device__ __forceinline__ int cost_bucket(float c) {
  int q = (int)(log2f(c + 1.0f) * 128.0f);
  q = q < 0 ? 0 : (q > 4095 ? 4095 : q);
  return 4095 - q;     // heaviest first
}
__global__ void cost_kernel(int ncell, const double* __restrict__ stats, float* __restrict__ cost, int* __restrict__ hist) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncell) {
    const float v = (float)(10.0 * stats[(size_t)3 * ncell + c] + stats[(size_t)5 * ncell + c] +
                            0.5 * stats[(size_t)1 * ncell + c]);
    cost[c] = v;
    atomicAdd(&hist[cost_bucket(v)], 1);
  }
}
__global__ void __launch_bounds__(1024) cost_scan_kernel(int* __restrict__ hist) {
  __shared__ int part[1024];
  const int t = threadIdx.x;
  int v[4], s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { v[k] = hist[4 * t + k]; s += v[k]; }
  part[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int add = (t >= o) ? part[t - o] : 0;
    __syncthreads();
    part[t] += add;
    __syncthreads();
  }
  int base = part[t] - s;
#pragma unroll
  for (int k = 0; k < 4; ++k) { hist[4 * t + k] = base; base += v[k]; }
}
__global__ void cost_scatter_kernel(int ncell, const float* __restrict__ cost, int* __restrict__ hist, int* __restrict__ order) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncell) order[atomicAdd(&hist[cost_bucket(cost[c])], 1)] = c;
}
cudaError_t launch_cost_order(int ncell, const double* stats, float* cost, int* order, int* hist, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(hist, 0, 4096 * sizeof(int), st);
  if (e != cudaSuccess) return e;
  cost_kernel<<<(ncell + 255) / 256, 256, 0, st>>>(ncell, stats, cost, hist);
  cost_scan_kernel<<<1, 1024, 0, st>>>(hist);
  cost_scatter_kernel<<<(ncell + 255) / 256, 256, 0, st>>>(ncell, cost, hist, order);
  return cudaGetLastError();
}

int integrate_threads() { return NT; }

size_t integrate_smem_bytes(DevNet& net) {
  if (net.glu.on) {
    const size_t n = net.n, R = net.R;
    size_t x_f = R + (size_t)net.rhs.npartial + 32;
    const size_t xj0 = R + (size_t)net.jac[0].npartial + 32, xj1 = R + (size_t)net.jac[1].npartial + 32;
    if (xj0 > x_f) x_f = xj0;
    if (xj1 > x_f) x_f = xj1;
    const size_t xs = (size_t)net.ubE.npartial + (size_t)net.lcE.npartial + 64;
    if (xs > x_f) x_f = xs;
    if (2 * n > x_f) x_f = 2 * n;
    x_f = (x_f + 1) & ~(size_t)1;
    const size_t big = (227 * 1024 - 1024) / 8;
    const size_t doff = (3 * n + 32 + 2 * NW + 1) & ~(size_t)1;   // int4 descriptors (16-byte aligned)
    const size_t voff = doff + 2 * (size_t)net.glu.ndesc;
    const size_t xg = (size_t)(net.o_tl - net.o_ub);
    const size_t sinv = x_f, tab = sinv + (size_t)net.ss.nblkS * 33 * 32;
    const size_t xend = tab + ((size_t)net.ss.blob_words + 1) / 2;
    if (voff + net.nstore + 2 <= big && xend <= xg && net.o_ub == net.n_hh &&
        net.ubE.nblk <= 2 * NW && net.lcE.nblk <= 2 * NW) {
      net.glu.voff = (int)voff; net.glu.doff = (int)doff; net.ss.xlow = (int)x_f; net.ss.sinv = (int)sinv; net.ss.tab = (int)tab;
    } else net.glu.on = 0;
  }
  return make_layout(net).total * sizeof(double);
}

size_t integrate_ws_doubles(const DevNet& net) {
  size_t w = (size_t)net.nstore + net.ubE.nval + net.lcE.nval + net.R + net.n_hh + net.n_ub;
  w = (w + 1) & ~(size_t)1;
  w += (size_t)net.n;
  return (w + 15) & ~(size_t)15;
}

// per-device launch serialisation of the single constant-memory descriptor (see c_net)
struct DevSerial { std::mutex mu; cudaEvent_t done = nullptr; unsigned long long owner = 0; };
static DevSerial g_serial[64];

cudaError_t launch_integrate(const DevNet& net, unsigned long long net_id, int device, const BatchArgs& args,
                             int nblocks, size_t smem, cudaStream_t stream) {
  if (net.nt > 16 * MAXTL || (net.nt & 15) || net.ldt != net.nt + 1 || device < 0 || device >= 64) return cudaErrorInvalidValue;
  DevSerial& S = g_serial[device];
  std::lock_guard<std::mutex> lk(S.mu);
  cudaError_t e;
  if (!S.done) { e = cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming); if (e != cudaSuccess) return e; }
  else { e = cudaStreamWaitEvent(stream, S.done, 0); if (e != cudaSuccess) return e; }
  if (S.owner != net_id) {
    // pageable source: the copy is staged before the call returns, the write is in stream order
    e = cudaMemcpyToSymbolAsync(c_net, &net, sizeof(DevNet), 0, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    S.owner = net_id;
  }
  const Layout L = make_layout(net);
  const bool all = L.glu != 0;
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e2 != cudaSuccess) return e2;
    kern<<<nblocks, NT, smem, stream>>>(args);
    return cudaSuccess;
  };
  // elements per thread of the state vectors: n <= EPT * NT (build_host_net limits N to 1000)
#if RACG_NT == 512
  if (net.n <= NT) e = all ? go(integrate_kernel<1, true>) : go(integrate_kernel<1, false>);
  else if (net.n <= 2 * NT) e = all ? go(integrate_kernel<2, true>) : go(integrate_kernel<2, false>);
#else
  if (net.n <= 2 * NT) e = all ? go(integrate_kernel<2, true>) : go(integrate_kernel<2, false>);
  else if (net.n <= 3 * NT) e = all ? go(integrate_kernel<3, true>) : go(integrate_kernel<3, false>);
  else if (net.n <= 4 * NT) e = go(integrate_kernel<4, false>);
#endif
  else return cudaErrorInvalidValue;
  if (e != cudaSuccess) return e;
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return cudaEventRecord(S.done, stream);
}

}  // namespace racg
