// racg_integrate.cu -- the persistent per-cell stiff integrator of libracg (sm_100a).
//
// One CTA integrates one grid cell at a time (north_star (c)); a persistent grid
// pulls cells from an atomic work queue (K7) so that cells of very different
// stiffness balance.  All length-N vectors of the BDF controller live in shared
// memory; the Jacobian and the LU factors of the cell live in a per-CTA workspace
// that stays L2-resident.  What it restates, per cell, is the reference's
//   chem_evol_solve loop            src/chemistry.f90:391-588
//   DLSODES driver (MF=21, ITASK=4) src/opkdmain.f:3069-3588
//   DSTODE / DPRJS / DSOLSS         src/opkda1.f:629-1126, 1664-1942
//   DEWSET / DVNORM / DINTDY        src/opkda1.f:1127-1209, 174-281
//   chem_ode_f / chem_ode_jac       src/disk.f90:4569-4659, 4746-4903 (evolT=.false.)
//   chem_cal_rates                  src/chemistry.f90:591-966
// with YSMP's sparse LU replaced by a fixed-pattern LU on the host-computed
// ordering: sparse "head" rows eliminated level by level (one warp per row) and a
// dense "tail" Schur complement factorised in shared memory.
#include <cuda_runtime.h>
#include <cstdio>
#include "racg_dev.cuh"
#include "racg_rates.cuh"

namespace racg {

constexpr int NT = 256;           // threads per CTA
constexpr int NW = NT / 32;       // warps per CTA

enum Phase { PH_RATES = 0, PH_F, PH_JAC, PH_FACT_HEAD, PH_FACT_SCHUR, PH_FACT_TAIL, PH_SOLVE,
             PH_VEC, PH_IO, PH_TOTAL, PH_NCELL, PH_COUNT };

struct Smem {
  double* yh;     // [6][n]
  double* y;      // [n]
  double* savf;   // [n]
  double* acor;   // [n]
  double* ewt;    // [n]
  double* xb;     // [n]  solve vector (permuted space)
  double* dinv;   // [n]
  double* par;    // [32]
  double* red;    // [NW*2]
  double* X;      // overlay region
};

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

// DVNORM over the species (the T slot contributes 0 to every vector we norm except
// YH(:,1); N in the denominator is NEQ as in the reference)
__device__ __forceinline__ double wrms(const double* v, const double* w, int n, int NEQ, double* red) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += NT) { double a = v[i] * w[i]; s += a * a; }
  return sqrt(block_sum(s, red) / (double)NEQ);
}

// ---------------------------------------------------------------------------
// generic segmented-ELL gather: out[target] = sum coef * src[idx]
template <bool GLOBAL_OUT>
__device__ __forceinline__ void run_gather(const GatherDev& g, const double* src, double* out,
                                           double* partial) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int b = w; b < g.nblk; b += NW) {
    const int off = g.blk_off[b], width = g.blk_width[b];
    const uint32_t* e = g.ent + off + l;
    double acc = 0.0;
    for (int j = 0; j < width; ++j) {
      const uint32_t v = __ldg(e + j * 32);
      const int c = (int)(v >> 24) - 4;
      const double s = src[v & 0xffffffu];
      acc += (c != 0) ? (double)c * s : 0.0;
    }
    const int t = g.sub_target[b * 32 + l];
    if (t >= 0) out[t] = acc;
    else if (t <= -2) partial[-2 - t] = acc;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < g.ncombine; q += NT) {
    double s = 0.0;
    for (int p = g.comb_ptr[q]; p < g.comb_ptr[q + 1]; ++p) s += partial[p];
    out[g.comb_row[q]] = s;
  }
  __syncthreads();
}

// flux of reaction r (branches of chem_ode_f, src/disk.f90:4583-4643)
__device__ __forceinline__ double flux_of(const DevNet& net, uint32_t w, double k, const double* y,
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
    const double tmp1 = DS * net.sat_c[w >> 22];
    if (tmp1 <= 0.0) return k;
    const double tmp = y1 / tmp1;
    return (tmp <= 1e-4) ? k * tmp : k * (1.0 - exp(-tmp));
  }
  return 0.0;
}

// d flux / d y(r1) and d flux / d y(r2) (branches of chem_ode_jac, src/disk.f90:4765-4866)
__device__ __forceinline__ void dflux_of(const DevNet& net, uint32_t w, double k, const double* y,
                                         double DS, double& d0, double& d1) {
  const int kind = (w >> 20) & 3;
  const int r1 = w & 1023, r2 = (w >> 10) & 1023;
  d0 = 0.0; d1 = 0.0;
  if (kind == FK_ONE) { d0 = k; return; }
  if (kind == FK_TWO) {
    const double y1 = y[r1], y2 = y[r2];
    const bool flip = (y1 < 0.0 && y2 < 0.0);
    if (r1 != r2) { d0 = k * y2; d1 = k * y1; }
    else d0 = 2.0 * k * y2;
    if (flip) { d0 = -d0; d1 = -d1; }
    return;
  }
  if (kind == FK_SAT) {
    const double tmp2 = DS * net.sat_c[w >> 22];
    if (tmp2 <= 0.0) { d0 = 0.0; return; }
    const double tmp1 = 1.0 / tmp2;
    const double tmp = y[r1] * tmp1;
    d0 = (tmp <= 1e-4) ? k * tmp1 : k * tmp1 * exp(-tmp);
  }
}

struct Ws {          // per-CTA global workspace
  double* J;         // [nJ] Jacobian in storage order (sparse slots, then dense tail row-major)
  double* LU;        // [nslots] sparse LU values
  double* Dt;        // [nt*nt] factored tail, column-major
  double* ksave;     // [R]
  double* rtol;      // [NEQ]
  double* atol;      // [NEQ]
};

// chem_ode_f: out = S * flux(k, yv).  kx = rates in smem, fx = flux scratch, px = partials
__device__ __forceinline__ void eval_f(const DevNet& net, const double* kx, double* fx, double* px,
                                       const double* yv, double* out, double DS) {
  for (int r = threadIdx.x; r < net.R; r += NT) fx[r] = flux_of(net, __ldg(net.fw + r), kx[r], yv, DS);
  for (int i = threadIdx.x; i < net.n; i += NT) out[i] = 0.0;
  __syncthreads();
  run_gather<false>(net.rhs, fx, out, px);
}

// chem_ode_jac for all columns at once -> ws.J
__device__ __forceinline__ void eval_jac(const DevNet& net, const double* kx, double* dfx, double* px,
                                         const double* yv, double* J, double DS) {
  for (int r = threadIdx.x; r < net.R; r += NT) {
    double d0, d1;
    dflux_of(net, __ldg(net.fw + r), kx[r], yv, DS, d0, d1);
    dfx[2 * r] = d0; dfx[2 * r + 1] = d1;
  }
  __syncthreads();
  run_gather<true>(net.jac, dfx, J, px);
}

// P = I - hl0*J, numeric LU.  Returns (to all threads) 0 ok / 1 zero pivot.
__device__ int factor(const DevNet& net, const Ws& ws, Smem& sm, double con, int* flag,
                      unsigned long long* ph) {
  const int n = net.n, nh = net.nh, nt = net.nt;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  double* wrow = sm.X + (size_t)w * n;             // per-warp dense work row
  double* S = sm.X + (size_t)NW * n;               // [nt][nt] row-major Schur complement
  if (threadIdx.x == 0) *flag = 0;
  __syncthreads();
  long long t0 = clock64();
  auto pval = [&](int slot, bool diag) -> double {   // WK = J*CON (+1 on the diagonal), src/opkda1.f:1763-1764
    double v = ws.J[slot] * con;
    if (diag) v = v + 1.0;
    return v;
  };
  // ---- phase 1: head rows, level by level, one warp per row (up-looking)
  for (int lev = 0; lev < net.nflev; ++lev) {
    const int rb = net.flev_ptr[lev], re = net.flev_ptr[lev + 1];
    for (int ri = rb + w; ri < re; ri += NW) {
      const int i = net.flev_rows[ri];
      const int base = net.row_ptr[i], nl = net.row_nl[i], len = net.row_ptr[i + 1] - base;
      for (int q = l; q < len; q += 32) { const int c = net.col[base + q]; wrow[c] = pval(base + q, c == i); }
      __syncwarp();
      for (int q = 0; q < nl; ++q) {
        const int k = net.col[base + q];
        const double lv = wrow[k] * sm.dinv[k];
        if (l == 0) ws.LU[base + q] = lv;
        const int kb = net.row_ptr[k] + net.row_nl[k] + 1, klen = net.row_ptr[k + 1] - kb;
        for (int t = l; t < klen; t += 32) wrow[net.col[kb + t]] -= lv * ws.LU[kb + t];
        __syncwarp();
      }
      const double d = wrow[i];
      if (l == 0) {
        if (d == 0.0 || isnan(d)) *flag = 1;
        sm.dinv[i] = 1.0 / d;
        ws.LU[base + nl] = d;
      }
      for (int q = nl + 1 + l; q < len; q += 32) ws.LU[base + q] = wrow[net.col[base + q]];
      __syncwarp();
    }
    __syncthreads();
  }
  long long t1 = clock64();
  // ---- phase 2: tail rows against the head pivots (independent rows), Schur row into S
  for (int a = w; a < nt; a += NW) {
    const int i = nh + a;
    const int base = net.row_ptr[i], nl = net.row_ptr[i + 1] - base;
    for (int q = l; q < nl; q += 32) wrow[net.col[base + q]] = pval(base + q, false);
    for (int b = l; b < nt; b += 32) wrow[nh + b] = pval(net.nslots + a * nt + b, a == b);
    __syncwarp();
    for (int q = 0; q < nl; ++q) {
      const int k = net.col[base + q];
      const double lv = wrow[k] * sm.dinv[k];
      if (l == 0) ws.LU[base + q] = lv;
      const int kb = net.row_ptr[k] + net.row_nl[k] + 1, klen = net.row_ptr[k + 1] - kb;
      for (int t = l; t < klen; t += 32) wrow[net.col[kb + t]] -= lv * ws.LU[kb + t];
      __syncwarp();
    }
    for (int b = l; b < nt; b += 32) S[a * nt + b] = wrow[nh + b];
    __syncwarp();
  }
  __syncthreads();
  long long t2 = clock64();
  // ---- phase 3: dense right-looking LU of S in shared memory (no pivoting)
  for (int k = 0; k < nt; ++k) {
    const double d = S[k * nt + k];
    if (d == 0.0 || isnan(d)) { if (threadIdx.x == 0) *flag = 1; }
    const double inv = 1.0 / d;
    if (threadIdx.x == 0) sm.dinv[nh + k] = inv;
    __syncthreads();
    for (int i = k + 1 + threadIdx.x; i < nt; i += NT) S[i * nt + k] *= inv;
    __syncthreads();
    const int m = nt - k - 1;
    for (int e = threadIdx.x; e < m * m; e += NT) {
      const int i = k + 1 + e / m, j = k + 1 + e % m;
      S[i * nt + j] -= S[i * nt + k] * S[k * nt + j];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nt * nt; e += NT) {
    const int j = e / nt, i = e % nt;
    ws.Dt[e] = S[i * nt + j];      // column-major copy
  }
  __syncthreads();
  const int res = *flag;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t3 = clock64();
    ph[PH_FACT_HEAD] += t1 - t0; ph[PH_FACT_SCHUR] += t2 - t1; ph[PH_FACT_TAIL] += t3 - t2;
  }
  return res;
}

// DSOLSS: x <- P^{-1} x, x = sm.y in original species order
// (wiped: P is pw*I, see DPRJS below)
__device__ void solve(const DevNet& net, const Ws& ws, Smem& sm, bool wiped, double pw) {
  const int n = net.n, nh = net.nh, nt = net.nt;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (wiped) {
    for (int i = threadIdx.x; i < n; i += NT) sm.y[i] = sm.y[i] / pw;
    __syncthreads();
    return;
  }
  for (int i = threadIdx.x; i < n; i += NT) sm.xb[i] = sm.y[net.perm[i]];
  __syncthreads();
  // forward, head (level 0 rows have no L entries)
  for (int lev = 1; lev < net.nflev; ++lev) {
    const int rb = net.flev_ptr[lev], re = net.flev_ptr[lev + 1];
    for (int ri = rb + threadIdx.x; ri < re; ri += NT) {
      const int i = net.flev_rows[ri];
      const int base = net.row_ptr[i], nl = net.row_nl[i];
      double s = sm.xb[i];
      for (int q = 0; q < nl; ++q) s -= ws.LU[base + q] * sm.xb[net.col[base + q]];
      sm.xb[i] = s;
    }
    __syncthreads();
  }
  // forward, tail rows against head unknowns (independent rows; warp per row)
  for (int a = w; a < nt; a += NW) {
    const int i = nh + a;
    const int base = net.row_ptr[i], nl = net.row_ptr[i + 1] - base;
    double s = 0.0;
    for (int q = l; q < nl; q += 32) s += ws.LU[base + q] * sm.xb[net.col[base + q]];
    s = warp_sum(s);
    if (l == 0) sm.xb[i] -= s;
  }
  __syncthreads();
  // dense tail: unit-lower forward then upper backward, one warp, column oriented
  if (w == 0) {
    double* xt = sm.xb + nh;
    for (int j = 0; j < nt; ++j) {
      const double xj = xt[j];
      const double* cj = ws.Dt + (size_t)j * nt;
      for (int i = j + 1 + l; i < nt; i += 32) xt[i] -= cj[i] * xj;
      __syncwarp();
    }
    for (int j = nt - 1; j >= 0; --j) {
      const double xj = xt[j] * sm.dinv[nh + j];
      __syncwarp();
      if (l == 0) xt[j] = xj;
      const double* cj = ws.Dt + (size_t)j * nt;
      for (int i = l; i < j; i += 32) xt[i] -= cj[i] * xj;
      __syncwarp();
    }
  }
  __syncthreads();
  // backward, head rows
  for (int lev = 0; lev < net.nsu; ++lev) {
    const int rb = net.su_ptr[lev], re = net.su_ptr[lev + 1];
    for (int ri = rb + threadIdx.x; ri < re; ri += NT) {
      const int i = net.su_rows[ri];
      const int ub = net.row_ptr[i] + net.row_nl[i] + 1, ue = net.row_ptr[i + 1];
      double s = sm.xb[i];
      for (int q = ub; q < ue; ++q) s -= ws.LU[q] * sm.xb[net.col[q]];
      sm.xb[i] = s * sm.dinv[i];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += NT) sm.y[net.perm[i]] = sm.xb[i];
  __syncthreads();
}

// ---------------------------------------------------------------------------
// labels of the DSTODE state machine (src/opkda1.f:746-1124)
enum { L100, L150, L160, L170, L175, L200, L220, L250, L270, L410, L430, L450, L500, L520, L540,
       L610, L620, L630, L640, L660, L670, L680, L690, L700, L720 };

struct Lsodes {      // COMMON /DLS001/ + /DLSS01/ (uniform across the CTA, held in registers)
  double CONIT, CRATE, EL[7], HOLD, RMAX, CCMAX, EL0, H, HMIN, HMXI, HU, RC, TN, UROUND;
  double CON0, CONMIN, CCMXJ, PSMALL, RBIG, TCRIT;
  int MXSTEP, NSLAST, IALTH, IPUP, LMAX, NSLP, ICF, IERPJ, IERSL, JCUR, JSTART, KFLAG, L;
  int MAXORD, MAXCOR, MSBP, MXNCF, NQ, NST, NFE, NJE, NQU, MSBJ, NSLJ, NLU, IMXER, IPLOST, INIT;
  int wiped;         // saved P zeroed by ISTATE=1/3 preprocessing and not yet rebuilt from a fresh J
  double pw;         // while wiped: P = pw * I
  bool IHIT;
  long long n_solve, n_cfail, n_efail;
};

__global__ void __launch_bounds__(NT, 1)
integrate_kernel(const DevNet net, const BatchArgs args) {
  extern __shared__ __align__(16) double smem_raw[];
  __shared__ int s_cell, s_flag;
  const int n = net.n, NEQ = net.NEQ, R = net.R;
  const int tid = threadIdx.x;
  Smem sm;
  {
    double* p = smem_raw;
    sm.yh = p; p += 6 * n; sm.y = p; p += n; sm.savf = p; p += n; sm.acor = p; p += n;
    sm.ewt = p; p += n; sm.xb = p; p += n; sm.dinv = p; p += n; sm.par = p; p += 32;
    sm.red = p; p += 2 * NW; sm.X = p;
  }
  double* kx = sm.X;                 // rates [R]
  double* fx = sm.X + R;             // flux [R] / dflux [2R]
  Ws ws;
  {
    double* p = args.ws + (size_t)blockIdx.x * args.ws_stride;
    ws.J = p; p += net.nJ; ws.LU = p; p += net.nslots; ws.Dt = p; p += net.nt * net.nt;
    ws.ksave = p; p += R; ws.rtol = p; p += NEQ; ws.atol = p; p += NEQ;
  }
  unsigned long long ph[PH_COUNT];
  for (int k = 0; k < PH_COUNT; ++k) ph[k] = 0;
  const long long tk0 = clock64();
  const int ncell = args.ncell;
#define YH(i, j) sm.yh[((j) - 1) * n + (i)]
#define VEC(i) for (int i = tid; i < n; i += NT)
  for (;;) {
    __syncthreads();
    if (tid == 0) s_cell = atomicAdd(args.queue, 1);
    __syncthreads();
    const int cell = s_cell;
    if (cell >= ncell) break;
    long long tc = clock64();
    // ---- load the cell
    if (tid < RACG_NPAR) sm.par[tid] = args.cellpar[(size_t)tid * ncell + cell];
    VEC(i) sm.y[i] = args.y0[(size_t)i * ncell + cell];
    const double Tslot = args.y0[(size_t)(NEQ - 1) * ncell + cell];
    __syncthreads();
    const double DS = sm.par[RACG_P_ratioDust2HnucNum] * sm.par[RACG_P_SitesPerGrain];
    // tolerances: given, or chem_set_solver_flags_alt(j) (src/chemistry.f90:205-268)
    if (args.rtol) {
      for (int i = tid; i < NEQ; i += NT) {
        ws.rtol[i] = args.rtol[(size_t)i * ncell + cell];
        ws.atol[i] = args.atol[(size_t)i * ncell + cell];
      }
    } else {
      const int j = args.sp.tol_policy_j;
      const double RT = args.sp.RTOL, AT = args.sp.ATOL, D = sm.par[RACG_P_ratioDust2HnucNum];
      double r, a, rT, aT;
      if (j == 1) { r = RT; a = AT; rT = 1e-3; aT = 1e-1; }
      else if (j == 2) { r = fmin(RT * 1e1, 1e-4); a = fmin(AT * 1e5, 1e-25); rT = 1e-2; aT = 1e-1; }
      else if (j == 3) { r = fmin(RT * 1e2, 1e-4); a = fmin(AT * 1e10, 1e-20); rT = 1e-3; aT = 1.0; }
      else if (j == 4) { r = fmin(RT * 1e2, 1e-4); a = fmin(AT * 1e10, 1e-18); rT = 1e-3; aT = 1.0; }
      else { r = fmin(RT * pow(2.0, (double)j), 1e-3); a = fmin(AT * pow(1e2, (double)j), 1e-15); rT = 1e-2; aT = 1.0; }
      for (int i = tid; i < NEQ; i += NT) { ws.rtol[i] = (i == NEQ - 1) ? rT : r; ws.atol[i] = (i == NEQ - 1) ? aT : a; }
      __syncthreads();
      if (tid < 10 && net.hc_idx[tid] >= 0) { ws.rtol[net.hc_idx[tid]] = fmax(RT, 1e-4); ws.atol[net.hc_idx[tid]] = fmax(AT, 1e-30); }
      __syncthreads();
      if (tid == 0 && net.iGrain0 >= 0) {
        const int g3[3] = {net.iGrain0, net.iGrainM, net.iGrainP};
        for (int q = 0; q < 3; ++q) if (g3[q] >= 0) { ws.rtol[g3[q]] = 1e-4; ws.atol[g3[q]] = fmax(D * 1e-6, 1e-30); }
      }
      __syncthreads();
      for (int q = tid; q < net.ngrain; q += NT) { ws.rtol[net.grain_idx[q]] = fmax(RT, 1e-3); ws.atol[net.grain_idx[q]] = fmax(AT, D * 1e-8); }
    }
    __syncthreads();
    // ---- K1: rate coefficients of this cell
    {
      CellCommon cc;
      cell_common(net.cfg, [&](int k) { return sm.par[k]; }, cc);
      for (int r = tid; r < R; r += NT) kx[r] = rate_coeff(net, cc, r);
      __syncthreads();
      for (int d = tid; d < net.ndup; d += NT) resolve_dupli(net, cc.Tgas, d, [&](int z) { kx[z] = 0.0; });
      __syncthreads();
      for (int r = tid; r < R; r += NT) ws.ksave[r] = kx[r];
    }
    if (tid == 0) { long long t = clock64(); ph[PH_RATES] += t - tc; }

    // ---- chem_evol_solve (src/chemistry.f90:391-588)
    const double t_max = args.tmax[cell], t_start = args.t0[cell];
    const double ratio = args.sp.ratio_tstep;
    int n_record = (int)ceil(log((t_max - t_start) / args.dt_first[cell] * (ratio - 1.0) + 1.0) / log(ratio)) + 1;
    const int n_record_formula = n_record;
    if (n_record > args.sp.nrec_max) n_record = args.sp.nrec_max;
    double t = t_start, t_step = args.dt_first[cell], tout = t + t_step;
    int NERR = 0, nerr_c = 0, quality = 0, ISTATE = 1, n_record_real = 1;
    long long aNST = 0, aNFE = 0, aNJE = 0, aNLU = 0, nrestart = 0;
    Lsodes s;
    s.NST = s.NFE = s.NJE = s.NLU = s.NQU = 0; s.HU = 0.0; s.INIT = 0; s.IMXER = 0;
    s.n_solve = s.n_cfail = s.n_efail = 0; s.wiped = 0; s.pw = 0.0;
    s.NQ = 1; s.L = 2; s.H = 0.0; s.TN = t; s.IHIT = false;
    auto record_out = [&](int irec) {   // touts(i) = t; record(:,i) = y
      if (args.touts && tid == 0) args.touts[(size_t)(irec - 1) * ncell + cell] = t;
      if (args.record) {
        double* rp = args.record + (size_t)(irec - 1) * NEQ * ncell + cell;
        VEC(i) rp[(size_t)i * ncell] = sm.y[i];
        if (tid == 0) rp[(size_t)(NEQ - 1) * ncell] = Tslot;
      }
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
        enum { D_BLOCKC, D200, D245, D250, D270, D_STEP, D_AFTER, D345, D400, D420, D560, D580, D_RET };
        if (ISTATE == 1 && TOUT == t) { dl = D_RET; }
        else if (ISTATE == 2) dl = D200;
        else {
          // Block B
          s.MAXORD = 5; s.MXSTEP = args.sp.mxstep_per_interval; if (s.MXSTEP == 0) s.MXSTEP = 500;
          s.HMXI = (t_max > 0.0) ? 1.0 / t_max : 0.0; s.HMIN = 0.0;
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
              VEC(i) YH(i, 1) = sm.y[i];
              __syncthreads();
              { long long ta = clock64(); eval_f(net, kx, fx, fx + R, sm.y, &YH(0, 2), DS); if (tid == 0) ph[PH_F] += clock64() - ta; }
              s.NFE = 1;
              bool bad = false;
              VEC(i) { const double e = ws.rtol[i] * fabs(YH(i, 1)) + ws.atol[i]; if (e <= 0.0) bad = true; sm.ewt[i] = 1.0 / e; }
              if (__syncthreads_or(bad)) { ISTATE = -3; dl = D_RET; break; }
              s.TCRIT = t_max;
              if ((s.TCRIT - TOUT) * (TOUT - t) < 0.0) { ISTATE = -3; dl = D_RET; break; }
              s.UROUND = 2.220446049250313e-16;
              s.JSTART = 0; s.MSBJ = 50; s.NSLJ = 0; s.CCMXJ = 0.2; s.PSMALL = 1000.0 * s.UROUND;
              s.RBIG = 0.01 / s.PSMALL; s.NJE = 0; s.NLU = 0; s.NSLAST = 0; s.HU = 0.0; s.NQU = 0;
              s.CCMAX = 0.3; s.MAXCOR = 3; s.MSBP = 20; s.MXNCF = 10; s.IPLOST = 0; s.CON0 = 0.0; s.CONMIN = 0.0;
              const double TDIST = fabs(TOUT - t), W0 = fmax(fabs(t), fabs(TOUT));
              if (TDIST < 2.0 * s.UROUND * W0) { ISTATE = -3; dl = D_RET; break; }
              double TOL = 0.0;
              for (int i = tid; i < NEQ; i += NT) TOL = fmax(TOL, ws.rtol[i]);
              {  // block max
                for (int o = 16; o > 0; o >>= 1) TOL = fmax(TOL, __shfl_xor_sync(0xffffffffu, TOL, o));
                __syncthreads();
                if ((tid & 31) == 0) sm.red[tid >> 5] = TOL;
                __syncthreads();
                TOL = sm.red[0];
                for (int k = 1; k < NW; ++k) TOL = fmax(TOL, sm.red[k]);
              }
              if (TOL <= 0.0) {
                double tl = 0.0;
                VEC(i) { const double ay = fabs(sm.y[i]); if (ay != 0.0) tl = fmax(tl, ws.atol[i] / ay); }
                for (int o = 16; o > 0; o >>= 1) tl = fmax(tl, __shfl_xor_sync(0xffffffffu, tl, o));
                __syncthreads();
                if ((tid & 31) == 0) sm.red[tid >> 5] = tl;
                __syncthreads();
                for (int k = 0; k < NW; ++k) TOL = fmax(TOL, sm.red[k]);
                if (Tslot != 0.0) TOL = fmax(TOL, ws.atol[NEQ - 1] / fabs(Tslot));
              }
              TOL = fmax(TOL, 100.0 * s.UROUND);
              TOL = fmin(TOL, 0.001);
              double SUM = wrms(&YH(0, 2), sm.ewt, n, NEQ, sm.red);
              SUM = 1.0 / (TOL * W0 * W0) + TOL * SUM * SUM;
              H0 = 1.0 / sqrt(SUM);
              H0 = fmin(H0, TDIST);
              H0 = copysign(H0, TOUT - t);
              const double RH = fabs(H0) * s.HMXI;
              if (RH > 1.0) H0 = H0 / RH;
              s.H = H0;
              VEC(i) YH(i, 2) = H0 * YH(i, 2);
              __syncthreads();
              dl = D270;
              break;
            }
            case D200: {
              s.NSLAST = s.NST;
              s.TCRIT = t_max;
              if ((s.TN - s.TCRIT) * s.H > 0.0) { ISTATE = -3; dl = D_RET; break; }
              if ((s.TCRIT - TOUT) * s.H < 0.0) { ISTATE = -3; dl = D_RET; break; }
              if ((s.TN - TOUT) * s.H < 0.0) { dl = D245; break; }
              dl = D_AFTER + 100;   // interpolate (handled below)
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
              VEC(i) { const double e = ws.rtol[i] * fabs(YH(i, 1)) + ws.atol[i]; if (e <= 0.0) bad = true; sm.ewt[i] = 1.0 / e; }
              if (__syncthreads_or(bad)) { ISTATE = -6; dl = D580; break; }
              dl = D270;
              break;
            }
            case D270: {
              double sq = 0.0;
              VEC(i) { const double a = YH(i, 1) * sm.ewt[i]; sq += a * a; }
              sq = block_sum(sq, sm.red);
              { const double eT = ws.rtol[NEQ - 1] * fabs(Tslot) + ws.atol[NEQ - 1]; const double a = Tslot / eT; sq += a * a; }
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
              long long tv = clock64();
              int pc;
              double DCON, DDN, DEL = 0.0, DELP = 0.0, DSM = 0.0, DUP, R_, RH = 0.0, RHDN, RHSM, RHUP = 0.0, TOLD;
              int IREDO = 0, IRET = 0, M = 0, NCF = 0, NEWQ = 0;
              s.KFLAG = 0; TOLD = s.TN; s.IERPJ = 0; s.IERSL = 0; s.JCUR = 0; s.ICF = 0;
              if (s.JSTART > 0) pc = L200;
              else if (s.JSTART == -1) pc = L100;
              else if (s.JSTART == -2) pc = L160;
              else {
                s.LMAX = s.MAXORD + 1; s.NQ = 1; s.L = 2; s.IALTH = 2; s.RMAX = 10000.0; s.RC = 0.0;
                s.EL0 = 1.0; s.CRATE = 0.7; s.HOLD = s.H; s.NSLP = 0; s.IPUP = 1; IRET = 3;
                pc = L150;
              }
              for (;;) {
                if (pc == L720) break;
                switch (pc) {
                  case L100:
                    s.IPUP = 1; s.LMAX = s.MAXORD + 1;
                    if (s.IALTH == 1) s.IALTH = 2;
                    pc = L160;
                    break;
                  case L150:
                    for (int i = 1; i <= s.L; ++i) s.EL[i] = net.el[s.NQ][i];
                    s.RC = s.RC * s.EL[1] / s.EL0;
                    s.EL0 = s.EL[1];
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
                    for (int j = 2; j <= s.L; ++j) { Rj = Rj * RH; VEC(i) YH(i, j) = YH(i, j) * Rj; }
                    __syncthreads();
                    s.H = s.H * RH; s.RC = s.RC * RH; s.IALTH = s.L;
                    pc = (IREDO == 0) ? L690 : L200;
                    break;
                  }
                  case L200: {
                    if (fabs(s.RC - 1.0) > s.CCMAX) s.IPUP = 1;
                    if (s.NST >= s.NSLP + s.MSBP) s.IPUP = 1;
                    s.TN = s.TN + s.H;
                    // Pascal-triangle prediction: per element, same operation order as the
                    // flat YH1 sweep of src/opkda1.f:868-874
                    VEC(i) {
                      for (int JB = 1; JB <= s.NQ; ++JB)
                        for (int j = s.NQ + 1 - JB; j <= s.NQ; ++j) YH(i, j) = YH(i, j) + YH(i, j + 1);
                    }
                    __syncthreads();
                    pc = L220;
                    break;
                  }
                  case L220: {
                    M = 0;
                    VEC(i) sm.y[i] = YH(i, 1);
                    __syncthreads();
                    if (tid == 0) ph[PH_VEC] += clock64() - tv;
                    { long long ta = clock64(); eval_f(net, kx, fx, fx + R, sm.y, sm.savf, DS); if (tid == 0) ph[PH_F] += clock64() - ta; }
                    tv = clock64();
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
                        if (tid == 0) ph[PH_VEC] += clock64() - tv;
                        long long ta = clock64();
                        eval_jac(net, kx, fx, fx + 2 * R, sm.y, ws.J, DS);
                        if (tid == 0) ph[PH_JAC] += clock64() - ta;
                        tv = clock64();
                        s.wiped = 0;
                      }
                      s.NLU = s.NLU + 1;
                      int flag;
                      if (s.wiped) {
                        flag = (s.pw == 0.0 || isnan(s.pw)) ? 1 : 0;
                      } else {
                        if (tid == 0) ph[PH_VEC] += clock64() - tv;
                        flag = factor(net, ws, sm, CON, &s_flag, ph);
                        // reload the rates that the factorisation scratch overwrote
                        for (int r = tid; r < R; r += NT) kx[r] = ws.ksave[r];
                        __syncthreads();
                        tv = clock64();
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
                    VEC(i) sm.acor[i] = 0.0;
                  case L270: {
                    VEC(i) sm.y[i] = s.H * sm.savf[i] - (YH(i, 2) + sm.acor[i]);
                    __syncthreads();
                    if (tid == 0) ph[PH_VEC] += clock64() - tv;
                    { long long ta = clock64(); solve(net, ws, sm, s.wiped != 0, s.pw); if (tid == 0) ph[PH_SOLVE] += clock64() - ta; }
                    tv = clock64();
                    s.n_solve++;
                    DEL = wrms(sm.y, sm.ewt, n, NEQ, sm.red);
                    VEC(i) { sm.acor[i] = sm.acor[i] + sm.y[i]; sm.y[i] = YH(i, 1) + s.EL[1] * sm.acor[i]; }
                    __syncthreads();
                    if (M != 0) s.CRATE = fmax(0.2 * s.CRATE, DEL / DELP);
                    DCON = DEL * fmin(1.0, 1.5 * s.CRATE) / (net.tesco[s.NQ][2] * s.CONIT);
                    if (DCON <= 1.0) { pc = L450; break; }
                    M = M + 1;
                    if (M == s.MAXCOR) { pc = L410; break; }
                    if (M >= 2 && DEL > 2.0 * DELP) { pc = L410; break; }
                    DELP = DEL;
                    if (tid == 0) ph[PH_VEC] += clock64() - tv;
                    { long long ta = clock64(); eval_f(net, kx, fx, fx + R, sm.y, sm.savf, DS); if (tid == 0) ph[PH_F] += clock64() - ta; }
                    tv = clock64();
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
                    VEC(i) {
                      for (int JB = 1; JB <= s.NQ; ++JB)
                        for (int j = s.NQ + 1 - JB; j <= s.NQ; ++j) YH(i, j) = YH(i, j) - YH(i, j + 1);
                    }
                    __syncthreads();
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
                    if (M > 0) DSM = wrms(sm.acor, sm.ewt, n, NEQ, sm.red) / net.tesco[s.NQ][2];
                    if (DSM > 1.0) { pc = L500; break; }
                    s.KFLAG = 0; IREDO = 0; s.NST = s.NST + 1; s.HU = s.H; s.NQU = s.NQ;
                    VEC(i) { for (int j = 1; j <= s.L; ++j) YH(i, j) = YH(i, j) + s.EL[j] * sm.acor[i]; }
                    __syncthreads();
                    s.IALTH = s.IALTH - 1;
                    if (s.IALTH == 0) { pc = L520; break; }
                    if (s.IALTH > 1) { pc = L700; break; }
                    if (s.L == s.LMAX) { pc = L700; break; }
                    VEC(i) YH(i, s.LMAX) = sm.acor[i];
                    __syncthreads();
                    pc = L700;
                    break;
                  }
                  case L500: {
                    s.KFLAG = s.KFLAG - 1; s.n_efail++; s.TN = TOLD;
                    VEC(i) {
                      for (int JB = 1; JB <= s.NQ; ++JB)
                        for (int j = s.NQ + 1 - JB; j <= s.NQ; ++j) YH(i, j) = YH(i, j) - YH(i, j + 1);
                    }
                    __syncthreads();
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
                    VEC(i) sm.savf[i] = sm.acor[i] - YH(i, s.LMAX);
                    DUP = wrms(sm.savf, sm.ewt, n, NEQ, sm.red) / net.tesco[s.NQ][3];
                    RHUP = 1.0 / (1.4 * pow(DUP, 1.0 / (s.L + 1)) + 0.0000014);
                  }
                  case L540: {
                    RHSM = 1.0 / (1.2 * pow(DSM, 1.0 / s.L) + 0.0000012);
                    RHDN = 0.0;
                    if (s.NQ != 1) {
                      DDN = wrms(&YH(0, s.L), sm.ewt, n, NEQ, sm.red) / net.tesco[s.NQ][1];
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
                      R_ = s.EL[s.L] / s.L;
                      VEC(i) YH(i, NEWQ + 1) = sm.acor[i] * R_;
                      __syncthreads();
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
                    VEC(i) sm.y[i] = YH(i, 1);
                    __syncthreads();
                    if (tid == 0) ph[PH_VEC] += clock64() - tv;
                    { long long ta = clock64(); eval_f(net, kx, fx, fx + R, sm.y, sm.savf, DS); if (tid == 0) ph[PH_F] += clock64() - ta; }
                    tv = clock64();
                    s.NFE = s.NFE + 1;
                    VEC(i) YH(i, 2) = s.H * sm.savf[i];
                    __syncthreads();
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
                    VEC(i) sm.acor[i] = sm.acor[i] * R_;
                    __syncthreads();
                    pc = L720;
                    break;
                  }
                }
              }
              s.HOLD = s.H; s.JSTART = 1;
              if (tid == 0) ph[PH_VEC] += clock64() - tv;
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
              dl = D_AFTER + 100;
              break;
            }
            case D_AFTER + 100: {   // DINTDY(TOUT, 0, ...) ; T = TOUT ; goto 420
              const double S_ = (TOUT - s.TN) / s.H;
              VEC(i) {
                double v = YH(i, s.L);
                for (int j = s.NQ; j >= 1; --j) v = YH(i, j) + S_ * v;
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
              VEC(i) sm.y[i] = YH(i, 1);
              __syncthreads();
              t = s.TN;
              if (s.IHIT) t = s.TCRIT;
              dl = D420;
              break;
            }
            case D420: ISTATE = 2; dl = D_RET; break;
            case D560: {   // IMXER = first index of max |ACOR*EWT| (T slot contributes 0)
              double big = 0.0; int im = 0x7fffffff;
              VEC(i) { const double sz = fabs(sm.acor[i] * sm.ewt[i]); if (sz > big) { big = sz; im = i; } }
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
              VEC(i) sm.y[i] = YH(i, 1);
              __syncthreads();
              t = s.TN;
              dl = D_RET;
              break;
            }
          }
        }
      }
      // =================== back in chem_evol_solve ===================
      record_out(irec);
      n_record_real = irec;
      if (t >= t_max) break;
      if (ISTATE < 0) {
        NERR += 1; nerr_c += 1;
        if (ISTATE == -4 || ISTATE == -5) {   // ode_solver_error_handling, src/chemistry.f90:326-337
          const int idx = s.IMXER - 1;
          if (tid == 0) {
            ws.rtol[idx] = fmin(ws.rtol[idx] * 10.0, 1e-3);
            ws.atol[idx] = fmin(ws.atol[idx] * 100.0, 1e-20);
          }
          __syncthreads();
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
    for (int r2 = n_record_real + 1; r2 <= args.sp.nrec_max; ++r2) record_out(r2);
    if (NERR > (int)(0.1f * (float)n_record_formula)) quality += 1;
    if (t <= 0.5 * t_max) quality += 2;
    // ---- write results
    VEC(i) args.y_final[(size_t)i * ncell + cell] = sm.y[i];
    if (tid == 0) {
      args.y_final[(size_t)(NEQ - 1) * ncell + cell] = Tslot;
      args.t_final[cell] = t; args.nrec_real[cell] = n_record_real; args.istate[cell] = ISTATE;
      args.quality[cell] = quality;
      double* st = args.stats + cell;
      const double sv[RACG_NSTAT] = {(double)aNST, (double)aNFE, (double)aNJE, (double)aNLU, (double)s.NQU,
                                     (double)s.n_solve, (double)NERR, (double)nrestart, (double)s.n_cfail,
                                     (double)s.n_efail, (double)n_record_real, (double)ISTATE, s.HU, 0, 0, 0};
      for (int k = 0; k < RACG_NSTAT; ++k) st[(size_t)k * ncell] = sv[k];
      ph[PH_NCELL] += 1;
    }
  }
  if (tid == 0 && args.phase) {
    ph[PH_TOTAL] = clock64() - tk0;
    for (int k = 0; k < PH_COUNT; ++k) atomicAdd(&args.phase[k], ph[k]);
  }
#undef YH
#undef VEC
}

size_t integrate_smem_bytes(const DevNet& net, int npart_rhs, int npart_jac) {
  const size_t n = net.n, R = net.R, nt = net.nt;
  size_t V = 12 * n + 32 + 2 * NW;
  size_t X = R + R + (size_t)npart_rhs;
  X = X > R + 2 * R + (size_t)npart_jac ? X : R + 2 * R + (size_t)npart_jac;
  size_t X3 = (size_t)NW * n + nt * nt;
  X = X > X3 ? X : X3;
  return (V + X) * sizeof(double);
}

size_t integrate_ws_doubles(const DevNet& net) {
  size_t w = (size_t)net.nJ + net.nslots + (size_t)net.nt * net.nt + net.R + 2 * (size_t)net.NEQ;
  return (w + 15) & ~(size_t)15;
}

cudaError_t launch_integrate(const DevNet& net, const BatchArgs& args, int nblocks, size_t smem,
                             cudaStream_t stream) {
  static bool attr_set = false;
  cudaError_t e = cudaFuncSetAttribute(integrate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  (void)attr_set;
  integrate_kernel<<<nblocks, NT, smem, stream>>>(net, args);
  return cudaGetLastError();
}

}  // namespace racg
