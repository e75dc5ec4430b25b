// racg_capi.cu -- the extern "C" boundary of libracg.so (declared in include/racg.h).
// Plain pointers and sizes only; no CPU fallback: every compute entry point
// launches sm_100a kernels or fails with RACG_ERR_CUDA.
#include <cuda_runtime.h>
#include <cmath>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>
#include "racg_dev.cuh"
#include "racg_host.hpp"

namespace racg {
// racg_integrate.cu
size_t integrate_smem_bytes(DevNet& net);
size_t integrate_ws_doubles(const DevNet& net);
cudaError_t launch_integrate(const DevNet& net, const BatchArgs& args, int nblocks, size_t smem, cudaStream_t stream);
void launch_cost(int ncell, const double* stats, float* cost, cudaStream_t st);
// racg_batch.cu
cudaError_t launch_rates(const DevNet& net, int ncell, const double* cellpar, double* rates, cudaStream_t st);
cudaError_t launch_rhs(const DevNet& net, int ncell, const double* cellpar, const double* y, const double* rates,
                       double* ydot, int nsm, cudaStream_t st);
cudaError_t launch_jac(const DevNet& net, const JacColTables& jc, int ncell, const double* cellpar, const double* y,
                       const double* rates, double* pd, int nsm, cudaStream_t st);
}  // namespace racg

using namespace racg;

static thread_local std::string g_err;
static int fail(int code, const std::string& s) { g_err = s; return code; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(RACG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

struct racg_handle {
  int device = 0, nsm = 0;
  HostNet hn;
  DevNet dn;
  JacColTables jc;
  std::vector<void*> allocs;
  size_t smem_int = 0, ws_stride = 0;
  int nblocks = 0;
  double* d_ws = nullptr;
  int* d_queue = nullptr;
  unsigned long long* d_phase = nullptr;
  long launches = 0;
  // warm scheduling: per-cell cost of the previous batch and the queue order derived from it
  float* d_cost = nullptr; int* d_order = nullptr; int cost_cap = 0, cost_n = 0;
  std::vector<float> h_cost; std::vector<int> h_order;
  double* dbg_J = nullptr;   // set only inside racg_debug_fjac
  double dbg_con = 0.0;
};

template <typename T>
static int upload(racg_handle* h, const std::vector<T>& v, const T** out) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  CK(cudaMalloc(&p, bytes));
  h->allocs.push_back(p);
  if (!v.empty()) CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = (const T*)p;
  return 0;
}

static int upload_gather(racg_handle* h, const Gather& g, GatherDev& d) {
  d.nblk = g.nblk; d.npartial = g.npartial; d.ncombine = g.ncombine;
  int rc;
  if ((rc = upload(h, g.blk_off, &d.blk_off))) return rc;
  if ((rc = upload(h, g.blk_width, &d.blk_width))) return rc;
  if ((rc = upload(h, g.sub_target, &d.sub_target))) return rc;
  if ((rc = upload(h, g.ent, &d.ent))) return rc;
  if ((rc = upload(h, g.comb_row, &d.comb_row))) return rc;
  if ((rc = upload(h, g.comb_ptr, &d.comb_ptr))) return rc;
  d.sub_add = nullptr; d.comb_add = nullptr;
  if (!g.sub_add.empty()) {
    if ((rc = upload(h, g.sub_add, &d.sub_add))) return rc;
    if ((rc = upload(h, g.comb_add, &d.comb_add))) return rc;
  }
  return 0;
}

static int upload_ell(racg_handle* h, const HostNet::Ell& e, EllDev& d) {
  d.nblk = e.nblk; d.npartial = e.npartial; d.ncombine = e.ncombine; d.nval = e.nval;
  int rc;
  if ((rc = upload(h, e.blk_off, &d.blk_off))) return rc;
  if ((rc = upload(h, e.blk_width, &d.blk_width))) return rc;
  if ((rc = upload(h, e.sub_target, &d.sub_target))) return rc;
  if ((rc = upload(h, e.col, &d.col))) return rc;
  if ((rc = upload(h, e.comb_row, &d.comb_row))) return rc;
  if ((rc = upload(h, e.comb_ptr, &d.comb_ptr))) return rc;
  return 0;
}

// DCFODE, METH = 2 (src/opkda1.f:146-171)
static void bdf_coefficients(DevNet& dn) {
  double pc[13];
  memset(dn.el, 0, sizeof(dn.el)); memset(dn.tesco, 0, sizeof(dn.tesco));
  pc[1] = 1.0;
  double rq1fac = 1.0;
  for (int nq = 1; nq <= 5; ++nq) {
    const double fnq = nq;
    const int nqp1 = nq + 1;
    pc[nqp1] = 0.0;
    for (int ib = 1; ib <= nq; ++ib) { int i = nq + 2 - ib; pc[i] = pc[i - 1] + fnq * pc[i]; }
    pc[1] = fnq * pc[1];
    for (int i = 1; i <= nqp1; ++i) dn.el[nq][i] = pc[i] / pc[2];
    dn.el[nq][2] = 1.0;
    dn.tesco[nq][1] = rq1fac;
    dn.tesco[nq][2] = nqp1 / dn.el[nq][1];
    dn.tesco[nq][3] = (nq + 2) / dn.el[nq][1];
    rq1fac = rq1fac / fnq;
  }
}

// ---- host-pointer entry points: copy in, launch, copy out ----
struct DevBuf {
  std::vector<void*> p;
  ~DevBuf() { for (void* q : p) cudaFree(q); }
  template <typename T> T* get(size_t nelem) { void* q = nullptr; if (cudaMalloc(&q, std::max<size_t>(nelem, 1) * sizeof(T)) != cudaSuccess) return nullptr; p.push_back(q); return (T*)q; }
};

extern "C" {

const char* racg_last_error(void) { return g_err.c_str(); }

void racg_default_cfg(racg_cfg* c) {
  // chemsol_params defaults (src/chemistry.f90:126-131) and phy_const
  // (src/sub_global_variables.f90:3-90; src/chemistry.f90:179-181)
  c->Diff2DesorRatio = 0.5; c->special_gH_E_diff = 225.0;
  c->H2_form_use_moeq = 0; c->use_special_gH_mobi = 0; c->update_gH_params_realtime = 0; c->evol_dust_size = 0;
  c->phy_Pi = 3.1415926535897932384626433; c->phy_elementaryCharge_SI = 1.602176487e-19;
  c->phy_CoulombConst_SI = 8.9875517873681764e9; c->phy_mProton_CGS = 1.67262158e-24;
  c->phy_kBoltzmann_SI = 1.3806503e-23; c->phy_kBoltzmann_CGS = 1.3806503e-16;
  c->phy_hbarPlanck_CGS = 1.054571628e-27; c->phy_SecondsPerYear = 3600.0 * 24.0 * 365.0;
  c->phy_Habing_photon_flux_CGS = 6e7; c->phy_UVext2Av = 2.6;
  c->const_cosmicray_attenuate_N = 5.75e25; c->const_cosmicRay_intensity_0 = 1.36e-17;
  c->CosmicDesorpPreFactor = 3.16e-19; c->CosmicDesorpGrainT = 70.0;
}

int racg_set_device(int device) {
  CK(cudaSetDevice(device));
  return 0;
}

int racg_network_create(racg_handle** out, int R, int N, const int* reac, const int* prod,
                        const int* n_reac, const int* n_prod, const int* itype, const double* ABC,
                        const double* T_range, const char* ctype, const char* names,
                        const int* elements, const double* mass_num, const double* vib_freq,
                        const double* Edesorb, const int* dupli_ptr, const int* dupli_list,
                        const racg_cfg* cfg) {
  if (!out || !reac || !prod || !n_reac || !n_prod || !itype || !ABC || !T_range || !ctype || !names ||
      !elements || !mass_num || !vib_freq || !Edesorb || !dupli_ptr || !dupli_list || !cfg)
    return fail(RACG_ERR_ARG, "null argument");
  racg_handle* h = new racg_handle();
  if (!build_host_net(h->hn, R, N, reac, prod, n_reac, n_prod, itype, ABC, T_range, ctype, names, elements,
                      mass_num, vib_freq, Edesorb, dupli_ptr, dupli_list, cfg)) {
    std::string e = h->hn.error;
    int code = (e.find("not supported") != std::string::npos) ? RACG_ERR_UNSUPPORTED : RACG_ERR_NETWORK;
    delete h;
    return fail(code, e);
  }
  *out = h;   // the host part (pattern, ordering) is usable even without a GPU
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    h->device = -1;
    g_err = "no CUDA device: handle is host-only (pattern/ordering queries work, compute calls fail)";
    return 0;
  }
  CK(cudaGetDevice(&h->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->device));
  h->nsm = prop.multiProcessorCount;
  HostNet& hn = h->hn;
  DevNet& dn = h->dn;
  memset(&dn, 0, sizeof(dn));
  dn.R = hn.R; dn.N = hn.N; dn.NEQ = hn.NEQ; dn.n = hn.n; dn.nh = hn.nh; dn.nt = hn.nt;
  dn.nsat = hn.nsat; dn.NNZ = hn.NNZ;
  dn.n_hh = hn.n_hh; dn.n_ub = hn.n_ub; dn.n_lc = hn.n_lc; dn.o_ub = hn.o_ub; dn.o_lc = hn.o_lc;
  dn.o_tl = hn.o_tl; dn.ldt = hn.ldt; dn.nstore = hn.nstore;
  dn.cfg = hn.cfg;
  bdf_coefficients(dn);
  int rc;
#define UP(vec, field) if ((rc = upload(h, vec, &dn.field))) return rc
  UP(hn.rcode, rcode); UP(hn.rA, rA); UP(hn.rB, rB); UP(hn.rC, rC); UP(hn.rTlo, rTlo); UP(hn.rThi, rThi);
  UP(hn.rX, rX);
  {
    std::vector<int> dreac, dptr(1, 0), dlist;
    for (int i = 0; i < hn.R; ++i) {
      if (hn.dupli_ptr[i + 1] == hn.dupli_ptr[i]) continue;
      dreac.push_back(i);
      for (int q = hn.dupli_ptr[i]; q < hn.dupli_ptr[i + 1]; ++q) dlist.push_back(hn.dupli_list[q] - 1);
      dptr.push_back((int)dlist.size());
    }
    dn.ndup = (int)dreac.size();
    UP(dreac, dup_reac); UP(dptr, dup_ptr); UP(dlist, dup_list);
  }
  UP(hn.fw, fw); UP(hn.sat_c, sat_c);
  if ((rc = upload_gather(h, hn.rhs, dn.rhs))) return rc;
  if ((rc = upload_gather(h, hn.jac[0], dn.jac[0]))) return rc;
  if ((rc = upload_gather(h, hn.jac[1], dn.jac[1]))) return rc;
  UP(hn.hh_ptr, hh_ptr); UP(hn.hh_nl, hh_nl); UP(hn.hh_col, hh_col);
  UP(hn.ub_col, ub_col); UP(hn.ub_ellpos, ub_ellpos);
  UP(hn.lc_ptr, lc_ptr); UP(hn.lc_col, lc_col); UP(hn.lc_ellpos, lc_ellpos);
  if ((rc = upload_ell(h, hn.ubE, dn.ubE))) return rc;
  if ((rc = upload_ell(h, hn.lcE, dn.lcE))) return rc;
  UP(hn.perm, perm); UP(hn.tail_order, tail_order);
  {
    const int* p;
    if ((rc = upload(h, hn.pivmeta, &p))) return rc; dn.pivmeta = (const int4*)p;
    if ((rc = upload(h, hn.fmeta, &p))) return rc; dn.fmeta = (const int4*)p;
    if ((rc = upload(h, hn.bmeta, &p))) return rc; dn.bmeta = (const int4*)p;
  }
  dn.flev_nfat_rows = hn.flev_ptr[hn.nfat_f]; dn.su_nfat_rows = hn.su_ptr[hn.nfat_b];
  dn.nflev = (int)hn.flev_ptr.size() - 1; dn.nfat_f = hn.nfat_f; UP(hn.flev_ptr, flev_ptr); UP(hn.flev_rows, flev_rows);
  dn.nsu = (int)hn.su_ptr.size() - 1; dn.nfat_b = hn.nfat_b; UP(hn.su_ptr, su_ptr); UP(hn.su_rows, su_rows);
  {
    const HostNet::LevelLU& g = hn.glu;
    const HostNet::SolveSched& ss = hn.ss;
    dn.glu.on = 0;
    if (g.nlev > 0 && ss.nent > 0 && !getenv("RACG_NO_GLU")) {
      dn.glu.on = 1; dn.glu.nlev = g.nlev; dn.glu.zpos = g.zpos;
      // bits: 1 = L blocks, 2 = U blocks of the dense tail, 8 = U blocks of the S rows solved by
      // substitution instead of explicit inverses.  Default 0 (all inverses: fastest solves); a cell
      // on which that defeats the corrector is integrated again with bit 2 flipped (see the kernel).
      dn.glu.subst = getenv("RACG_SUBST") ? atoi(getenv("RACG_SUBST")) : 0;
      UP(g.piv, glu.piv); UP(g.mul, glu.mul); UP(g.ent, glu.ent); UP(g.tgt, glu.tgt);
      std::vector<int> desc(g.lvl);
      desc.insert(desc.end(), g.grp.begin(), g.grp.end());
      desc.insert(desc.end(), ss.st.begin(), ss.st.end());
      desc.insert(desc.end(), g.r1.begin(), g.r1.end());
      dn.glu.nst = (int)ss.st.size() / 4;
      UP(g.r1tgt, glu.r1tgt);
      dn.glu.ngrp = (int)g.grp.size() / 4; dn.glu.ndesc = (int)desc.size() / 4;
      { const int* p; if ((rc = upload(h, desc, &p))) return rc; dn.glu.desc = (const int4*)p; }
      dn.ss.nf = ss.nf; dn.ss.nb = ss.nb; dn.ss.nblkS = ss.nblkS; dn.ss.next = (int)ss.ext.size();
      dn.ss.nent = ss.nent; dn.ss.nrp = ss.nrp; dn.ss.nrows = ss.nrows; dn.ss.blob_words = (int)ss.blob.size();
      UP(ss.blob, ss.blob); UP(ss.ext, ss.ext);
    }
  }
  dn.iH = hn.iH; dn.iE = hn.iE; dn.igH = hn.igH; dn.igH2 = hn.igH2; dn.igH2O = hn.igH2O;
  dn.iGrain0 = hn.iGrain0; dn.iGrainM = hn.iGrainM; dn.iGrainP = hn.iGrainP;
  UP(hn.hc_idx, hc_idx);
  dn.ngrain = (int)hn.grain_idx.size(); UP(hn.grain_idx, grain_idx);
  UP(hn.csc_to_store, csc_to_store);
  {
    JacColTables& jc = h->jc;
    const HostNet::JacCols& s = hn.jc;
    jc.ngroups = s.ngroups; jc.max_pairs = s.max_pairs; jc.nzero = (int)s.zero_slots.size();
    if ((rc = upload(h, s.grp_pair_ptr, &jc.grp_pair_ptr))) return rc;
    if ((rc = upload(h, s.pair, &jc.pair))) return rc;
    if ((rc = upload(h, s.grp_slot_ptr, &jc.grp_slot_ptr))) return rc;
    if ((rc = upload(h, s.grp_accum, &jc.grp_accum))) return rc;
    if ((rc = upload(h, s.slot_id, &jc.slot_id))) return rc;
    if ((rc = upload(h, s.slot_ent_ptr, &jc.slot_ent_ptr))) return rc;
    if ((rc = upload(h, s.ent, &jc.ent))) return rc;
    if ((rc = upload(h, s.zero_slots, &jc.zero_slots))) return rc;
  }
#undef UP
  // integrator: one persistent CTA per SM, L2-resident workspace per CTA
  h->smem_int = integrate_smem_bytes(dn);   // also plans the scratch region of the level-parallel mode
  if (h->smem_int > (size_t)prop.sharedMemPerBlockOptin)
    return fail(RACG_ERR_UNSUPPORTED, "network too large for the shared-memory layout of the integrator: " +
                                      std::to_string(h->smem_int) + " B needed");
  h->ws_stride = integrate_ws_doubles(dn);
  h->nblocks = h->nsm;
  CK(cudaMalloc(&h->d_ws, h->ws_stride * sizeof(double) * h->nblocks));
  CK(cudaMemset(h->d_ws, 0, h->ws_stride * sizeof(double) * h->nblocks));
  CK(cudaMalloc(&h->d_queue, sizeof(int)));
  CK(cudaMalloc(&h->d_phase, RACG_NPHASE * sizeof(unsigned long long)));
  CK(cudaMemset(h->d_phase, 0, RACG_NPHASE * sizeof(unsigned long long)));
  return 0;
}

int racg_destroy(racg_handle* h) {
  if (!h) return 0;
  if (h->device >= 0) {
    for (void* p : h->allocs) cudaFree(p);
    cudaFree(h->d_ws); cudaFree(h->d_queue); cudaFree(h->d_phase);
    if (h->d_cost) { cudaFree(h->d_cost); cudaFree(h->d_order); }
  }
  delete h;
  return 0;
}

int racg_network_sizes(const racg_handle* h, int* s) {
  if (!h || !s) return fail(RACG_ERR_ARG, "null argument");
  const HostNet& hn = h->hn;
  s[0] = hn.R; s[1] = hn.N; s[2] = hn.NEQ; s[3] = hn.NNZ; s[4] = hn.NNZ_diag; s[5] = hn.nnz_lu;
  s[6] = hn.nt; s[7] = (int)hn.flev_ptr.size() - 1;
  return 0;
}

int racg_network_pattern(const racg_handle* h, int* ia, int* ja) {
  if (!h || !ia || !ja) return fail(RACG_ERR_ARG, "null argument");
  memcpy(ia, h->hn.ia.data(), sizeof(int) * (h->hn.NEQ + 1));
  memcpy(ja, h->hn.ja.data(), sizeof(int) * h->hn.NNZ);
  return 0;
}

int racg_network_ordering(const racg_handle* h, int* perm) {
  if (!h || !perm) return fail(RACG_ERR_ARG, "null argument");
  for (int i = 0; i < h->hn.n; ++i) perm[i] = h->hn.perm[i] + 1;
  return 0;
}

// chem_set_solver_flags_alt(j), src/chemistry.f90:205-268 -- host arithmetic on host
// arrays (this is configuration, not the hot path; the integrator applies the same
// policy on the device when rtol/atol are NULL).
int racg_solver_flags_alt(const racg_handle* h, int j, double RTOL, double ATOL, int ncell,
                          const double* cellpar, double* rtol, double* atol) {
  if (!h || !cellpar || !rtol || !atol || ncell < 0) return fail(RACG_ERR_ARG, "bad argument");
  const HostNet& hn = h->hn;
  const int NEQ = hn.NEQ, N = hn.N;
  double r, a, rT, aT;
  switch (j) {
    case 1: r = RTOL; a = ATOL; rT = 1e-3; aT = 1e-1; break;
    case 2: r = fmin(RTOL * 1e1, 1e-4); a = fmin(ATOL * 1e5, 1e-25); rT = 1e-2; aT = 1e-1; break;
    case 3: r = fmin(RTOL * 1e2, 1e-4); a = fmin(ATOL * 1e10, 1e-20); rT = 1e-3; aT = 1.0; break;
    case 4: r = fmin(RTOL * 1e2, 1e-4); a = fmin(ATOL * 1e10, 1e-18); rT = 1e-3; aT = 1.0; break;
    default: r = fmin(RTOL * pow(2.0, j), 1e-3); a = fmin(ATOL * pow(1e2, j), 1e-15); rT = 1e-2; aT = 1.0;
  }
  for (int c = 0; c < ncell; ++c) {
    const double D = cellpar[(size_t)RACG_P_ratioDust2HnucNum * ncell + c];
    auto RT = [&](int i) -> double& { return rtol[(size_t)i * ncell + c]; };
    auto AT = [&](int i) -> double& { return atol[(size_t)i * ncell + c]; };
    for (int i = 0; i < NEQ; ++i) { RT(i) = r; AT(i) = a; }
    RT(N) = rT; AT(N) = aT;
    for (int s : hn.hc_idx) if (s >= 0) { RT(s) = fmax(RTOL, 1e-4); AT(s) = fmax(ATOL, 1e-30); }
    if (hn.iGrain0 >= 0)
      for (int s : {hn.iGrain0, hn.iGrainM, hn.iGrainP}) if (s >= 0) { RT(s) = 1e-4; AT(s) = fmax(D * 1e-6, 1e-30); }
    for (int s : hn.grain_idx) { RT(s) = fmax(RTOL, 1e-3); AT(s) = fmax(ATOL, D * 1e-8); }
  }
  return 0;
}

static int need_gpu(const racg_handle* h) {
  if (!h) return fail(RACG_ERR_ARG, "null handle");
  if (h->device < 0) return fail(RACG_ERR_CUDA, "libracg has no CPU fallback: no CUDA device was visible when the handle was created");
  return 0;
}

int racg_rates_dev(racg_handle* h, int ncell, const double* cellpar, double* rates, void* stream) {
  int rc = need_gpu(h); if (rc) return rc;
  if (ncell <= 0) return 0;
  CK(launch_rates(h->dn, ncell, cellpar, rates, (cudaStream_t)stream));
  h->launches += 1;
  return 0;
}

int racg_rhs_jac_dev(racg_handle* h, int ncell, const double* cellpar, const double* y, const double* rates,
                     double* ydot, double* pd, void* stream) {
  int rc = need_gpu(h); if (rc) return rc;
  if (ncell <= 0) return 0;
  if (ydot) { CK(launch_rhs(h->dn, ncell, cellpar, y, rates, ydot, h->nsm, (cudaStream_t)stream)); h->launches += 1; }
  if (pd) { CK(launch_jac(h->dn, h->jc, ncell, cellpar, y, rates, pd, h->nsm, (cudaStream_t)stream)); h->launches += 1; }
  return 0;
}

int racg_solve_batch_dev(racg_handle* h, int ncell, const double* cellpar, const double* y0,
                         const double* rtol, const double* atol, const double* t0, const double* tmax,
                         const double* dt_first, const racg_solve_params* sp, double* y_final,
                         double* t_final, double* touts, double* record, int* nrec_real, int* istate,
                         int* quality, double* stats, void* stream) {
  int rc = need_gpu(h); if (rc) return rc;
  if (!cellpar || !y0 || !t0 || !tmax || !dt_first || !sp || !y_final || !t_final || !nrec_real || !istate ||
      !quality || !stats) return fail(RACG_ERR_ARG, "null argument");
  if ((rtol == nullptr) != (atol == nullptr)) return fail(RACG_ERR_ARG, "rtol and atol must both be given or both NULL");
  if (sp->nrec_max < 2 || sp->ratio_tstep <= 1.0 || sp->steps_reset_solver < 1 || sp->mxstep_per_interval < 0)
    return fail(RACG_ERR_ARG, "bad solve parameters");
  if (ncell <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  BatchArgs a;
  memset(&a, 0, sizeof(a));
  a.ncell = ncell; a.cellpar = cellpar; a.y0 = y0; a.rtol = rtol; a.atol = atol; a.t0 = t0; a.tmax = tmax;
  a.dt_first = dt_first; a.sp = *sp; a.y_final = y_final; a.t_final = t_final; a.touts = touts; a.record = record;
  a.nrec_real = nrec_real; a.istate = istate; a.quality = quality; a.stats = stats;
  a.queue = h->d_queue; a.ws = h->d_ws; a.ws_stride = h->ws_stride; a.phase = h->d_phase;
  a.dbg_J = h->dbg_J; a.dbg_con = h->dbg_con;
  if (h->cost_cap < ncell) {
    if (h->d_cost) { cudaFree(h->d_cost); cudaFree(h->d_order); }
    CK(cudaMalloc(&h->d_cost, sizeof(float) * ncell)); CK(cudaMalloc(&h->d_order, sizeof(int) * ncell));
    h->cost_cap = ncell; h->cost_n = 0;
  }
  if (h->cost_n == ncell && !h->dbg_J && !getenv("RACG_NO_WARM_ORDER")) {
    // same batch size as the previous call: serve the queue heaviest first (order only, results
    // do not depend on it)
    h->h_cost.resize(ncell); h->h_order.resize(ncell);
    CK(cudaMemcpyAsync(h->h_cost.data(), h->d_cost, sizeof(float) * ncell, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int i = 0; i < ncell; ++i) h->h_order[i] = i;
    const float* cst = h->h_cost.data();
    std::stable_sort(h->h_order.begin(), h->h_order.end(), [cst](int x, int y) { return cst[x] > cst[y]; });
    CK(cudaMemcpyAsync(h->d_order, h->h_order.data(), sizeof(int) * ncell, cudaMemcpyHostToDevice, st));
    a.order = h->d_order;
  }
  CK(cudaMemsetAsync(h->d_queue, 0, sizeof(int), st));
  CK(cudaMemsetAsync(h->d_phase, 0, RACG_NPHASE * sizeof(unsigned long long), st));
  int nblocks = ncell < h->nblocks ? ncell : h->nblocks;
  CK(launch_integrate(h->dn, a, nblocks, h->smem_int, st));
  h->launches += 1;
  if (!h->dbg_J) {
    launch_cost(ncell, stats, h->d_cost, st);
    CK(cudaGetLastError());
    h->launches += 1;
    h->cost_n = ncell;
  }
  return 0;
}

#define ALLOC(T, name, nelem) T* name = buf.get<T>(nelem); if (!name) return fail(RACG_ERR_CUDA, "cudaMalloc failed")

int racg_rates(racg_handle* h, int ncell, const double* cellpar, double* rates) {
  int rc = need_gpu(h); if (rc) return rc;
  if (!cellpar || !rates || ncell < 0) return fail(RACG_ERR_ARG, "bad argument");
  if (ncell == 0) return 0;
  DevBuf buf;
  ALLOC(double, d_par, (size_t)RACG_NPAR * ncell);
  ALLOC(double, d_k, (size_t)h->hn.R * ncell);
  CK(cudaMemcpy(d_par, cellpar, sizeof(double) * RACG_NPAR * ncell, cudaMemcpyHostToDevice));
  if ((rc = racg_rates_dev(h, ncell, d_par, d_k, nullptr))) return rc;
  CK(cudaMemcpy(rates, d_k, sizeof(double) * h->hn.R * ncell, cudaMemcpyDeviceToHost));
  return 0;
}

int racg_rhs_jac(racg_handle* h, int ncell, const double* cellpar, const double* y, const double* rates,
                 double* ydot, double* pd) {
  int rc = need_gpu(h); if (rc) return rc;
  if (!cellpar || !y || !rates || ncell < 0) return fail(RACG_ERR_ARG, "bad argument");
  if (ncell == 0) return 0;
  const HostNet& hn = h->hn;
  DevBuf buf;
  ALLOC(double, d_par, (size_t)RACG_NPAR * ncell);
  ALLOC(double, d_y, (size_t)hn.NEQ * ncell);
  ALLOC(double, d_k, (size_t)hn.R * ncell);
  double *d_yd = nullptr, *d_pd = nullptr;
  if (ydot) { d_yd = buf.get<double>((size_t)hn.NEQ * ncell); if (!d_yd) return fail(RACG_ERR_CUDA, "cudaMalloc failed"); }
  if (pd) { d_pd = buf.get<double>((size_t)hn.NNZ * ncell); if (!d_pd) return fail(RACG_ERR_CUDA, "cudaMalloc failed"); }
  CK(cudaMemcpy(d_par, cellpar, sizeof(double) * RACG_NPAR * ncell, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_y, y, sizeof(double) * hn.NEQ * ncell, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_k, rates, sizeof(double) * hn.R * ncell, cudaMemcpyHostToDevice));
  if ((rc = racg_rhs_jac_dev(h, ncell, d_par, d_y, d_k, d_yd, d_pd, nullptr))) return rc;
  if (ydot) CK(cudaMemcpy(ydot, d_yd, sizeof(double) * hn.NEQ * ncell, cudaMemcpyDeviceToHost));
  if (pd) CK(cudaMemcpy(pd, d_pd, sizeof(double) * hn.NNZ * ncell, cudaMemcpyDeviceToHost));
  return 0;
}

int racg_solve_batch(racg_handle* h, int ncell, const double* cellpar, const double* y0, const double* rtol,
                     const double* atol, const double* t0, const double* tmax, const double* dt_first,
                     const racg_solve_params* sp, double* y_final, double* t_final, double* touts,
                     double* record, int* nrec_real, int* istate, int* quality, double* stats) {
  int rc = need_gpu(h); if (rc) return rc;
  if (!cellpar || !y0 || !t0 || !tmax || !dt_first || !sp || !y_final || !t_final || !nrec_real || !istate ||
      !quality || !stats || ncell < 0) return fail(RACG_ERR_ARG, "bad argument");
  if (ncell == 0) return 0;
  const HostNet& hn = h->hn;
  const size_t NEQ = hn.NEQ, nc = ncell, nrec = sp->nrec_max;
  DevBuf buf;
  ALLOC(double, d_par, RACG_NPAR * nc);
  ALLOC(double, d_y0, NEQ * nc);
  ALLOC(double, d_t, 3 * nc);
  ALLOC(double, d_yf, NEQ * nc);
  ALLOC(double, d_tf, nc);
  ALLOC(int, d_i, 3 * nc);
  ALLOC(double, d_st, RACG_NSTAT * nc);
  double *d_rt = nullptr, *d_at = nullptr, *d_touts = nullptr, *d_rec = nullptr;
  if (rtol && atol) {
    d_rt = buf.get<double>(NEQ * nc); d_at = buf.get<double>(NEQ * nc);
    if (!d_rt || !d_at) return fail(RACG_ERR_CUDA, "cudaMalloc failed");
    CK(cudaMemcpy(d_rt, rtol, sizeof(double) * NEQ * nc, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_at, atol, sizeof(double) * NEQ * nc, cudaMemcpyHostToDevice));
  }
  if (touts) { d_touts = buf.get<double>(nrec * nc); if (!d_touts) return fail(RACG_ERR_CUDA, "cudaMalloc failed"); }
  if (record) { d_rec = buf.get<double>(nrec * NEQ * nc); if (!d_rec) return fail(RACG_ERR_CUDA, "cudaMalloc failed (record)"); }
  CK(cudaMemcpy(d_par, cellpar, sizeof(double) * RACG_NPAR * nc, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_y0, y0, sizeof(double) * NEQ * nc, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_t, t0, sizeof(double) * nc, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_t + nc, tmax, sizeof(double) * nc, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_t + 2 * nc, dt_first, sizeof(double) * nc, cudaMemcpyHostToDevice));
  if ((rc = racg_solve_batch_dev(h, ncell, d_par, d_y0, d_rt, d_at, d_t, d_t + nc, d_t + 2 * nc, sp, d_yf, d_tf,
                                 d_touts, d_rec, d_i, d_i + nc, d_i + 2 * nc, d_st, nullptr))) return rc;
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(y_final, d_yf, sizeof(double) * NEQ * nc, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(t_final, d_tf, sizeof(double) * nc, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(nrec_real, d_i, sizeof(int) * nc, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(istate, d_i + nc, sizeof(int) * nc, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(quality, d_i + 2 * nc, sizeof(int) * nc, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(stats, d_st, sizeof(double) * RACG_NSTAT * nc, cudaMemcpyDeviceToHost));
  if (touts) CK(cudaMemcpy(touts, d_touts, sizeof(double) * nrec * nc, cudaMemcpyDeviceToHost));
  if (record) CK(cudaMemcpy(record, d_rec, sizeof(double) * nrec * NEQ * nc, cudaMemcpyDeviceToHost));
  return 0;
}

int racg_debug_fjac(racg_handle* h, int ncell, const double* cellpar, const double* y, double* f,
                    double* jstore, int* csc_to_store, int* nstore, double con) {
  int rc = need_gpu(h); if (rc) return rc;
  const HostNet& hn = h->hn;
  if (nstore) *nstore = hn.nstore;
  if (csc_to_store) memcpy(csc_to_store, hn.csc_to_store.data(), sizeof(int) * hn.NNZ);
  if (!cellpar || !y || !f || !jstore || ncell <= 0) return 0;
  const size_t nc = ncell, NEQ = hn.NEQ;
  DevBuf buf;
  ALLOC(double, d_J, (size_t)hn.nstore * nc);
  std::vector<double> t0(nc, 0.0), tm(nc, 1.0), dt(nc, 1e-8), tf(nc), st((size_t)RACG_NSTAT * nc);
  std::vector<int> ii(3 * nc);
  racg_solve_params sp; sp.ratio_tstep = 1.1; sp.mxstep_per_interval = 10; sp.steps_reset_solver = 50;
  sp.nrec_max = 2; sp.tol_policy_j = 1; sp.RTOL = 1e-4; sp.ATOL = 1e-30;
  h->dbg_J = d_J; h->dbg_con = con;
  rc = racg_solve_batch(h, ncell, cellpar, y, nullptr, nullptr, t0.data(), tm.data(), dt.data(), &sp, f, tf.data(),
                        nullptr, nullptr, ii.data(), ii.data() + nc, ii.data() + 2 * nc, st.data());
  h->dbg_J = nullptr; h->dbg_con = 0.0;
  if (rc) return rc;
  CK(cudaMemcpy(jstore, d_J, sizeof(double) * hn.nstore * nc, cudaMemcpyDeviceToHost));
  return 0;
}

int racg_selfcheck(const racg_handle* h) {
  if (!h) return fail(RACG_ERR_ARG, "null argument");
  std::string err;
  if (!selfcheck_schedules(h->hn, err)) return fail(RACG_ERR_NETWORK, "schedule self-check: " + err);
  return 0;
}

long racg_launch_count(const racg_handle* h) { return h ? h->launches : 0; }

int racg_phase_cycles(racg_handle* h, double* out) {
  int rc = need_gpu(h); if (rc) return rc;
  unsigned long long v[RACG_NPHASE];
  CK(cudaMemcpy(v, h->d_phase, sizeof(v), cudaMemcpyDeviceToHost));
  for (int k = 0; k < RACG_NPHASE; ++k) out[k] = (double)v[k];
  return 0;
}

}  // extern "C"
