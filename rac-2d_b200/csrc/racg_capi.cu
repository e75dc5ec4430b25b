// racg_capi.cu -- the extern "C" boundary of libracg.so (declared in include/racg.h).
// Plain pointers and sizes only; no CPU fallback: every compute entry point
// launches sm_100a kernels or fails with RACG_ERR_CUDA.
//
// A handle owns one DevCtx per GPU it was replicated to (racg_use_devices): the network
// tables, the integrator's L2 workspace, a stream, pinned staging buffers and a constant-
// memory slot.  The host-pointer racg_solve_batch shards the batch over the handle's
// devices (independent cells, no collective: north_star (e)) and gathers the results into
// the caller's arrays; the device-pointer variants run on the handle's first device.
#include <cuda_runtime.h>
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>
#include "racg_dev.cuh"
#include "racg_host.hpp"

namespace racg {
// racg_integrate.cu
int integrate_threads();
size_t integrate_smem_bytes(DevNet& net);
size_t integrate_ws_doubles(const DevNet& net);
cudaError_t launch_integrate(const DevNet& net, unsigned long long net_id, int device, const BatchArgs& args, int nblocks,
                             size_t smem, cudaStream_t stream);
cudaError_t launch_cost_order(int ncell, const double* stats, float* cost, int* order, int* hist, cudaStream_t st);
cudaError_t launch_cost(int ncell, const double* stats, float* cost, cudaStream_t st);
// racg_batch.cu
extern int g_k3_variant;
cudaError_t launch_rates(const DevNet& net, int ncell, const double* cellpar, double* rates, cudaStream_t st);
cudaError_t launch_rhs(const DevNet& net, const RhsChunkDev& rc, int ncell, const double* cellpar, const double* y,
                       const double* rates, double* ydot, int nsm, cudaStream_t st);
cudaError_t launch_jac(const DevNet& net, const JacColTables& jc, int ncell, const double* cellpar, const double* y,
                       const double* rates, double* pd, int nsm, cudaStream_t st);
}  // namespace racg

using namespace racg;

static thread_local std::string g_err;
static int fail(int code, const std::string& s) { g_err = s; return code; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(RACG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

// identity of a device context's descriptor (never reused): tells launch_integrate whether the
// device's constant memory already holds it
static std::mutex g_id_mu;
static unsigned long long g_next_id = 1;

struct DevCtx {
  int device = -1, nsm = 0;
  unsigned long long net_id = 0;   // changes whenever dn changes
  DevNet dn;
  JacColTables jc;
  RhsChunkDev rhsc;
  std::vector<void*> allocs;
  size_t smem_int = 0, ws_stride = 0;
  int nblocks = 0;
  double* d_ws = nullptr;
  int* d_queue = nullptr;
  unsigned long long* d_phase = nullptr;
  cudaStream_t stream = nullptr;   // used by the host-pointer entry points
  // warm scheduling (device side): per-cell cost of the previous batch and the queue order derived from it
  float* d_cost = nullptr; int* d_order = nullptr; int* d_hist = nullptr; int cost_cap = 0, cost_n = 0;
  // staging of the host-pointer solve: grow-only device arena + pinned host mirror
  char* d_arena = nullptr; size_t d_arena_bytes = 0;
  char* h_arena = nullptr; size_t h_arena_bytes = 0;
  std::vector<int> cells;          // global cell ids of this device's shard, in queue order
};

struct racg_handle {
  HostNet hn;
  std::vector<DevCtx*> dev;        // dev[0]: the device current at racg_network_create
  long launches = 0;
  // options (racg_set_option)
  int warm_order = 1, level_lu = 1;
  // warm scheduling (host side): cost of every cell of the previous host-pointer batch
  std::vector<float> last_cost;
  double* dbg_J = nullptr;   // set only inside racg_debug_fjac
  double dbg_con = 0.0;
};

template <typename T>
static int upload(DevCtx* c, const std::vector<T>& v, const T** out) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  CK(cudaMalloc(&p, bytes));
  c->allocs.push_back(p);
  if (!v.empty()) CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = (const T*)p;
  return 0;
}

static int upload_gather(DevCtx* c, const Gather& g, GatherDev& d) {
  d.nblk = g.nblk; d.npartial = g.npartial; d.ncombine = g.ncombine;
  int rc;
  if ((rc = upload(c, g.blk_off, &d.blk_off))) return rc;
  if ((rc = upload(c, g.blk_width, &d.blk_width))) return rc;
  if ((rc = upload(c, g.sub_target, &d.sub_target))) return rc;
  if ((rc = upload(c, g.ent, &d.ent))) return rc;
  if ((rc = upload(c, g.comb_row, &d.comb_row))) return rc;
  if ((rc = upload(c, g.comb_ptr, &d.comb_ptr))) return rc;
  d.sub_add = nullptr; d.comb_add = nullptr;
  if (!g.sub_add.empty()) {
    if ((rc = upload(c, g.sub_add, &d.sub_add))) return rc;
    if ((rc = upload(c, g.comb_add, &d.comb_add))) return rc;
  }
  return 0;
}

static int upload_ell(DevCtx* c, const HostNet::Ell& e, EllDev& d) {
  d.nblk = e.nblk; d.npartial = e.npartial; d.ncombine = e.ncombine; d.nval = e.nval;
  int rc;
  if ((rc = upload(c, e.blk_off, &d.blk_off))) return rc;
  if ((rc = upload(c, e.blk_width, &d.blk_width))) return rc;
  if ((rc = upload(c, e.sub_target, &d.sub_target))) return rc;
  if ((rc = upload(c, e.col, &d.col))) return rc;
  if ((rc = upload(c, e.comb_row, &d.comb_row))) return rc;
  if ((rc = upload(c, e.comb_ptr, &d.comb_ptr))) return rc;
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

// the deterministic clock that replaces cpu_time in the reference's budgets (include/racg.h)
static void model_runtime_coefs(const HostNet& hn, double* c) {
  c[0] = 1.04e-8 * hn.R;
  c[1] = 6.45e-9 * (double)hn.NEQ * hn.R;
  c[2] = 1.41e-7 * hn.NNZ;
  c[3] = 3.0e-9 * hn.NNZ;
  c[4] = 6.4e-8 * hn.NEQ;
}

static void destroy_ctx(DevCtx* c) {
  if (!c) return;
  if (c->device >= 0) {
    cudaSetDevice(c->device);
    for (void* p : c->allocs) cudaFree(p);
    cudaFree(c->d_ws); cudaFree(c->d_queue); cudaFree(c->d_phase);
    if (c->d_cost) { cudaFree(c->d_cost); cudaFree(c->d_order); cudaFree(c->d_hist); }
    if (c->d_arena) cudaFree(c->d_arena);
    if (c->h_arena) cudaFreeHost(c->h_arena);
    if (c->stream) cudaStreamDestroy(c->stream);
  }
  delete c;
}

// layout decisions that depend on the handle's options, then the constant-memory slot
static int finish_ctx(racg_handle* h, DevCtx* c) {
  const HostNet& hn = h->hn;
  DevNet& dn = c->dn;
  dn.glu.on = (hn.glu.nlev > 0 && hn.ss.nent > 0 && h->level_lu) ? 1 : 0;
  c->smem_int = integrate_smem_bytes(dn);   // also plans the scratch region of the level-parallel mode
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  if (c->smem_int > (size_t)prop.sharedMemPerBlockOptin)
    return fail(RACG_ERR_UNSUPPORTED, "network too large for the shared-memory layout of the integrator: " +
                                      std::to_string(c->smem_int) + " B needed");
  { std::lock_guard<std::mutex> lk(g_id_mu); c->net_id = g_next_id++; }
  return 0;
}

// replicate the network on one GPU: tables, workspace, stream, constant slot
static int create_ctx(racg_handle* h, int device, DevCtx** out) {
  DevCtx* c = new DevCtx();
  *out = c;
  c->device = device;
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  c->nsm = prop.multiProcessorCount;
  if (device >= 64) return fail(RACG_ERR_ARG, "device index above 63");
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  const HostNet& hn = h->hn;
  DevNet& dn = c->dn;
  memset(&dn, 0, sizeof(dn));
  dn.R = hn.R; dn.N = hn.N; dn.NEQ = hn.NEQ; dn.n = hn.n; dn.nh = hn.nh; dn.nt = hn.nt;
  dn.nsat = hn.nsat; dn.NNZ = hn.NNZ;
  dn.n_hh = hn.n_hh; dn.n_ub = hn.n_ub; dn.n_lc = hn.n_lc; dn.o_ub = hn.o_ub; dn.o_lc = hn.o_lc;
  dn.o_tl = hn.o_tl; dn.ldt = hn.ldt; dn.nstore = hn.nstore;
  dn.cfg = hn.cfg;
  bdf_coefficients(dn);
  model_runtime_coefs(hn, dn.rt_coef);
  int rc;
#define UP(vec, field) if ((rc = upload(c, vec, &dn.field))) return rc
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
  if ((rc = upload_gather(c, hn.rhs, dn.rhs))) return rc;
  if ((rc = upload_gather(c, hn.jac[0], dn.jac[0]))) return rc;
  if ((rc = upload_gather(c, hn.jac[1], dn.jac[1]))) return rc;
  UP(hn.hh_ptr, hh_ptr); UP(hn.hh_nl, hh_nl); UP(hn.hh_col, hh_col);
  UP(hn.ub_col, ub_col); UP(hn.ub_ellpos, ub_ellpos);
  UP(hn.lc_ptr, lc_ptr); UP(hn.lc_col, lc_col); UP(hn.lc_ellpos, lc_ellpos);
  if ((rc = upload_ell(c, hn.ubE, dn.ubE))) return rc;
  if ((rc = upload_ell(c, hn.lcE, dn.lcE))) return rc;
  UP(hn.perm, perm); UP(hn.tail_order, tail_order);
  {
    const int* p;
    if ((rc = upload(c, hn.pivmeta, &p))) return rc; dn.pivmeta = (const int4*)p;
    if ((rc = upload(c, hn.fmeta, &p))) return rc; dn.fmeta = (const int4*)p;
    if ((rc = upload(c, hn.bmeta, &p))) return rc; dn.bmeta = (const int4*)p;
  }
  dn.flev_nfat_rows = hn.flev_ptr[hn.nfat_f]; dn.su_nfat_rows = hn.su_ptr[hn.nfat_b];
  dn.nflev = (int)hn.flev_ptr.size() - 1; dn.nfat_f = hn.nfat_f; UP(hn.flev_ptr, flev_ptr); UP(hn.flev_rows, flev_rows);
  dn.nsu = (int)hn.su_ptr.size() - 1; dn.nfat_b = hn.nfat_b; UP(hn.su_ptr, su_ptr); UP(hn.su_rows, su_rows);
  {
    const HostNet::LevelLU& g = hn.glu;
    const HostNet::SolveSched& ss = hn.ss;
    dn.glu.on = 0;
    if (g.nlev > 0 && ss.nent > 0) {
      dn.glu.nlev = g.nlev; dn.glu.zpos = g.zpos;
      UP(g.piv, glu.piv); UP(g.mul, glu.mul); UP(g.ent, glu.ent); UP(g.tgt, glu.tgt);
      std::vector<int> desc(g.lvl);
      desc.insert(desc.end(), g.grp.begin(), g.grp.end());
      desc.insert(desc.end(), ss.st.begin(), ss.st.end());
      desc.insert(desc.end(), g.r1.begin(), g.r1.end());
      dn.glu.nst = (int)ss.st.size() / 4;
      UP(g.r1tgt, glu.r1tgt);
      dn.glu.ngrp = (int)g.grp.size() / 4; dn.glu.ndesc = (int)desc.size() / 4;
      { const int* p; if ((rc = upload(c, desc, &p))) return rc; dn.glu.desc = (const int4*)p; }
      dn.ss.nf = ss.nf; dn.ss.nb = ss.nb; dn.ss.nblkS = ss.nblkS; dn.ss.next = (int)ss.ext.size();
      dn.ss.nent = ss.nent; dn.ss.nrp = ss.nrp; dn.ss.nrows = ss.nrows; dn.ss.blob_words = (int)ss.blob.size();
      UP(ss.blob, ss.blob); UP(ss.ext, ss.ext);
    }
  }
  dn.iH = hn.iH; dn.iE = hn.iE; dn.igH = hn.igH; dn.igH2 = hn.igH2; dn.igH2O = hn.igH2O;
  dn.iGrain0 = hn.iGrain0; dn.iGrainM = hn.iGrainM; dn.iGrainP = hn.iGrainP; dn.iH2 = hn.iH2;
  dn.h2form_reac = hn.h2form_reac;
  UP(hn.hc_idx, hc_idx);
  dn.ngrain = (int)hn.grain_idx.size(); UP(hn.grain_idx, grain_idx);
  UP(hn.csc_to_store, csc_to_store);
  {
    RhsChunkDev& r = c->rhsc;
    const HostNet::RhsChunks& s = hn.rhsc;
    r.RC = s.RC; r.nchunk = s.nchunk; r.spw = s.spw; r.max_len4 = s.max_len4;
    if ((rc = upload(c, s.slot_species, &r.slot_species))) return rc;
    if ((rc = upload(c, s.off, &r.off))) return rc;
    if ((rc = upload(c, s.nrun, &r.nrun))) return rc;
    if ((rc = upload(c, s.len4, &r.len4))) return rc;
    if ((rc = upload(c, s.stream, &r.stream))) return rc;
    if ((rc = upload(c, s.fl_off, &r.fl_off))) return rc;
    if ((rc = upload(c, s.flux, &r.flux))) return rc;
  }
  {
    JacColTables& jc = c->jc;
    const HostNet::JacCols& s = hn.jc;
    jc.ngroups = s.ngroups; jc.max_pairs = s.max_pairs; jc.nzero = (int)s.zero_slots.size();
    if ((rc = upload(c, s.grp_pair_ptr, &jc.grp_pair_ptr))) return rc;
    if ((rc = upload(c, s.pair, &jc.pair))) return rc;
    if ((rc = upload(c, s.grp_slot_ptr, &jc.grp_slot_ptr))) return rc;
    if ((rc = upload(c, s.grp_accum, &jc.grp_accum))) return rc;
    if ((rc = upload(c, s.slot_id, &jc.slot_id))) return rc;
    if ((rc = upload(c, s.slot_ent_ptr, &jc.slot_ent_ptr))) return rc;
    if ((rc = upload(c, s.ent, &jc.ent))) return rc;
    if ((rc = upload(c, s.zero_slots, &jc.zero_slots))) return rc;
    if ((rc = upload(c, s.pairw, &jc.pairw))) return rc;
    if ((rc = upload(c, s.slotw, &jc.slotw))) return rc;
    if ((rc = upload(c, s.ent4, &jc.ent4))) return rc;
    if ((rc = upload(c, s.grp_two_ptr, &jc.grp_two_ptr))) return rc;
    if ((rc = upload(c, s.grp_sat_ptr, &jc.grp_sat_ptr))) return rc;
  }
#undef UP
  if ((rc = finish_ctx(h, c))) return rc;
  // integrator: one persistent CTA per SM, L2-resident workspace per CTA (sized for either mode)
  c->ws_stride = integrate_ws_doubles(dn);
  c->nblocks = c->nsm;
  CK(cudaMalloc(&c->d_ws, c->ws_stride * sizeof(double) * c->nblocks));
  CK(cudaMemset(c->d_ws, 0, c->ws_stride * sizeof(double) * c->nblocks));
  CK(cudaMalloc(&c->d_queue, sizeof(int)));
  CK(cudaMalloc(&c->d_phase, RACG_NPHASE * sizeof(unsigned long long)));
  CK(cudaMemset(c->d_phase, 0, RACG_NPHASE * sizeof(unsigned long long)));
  return 0;
}

static int need_gpu(const racg_handle* h) {
  if (!h) return fail(RACG_ERR_ARG, "null handle");
  if (h->dev.empty()) return fail(RACG_ERR_CUDA, "libracg has no CPU fallback: no CUDA device was visible when the handle was created");
  return 0;
}

// every compute entry point runs on the device(s) of the handle, whatever device the calling
// thread had selected; the caller's selection is restored on return
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int grow(DevCtx* c, size_t dbytes, size_t hbytes) {
  if (dbytes > c->d_arena_bytes) {
    if (c->d_arena) cudaFree(c->d_arena);
    c->d_arena = nullptr; c->d_arena_bytes = 0;
    const size_t want = dbytes + dbytes / 8;
    CK(cudaMalloc((void**)&c->d_arena, want));
    c->d_arena_bytes = want;
  }
  if (hbytes > c->h_arena_bytes) {
    if (c->h_arena) cudaFreeHost(c->h_arena);
    c->h_arena = nullptr; c->h_arena_bytes = 0;
    const size_t want = hbytes + hbytes / 8;
    CK(cudaMallocHost((void**)&c->h_arena, want));
    c->h_arena_bytes = want;
  }
  return 0;
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// cost of a cell in units of one triangular solve (same weights as cost_kernel)
static inline float cell_cost(const double* stats, size_t ncell, size_t c) {
  return (float)(10.0 * stats[3 * ncell + c] + stats[5 * ncell + c] + 0.5 * stats[1 * ncell + c]);
}

static int solve_on_ctx(racg_handle* h, DevCtx* c, const BatchArgs& a0, bool device_order, cudaStream_t st) {
  BatchArgs a = a0;
  const int ncell = a.ncell;
  a.queue = c->d_queue; a.ws = c->d_ws; a.ws_stride = c->ws_stride; a.phase = c->d_phase;
  a.dbg_J = h->dbg_J; a.dbg_con = h->dbg_con;
  if (device_order) {
    if (c->cost_cap < ncell) {
      if (c->d_cost) { cudaFree(c->d_cost); cudaFree(c->d_order); cudaFree(c->d_hist); c->d_cost = nullptr; }
      CK(cudaMalloc(&c->d_cost, sizeof(float) * ncell)); CK(cudaMalloc(&c->d_order, sizeof(int) * ncell));
      CK(cudaMalloc(&c->d_hist, sizeof(int) * 4096));
      c->cost_cap = ncell; c->cost_n = 0;
    }
    // same batch size as the previous call (the disk code re-integrates the same grid every
    // structure iteration): serve the queue heaviest first.  Order only; results do not depend
    // on it.  The order was built on the device right after the previous batch (no host sync).
    if (c->cost_n == ncell && !h->dbg_J && h->warm_order) a.order = c->d_order;
  }
  CK(cudaMemsetAsync(c->d_queue, 0, sizeof(int), st));
  CK(cudaMemsetAsync(c->d_phase, 0, RACG_NPHASE * sizeof(unsigned long long), st));
  const int nblocks = ncell < c->nblocks ? ncell : c->nblocks;
  CK(launch_integrate(c->dn, c->net_id, c->device, a, nblocks, c->smem_int, st));
  h->launches += 1;
  if (device_order && !h->dbg_J) {
    CK(launch_cost_order(ncell, a.stats, c->d_cost, c->d_order, c->d_hist, st));
    h->launches += 3;
    c->cost_n = ncell;
  }
  return 0;
}

// small RAII arena for the K1-K3 host-pointer entry points
struct DevBuf {
  std::vector<void*> p;
  ~DevBuf() { for (void* q : p) cudaFree(q); }
  template <typename T> T* get(size_t nelem) { void* q = nullptr; if (cudaMalloc(&q, std::max<size_t>(nelem, 1) * sizeof(T)) != cudaSuccess) return nullptr; p.push_back(q); return (T*)q; }
};
#define ALLOC(T, name, nelem) T* name = buf.get<T>(nelem); if (!name) return fail(RACG_ERR_CUDA, "cudaMalloc failed")

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
  *out = nullptr;
  racg_handle* h = new racg_handle();
  if (!build_host_net(h->hn, R, N, reac, prod, n_reac, n_prod, itype, ABC, T_range, ctype, names, elements,
                      mass_num, vib_freq, Edesorb, dupli_ptr, dupli_list, cfg, integrate_threads())) {
    std::string e = h->hn.error;
    int code = (e.find("not supported") != std::string::npos) ? RACG_ERR_UNSUPPORTED : RACG_ERR_NETWORK;
    delete h;
    return fail(code, e);
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    *out = h;   // the host part (pattern, ordering, tolerances) is usable without a GPU
    g_err = "no CUDA device: handle is host-only (pattern/ordering queries work, compute calls fail)";
    return 0;
  }
  int device = 0;
  CK(cudaGetDevice(&device));
  DevCtx* c = nullptr;
  int rc = create_ctx(h, device, &c);
  if (rc) {   // no half-initialised handle leaves this function
    destroy_ctx(c);
    delete h;
    return rc;
  }
  h->dev.push_back(c);
  *out = h;
  return 0;
}

int racg_use_devices(racg_handle* h, int ndev, const int* devices) {
  int rc = need_gpu(h); if (rc) return rc;
  DeviceGuard guard;
  std::vector<int> want;
  if (ndev <= 0) {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    for (int d = 0; d < n; ++d) want.push_back(d);
  } else {
    if (!devices) return fail(RACG_ERR_ARG, "null device list");
    want.assign(devices, devices + ndev);
  }
  for (int d : want) {
    bool have = false;
    for (DevCtx* c : h->dev) if (c->device == d) have = true;
    if (have) continue;
    DevCtx* c = nullptr;
    rc = create_ctx(h, d, &c);
    if (rc) { destroy_ctx(c); return rc; }
    h->dev.push_back(c);
  }
  // drop the contexts that are not wanted any more (never the last one)
  for (size_t k = 0; k < h->dev.size();) {
    if (std::find(want.begin(), want.end(), h->dev[k]->device) == want.end() && h->dev.size() > 1) {
      destroy_ctx(h->dev[k]);
      h->dev.erase(h->dev.begin() + k);
    } else ++k;
  }
  h->last_cost.clear();
  return 0;
}

int racg_device_count(const racg_handle* h) { return h ? (int)h->dev.size() : 0; }

int racg_set_option(racg_handle* h, const char* name, double value) {
  if (!h || !name) return fail(RACG_ERR_ARG, "null argument");
  const std::string k = name;
  if (k == "warm_order") h->warm_order = value != 0.0;
  else if (k == "level_lu") h->level_lu = value != 0.0;
  else if (k == "k3_variant") { g_k3_variant = (int)value; return 0; }   // process-wide, diagnostics
  else return fail(RACG_ERR_ARG, "unknown option: " + k);
  if (k != "warm_order" && !h->dev.empty()) {
    DeviceGuard guard;
    for (DevCtx* c : h->dev) { CK(cudaSetDevice(c->device)); CK(cudaDeviceSynchronize()); int rc = finish_ctx(h, c); if (rc) return rc; }
  }
  return 0;
}

int racg_destroy(racg_handle* h) {
  if (!h) return 0;
  DeviceGuard guard;
  for (DevCtx* c : h->dev) destroy_ctx(c);
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

int racg_network_describe(const racg_handle* h, char* buf, int len) {
  if (!h || !buf || len <= 0) return fail(RACG_ERR_ARG, "bad argument");
  const std::string s = describe_host_net(h->hn) +
      (h->dev.empty() ? std::string("devices: none\n")
                      : "devices: " + std::to_string(h->dev.size()) + ", integrator " + std::to_string(integrate_threads()) +
                        " threads per cell, " + (h->dev[0]->dn.glu.on ? "level-parallel LU (factor in shared memory)" : "generic LU (head rows in L2)") +
                        ", " + std::to_string(h->dev[0]->smem_int) + " B shared memory\n");
  snprintf(buf, (size_t)len, "%s", s.c_str());
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

int racg_model_runtime_coefs(const racg_handle* h, double* coef) {
  if (!h || !coef) return fail(RACG_ERR_ARG, "null argument");
  model_runtime_coefs(h->hn, coef);
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

int racg_rates_dev(racg_handle* h, int ncell, const double* cellpar, double* rates, void* stream) {
  int rc = need_gpu(h); if (rc) return rc;
  if (ncell <= 0) return 0;
  DeviceGuard guard;
  DevCtx* c = h->dev[0];
  CK(cudaSetDevice(c->device));
  CK(launch_rates(c->dn, ncell, cellpar, rates, (cudaStream_t)stream));
  h->launches += 1;
  return 0;
}

int racg_rhs_jac_dev(racg_handle* h, int ncell, const double* cellpar, const double* y, const double* rates,
                     double* ydot, double* pd, void* stream) {
  int rc = need_gpu(h); if (rc) return rc;
  if (ncell <= 0) return 0;
  DeviceGuard guard;
  DevCtx* c = h->dev[0];
  CK(cudaSetDevice(c->device));
  if (ydot) { CK(launch_rhs(c->dn, c->rhsc, ncell, cellpar, y, rates, ydot, c->nsm, (cudaStream_t)stream)); h->launches += 1; }
  if (pd) { CK(launch_jac(c->dn, c->jc, ncell, cellpar, y, rates, pd, c->nsm, (cudaStream_t)stream)); h->launches += 1; }
  return 0;
}

static int check_solve_params(const racg_solve_params* sp, bool want_records) {
  if (sp->ratio_tstep <= 1.0 || sp->steps_reset_solver < 1 || sp->mxstep_per_interval < 0 || sp->nrec_max < 0)
    return fail(RACG_ERR_ARG, "bad solve parameters");
  if (want_records && sp->nrec_max < 2) return fail(RACG_ERR_ARG, "touts/record given but nrec_max < 2");
  if (std::isnan(sp->max_runtime_allowed)) return fail(RACG_ERR_ARG, "max_runtime_allowed is NaN");
  return 0;
}

// Asynchronous on `stream` (no host synchronisation): calls on one handle must be issued in
// stream order, the handle's workspace is shared by them.
int racg_solve_batch_dev(racg_handle* h, int ncell, const double* cellpar, const double* y0,
                         const double* rtol, const double* atol, const double* t0, const double* tmax,
                         const double* dt_first, const racg_solve_params* sp, double* y_final,
                         double* t_final, double* touts, double* record, int* nrec_real, int* istate,
                         int* quality, double* stats, void* stream) {
  int rc = need_gpu(h); if (rc) return rc;
  if (!cellpar || !y0 || !t0 || !tmax || !dt_first || !sp || !y_final || !t_final || !nrec_real || !istate ||
      !quality || !stats) return fail(RACG_ERR_ARG, "null argument");
  if ((rtol == nullptr) != (atol == nullptr)) return fail(RACG_ERR_ARG, "rtol and atol must both be given or both NULL");
  if ((rc = check_solve_params(sp, touts || record))) return rc;
  if (ncell <= 0) return 0;
  DeviceGuard guard;
  DevCtx* c = h->dev[0];
  CK(cudaSetDevice(c->device));
  BatchArgs a;
  memset(&a, 0, sizeof(a));
  a.ncell = ncell; a.cellpar = cellpar; a.y0 = y0; a.rtol = rtol; a.atol = atol; a.t0 = t0; a.tmax = tmax;
  a.dt_first = dt_first; a.sp = *sp; a.y_final = y_final; a.t_final = t_final; a.touts = touts; a.record = record;
  a.nrec_real = nrec_real; a.istate = istate; a.quality = quality; a.stats = stats;
  return solve_on_ctx(h, c, a, true, (cudaStream_t)stream);
}


int racg_rates(racg_handle* h, int ncell, const double* cellpar, double* rates) {
  int rc = need_gpu(h); if (rc) return rc;
  if (!cellpar || !rates || ncell < 0) return fail(RACG_ERR_ARG, "bad argument");
  if (ncell == 0) return 0;
  DeviceGuard guard;
  CK(cudaSetDevice(h->dev[0]->device));
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
  DeviceGuard guard;
  CK(cudaSetDevice(h->dev[0]->device));
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

// The host-pointer solve: shards the batch over the handle's devices.  Per device: gather the
// shard's columns into pinned staging (the caller's arrays are [item][cell] with leading
// dimension ncell; a shard is an arbitrary subset of cells), one async H2D, the persistent
// integrator on the device's stream, one async D2H; then the results are scattered back into
// the caller's arrays.  Cells are dealt by descending cost of the previous batch of the same
// size (greedy longest-processing-time), else round-robin.
struct Harvest { double* y_good; double* t_good; int* isav; double* side; };

static int solve_host(racg_handle* h, int ncell, const double* cellpar, const double* y0, const double* rtol,
                      const double* atol, const double* t0, const double* tmax, const double* dt_first,
                      const racg_solve_params* sp, double* y_final, double* t_final, double* touts,
                      double* record, int* nrec_real, int* istate, int* quality, double* stats, const Harvest* hv) {
  int rc = need_gpu(h); if (rc) return rc;
  if (!cellpar || !y0 || !t0 || !tmax || !dt_first || !sp || !y_final || !t_final || !nrec_real || !istate ||
      !quality || !stats || ncell < 0) return fail(RACG_ERR_ARG, "bad argument");
  if ((rtol == nullptr) != (atol == nullptr)) return fail(RACG_ERR_ARG, "rtol and atol must both be given or both NULL");
  if ((rc = check_solve_params(sp, touts || record))) return rc;
  if (ncell == 0) return 0;
  const HostNet& hn = h->hn;
  const size_t NEQ = hn.NEQ, nc = ncell, nrec = sp->nrec_max;
  if (touts || record) {   // capacity check: every cell's own n_record must fit
    for (size_t c = 0; c < nc; ++c) {
      const int need = (int)ceil(log((tmax[c] - t0[c]) / dt_first[c] * (sp->ratio_tstep - 1.0) + 1.0) / log(sp->ratio_tstep)) + 1;
      if (need > sp->nrec_max)
        return fail(RACG_ERR_ARG, "nrec_max = " + std::to_string(sp->nrec_max) + " is smaller than n_record = " +
                                  std::to_string(need) + " of cell " + std::to_string(c + 1));
    }
  }
  DeviceGuard guard;
  const int nd = (int)h->dev.size();
  // ---- deal the cells
  const bool warm = h->warm_order && !h->dbg_J && h->last_cost.size() == nc;
  for (DevCtx* c : h->dev) c->cells.clear();
  if (warm) {
    std::vector<int> ord(nc);
    std::iota(ord.begin(), ord.end(), 0);
    const float* cst = h->last_cost.data();
    std::stable_sort(ord.begin(), ord.end(), [cst](int x, int y) { return cst[x] > cst[y]; });
    std::vector<double> load(nd, 0.0);
    for (int c : ord) {
      int best = 0;
      for (int d = 1; d < nd; ++d) if (load[d] < load[best]) best = d;
      h->dev[best]->cells.push_back(c);      // descending cost inside every shard as well
      load[best] += cst[c];
    }
  } else {
    for (size_t c = 0; c < nc; ++c) h->dev[c % nd]->cells.push_back((int)c);
  }
  // ---- per device: stage, copy in, launch
  struct Lay { size_t par, y0, rt, at, t, in_end, yf, tf, ii, st, yg, tg, sd, touts, rec, out_end; };
  std::vector<Lay> lays(nd);
  for (int d = 0; d < nd; ++d) {
    DevCtx* c = h->dev[d];
    const size_t nl = c->cells.size();
    if (nl == 0) continue;
    Lay& L = lays[d];
    size_t o = 0;
    L.par = o; o = al256(o + 8 * RACG_NPAR * nl);
    L.y0 = o; o = al256(o + 8 * NEQ * nl);
    L.rt = o; if (rtol) o = al256(o + 8 * NEQ * nl);
    L.at = o; if (rtol) o = al256(o + 8 * NEQ * nl);
    L.t = o; o = al256(o + 8 * 3 * nl);
    L.in_end = o;
    L.yf = o; o = al256(o + 8 * NEQ * nl);
    L.tf = o; o = al256(o + 8 * nl);
    L.ii = o; o = al256(o + 4 * 3 * nl);
    L.st = o; o = al256(o + 8 * RACG_NSTAT * nl);
    L.yg = o; if (hv) o = al256(o + 8 * NEQ * nl);
    L.tg = o; if (hv) o = al256(o + 8 * nl + 4 * nl);
    L.sd = o; if (hv) o = al256(o + 8 * 2 * nl);
    L.touts = o; if (touts) o = al256(o + 8 * nrec * nl);
    L.rec = o; if (record) o = al256(o + 8 * nrec * NEQ * nl);
    L.out_end = o;
    CK(cudaSetDevice(c->device));
    // order array (identity: the shard is already in queue order) is not needed; pinned mirror
    // holds inputs and all outputs except `record` (copied straight into a bounce per item row)
    const size_t hbytes = record ? L.rec : L.out_end;
    if ((rc = grow(c, L.out_end, hbytes))) return rc;
    char* hb = c->h_arena;
    auto gather = [&](size_t off, const double* src, size_t nitem) {
      double* dst = (double*)(hb + off);
      for (size_t i = 0; i < nitem; ++i) {
        const double* s = src + i * nc;
        double* q = dst + i * nl;
        for (size_t k = 0; k < nl; ++k) q[k] = s[c->cells[k]];
      }
    };
    gather(L.par, cellpar, RACG_NPAR);
    gather(L.y0, y0, NEQ);
    if (rtol) { gather(L.rt, rtol, NEQ); gather(L.at, atol, NEQ); }
    {
      double* q = (double*)(hb + L.t);
      for (size_t k = 0; k < nl; ++k) { const int g = c->cells[k]; q[k] = t0[g]; q[nl + k] = tmax[g]; q[2 * nl + k] = dt_first[g]; }
    }
    CK(cudaMemcpyAsync(c->d_arena, hb, L.in_end, cudaMemcpyHostToDevice, c->stream));
    BatchArgs a;
    memset(&a, 0, sizeof(a));
    char* db = c->d_arena;
    a.ncell = (int)nl; a.cellpar = (const double*)(db + L.par); a.y0 = (const double*)(db + L.y0);
    a.rtol = rtol ? (const double*)(db + L.rt) : nullptr; a.atol = rtol ? (const double*)(db + L.at) : nullptr;
    a.t0 = (const double*)(db + L.t); a.tmax = a.t0 + nl; a.dt_first = a.t0 + 2 * nl; a.sp = *sp;
    a.y_final = (double*)(db + L.yf); a.t_final = (double*)(db + L.tf);
    a.nrec_real = (int*)(db + L.ii); a.istate = a.nrec_real + nl; a.quality = a.nrec_real + 2 * nl;
    a.stats = (double*)(db + L.st);
    if (hv) {
      a.y_good = (double*)(db + L.yg); a.t_good = (double*)(db + L.tg); a.isav = (int*)(db + L.tg + 8 * nl);
      a.side = (double*)(db + L.sd);
    }
    a.touts = touts ? (double*)(db + L.touts) : nullptr; a.record = record ? (double*)(db + L.rec) : nullptr;
    if ((rc = solve_on_ctx(h, c, a, false, c->stream))) return rc;
    const size_t out_small = (touts ? al256(L.touts + 8 * nrec * nl) : L.touts) - L.yf;
    CK(cudaMemcpyAsync(hb + L.yf, db + L.yf, out_small, cudaMemcpyDeviceToHost, c->stream));
  }
  // ---- per device: wait, scatter back
  if (warm || h->warm_order) h->last_cost.assign(nc, 0.0f);
  for (int d = 0; d < nd; ++d) {
    DevCtx* c = h->dev[d];
    const size_t nl = c->cells.size();
    if (nl == 0) continue;
    const Lay& L = lays[d];
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    char* hb = c->h_arena;
    auto scatter = [&](size_t off, double* dst, size_t nitem) {
      const double* src = (const double*)(hb + off);
      for (size_t i = 0; i < nitem; ++i) {
        double* q = dst + i * nc;
        const double* s = src + i * nl;
        for (size_t k = 0; k < nl; ++k) q[c->cells[k]] = s[k];
      }
    };
    scatter(L.yf, y_final, NEQ);
    scatter(L.tf, t_final, 1);
    scatter(L.st, stats, RACG_NSTAT);
    if (hv) {
      scatter(L.yg, hv->y_good, NEQ);
      scatter(L.tg, hv->t_good, 1);
      scatter(L.sd, hv->side, 2);
      const int* sv = (const int*)(hb + L.tg + 8 * nl);
      for (size_t k = 0; k < nl; ++k) hv->isav[c->cells[k]] = sv[k];
    }
    if (touts) scatter(L.touts, touts, nrec);
    const int* ii = (const int*)(hb + L.ii);
    for (size_t k = 0; k < nl; ++k) {
      const int g = c->cells[k];
      nrec_real[g] = ii[k]; istate[g] = ii[nl + k]; quality[g] = ii[2 * nl + k];
    }
    if (record) {   // [nrec][NEQ][cell]: item rows of this shard -> the caller's rows
      std::vector<double> row(nl);
      const double* drec = (const double*)(c->d_arena + L.rec);
      if (nd == 1 && !warm) {
        CK(cudaMemcpy(record, drec, 8 * nrec * NEQ * nl, cudaMemcpyDeviceToHost));   // identity shard
      } else {
        for (size_t i = 0; i < nrec * NEQ; ++i) {
          CK(cudaMemcpy(row.data(), drec + i * nl, 8 * nl, cudaMemcpyDeviceToHost));
          double* q = record + i * nc;
          for (size_t k = 0; k < nl; ++k) q[c->cells[k]] = row[k];
        }
      }
    }
  }
  if (h->warm_order && !h->dbg_J) for (size_t c = 0; c < nc; ++c) h->last_cost[c] = cell_cost(stats, nc, c);
  return 0;
}

int racg_solve_batch(racg_handle* h, int ncell, const double* cellpar, const double* y0, const double* rtol,
                     const double* atol, const double* t0, const double* tmax, const double* dt_first,
                     const racg_solve_params* sp, double* y_final, double* t_final, double* touts,
                     double* record, int* nrec_real, int* istate, int* quality, double* stats) {
  return solve_host(h, ncell, cellpar, y0, rtol, atol, t0, tmax, dt_first, sp, y_final, t_final, touts, record,
                    nrec_real, istate, quality, stats, nullptr);
}

// The batch form of calc_this_cell's local-iteration loop (src/disk.f90:1651-1791) for
// evolT = .false.: tolerance ladder chem_set_solver_flags_alt(j), continuation from t_final with
// dt_first = max(dt_first0, 1e-3 t0) and rectify_abundances (src/disk.f90:2103-2146,
// src/chemistry.f90:2170-2201), harvest of the last record whose T and X(H2) are not NaN
// (src/disk.f90:1716-1733), exit rules 1703-1712, 1737-1740, 1785-1789.  Cells that need another
// local iteration are compacted into a smaller batch; a cell never returns to the host mid-ladder.
int racg_calc_batch(racg_handle* h, int ncell, const double* cellpar, const double* y0, const double* tmax,
                    double dt_first0, const racg_solve_params* sp, int nlocal_iter, double* abundances,
                    double* t_final, int* quality, int* istate, int* n_iter_used, double* R_H2_form_rate_coeff,
                    double* n_mol_on_grain, double* stats) {
  int rc = need_gpu(h); if (rc) return rc;
  if (!cellpar || !y0 || !tmax || !sp || !abundances || !t_final || !quality || !istate || !n_iter_used ||
      !stats || ncell < 0 || nlocal_iter < 1 || !(dt_first0 > 0.0)) return fail(RACG_ERR_ARG, "bad argument");
  if (ncell == 0) return 0;
  const HostNet& hn = h->hn;
  const size_t NEQ = hn.NEQ, N = hn.N, nc = ncell;
  for (size_t k = 0; k < NEQ * nc; ++k) abundances[k] = y0[k];
  for (size_t c = 0; c < nc; ++c) { t_final[c] = 0.0; quality[c] = 0; istate[c] = 0; n_iter_used[c] = 0; }
  for (size_t k = 0; k < (size_t)RACG_NSTAT * nc; ++k) stats[k] = 0.0;
  if (R_H2_form_rate_coeff) for (size_t c = 0; c < nc; ++c) R_H2_form_rate_coeff[c] = 0.0;
  if (n_mol_on_grain) for (size_t c = 0; c < nc; ++c) n_mol_on_grain[c] = 0.0;
  std::vector<int> act(nc);
  std::iota(act.begin(), act.end(), 0);
  racg_solve_params spj = *sp;
  spj.nrec_max = 0;
  std::vector<double> par, y, t0v, tmv, dtv, yf, tf, st, yg, tg, sd;
  std::vector<int> ii, sv;
  for (int j = 1; j <= nlocal_iter && !act.empty(); ++j) {
    const size_t na = act.size();
    par.resize(RACG_NPAR * na); y.resize(NEQ * na); t0v.resize(na); tmv.resize(na); dtv.resize(na);
    yf.resize(NEQ * na); tf.resize(na); st.resize((size_t)RACG_NSTAT * na); yg.resize(NEQ * na); tg.resize(na);
    sd.resize(2 * na); ii.resize(3 * na); sv.resize(na);
    for (size_t k = 0; k < na; ++k) {
      const size_t c = act[k];
      for (size_t i = 0; i < RACG_NPAR; ++i) par[i * na + k] = cellpar[i * nc + c];
      for (size_t i = 0; i < NEQ; ++i) y[i * na + k] = abundances[i * nc + c];
      if (j > 1 && hn.iE >= 0) {   // rectify_abundances: neutralise with electrons
        double q = 0.0;
        for (size_t i = 0; i < N; ++i) q += abundances[i * nc + c] * (double)hn.elements[(size_t)RACG_NELEM * i];
        y[(size_t)hn.iE * na + k] += q;
      }
      t0v[k] = (j == 1) ? 0.0 : t_final[c];
      dtv[k] = (j == 1) ? dt_first0 : fmax(dt_first0, t0v[k] * 1e-3);
      tmv[k] = tmax[c];
    }
    spj.tol_policy_j = j;
    Harvest hv{yg.data(), tg.data(), sv.data(), sd.data()};
    rc = solve_host(h, (int)na, par.data(), y.data(), nullptr, nullptr, t0v.data(), tmv.data(), dtv.data(), &spj,
                    yf.data(), tf.data(), nullptr, nullptr, ii.data(), ii.data() + na, ii.data() + 2 * na, st.data(), &hv);
    if (rc) return rc;
    std::vector<int> next;
    for (size_t k = 0; k < na; ++k) {
      const size_t c = act[k];
      n_iter_used[c] = j;
      for (int q = 0; q < RACG_NSTAT; ++q) {
        double& dst = stats[(size_t)q * nc + c];
        const double v = st[(size_t)q * na + k];
        if (q == 4 || q == 10 || q == 11 || q == 12 || q == 14) dst = v; else dst += v;   // counters accumulate over the ladder
      }
      istate[c] = ii[na + k];
      // 'Local iteration does not proceed' (src/disk.f90:1703-1712): touts(n_record_real) <= previous t_final
      if (j > 1 && tf[k] <= t_final[c]) continue;
      quality[c] = ii[2 * na + k];
      if (sv[k] <= 1) continue;                       // 'No useful data produced' (1737-1740)
      for (size_t i = 0; i < NEQ; ++i) abundances[i * nc + c] = yg[i * na + k];
      t_final[c] = tg[k];
      if (R_H2_form_rate_coeff) R_H2_form_rate_coeff[c] = sd[k];
      if (n_mol_on_grain) n_mol_on_grain[c] = sd[na + k];
      if (quality[c] == 0 || t_final[c] >= 0.5 * tmax[c]) continue;   // done (1785-1789)
      next.push_back((int)c);
    }
    act.swap(next);
    h->last_cost.clear();   // sub-batches of the ladder have no cost history
  }
  return 0;
}

int racg_debug_fjac(racg_handle* h, int ncell, const double* cellpar, const double* y, double* f,
                    double* jstore, int* csc_to_store, int* nstore, double con) {
  int rc = need_gpu(h); if (rc) return rc;
  const HostNet& hn = h->hn;
  if (nstore) *nstore = hn.nstore;
  if (csc_to_store) memcpy(csc_to_store, hn.csc_to_store.data(), sizeof(int) * hn.NNZ);
  if (!cellpar || !y || !f || !jstore || ncell <= 0) return 0;
  DeviceGuard guard;
  DevCtx* c = h->dev[0];
  CK(cudaSetDevice(c->device));
  const size_t nc = ncell, NEQ = hn.NEQ;
  DevBuf buf;
  ALLOC(double, d_J, (size_t)hn.nstore * nc);
  ALLOC(double, d_par, (size_t)RACG_NPAR * nc);
  ALLOC(double, d_y, NEQ * nc);
  ALLOC(double, d_f, NEQ * nc);
  ALLOC(double, d_t, 4 * nc);
  ALLOC(int, d_i, 3 * nc);
  ALLOC(double, d_st, (size_t)RACG_NSTAT * nc);
  std::vector<double> tt(3 * nc);
  for (size_t k = 0; k < nc; ++k) { tt[k] = 0.0; tt[nc + k] = 1.0; tt[2 * nc + k] = 1e-8; }
  CK(cudaMemcpy(d_par, cellpar, 8 * RACG_NPAR * nc, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_y, y, 8 * NEQ * nc, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_t, tt.data(), 8 * 3 * nc, cudaMemcpyHostToDevice));
  racg_solve_params sp; memset(&sp, 0, sizeof(sp));
  sp.ratio_tstep = 1.1; sp.mxstep_per_interval = 10; sp.steps_reset_solver = 50;
  sp.nrec_max = 0; sp.tol_policy_j = 1; sp.RTOL = 1e-4; sp.ATOL = 1e-30;
  h->dbg_J = d_J; h->dbg_con = con;
  rc = racg_solve_batch_dev(h, ncell, d_par, d_y, nullptr, nullptr, d_t, d_t + nc, d_t + 2 * nc, &sp, d_f, d_t + 3 * nc,
                            nullptr, nullptr, d_i, d_i + nc, d_i + 2 * nc, d_st, nullptr);
  h->dbg_J = nullptr; h->dbg_con = 0.0;
  if (rc) return rc;
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(f, d_f, 8 * NEQ * nc, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(jstore, d_J, sizeof(double) * hn.nstore * nc, cudaMemcpyDeviceToHost));
  return 0;
}

int racg_selfcheck(const racg_handle* h) {
  if (!h) return fail(RACG_ERR_ARG, "null argument");
  std::string err;
  if (!selfcheck_schedules(h->hn, err)) return fail(RACG_ERR_NETWORK, "schedule self-check: " + err);
  return 0;
}

int racg_selfcheck_damaged(const racg_handle* h, int mode) {
  if (!h) return fail(RACG_ERR_ARG, "null argument");
  HostNet copy = h->hn;
  HostNet::LevelLU& g = copy.glu;
  if (g.ent.empty()) return fail(RACG_ERR_ARG, "no level-parallel schedule to damage");
  if (mode == 1) g.ent[g.ent.size() / 2] ^= 1u;                       // wrong operand position
  else if (mode == 2 && g.r1tgt.size() > 8) g.r1tgt[5] = g.r1tgt[4];  // target written twice
  else return fail(RACG_ERR_ARG, "unknown damage mode");
  std::string err;
  if (!selfcheck_schedules(copy, err)) return fail(RACG_ERR_NETWORK, "schedule self-check: " + err);
  return 0;
}

// chemical_data_iter_NNNN.bin (back_cells_chemical_data, src/data_dump.f90:88-162): a direct-access
// unformatted file, record i (fixed length, no record markers) = abundances(1:nSpecies),
// col_den_toStar(1:ncd), col_den_toISM(1:ncd) of leaf i as native doubles; the file name follows
// the reference ('chemical_data_iter_', I0.4, '.bin'; iiter < 0: 'chemical_data.bin').
static std::string chemical_data_name(const char* dir, int iiter) {
  char nm[64];
  if (iiter >= 0) snprintf(nm, sizeof nm, "chemical_data_iter_%04d.bin", iiter); else snprintf(nm, sizeof nm, "chemical_data.bin");
  std::string d = dir ? dir : ".";
  if (!d.empty() && d.back() != '/') d += '/';
  return d + nm;
}

int racg_write_chemical_data(const char* dir, int iiter, int ncell, int nspecies, const double* abundances,
                             int ncd, const double* col_den_toStar, const double* col_den_toISM) {
  if (!abundances || ncell < 0 || nspecies <= 0 || ncd < 0 || (ncd > 0 && (!col_den_toStar || !col_den_toISM)))
    return fail(RACG_ERR_ARG, "bad argument");
  const std::string fn = chemical_data_name(dir, iiter);
  FILE* f = fopen(fn.c_str(), "wb");
  if (!f) return fail(RACG_ERR_ARG, "cannot open " + fn);
  std::vector<double> rec((size_t)nspecies + 2 * (size_t)ncd);
  for (int c = 0; c < ncell; ++c) {
    for (int i = 0; i < nspecies; ++i) rec[i] = abundances[(size_t)i * ncell + c];      // a(ncell, item)
    for (int k = 0; k < ncd; ++k) { rec[nspecies + k] = col_den_toStar[(size_t)c * ncd + k]; rec[nspecies + ncd + k] = col_den_toISM[(size_t)c * ncd + k]; }
    if (fwrite(rec.data(), sizeof(double), rec.size(), f) != rec.size()) { fclose(f); return fail(RACG_ERR_ARG, "write failed: " + fn); }
  }
  fclose(f);
  return 0;
}

int racg_read_chemical_data(const char* dir, int iiter, int ncell, int nspecies, double* abundances, int ncd,
                            double* col_den_toStar, double* col_den_toISM) {
  if (!abundances || ncell < 0 || nspecies <= 0 || ncd < 0 || (ncd > 0 && (!col_den_toStar || !col_den_toISM)))
    return fail(RACG_ERR_ARG, "bad argument");
  const std::string fn = chemical_data_name(dir, iiter);
  FILE* f = fopen(fn.c_str(), "rb");
  if (!f) return fail(RACG_ERR_ARG, "cannot open " + fn);
  std::vector<double> rec((size_t)nspecies + 2 * (size_t)ncd);
  for (int c = 0; c < ncell; ++c) {
    if (fread(rec.data(), sizeof(double), rec.size(), f) != rec.size()) { fclose(f); return fail(RACG_ERR_ARG, "short file: " + fn); }
    for (int i = 0; i < nspecies; ++i) abundances[(size_t)i * ncell + c] = rec[i];
    for (int k = 0; k < ncd; ++k) { col_den_toStar[(size_t)c * ncd + k] = rec[nspecies + k]; col_den_toISM[(size_t)c * ncd + k] = rec[nspecies + ncd + k]; }
  }
  fclose(f);
  return 0;
}

long racg_launch_count(const racg_handle* h) { return h ? h->launches : 0; }

int racg_phase_cycles(racg_handle* h, double* out) {
  int rc = need_gpu(h); if (rc) return rc;
  DeviceGuard guard;
  for (int k = 0; k < RACG_NPHASE; ++k) out[k] = 0.0;
  for (DevCtx* c : h->dev) {
    unsigned long long v[RACG_NPHASE];
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy(v, c->d_phase, sizeof(v), cudaMemcpyDeviceToHost));
    for (int k = 0; k < RACG_NPHASE; ++k) out[k] += (double)v[k];
  }
  return 0;
}

}  // extern "C"
