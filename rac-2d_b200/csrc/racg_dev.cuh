// racg_dev.cuh -- device-side table descriptors shared by the kernels of libracg.
#pragma once
#include <cstdint>
#include <vector_types.h>
#include "../../include/racg.h"

namespace racg {

struct GatherDev {
  int nblk, npartial, ncombine;
  const int* blk_off;        // [nblk+1]
  const int* blk_width;      // [nblk]
  const int* sub_target;     // [nblk*32]
  const uint32_t* ent;
  const int* comb_row;       // [ncombine]
  const int* comb_ptr;       // [ncombine+1]
  const uint8_t* sub_add;    // optional (null = always store)
  const uint8_t* comb_add;
};

// ELL copy of a coupling block for the solve's SpMV passes (values live in the
// per-CTA workspace at the same positions as col)
struct EllDev {
  int nblk, npartial, ncombine, nval;
  const int* blk_off; const int* blk_width; const int* sub_target;
  const uint16_t* col;
  const int* comb_row; const int* comb_ptr;
};

// level-parallel factorisation schedule (HostNet::LevelLU); the per-level and per-group
// descriptors sit in the constant-memory copy of DevNet so that the level loop never waits
// on a dependent global load
// level-parallel factorisation schedule (HostNet::LevelLU) and staged head solves
// (HostNet::SolveSched).  The int4 descriptors (levels, groups, stages) are staged into
// shared memory once per CTA: the level loop must never wait on a dependent global or
// constant-cache miss.
struct GluDev {
  int on, nlev, zpos, voff;
  int ngrp, ndesc, doff;      // descriptor block: [lvl (nlev+1) | grp (ngrp) | stages (nst) | rank-1 pairs]; doff = offset in smem (doubles)
  const uint32_t* piv; const uint32_t* mul; const uint32_t* ent; const uint16_t* tgt;
  const uint16_t* r1tgt;      // rank-1 levels: 4 target positions per (row chunk, column chunk, lane)
  int nst;                    // number of solve stages (descriptors follow the groups, then the rank-1 pairs)
  const int4* desc;           // lvl: {piv, mul, grp offsets, rank-1 index or -1}; grp: {width, nblk, ent_off, tgt_off}
};
struct SolveDev {
  int nf, nb, nblkS, next;
  int nent, nrp, nrows, blob_words;
  int xlow, sinv, tab;          // offsets (doubles) inside the scratch region X: see integrate_smem_bytes()
  const uint32_t* blob; const uint32_t* ext;
};

struct DevNet {
  int R, N, NEQ, n, nh, nt, nsat, NNZ;
  // rates
  const int* rcode;
  const double *rA, *rB, *rC, *rTlo, *rThi, *rX;
  int ndup;                  // reactions that have earlier twins
  const int* dup_reac;       // [ndup] reaction id (0-based)
  const int* dup_ptr;        // [ndup+1] into dup_list
  const int* dup_list;       // twin ids (0-based)
  // flux
  const uint32_t* fw;
  const double* sat_c;
  GatherDev rhs, jac[2];
  // LU storage layout (see racg_host.hpp)
  int n_hh, n_ub, n_lc, o_ub, o_lc, o_tl, ldt, nstore;
  const int* hh_ptr; const int* hh_nl; const uint16_t* hh_col;
  const uint16_t* ub_col; const int* ub_ellpos;   // [n_ub] tail-local column / position in the ELL copy
  const int* lc_ptr; const uint16_t* lc_col; const int* lc_ellpos;
  EllDev ubE, lcE;
  const int* perm;           // permuted -> original species
  int nflev, nfat_f; const int* flev_ptr; const int* flev_rows;
  int nsu, nfat_b;   const int* su_ptr;   const int* su_rows;
  int flev_nfat_rows, su_nfat_rows;   // rows in the fat levels (= flev_ptr[nfat_f], su_ptr[nfat_b])
  const int* tail_order;     // [nt]
  const int4* pivmeta;       // [nh]
  const int4* fmeta;         // [nh] forward-level order
  const int4* bmeta;         // [nh] backward-level order
  GluDev glu;
  SolveDev ss;
  // species of the sanity test (src/chemistry.f90:520-526), 0-based or -1
  int iH, iE, igH, igH2, igH2O, iGrain0, iGrainM, iGrainP, iH2;
  int h2form_reac;           // the reaction whose coefficient is chem_params%R_H2_form_rate_coeff (src/chemistry.f90:804, 891), or -1
  const int* hc_idx;         // [10]
  int ngrain; const int* grain_idx;
  // standalone K2/K3
  const int* csc_to_store;   // [NNZ]
  racg_cfg cfg;
  // BDF coefficients (DCFODE, METH=2): el[q][1..6], tesco[q][1..3], q = 1..5
  double el[6][8];
  double tesco[6][4];
  // deterministic work model (include/racg.h racg_model_runtime_coefs): c_f, c_jac, c_lu, c_solve, c_step
  double rt_coef[5];
};

struct BatchArgs {
  int ncell;
  const double* cellpar;     // [NPAR][ncell]
  const double* y0;          // [NEQ][ncell]
  const double* rtol;        // [NEQ][ncell] or null
  const double* atol;
  const double* t0; const double* tmax; const double* dt_first;   // [ncell]
  racg_solve_params sp;
  double* y_final;           // [NEQ][ncell]
  double* t_final;
  double* touts;             // [nrec_max][ncell] or null
  double* record;            // [nrec_max][NEQ][ncell] or null
  int* nrec_real; int* istate; int* quality;
  double* stats;             // [NSTAT][ncell]
  // optional harvest outputs (racg_calc_batch): last record with finite T and X(H2)
  double* y_good;            // [NEQ][ncell] or null
  double* t_good; int* isav; // [ncell]
  double* side;              // [2][ncell]: R_H2_form_rate_coeff, n_mol_on_grain; or null
  // scheduler + workspace
  int* queue;                // work-queue counter
  const int* order;          // optional: queue position -> cell (heaviest first), else identity
  double* ws;                // per-CTA workspace
  size_t ws_stride;          // doubles per CTA
  unsigned long long* phase; // [RACG_NPHASE] cycle counters
  // diagnostics (racg_debug_fjac): when set, the kernel evaluates f and J at y0 with its own
  // in-kernel routines, writes f to y_final and J (storage order) to dbg_J [nstore][ncell], and
  // skips the integration
  double* dbg_J;
  double dbg_con;            // != 0: additionally factor P = I + dbg_con*J and return P^-1 f in y_final
};

// streaming K2 schedule (HostNet::RhsChunks)
struct RhsChunkDev {
  int RC, nchunk, spw, max_len4;
  const int* slot_species; const uint32_t* off; const int* nrun; const int* len4; const uint32_t* stream;
  const int* fl_off; const uint32_t* flux;
};

// stand-alone K3 column-group schedule (racg_batch.cu)
struct JacColTables {
  int ngroups;
  const int* grp_pair_ptr;    // [ngroups+1] into pair list
  const uint32_t* pair;       // r | which<<16
  const int* grp_slot_ptr;    // [ngroups+1] into slot list
  const int* grp_accum;       // [ngroups] 1: add to what an earlier chunk of the same column stored
  const int* slot_id;         // CSC slot
  const int* slot_ent_ptr;    // [nslots_listed+1] into ent
  const uint32_t* ent;        // local pair index | (coef+4)<<24
  int max_pairs;
  int nzero; const int* zero_slots;   // CSC slots that are structurally zero for evolT=F (T row/col etc.)
  // repacked copies for jac_kernel_pipe (HostNet::JacCols::pairw / slotw / ent4)
  const uint32_t* pairw; const uint32_t* slotw; const uint32_t* ent4;
  const int* grp_two_ptr; const int* grp_sat_ptr;
};

}  // namespace racg
