// racg_dev.cuh -- device-side table descriptors shared by the kernels of libracg.
#pragma once
#include <cstdint>
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
};

struct DevNet {
  int R, N, NEQ, n, nh, nt, nslots, nJ, nsat, NNZ;
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
  GatherDev rhs, jac;
  // LU
  const int* row_ptr;
  const int* row_nl;
  const uint16_t* col;
  const int* perm;           // permuted -> original species
  int nflev; const int* flev_ptr; const int* flev_rows;
  int nsu;   const int* su_ptr;   const int* su_rows;
  // species of the sanity test (src/chemistry.f90:520-526), 0-based or -1
  int iH, iE, igH, igH2, igH2O, iGrain0, iGrainM, iGrainP;
  const int* hc_idx;         // [10]
  int ngrain; const int* grain_idx;
  // standalone K2/K3
  const int* csc_to_store;   // [NNZ]
  racg_cfg cfg;
  // BDF coefficients (DCFODE, METH=2): el[q][1..6], tesco[q][1..3], q = 1..5
  double el[6][8];
  double tesco[6][4];
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
  // scheduler + workspace
  int* queue;                // work-queue counter
  double* ws;                // per-CTA workspace
  size_t ws_stride;          // doubles per CTA
  unsigned long long* phase; // [16] cycle counters
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
};

}  // namespace racg
