// racg_host.hpp -- host-side network setup of libracg (product code, no CUDA).
// Done once per network (north_star (a)): pack the reference's reaction/species
// tables into SoA arrays, build the Jacobian pattern IA/JA bit-exactly as
// chem_make_sparse_structure / chem_prepare_solver_storage do
// (reference src/chemistry.f90:1858-1885, 1962-1971), choose a fill-reducing
// elimination order, compute the symbolic LU every cell shares, and lay out the
// gather schedules the kernels walk.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <array>
#include "../../include/racg.h"

namespace racg {

// rate-coefficient classes (one per branch of chem_cal_rates, src/chemistry.f90:680-934)
enum RateClass : int {
  RC_ZERO = 0, RC_ARRH = 1, RC_ARRH_STRICT = 2, RC_CR = 3, RC_CRPHOT = 4, RC_PHOTO = 5,
  RC_PHOTO_H2 = 6, RC_GRAIN_NP = 7, RC_GRAIN_N0 = 8, RC_LYA = 9, RC_H2FORM0 = 10,
  RC_ADSORB = 11, RC_DESORB = 12, RC_SURF_AA = 13, RC_SURF_AB = 14, RC_PHOTODES = 15
};
// bit layout of rcode: class | fss_kind<<8 | two_body_gas<<12 | sigcheck<<13
// flux kinds (branches of chem_ode_f, src/disk.f90:4583-4643)
enum FluxKind : int { FK_ONE = 0, FK_TWO = 1, FK_SAT = 2, FK_SKIP = 3 };

// "segmented ELL" gather schedule: out[row] = sum_e coef_e * src[idx_e].
// Rows are cut into sub-rows of <= SEG entries; 32 sub-rows form a block stored
// transposed (entry j of lane l at ent[blk_off + j*32 + l]) so a warp reads 128 B.
struct Gather {
  int nrows = 0, nblk = 0, npartial = 0, ncombine = 0;
  std::vector<int> blk_off;        // [nblk+1] offsets into ent (in entries)
  std::vector<int> blk_width;      // [nblk]
  std::vector<int> sub_target;     // [nblk*32] >=0: direct row id; <=-2: partial id = -2-v; -1: idle lane
  std::vector<uint32_t> ent;       // idx | (coef+4)<<24   (idx < 2^24)
  std::vector<int> comb_row;       // [ncombine] rows that own several sub-rows
  std::vector<int> comb_ptr;       // [ncombine+1] into partial ids (consecutive ids)
  long nent_real = 0;
};

struct HostNet {
  int R = 0, N = 0, NEQ = 0, NNZ = 0, NNZ_diag = 0;
  racg_cfg cfg;
  // reference tables (1-based values kept)
  std::vector<int> reac, prod, n_reac, n_prod, itype;
  std::vector<double> ABC, T_range;
  std::vector<std::string> names, ctype;
  std::vector<int> elements;
  std::vector<double> mass_num, vib_freq, Edesorb;
  std::vector<int> dupli_ptr, dupli_list;
  // pattern exactly as IWORK(31:) (1-based)
  std::vector<int> ia, ja;
  // special species (0-based, -1 absent)
  int iH2 = -1, iH = -1, iE = -1, igH = -1, igH2 = -1, igH2O = -1, iGrain0 = -1, iGrainM = -1,
      iGrainP = -1;
  std::vector<int> hc_idx;       // the 10 heating/cooling species of chem_idx_some_spe%idx
  std::vector<int> grain_idx;    // surface species
  // ---- rate tables (SoA over reactions) ----
  std::vector<int> rcode;
  std::vector<double> rA, rB, rC, rTlo, rThi, rX;  // rX: [6*R] class-specific constants
  // ---- flux words ----
  std::vector<uint32_t> fw;       // r1 | r2<<10 | kind<<20 | sat<<22
  std::vector<double> sat_c;      // per saturating reaction: extra factor on D*S
  int nsat = 0;
  // ---- net stoichiometry, RHS gather over species (original order) ----
  Gather rhs;
  // ---- symbolic LU of the species block (T row/col dropped: decoupled for evolT=F) ----
  int n = 0;                      // = N
  int nh = 0, nt = 0;             // head (sparse) / tail (dense) split, nh + nt = n
  std::vector<int> perm, iperm;   // permuted -> original species (0-based), inverse
  // head rows i < nh : CSR  [L_A | diag | U_A U_B], tail rows: CSR of L_C only
  std::vector<int> row_ptr;       // [n+1] into col/val slots ("LU slots")
  std::vector<int> row_nl;        // [n] number of L entries of the row (tail rows: all)
  std::vector<uint16_t> col;      // [nslots] permuted column ids
  int nslots = 0;                 // sparse LU slots (head rows + L_C)
  int nnz_lu = 0;                 // nnz(L+D+U) of the whole pattern (tail counted by its pattern)
  // levels
  std::vector<int> flev_ptr, flev_rows;   // head factorisation levels (rows < nh)
  std::vector<int> sl_ptr, sl_rows;       // forward-solve levels over head rows
  std::vector<int> su_ptr, su_rows;       // backward-solve levels over head rows
  // ---- Jacobian gather into: sparse LU slots [0,nslots) and dense tail nslots + a*nt + b ----
  Gather jac;
  // map from the user's CSC slot (ia/ja) to the J storage index (for racg_rhs_jac parity) or -1
  std::vector<int> csc_to_store;
  // standalone K2/K3 tables: per reaction (r1,r2,p[4] 0-based or -1, kind) and CSC slots
  std::vector<int> rx_species;    // [6*R] r1,r2,p1..p4 (0-based, -1 none)
  std::vector<int> rx_slots;      // [12*R] CSC slot of (participant k, reactant q) or -1
  // ---- stand-alone K3 (column-group schedule, see racg_batch.cu) ----
  struct JacCols {
    int ngroups = 0, max_pairs = 0;
    std::vector<int> grp_pair_ptr, grp_slot_ptr, grp_accum;
    std::vector<uint32_t> pair;        // r | which<<16
    std::vector<int> slot_id, slot_ent_ptr;
    std::vector<uint32_t> ent;         // local pair idx | (coef+4)<<24
    std::vector<int> zero_slots;
  } jc;
  std::string error;
};

bool build_host_net(HostNet& hn, int R, int N, const int* reac, const int* prod, const int* n_reac,
                    const int* n_prod, const int* itype, const double* ABC, const double* T_range,
                    const char* ctype, const char* names, const int* elements,
                    const double* mass_num, const double* vib_freq, const double* Edesorb,
                    const int* dupli_ptr, const int* dupli_list, const racg_cfg* cfg);

}  // namespace racg
