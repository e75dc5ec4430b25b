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
  std::vector<uint8_t> sub_add, comb_add;  // optional: 1 = add to the target instead of storing
  long nent_real = 0;
};

struct HostNet {
  int R = 0, N = 0, NEQ = 0, NNZ = 0, NNZ_diag = 0;
  int nthreads = 256;             // threads per cell of the integrator the solve stages are laid out for
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
  int h2form_reac = -1;          // last reaction that sets chem_params%R_H2_form_rate_coeff (0-based)
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
  int nnz_lu = 0;                 // nnz(L+D+U) of the symbolic factorisation
  // value-storage index space shared by J and LU:
  //   [0,n_hh)        head x head block, CSR by head row: [L_A | diag | U_A]
  //   [o_ub,+n_ub)    U_B: head rows x tail columns, CSR
  //   [o_lc,+n_lc)    L_C: tail rows x head columns, CSR
  //   [o_tl,+ldt*nt)  dense tail block, column-major, leading dimension ldt
  int n_hh = 0, n_ub = 0, n_lc = 0, o_ub = 0, o_lc = 0, o_tl = 0, ldt = 0, nstore = 0;
  std::vector<int> hh_ptr, hh_nl; std::vector<uint16_t> hh_col;
  std::vector<int> ub_ptr; std::vector<uint16_t> ub_col;   // tail-local column
  std::vector<int> lc_ptr; std::vector<uint16_t> lc_col;   // head column
  // ELL (32 sub-rows per block, transposed) copies of U_B / L_C for the solve passes
  struct Ell {
    int nblk = 0, npartial = 0, ncombine = 0, nval = 0;
    std::vector<int> blk_off, blk_width, sub_target, comb_row, comb_ptr;
    std::vector<uint16_t> col;
  } ubE, lcE;
  std::vector<int> ub_ellpos, lc_ellpos;     // CSR slot -> ELL position
  // levels over head rows
  std::vector<int> flev_ptr, flev_rows, head_level;   // forward / factorisation (L_A)
  std::vector<int> su_ptr, su_rows;                   // backward (U_A)
  int nfat_f = 0, nfat_b = 0;     // levels >= nfat have <= 32 rows (handled by one warp)
  std::vector<int> tail_order;    // tail rows, longest L_C row first
  // packed metadata (int4 each) so that the sparse phases never chase pointers:
  //   pivmeta[k]  = {start of U_A(k,:) in hh, its length, start of U_B(k,:), its length}
  //   fmeta[pos]  = {row, start of L_A(row,:) in hh, its length, 0}   in forward-level order
  //   bmeta[pos]  = {row, start of U_A(row,:) in hh, its length, 0}   in backward-level order
  std::vector<int> pivmeta, fmeta, bmeta;
  // ---- level-parallel numeric factorisation ("gather LU", racg_integrate.cu factor_glu):
  // head pivots are grouped into levels whose members neither update nor read one another;
  // for one level every multiplier l(i,k) = a(i,k)/a(k,k) and then every updated entry
  // a(i,j) -= sum_k l(i,k) u(k,j) is independent, so the schedule lists per target the
  // (position of l, position of u) pairs, packed as 32-lane ELL blocks.
  struct LevelLU {
    int nlev = 0, zpos = 0;
    long npairs = 0;
    std::vector<int> lvl;                               // int4 per level (+1 sentinel): {piv, mul, grp offsets, 0}
    std::vector<int> grp;                               // int4 per group: {width, nblk, ent_off, tgt_off}
    // single-pivot levels ("rank-1 levels": the nearly dense chain at the end of the head) are
    // done as an outer product rows(k) x cols(k) instead: lvl.w = index of the level's
    // descriptor pair in r1 (-1 otherwise): {ua0, nua4, ub0, nub4}, {tgt_off, nr, nj4, 0};
    // rows = the level's multipliers in order, cols = U(k,:) in storage order (head part and
    // tail part each padded to a multiple of 4); r1tgt holds 4 target positions per
    // (row chunk of 32, column chunk of 4, lane), padded with the trash slot zpos + 1
    std::vector<int> r1;
    std::vector<uint16_t> r1tgt;
    std::vector<uint32_t> piv;                          // diag position | pivot row << 16
    std::vector<uint32_t> mul;                          // position of a(i,k) | position of a(k,k) << 16
    std::vector<uint16_t> tgt;                          // [32 per block] target position, 0xFFFF = idle
    std::vector<uint32_t> ent;                          // pos_l | pos_u << 16 (padding: zpos twice);
                                                        // entry j of lane l of block b of a group at
                                                        // ent_off + (b*width + j)*32 + l
  } glu;
  // ---- staged triangular solves of the head block (racg_integrate.cu solve_glu).  The last
  // <= 96 head rows (deepest levels: a nearly dense chain) form 32-row blocks S whose diagonal
  // blocks are inverted explicitly after every factorisation; everything else goes level by
  // level.  A stage is one pass of the CTA's threads: row r = tid / lpr gathers
  // sum V[pos] * x[col] over its entries (lane tid % lpr takes every lpr-th one); the tables
  // are small enough to live in shared memory between factorisations.
  struct SolveSched {
    int nf = 0, nb = 0, nblkS = 0, nent = 0, nrp = 0, nrows = 0;
    std::vector<int> st;           // int4 per stage: {kind | log2(lpr)<<8, nrows | block<<16, row_off, rp_off}
                                   // kind 0: x = b - acc; 1: x = (b - acc) / pivot; 2: x_blk = Inv[block] (b - acc)
    std::vector<uint32_t> blob;    // [nent] entries pos | col<<16, then row pointers (u16 pairs), then row ids
    std::vector<uint32_t> ext;     // S diagonal blocks: pos | tile offset << 16 (tile = b*33*32 + c*33 + r)
  } ss;

  // ---- stand-alone K2, streaming variant (racg_batch.cu rhs_stream_kernel): reactions in chunks of
  // RC rows whose rate coefficients arrive by TMA tile loads; the 32 warps of a CTA each own up to
  // SPW species of a 32-cell tile (accumulators in registers).
  //   flux lists   per chunk the reactions sorted by flux kind (FK_ONE, FK_TWO, FK_SAT):
  //                word = chunk-local row | r1 << 9 | (r2 or saturation index) << 19
  //   run lists    per (warp, chunk): header words (slot | n_minus_groups << 5 | n_plus_groups << 18,
  //                padded to a multiple of 4), then the entries: byte offsets of the chunk's rows
  //                (row pitch 256 B; padding = the zero row RC) in groups of 4, per run first the
  //                consumed (-) then the produced (+) terms of the species inside the chunk, each in
  //                reaction order; a coefficient of magnitude m is m entries (as the reference's
  //                loop subtracts/adds the flux once per occurrence, src/disk.f90:4644-4650)
  struct RhsChunks {
    int RC = 0, nchunk = 0, nwarp = 32, spw = 0;
    std::vector<int> slot_species;      // [nwarp*spw] species id or -1
    std::vector<uint32_t> off;          // [nwarp*nchunk] offset of the run list in stream (multiple of 4)
    std::vector<int> nrun;              // [nwarp*nchunk]
    std::vector<int> len4;              // [nwarp*nchunk] length of the list (headers + entries) in 16-byte groups
    int max_len4 = 0;
    std::vector<uint32_t> stream;
    std::vector<int> fl_off;            // [nchunk*4] start of the ONE / TWO / SAT lists of the chunk, end
    std::vector<uint32_t> flux;
  } rhsc;
  // ---- Jacobian gather into the storage index space, two passes (d/dy_r1, d/dy_r2) ----
  Gather jac[2];
  // map from the user's CSC slot (ia/ja) to the storage index or -1
  std::vector<int> csc_to_store;
  std::vector<int> rx_species;    // [6*R] r1,r2,p1..p4 (0-based, -1 none)
  // ---- stand-alone K3 (column-group schedule, see racg_batch.cu) ----
  struct JacCols {
    int ngroups = 0, max_pairs = 0;
    std::vector<int> grp_pair_ptr, grp_slot_ptr, grp_accum;
    std::vector<uint32_t> pair;        // r | which<<16
    std::vector<int> slot_id, slot_ent_ptr;
    std::vector<uint32_t> ent;         // local pair idx | (coef+4)<<24
    std::vector<int> zero_slots;
    // the same schedule repacked for jac_kernel_pipe (one load per item, no dependent index chains):
    std::vector<int> grp_two_ptr, grp_sat_ptr;   // [ngroups] first two-body / first saturating pair of the group (pairs sorted by kind)
    std::vector<uint32_t> pairw;       // 2 words per pair: kind-specific operand rows (see build), reaction
    std::vector<uint32_t> slotw;       // 2 words per listed slot, heaviest first within a group: CSC slot, ent4 offset | n4<<24
    std::vector<uint32_t> ent4;        // 4 words per entry group; entry = pair index | high 16 bits of the coefficient (a double) << 16; padding = +1 x the zero row (index max_pairs)
  } jc;
  std::string error;
};

bool build_host_net(HostNet& hn, int R, int N, const int* reac, const int* prod, const int* n_reac,
                    const int* n_prod, const int* itype, const double* ABC, const double* T_range,
                    const char* ctype, const char* names, const int* elements,
                    const double* mass_num, const double* vib_freq, const double* Edesorb,
                    const int* dupli_ptr, const int* dupli_list, const racg_cfg* cfg, int nthreads);
std::string describe_host_net(const HostNet& hn);

// Host-side consistency check of the factorisation / solve schedules against the symbolic
// pattern they were derived from (used by the CPU tests; returns false and a message on the
// first inconsistency): every update a(i,j) -= l(i,k) u(k,j) of the head pivots appears exactly
// once and in the level of its pivot, the targets of one level are distinct and never operands
// of that level, every multiplier is listed once, and the staged solves cover every L and U
// entry of the head block exactly once.
bool selfcheck_schedules(const HostNet& hn, std::string& err);

}  // namespace racg
