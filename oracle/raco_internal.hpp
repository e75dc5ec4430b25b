// raco_internal.hpp -- internal structures of the CPU oracle (test infrastructure only;
// see raco.h).  Physical constants restate src/sub_global_variables.f90:3-90 and
// src/chemistry.f90:179-181 digit for digit.
#pragma once
#include <string>
#include <vector>
#include <array>
#include <cmath>
#include <cstring>
#include <functional>
#include "raco.h"

namespace raco {

// src/sub_global_variables.f90
constexpr double phy_Pi = 3.1415926535897932384626433;
constexpr double phy_elementaryCharge_SI = 1.602176487e-19;
constexpr double phy_CoulombConst_SI = 8.9875517873681764e9;
constexpr double phy_mProton_CGS = 1.67262158e-24;
constexpr double phy_kBoltzmann_SI = 1.3806503e-23;
constexpr double phy_kBoltzmann_CGS = 1.3806503e-16;
constexpr double phy_hbarPlanck_CGS = 1.054571628e-27;
constexpr double phy_SecondsPerYear = 3600.0 * 24.0 * 365.0;
constexpr double phy_Habing_photon_flux_CGS = 6e7;
constexpr double phy_UVext2Av = 2.6;
constexpr double const_cosmicray_attenuate_N = 5.75e25;
constexpr double const_SitesDensity_CGS = 1e15;
// src/chemistry.f90:179-181
constexpr double const_cosmicRay_intensity_0 = 1.36e-17;
constexpr double CosmicDesorpPreFactor = 3.16e-19;
constexpr double CosmicDesorpGrainT = 70.0;

struct Net {
  int R = 0, N = 0, NEQ = 0, NNZ = 0, nGrain = 0;
  // all index VALUES are 1-based as in the Fortran tables (0 = empty slot)
  std::vector<std::array<std::string, 3>> reac_names;
  std::vector<std::array<std::string, 4>> prod_names;
  std::vector<int> reac, prod;          // [3*R], [4*R] column-major like reac(3,R)
  std::vector<int> n_reac, n_prod, itype;
  std::vector<double> ABC, T_range;     // [3*R], [2*R]
  std::vector<std::string> ctype;       // 2 chars
  std::vector<std::vector<int>> dupli;  // earlier twins (1-based)
  std::vector<std::string> names;       // species, trimmed
  std::vector<int> elements;            // [20*N]
  std::vector<double> mass_num, vib_freq, Edesorb;
  std::vector<int> counterpart;         // idx_gasgrain_counterpart, -1 if none
  std::vector<int> grain_idx;           // idxGrainSpecies (1-based)
  int special[32] = {0};
  std::vector<int> ia, ja;              // 1-based CSC as in IWORK(31:)
  // derived helpers (0-based)
  std::vector<int> slot_of;             // dense NEQ*NEQ -> CSC slot or -1 (col-major: row + col*NEQ)
  std::vector<char> first_is_H2, first_is_gH, fss_kind; // per reaction
};

// special[] positions
enum { S_H2 = 0, S_HI, S_E, S_CI, S_CII, S_OI, S_O2, S_CO, S_H2O, S_OH,
       S_Hplus, S_Heplus, S_gH, S_gH2, S_Grain0, S_GrainM, S_GrainP, S_gH2O,
       S_gCO, S_gCO2, S_gN2, S_NII, S_SiII, S_FeII, S_NI };

void set_error(const std::string& s);

// ---- sparse LDU without pivoting at a fixed pattern (the role of YSMP's
// ODRV/CDRV in src/opkda1.f:1994-3801) ----
struct SparseLU {
  int n = 0;
  std::vector<int> perm, iperm;     // new->old, old->new
  // permuted matrix A' = P A P^T stored by rows (CSR) with sorted columns,
  // L strictly lower (unit diag), U strictly upper, D diagonal (stored inverted)
  std::vector<int> a_ptr, a_col, a_src; // a_src: CSC slot in the user's (ia,ja) or -1 for added diag
  std::vector<int> l_ptr, l_col, u_ptr, u_col;
  std::vector<double> l_val, u_val, dinv;
  int nnz_a = 0, nzl = 0, nzu = 0;
  void analyse(int n, const std::vector<int>& ia1, const std::vector<int>& ja1); // 1-based CSC
  // numeric: P given as CSC values (user's slot order) + implicit value for added diagonals
  // returns 0 ok, k>0 = zero pivot at (permuted) row k
  int factor(const double* pval, double added_diag_value);
  void solve(double* x) const;      // in place, user ordering
};

// ---- DLSODES restatement ----
struct Lsodes {
  int n = 0;
  std::vector<int> ia, ja;          // 1-based user pattern (without added diagonals)
  std::function<void(double, const double*, double*)> f;
  std::function<void(double, const double*, int, double*)> jac_col; // column j (1-based)
  // optional whole-Jacobian callback: fills CSC values pd[nnz_user]
  std::function<void(double, const double*, double*)> jac_csc;
  SparseLU lu;
  bool analysed = false;
  // state corresponding to COMMON /DLS001/, /DLSS01/
  double CONIT, CRATE, EL[14], ELCO[14][13], HOLD, RMAX, TESCO[4][13];
  double CCMAX, EL0, H, HMIN, HMXI, HU = 0, RC, TN, UROUND;
  int INIT = 0, MXSTEP, MXHNIL, NHNIL, NSLAST, NYH;
  int IALTH, IPUP, LMAX, MEO, NQNYH, NSLP;
  int ICF, IERPJ, IERSL, JCUR, JSTART, KFLAG, L;
  int METH, MITER, MAXORD, MAXCOR, MSBP, MXNCF, N, NQ = 0, NST = 0, NFE = 0, NJE = 0, NQU = 0;
  double CON0, CONMIN, CCMXJ, PSMALL, RBIG;
  int MSBJ, NSLJ, NLU = 0, IMXER = 0, iplost = 0;
  double Padd = 0.0;   // value of P on the diagonals DPREP appends to the user's pattern
  long n_solve = 0, n_cfail = 0, n_efail = 0;
  bool IHIT = false;
  double TCRIT = 0;
  std::vector<double> YH, EWT, SAVF, ACOR, FTEM, Jval, Pval, Ywork;
  int call(double* y, double* t, double tout, const double* rtol, const double* atol,
           int itask, int istate, int maxord, int mxstep, double hmax, double tcrit);
 private:
  void dcfode();
  void dstode(double* y);
  void dprjs(double* y);
  void dsolss(double* x);
  void dintdy(double t, double* dky) const;
  double dvnorm(const double* v, const double* w) const;
  void dewset(const double* rtol, const double* atol);
  bool ewt_invert_ok();
};

int cal_rates(const Net&, const raco_cfg&, const double* par, double* rates);
void ode_f(const Net&, const raco_cfg&, const double* par, const double* rates,
           const double* y, double* ydot);
void ode_jac_col(const Net&, const raco_cfg&, const double* par, const double* rates,
                 const double* y, int j, double* pdj);
void ode_jac_csc(const Net&, const raco_cfg&, const double* par, const double* rates,
                 const double* y, double* pd);

}  // namespace raco

struct raco_net { raco::Net net; };
struct raco_lsodes { raco::Lsodes s; };
