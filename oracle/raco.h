/*
 * raco.h -- C API of the CPU ORACLE for the RAC-2D per-cell stiff-chemistry path.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a CPU restatement (C++17, g++) of
 * the reference's algorithm for the hot path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * library (rac-2d_b200/csrc -> libracg.so) never links, includes or calls it.
 *
 * The reference is Fortran (no Fortran compiler exists in this image),
 * so the reference itself cannot be built into oracle/_ref: the oracle is a PORT.
 *
 * PARITY PINNING
 *   pinned   : DLSODES documentation example (reference src/opkdmain.f:1919-2133):
 *              Y(t) table + NNZ=44 (tests/golden/dlsodes_example.json);
 *              parser golden values of SURVEY App. B / App. E (species indices,
 *              R/N/NNZ, itype histograms), tests/golden/network_*.json.
 *   UNPINNED : trajectories on the real rate06/rate12 networks -- the reference
 *              holds no test, fixture or golden vector for them ("parity
 *              unpinned"); they are cross-checked against scipy's independent
 *              LSODA/BDF integrators driven by this oracle's f/J at tight
 *              tolerance (tests/test_oracle_integrator.py).
 *
 * Each function cites the reference file:line it restates.
 */
#ifndef RACO_H
#define RACO_H

#ifdef __cplusplus
extern "C" {
#endif

#define RACO_NPAR 32          /* doubles per cell-parameter record (layout below) */
#define RACO_NAME_LEN 12      /* const_len_species_name, src/chemistry.f90:11 */
#define RACO_NELEM 20         /* const_nElement, src/chemistry.f90:20 */

/* cell-parameter record: the subset of type_cell_rz_phy_basic
 * (src/data_struct.f90:316-442) that chem_cal_rates/f/J read. */
enum raco_par {
  RACO_P_Tgas = 0, RACO_P_Tdust, RACO_P_n_gas, RACO_P_GrainRadius_CGS,
  RACO_P_sigdust_ave, RACO_P_ndust_tot, RACO_P_ratioDust2HnucNum,
  RACO_P_SitesPerGrain, RACO_P_zeta_cosmicray_H2, RACO_P_zeta_Xray_H2,
  RACO_P_Ncol_toISM, RACO_P_omega_albedo, RACO_P_G0_UV_toISM,
  RACO_P_G0_UV_toStar, RACO_P_G0_UV_H2phd, RACO_P_G0_UV_toStar_photoDesorb,
  RACO_P_Av_toISM, RACO_P_Av_toStar, RACO_P_phflux_Lya,
  RACO_P_fss_toISM_H2, RACO_P_fss_toISM_CO, RACO_P_fss_toISM_H2O,
  RACO_P_fss_toISM_OH, RACO_P_fss_toStar_H2, RACO_P_fss_toStar_CO,
  RACO_P_fss_toStar_H2O, RACO_P_fss_toStar_OH
};

/* chemsol_params scalars used by the path (src/chemistry.f90:107-135). */
typedef struct raco_cfg {
  double Diff2DesorRatio;      /* 0.5 */
  double special_gH_E_diff;    /* 225 */
  int H2_form_use_moeq;        /* 0 */
  int use_special_gH_mobi;     /* 0 */
  int update_gH_params_realtime; /* 0 (only 0 supported) */
  int jac_mode;                /* 0: O(R) scatter Jacobian; 1: reference-faithful
                                  column-by-column full reaction scan (F6) */
} raco_cfg;

typedef struct raco_net raco_net;

/* ---- network setup (src/chemistry.f90:1427-1454,1364-1424,1221-1360,
 *      1188-1217,1089-1185,1858-1885,1943-1973) ---- */
raco_net* raco_net_load(const char* network_file);
void raco_net_free(raco_net*);
const char* raco_last_error(void);
/* sizes[0..7] = R, N, NEQ, NNZ(mask), nGrainSpecies, n_dupli_total, nnz_after_diag, nnz_ldu */
void raco_net_sizes(const raco_net*, int* sizes);
void raco_net_tables(const raco_net*, int* reac /*3*R*/, int* prod /*4*R*/,
                     int* n_reac, int* n_prod, int* itype, double* ABC /*3*R*/,
                     double* T_range /*2*R*/, char* ctype /*2*R*/);
void raco_net_species(const raco_net*, char* names /*12*N*/, int* elements /*20*N*/,
                      double* mass_num, double* vib_freq, double* Edesorb,
                      int* idx_counterpart);
/* dupli_ptr[R+1], dupli_list[n_dupli_total] (1-based reaction ids) */
void raco_net_dupli(const raco_net*, int* dupli_ptr, int* dupli_list);
/* special[0..31]: idx(10) of H2,H,E-,C,C+,O,O2,CO,H2O,OH then
 * i_Hplus,i_Heplus,i_gH,i_gH2,i_Grain0,i_GrainM,i_GrainP,i_gH2O,i_gCO,i_gCO2,
 * i_gN2,i_NII,i_SiII,i_FeII,i_NI  (1-based, 0 = absent) */
void raco_net_special(const raco_net*, int* special);
void raco_net_grain_species(const raco_net*, int* idx /*nGrainSpecies*/);
/* IA(NEQ+1), JA(NNZ): exactly IWORK(31:) of chem_prepare_solver_storage */
void raco_net_pattern(const raco_net*, int* ia, int* ja);
/* y0[N]; returns 0 ok (src/chemistry.f90:1978-2024) */
int raco_load_initial_abundances(const raco_net*, const char* file, double* y0);

/* ---- per-cell arithmetic ---- */
/* chem_cal_rates, src/chemistry.f90:591-966.  rates[R] in yr^-1. returns 0 ok */
int raco_cal_rates(const raco_net*, const raco_cfg*, const double* par, double* rates);
/* chem_ode_f fixed-T branch, src/disk.f90:4569-4659 */
void raco_ode_f(const raco_net*, const raco_cfg*, const double* par,
                const double* rates, const double* y, double* ydot);
/* chem_ode_jac fixed-T branch, src/disk.f90:4746-4903; column j (1-based) */
void raco_ode_jac_col(const raco_net*, const raco_cfg*, const double* par,
                      const double* rates, const double* y, int j, double* pdj);
/* the same Jacobian scattered in one O(R) sweep into the CSC slots of
 * raco_net_pattern (pd[NNZ]) */
void raco_ode_jac_csc(const raco_net*, const raco_cfg*, const double* par,
                      const double* rates, const double* y, double* pd);
/* rounding scales for parity tests: the same sums with |term| */
void raco_ode_f_abs(const raco_net*, const raco_cfg*, const double* par, const double* rates,
                    const double* y, double* out);
void raco_ode_jac_csc_abs(const raco_net*, const raco_cfg*, const double* par, const double* rates,
                          const double* y, double* pd);
/* chem_set_solver_flags_alt(j), src/chemistry.f90:205-268 */
void raco_set_solver_flags_alt(const raco_net*, int j, double RTOL, double ATOL,
                               double ratioDust2HnucNum, double* rtols, double* atols);

/* ---- chem_evol_solve for one cell, src/chemistry.f90:391-588 ---- */
typedef struct raco_solve_opts {
  double t0, t_max, dt_first_step, ratio_tstep;
  int mxstep_per_interval;   /* IWORK(6) */
  int steps_reset_solver;
  int n_record;              /* capacity of touts/record (>= computed n_record) */
  double max_runtime_allowed; /* chemsol_params%max_runtime_allowed (src/chemistry.f90:116) in MODEL
                                 seconds (raco_model_runtime_coefs), <= 0: budgets disabled */
} raco_solve_opts;

/* Deterministic stand-in for the reference's cpu_time clock (src/sub_trivials.f90:25-42;
 * used by src/chemistry.f90:438, 480-491): seconds = c_f*NFE + c_jac*NJE + c_lu*NLU +
 * c_solve*n_solve + c_step*NST with per-operation costs of the reference algorithm on one
 * host core, scaled by network size: c_f = 1.04e-8*R, c_jac = 6.45e-9*NEQ*R (column-wise
 * chem_ode_jac), c_lu = 1.41e-7*NNZ, c_solve = 3.0e-9*NNZ, c_step = 6.4e-8*NEQ (fitted to this
 * oracle, jac_mode = 1, built -O2, 300 cells of the config-2 stream; the constants are
 * part of the model and do not follow the build flags).  coef[5]. */
void raco_model_runtime_coefs(int R, int NEQ, int NNZ, double* coef);

/* stats[0..15]: NST,NFE,NJE,NLU,NQU(last),n_solve,n_err,n_restart,n_cfail,n_efail,
 *               n_record_real, last istate, ... accumulated over the whole cell */
int raco_evol_solve(const raco_net*, const raco_cfg*, const double* par,
                    const raco_solve_opts*, double* y /*NEQ in/out*/,
                    double* rtols /*NEQ, mutated*/, double* atols /*NEQ, mutated*/,
                    double* touts /*n_record*/, double* record /*NEQ*n_record, may be NULL*/,
                    double* t_final, int* n_record_real, int* istate_last, int* quality,
                    double* stats /*16*/);
/* n_record formula, src/chemistry.f90:1894-1899 */
int raco_n_record(double t0, double t_max, double dt_first_step, double ratio);

/* batch over cells with OpenMP-style std::thread pool (CPU baseline).
 * par[ncell*RACO_NPAR] (row per cell), y0[ncell*NEQ] (row per cell).
 * Tolerances by policy j (chem_set_solver_flags_alt). */
int raco_evol_solve_batch(const raco_net*, const raco_cfg*, int ncell, const double* par,
                          const double* y0, int tol_policy_j, double RTOL, double ATOL,
                          const raco_solve_opts*, int nthreads,
                          double* y_final, double* t_final, int* istate, int* quality,
                          double* stats /*ncell*16*/);

/* ---- generic DLSODES restatement (MF=21/121 semantics) on a user problem,
 *      used for the documentation-example KAT, src/opkdmain.f:1919-2133 ---- */
typedef void (*raco_f_cb)(int neq, double t, const double* y, double* ydot, void* ctx);
typedef void (*raco_jac_cb)(int neq, double t, const double* y, int j /*1-based*/,
                            double* pdj, void* ctx);
typedef struct raco_lsodes raco_lsodes;
/* ia/ja: 1-based CSC pattern as DLSODES takes in IWORK(31:) (MOSS=0). */
raco_lsodes* raco_lsodes_create(int neq, const int* ia, const int* ja,
                                raco_f_cb f, raco_jac_cb jac, void* ctx);
void raco_lsodes_free(raco_lsodes*);
/* mirrors CALL DLSODES(...) with ITOL=4 vectors; itask 1 or 4; iopt inputs:
 * maxord, mxstep, hmax, tcrit.  Returns istate. */
int raco_lsodes_call(raco_lsodes*, double* y, double* t, double tout,
                     const double* rtol, const double* atol, int itask, int istate,
                     int maxord, int mxstep, double hmax, double tcrit);
/* iwork-like counters: out[0..9]=NST,NFE,NJE,NQU,NQ,IMXER,NNZ,NLU,NZL,NZU; hu */
void raco_lsodes_stats(const raco_lsodes*, int* out, double* hu);

#ifdef __cplusplus
}
#endif
#endif
