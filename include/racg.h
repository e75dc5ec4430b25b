/*
 * racg.h -- C-ABI of libracg.so: the B200 (sm_100a) replacement for RAC-2D's
 * per-cell stiff-chemistry solve.
 *
 * The reference has no FFI seam; the cut is the zero-argument
 *     call chem_evol_solve              (reference src/disk.f90:1686)
 * inside calc_this_cell (src/disk.f90:1629-1801), whose inputs/outputs are the
 * module globals of src/chemistry.f90:158-170 and whose only callees are DLSODES
 * (src/opkdmain.f:1756) with the callbacks chem_ode_f / chem_ode_jac
 * (src/disk.f90:4569, 4746).  This header declares what a Fortran host binds with
 * ISO_C_BINDING instead (see INTEGRATION.md for the interface module).
 *
 * Conventions: every array is caller-owned, column-major as Fortran lays it out,
 * index VALUES are 1-based exactly as in the reference tables; scalars by value.
 * Every function returns 0 on success and a negative code on error (never
 * aborts, never falls back to a CPU path); racg_last_error() gives the text.
 * A handle is created on the CUDA device current at creation and can be replicated
 * to more GPUs of the node with racg_use_devices(); every compute entry point selects
 * the handle's device(s) itself and restores the caller's device on return.  No state
 * is shared between handles (each owns its tables, workspace, stream and constant-memory
 * slot); calls on ONE handle must not overlap (one host thread per handle, and for the
 * *_dev variants one stream per handle, in stream order).  The library reads no
 * environment variables.
 */
#ifndef RACG_H
#define RACG_H

#ifdef __cplusplus
extern "C" {
#endif

#define RACG_NPAR 32     /* doubles per cell record (enum racg_par) */
#define RACG_NSTAT 16    /* doubles per cell in stats */
#define RACG_NPHASE 32   /* cycle counters returned by racg_phase_cycles */
#define RACG_NAME_LEN 12 /* const_len_species_name, src/chemistry.f90:11 */
#define RACG_NELEM 20    /* const_nElement, src/chemistry.f90:20 */

/* error codes */
#define RACG_OK 0
#define RACG_ERR_ARG (-1)
#define RACG_ERR_CUDA (-2)
#define RACG_ERR_NETWORK (-3)   /* what the reference turns into error_stop */
#define RACG_ERR_UNSUPPORTED (-4)

/* Columns of cellpar(ncell, RACG_NPAR): the fields of type_cell_rz_phy_basic
 * (src/data_struct.f90:316-442) that chem_cal_rates (src/chemistry.f90:591-966),
 * chem_ode_f/jac and chem_set_solver_flags_alt read.  0-based. */
enum racg_par {
  RACG_P_Tgas = 0, RACG_P_Tdust, RACG_P_n_gas, RACG_P_GrainRadius_CGS,
  RACG_P_sigdust_ave, RACG_P_ndust_tot, RACG_P_ratioDust2HnucNum,
  RACG_P_SitesPerGrain, RACG_P_zeta_cosmicray_H2, RACG_P_zeta_Xray_H2,
  RACG_P_Ncol_toISM, RACG_P_omega_albedo, RACG_P_G0_UV_toISM,
  RACG_P_G0_UV_toStar, RACG_P_G0_UV_H2phd, RACG_P_G0_UV_toStar_photoDesorb,
  RACG_P_Av_toISM, RACG_P_Av_toStar, RACG_P_phflux_Lya,
  RACG_P_fss_toISM_H2, RACG_P_fss_toISM_CO, RACG_P_fss_toISM_H2O,
  RACG_P_fss_toISM_OH, RACG_P_fss_toStar_H2, RACG_P_fss_toStar_CO,
  RACG_P_fss_toStar_H2O, RACG_P_fss_toStar_OH, RACG_P_COUNT
};

/* The scalars of chemsol_params (src/chemistry.f90:107-135) and of phy_const
 * (src/sub_global_variables.f90) the path uses.  The Fortran host fills the
 * constants from its own modules so that the reference's digits are used
 * (SURVEY App. F.7); racg_default_cfg() fills the same values. */
typedef struct racg_cfg {
  double Diff2DesorRatio;               /* chemsol_params, default 0.5 */
  double special_gH_E_diff;             /* 225 */
  int H2_form_use_moeq;                 /* only 0 supported */
  int use_special_gH_mobi;              /* 0 / 1 */
  int update_gH_params_realtime;        /* only 0 supported */
  int evol_dust_size;                   /* only 0 supported */
  double phy_Pi, phy_elementaryCharge_SI, phy_CoulombConst_SI, phy_mProton_CGS,
         phy_kBoltzmann_SI, phy_kBoltzmann_CGS, phy_hbarPlanck_CGS,
         phy_SecondsPerYear, phy_Habing_photon_flux_CGS, phy_UVext2Av,
         const_cosmicray_attenuate_N, const_cosmicRay_intensity_0,
         CosmicDesorpPreFactor, CosmicDesorpGrainT;
} racg_cfg;

typedef struct racg_handle racg_handle;

const char* racg_last_error(void);
void racg_default_cfg(racg_cfg* cfg);
/* selects the CUDA device for subsequent racg_network_create calls of this thread */
int racg_set_device(int device);

/* Replaces the once-per-run setup chem_make_sparse_structure +
 * chem_prepare_solver_storage + DLSODES' DIPREP/DPREP (src/chemistry.f90:1858-1885,
 * 1943-1973; src/opkda1.f:1210-1540): takes the already-parsed tables of
 * type_chemical_evol_reactions / _species (src/chemistry.f90:72-104) from the
 * Fortran loaders, packs them into SoA device arrays, builds IA/JA, computes a
 * fill-reducing ordering and the symbolic LU every cell shares, uploads.
 *   reac(3,R), prod(4,R)       int, 1-based species ids, 0 = empty
 *   n_reac(R), n_prod(R), itype(R)
 *   ABC(3,R), T_range(2,R)
 *   ctype(2,R)                 character(len=2) ctype(R)
 *   names(12,N)                character(len=12) names(N), blank padded
 *   elements(20,N), mass_num(N), vib_freq(N), Edesorb(N)
 *   dupli_ptr(R+1) 0-based offsets into dupli_list(*) (1-based reaction ids):
 *                              chem_net%dupli(i)%list flattened */
int racg_network_create(racg_handle** h, int R, int N, const int* reac, const int* prod,
                        const int* n_reac, const int* n_prod, const int* itype,
                        const double* ABC, const double* T_range, const char* ctype,
                        const char* names, const int* elements, const double* mass_num,
                        const double* vib_freq, const double* Edesorb, const int* dupli_ptr,
                        const int* dupli_list, const racg_cfg* cfg);
int racg_destroy(racg_handle* h);

/* Multi-GPU (north_star (e): cells are independent, shards never exchange data): replicate the
 * network tables and workspaces on the listed CUDA devices (ndev <= 0: every visible device).
 * Afterwards the host-pointer racg_solve_batch deals the cells of a batch to these devices
 * (by descending cost of the previous batch of the same size, else round-robin), runs the
 * shards concurrently from the calling thread and gathers the results into the caller's arrays.
 * This is what the serial cell loop of the Fortran host (src/disk.f90:864-938) needs to use
 * all GPUs of a node through one call.  The *_dev variants always run on the first device. */
int racg_use_devices(racg_handle* h, int ndev, const int* devices);
int racg_device_count(const racg_handle* h);
/* options: "warm_order" (default 1: a batch with as many cells as the previous one is queued
 * heaviest-first from that batch's per-cell cost; order only, results do not depend on it),
 * "level_lu" (default 1; 0 forces the generic factorisation path, diagnostics) */
int racg_set_option(racg_handle* h, const char* name, double value);
/* human-readable summary of the shared symbolic factorisation and schedules */
int racg_network_describe(const racg_handle* h, char* buf, int len);

/* sizes[0..7] = R, N, NEQ, NNZ (= chemsol_params%NNZ), NNZ after DPREP's diagonals,
 * nnz(L+D+U) of the shared symbolic factorisation (species block), dense-tail size,
 * number of factorisation levels */
int racg_network_sizes(const racg_handle* h, int* sizes);
/* IA(NEQ+1), JA(NNZ) exactly as chem_prepare_solver_storage stores them in
 * IWORK(31:31+NEQ) and IWORK(32+NEQ:31+NEQ+NNZ) (src/chemistry.f90:1962-1971):
 * 1-based, column-major, rows ascending, without DPREP's added diagonals. */
int racg_network_pattern(const racg_handle* h, int* ia, int* ja);
/* the elimination order (perm(N), 1-based species ids, first eliminated first) */
int racg_network_ordering(const racg_handle* h, int* perm);

/* K1: chem_cal_rates for a batch.  cellpar(ncell,NPAR) -> rates(ncell,R) [yr^-1] */
int racg_rates(racg_handle* h, int ncell, const double* cellpar, double* rates);
/* K2/K3: chem_ode_f and the whole chem_ode_jac matrix for a batch.
 * y(ncell,NEQ), rates(ncell,R) -> ydot(ncell,NEQ), pd(ncell,NNZ) in the slot order
 * of racg_network_pattern.  ydot or pd may be NULL. */
int racg_rhs_jac(racg_handle* h, int ncell, const double* cellpar, const double* y,
                 const double* rates, double* ydot, double* pd);

/* chem_set_solver_flags_alt(j) (src/chemistry.f90:205-268) for a batch:
 * rtol(ncell,NEQ), atol(ncell,NEQ) */
int racg_solver_flags_alt(const racg_handle* h, int j, double RTOL, double ATOL, int ncell,
                          const double* cellpar, double* rtol, double* atol);

typedef struct racg_solve_params {
  double ratio_tstep;        /* chemsol_params%ratio_tstep */
  int mxstep_per_interval;   /* IWORK(6) */
  int steps_reset_solver;    /* chemsol_params%steps_reset_solver */
  int nrec_max;              /* leading capacity of touts/record per cell (0 allowed when both are NULL);
                                capacity only: every cell runs its own n_record
                                (chem_evol_solve_prepare_run_once, src/chemistry.f90:1894-1899) and
                                records beyond nrec_max are not stored */
  int tol_policy_j;          /* used when rtol/atol are NULL: chem_set_solver_flags_alt(j) */
  double RTOL, ATOL;         /* chemsol_params%RTOL/ATOL for the policy */
  double max_runtime_allowed;/* chemsol_params%max_runtime_allowed (src/chemistry.f90:116, 438, 480-491)
                                in MODEL seconds, <= 0: no budget.  The reference measures cpu_time; here
                                the clock is the deterministic work model racg_model_runtime() so that the
                                same cell is cut at the same place on every machine (and in the oracle). */
} racg_solve_params;

/* The deterministic stand-in for the reference's cpu_time clock (src/sub_trivials.f90:25-42):
 * seconds = c_f*NFE + c_jac*NJE + c_lu*NLU + c_solve*n_solve + c_step*NST, constants = what the
 * reference algorithm (O(NEQ*R) Jacobian, scalar sparse LDU) costs per operation on one host core
 * per unit of network size.  coef[5] receives c_f, c_jac, c_lu, c_solve, c_step
 * for this network. */
int racg_model_runtime_coefs(const racg_handle* h, double* coef);

/* The batch replacement for the loop body around `call chem_evol_solve`
 * (src/disk.f90:1686; semantics of src/chemistry.f90:391-588 with evolT=.false.
 * and the cpu_time budgets replaced by the deterministic work model, see max_runtime_allowed).
 *   in : cellpar(ncell,NPAR), y0(ncell,NEQ) [y0(:,NEQ) = Tgas], t0/tmax/dt_first(ncell),
 *        rtol/atol(ncell,NEQ) or NULL (policy j)
 *   out: y_final(ncell,NEQ), t_final(ncell), nrec_real(ncell), istate(ncell)
 *        [DLSODES codes: 2 ok, -1..-7], quality(ncell) [bits 1,2,256,512 as
 *        src/chemistry.f90:506,528,577-582; bit 1024 = DLSODES returned -7 (sparse-solver
 *        failure), where the reference calls error_stop (src/chemistry.f90:375-380)],
 *        stats(ncell,RACG_NSTAT) =
 *        NST,NFE,NJE,NLU,NQU,n_solve,NERR,n_restart,n_cfail,n_efail,nrec_real,istate,HU,
 *        model runtime [s], premature-finish flag (1 = the work budget ended the cell), 0
 *        touts(ncell,nrec_max), record(ncell,NEQ,nrec_max): optional (NULL to skip). */
int racg_solve_batch(racg_handle* h, int ncell, const double* cellpar, const double* y0,
                     const double* rtol, const double* atol, const double* t0,
                     const double* tmax, const double* dt_first, const racg_solve_params* sp,
                     double* y_final, double* t_final, double* touts, double* record,
                     int* nrec_real, int* istate, int* quality, double* stats);

/* The batch form of calc_this_cell's local-iteration loop (src/disk.f90:1651-1791), evolT=.false.:
 * for j = 1..nlocal_iter: tolerances chem_set_solver_flags_alt(j) (from sp->RTOL/ATOL); for j > 1 the
 * run continues from the harvested state with t0 = t_final, dt_first = max(dt_first0, 1e-3*t0) and
 * rectify_abundances (charge neutralised with E-; src/disk.f90:2103-2146, src/chemistry.f90:2170-2201);
 * after every solve the LAST record whose T and X(H2) are not NaN is harvested (src/disk.f90:1716-1733);
 * a cell leaves the ladder when quality == 0 or t_final >= 0.5*tmax (1785-1789), when a continuation
 * does not proceed (1703-1712) or produces no useful record (1737-1740).  Cells still in the ladder are
 * re-submitted as a compacted batch inside this call.
 *   in : cellpar(ncell,NPAR), y0(ncell,NEQ), tmax(ncell), dt_first0, sp (ratio_tstep, mxstep, reset,
 *        RTOL, ATOL, max_runtime_allowed; nrec_max and tol_policy_j are ignored)
 *   out: abundances(ncell,NEQ) [= leaves%list(id)%p%abundances and Tgas], t_final(ncell) [par%t_final],
 *        quality(ncell), istate(ncell) of the last solve, n_iter_used(ncell),
 *        R_H2_form_rate_coeff(ncell) [src/chemistry.f90:804,891; s^-1 based unit as in the reference],
 *        n_mol_on_grain(ncell) [get_ice_coverage, src/chemistry.f90:989-1003] (either may be NULL),
 *        stats(ncell,RACG_NSTAT) accumulated over the ladder */
int racg_calc_batch(racg_handle* h, int ncell, const double* cellpar, const double* y0, const double* tmax,
                    double dt_first0, const racg_solve_params* sp, int nlocal_iter, double* abundances,
                    double* t_final, int* quality, int* istate, int* n_iter_used, double* R_H2_form_rate_coeff,
                    double* n_mol_on_grain, double* stats);

/* On-disk format of the reference's chemistry checkpoint (back_cells_chemical_data,
 * src/data_dump.f90:88-162): 'chemical_data_iter_NNNN.bin' (iiter < 0: 'chemical_data.bin') in
 * directory dir, a direct-access unformatted file whose record i holds abundances(1:nSpecies),
 * col_den_toStar(1:ncd), col_den_toISM(1:ncd) of leaf i as native doubles (record length
 * 8*(nSpecies+2*ncd) bytes, no record markers).  Written straight from the gathered batch arrays:
 * abundances(ncell, >= nspecies) as returned by racg_calc_batch / racg_solve_batch,
 * col_den_*(ncd, ncell) as the host keeps them.  Host only, no GPU needed. */
int racg_write_chemical_data(const char* dir, int iiter, int ncell, int nspecies, const double* abundances,
                             int ncd, const double* col_den_toStar, const double* col_den_toISM);
int racg_read_chemical_data(const char* dir, int iiter, int ncell, int nspecies, double* abundances, int ncd,
                            double* col_den_toStar, double* col_den_toISM);

/* ---- device-pointer variants (same layouts, buffers already in HBM; `stream` is a
 * cudaStream_t or NULL) for callers that keep the grid resident on the GPU ---- */
int racg_rates_dev(racg_handle* h, int ncell, const double* cellpar, double* rates, void* stream);
int racg_rhs_jac_dev(racg_handle* h, int ncell, const double* cellpar, const double* y,
                     const double* rates, double* ydot, double* pd, void* stream);
int racg_solve_batch_dev(racg_handle* h, int ncell, const double* cellpar, const double* y0,
                         const double* rtol, const double* atol, const double* t0,
                         const double* tmax, const double* dt_first,
                         const racg_solve_params* sp, double* y_final, double* t_final,
                         double* touts, double* record, int* nrec_real, int* istate,
                         int* quality, double* stats, void* stream);
/* diagnostics: f(y) and the Jacobian as the integrator's own in-kernel routines compute them
 * (not the stand-alone K2/K3 kernels).  y(ncell,NEQ) -> f(ncell,NEQ) [slot NEQ undefined],
 * jstore(ncell,nstore) in the library's LU storage order; csc_to_store(NNZ) maps the slots of
 * racg_network_pattern to that order (-1: structurally zero for evolT=.false.).  With
 * ncell <= 0 only *nstore and csc_to_store are filled.  con != 0: f receives instead the solution
 * x of (I + con*J) x = f(y) from the integrator's own factorisation and solve (con = -h*el0). */
int racg_debug_fjac(racg_handle* h, int ncell, const double* cellpar, const double* y, double* f,
                    double* jstore, int* csc_to_store, int* nstore, double con);
/* host-side consistency check of the level-parallel factorisation schedule and of the staged
 * solves against the symbolic pattern (no GPU needed): 0 = consistent, RACG_ERR_NETWORK with a
 * message otherwise */
int racg_selfcheck(const racg_handle* h);
/* the same check on a copy of the schedules damaged on purpose (mode 1: one operand position
 * flipped, 2: one target written twice): must return RACG_ERR_NETWORK (test of the checker) */
int racg_selfcheck_damaged(const racg_handle* h, int mode);
/* number of kernel launches issued through this handle so far */
long racg_launch_count(const racg_handle* h);
/* per-phase SM-cycle counters of the last racg_solve_batch* call, summed over CTAs:
 * out[0..RACG_NPHASE-1]; returns 0 */
int racg_phase_cycles(racg_handle* h, double* out);

#ifdef __cplusplus
}
#endif
#endif
