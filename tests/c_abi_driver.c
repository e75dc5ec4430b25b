/* c_abi_driver.c -- drives libracg.so through its C-ABI alone, the way the serial Fortran host
 * would (src/disk.f90:864-938 cell loop replaced by ONE racg_solve_batch call): load a network
 * with the host loaders, create the handle, replicate it on every visible GPU
 * (racg_use_devices), integrate a batch, and check that the multi-GPU result is bitwise the
 * single-GPU result.  Built and run by tests/test_gpu_c_abi.py:
 *     gcc tests/c_abi_driver.c -Iinclude -Lrac-2d_b200 -lracg -lm -o /tmp/c_abi_driver
 *     /tmp/c_abi_driver <network.dat> <initial_abundances.dat> <ncell>
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "racg.h"

/* host loaders exported by libracg.so (rac-2d_b200/csrc/chem_loader.cpp); in production the
 * Fortran loaders of src/chemistry.f90 fill the same tables */
typedef struct chem_host chem_host;
chem_host* chem_read_reactions(const char* fn, char* errbuf, int errlen);
void chem_host_free(chem_host* p);
void chem_host_sizes(const chem_host* p, int* R, int* N, int* ndupli);
void chem_host_tables(const chem_host* p, int* reac, int* prod, int* n_reac, int* n_prod, int* itype, double* ABC,
                      double* T_range, char* ctype, char* names, int* elements, double* mass_num, double* vib_freq,
                      double* Edesorb, int* dupli_ptr, int* dupli_list);
int chem_load_initial_abundances(const chem_host* p, const char* fn, double* y0);

#define CHECK(call) do { int rc_ = (call); if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, racg_last_error()); return 2; } } while (0)

static double urand(unsigned long long* s) {   /* splitmix64 */
  unsigned long long z = (*s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
  return (double)(z >> 11) / 9007199254740992.0;
}
static double logu(double u, double lo, double hi) { return pow(10.0, log10(lo) + u * (log10(hi) - log10(lo))); }

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s network.dat ic.dat ncell\n", argv[0]); return 1; }
  const int ncell = atoi(argv[3]);
  char err[256];
  chem_host* net = chem_read_reactions(argv[1], err, sizeof err);
  if (!net) { fprintf(stderr, "loader: %s\n", err); return 2; }
  int R, N, nd;
  chem_host_sizes(net, &R, &N, &nd);
  const int NEQ = N + 1;
  int *reac = malloc(sizeof(int) * 3 * R), *prod = malloc(sizeof(int) * 4 * R), *n_reac = malloc(sizeof(int) * R),
      *n_prod = malloc(sizeof(int) * R), *itype = malloc(sizeof(int) * R), *elements = malloc(sizeof(int) * RACG_NELEM * N),
      *dptr = malloc(sizeof(int) * (R + 1)), *dlist = malloc(sizeof(int) * (nd > 0 ? nd : 1));
  double *ABC = malloc(8 * 3 * R), *Tr = malloc(8 * 2 * R), *mass = malloc(8 * N), *vib = malloc(8 * N), *Ed = malloc(8 * N);
  char *ctype = malloc(2 * R), *names = malloc(RACG_NAME_LEN * N);
  chem_host_tables(net, reac, prod, n_reac, n_prod, itype, ABC, Tr, ctype, names, elements, mass, vib, Ed, dptr, dlist);
  double* y0s = malloc(8 * N);
  if (chem_load_initial_abundances(net, argv[2], y0s) != 0) { fprintf(stderr, "initial abundances\n"); return 2; }
  int iGrain0 = -1;
  for (int i = 0; i < N; ++i) if (!strncmp(names + RACG_NAME_LEN * i, "Grain0      ", RACG_NAME_LEN)) iGrain0 = i;

  racg_cfg cfg;
  racg_default_cfg(&cfg);
  racg_handle* h = NULL;
  CHECK(racg_network_create(&h, R, N, reac, prod, n_reac, n_prod, itype, ABC, Tr, ctype, names, elements, mass, vib, Ed,
                            dptr, dlist, &cfg));

  /* synthetic cells, Fortran layout a(ncell, item) */
  double* par = calloc((size_t)RACG_NPAR * ncell, 8);
  double* y0 = calloc((size_t)NEQ * ncell, 8);
  const double a = 1e-5, sig = cfg.phy_Pi * a * a, sites = 4.0 * sig * 1e15;
  const double D = 0.01 * (cfg.phy_mProton_CGS * 1.4) / (4.0 * cfg.phy_Pi / 3.0 * a * a * a * 2.0);
  unsigned long long seed = 20240613ULL;
#define P(k) par[(size_t)(k) * ncell + c]
  for (int c = 0; c < ncell; ++c) {
    const double ngas = logu(urand(&seed), 1e3, 1e13), Tg = logu(urand(&seed), 8.0, 1000.0);
    const double Td = fmin(fmax(Tg * pow(10.0, -0.5 * urand(&seed)), 5.0), 1500.0);
    const double G0 = logu(urand(&seed), 1e-2, 1e8), Avs = 20.0 * urand(&seed), Avi = 20.0 * urand(&seed);
    P(RACG_P_Tgas) = Tg; P(RACG_P_Tdust) = Td; P(RACG_P_n_gas) = ngas; P(RACG_P_GrainRadius_CGS) = a;
    P(RACG_P_sigdust_ave) = sig; P(RACG_P_ndust_tot) = ngas * D; P(RACG_P_ratioDust2HnucNum) = D;
    P(RACG_P_SitesPerGrain) = sites; P(RACG_P_zeta_cosmicray_H2) = 1.36e-17;
    P(RACG_P_zeta_Xray_H2) = logu(urand(&seed), 1e-19, 1e-11); P(RACG_P_Ncol_toISM) = Avi / 5.3e-22;
    P(RACG_P_omega_albedo) = 0.5; P(RACG_P_G0_UV_toISM) = 1.0; P(RACG_P_G0_UV_toStar) = G0;
    P(RACG_P_G0_UV_toStar_photoDesorb) = G0 * exp(-2.6 * Avs / 1.086); P(RACG_P_G0_UV_H2phd) = 0.3 * P(RACG_P_G0_UV_toStar_photoDesorb);
    P(RACG_P_Av_toISM) = Avi; P(RACG_P_Av_toStar) = Avs; P(RACG_P_phflux_Lya) = 0.0;
    for (int k = RACG_P_fss_toISM_H2; k <= RACG_P_fss_toStar_OH; ++k) P(k) = 1.0;
    P(RACG_P_fss_toISM_H2) = logu(urand(&seed), 1e-8, 1.0); P(RACG_P_fss_toStar_H2) = logu(urand(&seed), 1e-8, 1.0);
    P(RACG_P_fss_toISM_CO) = logu(urand(&seed), 1e-8, 1.0); P(RACG_P_fss_toStar_CO) = logu(urand(&seed), 1e-8, 1.0);
    for (int i = 0; i < N; ++i) y0[(size_t)i * ncell + c] = y0s[i];
    if (iGrain0 >= 0) y0[(size_t)iGrain0 * ncell + c] = D;
    y0[(size_t)N * ncell + c] = Tg;
  }
  double *t0 = calloc(ncell, 8), *tmax = malloc(8 * ncell), *dt = malloc(8 * ncell);
  for (int c = 0; c < ncell; ++c) { tmax[c] = 1e4; dt[c] = 1e-8; }
  racg_solve_params sp;
  memset(&sp, 0, sizeof sp);
  sp.ratio_tstep = 1.1; sp.mxstep_per_interval = 6000; sp.steps_reset_solver = 50; sp.nrec_max = 0;
  sp.tol_policy_j = 1; sp.RTOL = 1e-4; sp.ATOL = 1e-30; sp.max_runtime_allowed = 60.0;
  double *yf1 = malloc(8 * (size_t)NEQ * ncell), *yfN = malloc(8 * (size_t)NEQ * ncell), *tf = malloc(8 * ncell),
         *st = malloc(8 * (size_t)RACG_NSTAT * ncell);
  int *nrec = malloc(4 * ncell), *ist = malloc(4 * ncell), *q = malloc(4 * ncell);

  CHECK(racg_solve_batch(h, ncell, par, y0, NULL, NULL, t0, tmax, dt, &sp, yf1, tf, NULL, NULL, nrec, ist, q, st));
  int bad = 0;
  for (int c = 0; c < ncell; ++c) if (ist[c] != 2 || tf[c] != 1e4) ++bad;
  printf("1 device: %d cells, %d not at t_max with ISTATE 2\n", ncell, bad);

  CHECK(racg_use_devices(h, 0, NULL));                      /* every visible GPU */
  const int ndev = racg_device_count(h);
  for (int rep = 0; rep < 2; ++rep)                         /* second call: cost-aware dealing */
    CHECK(racg_solve_batch(h, ncell, par, y0, NULL, NULL, t0, tmax, dt, &sp, yfN, tf, NULL, NULL, nrec, ist, q, st));
  const int same = memcmp(yf1, yfN, 8 * (size_t)NEQ * ncell) == 0;
  printf("%d device(s): results %s the single-device results\n", ndev, same ? "bitwise equal to" : "DIFFER from");
  printf("launches through the handle: %ld\n", racg_launch_count(h));
  CHECK(racg_destroy(h));
  chem_host_free(net);
  if (bad > ncell / 20 || !same) { printf("C_ABI_DRIVER FAIL\n"); return 3; }
  printf("C_ABI_DRIVER OK ndev=%d\n", ndev);
  return 0;
}
