// raco_solve.cpp -- ORACLE (test infrastructure only, see raco.h): the per-cell
// output-time loop chem_evol_solve (src/chemistry.f90:391-588) with the error
// policy ode_solver_error_handling (272-387), for evolT = .false.  The cpu_time
// budgets (438, 480-491) make the reference non-deterministic: here the
// clock is a deterministic work model (raco_model_runtime_coefs) and the budget logic
// itself is restated literally; max_runtime_allowed <= 0 disables it.
#include "raco_internal.hpp"
#include <thread>
#include <atomic>
#include <algorithm>

using namespace raco;

extern "C" {

void raco_model_runtime_coefs(int R, int NEQ, int NNZ, double* coef) {
  coef[0] = 1.04e-8 * R;
  coef[1] = 6.45e-9 * (double)NEQ * R;
  coef[2] = 1.41e-7 * NNZ;
  coef[3] = 3.0e-9 * NNZ;
  coef[4] = 6.4e-8 * NEQ;
}

int raco_n_record(double t0, double t_max, double dt_first_step, double ratio) {
  // src/chemistry.f90:1894-1899
  return (int)std::ceil(std::log((t_max - t0) / dt_first_step * (ratio - 1.0) + 1.0) / std::log(ratio)) + 1;
}

int raco_evol_solve(const raco_net* h, const raco_cfg* cfg, const double* par,
                    const raco_solve_opts* o, double* y, double* rtols, double* atols,
                    double* touts, double* record, double* t_final, int* n_record_real_out,
                    int* istate_last, int* quality_out, double* stats) {
  const Net& n = h->net;
  const int NEQ = n.NEQ, N = n.N;
  std::vector<double> rates(n.R);
  int rc = cal_rates(n, *cfg, par, rates.data());
  if (rc != 0) return rc;
  Lsodes s;
  s.n = NEQ;
  s.ia = n.ia; s.ja = n.ja;
  s.f = [&](double, const double* yy, double* yd) { ode_f(n, *cfg, par, rates.data(), yy, yd); };
  s.jac_col = [&](double, const double* yy, int j, double* pdj) {
    ode_jac_col(n, *cfg, par, rates.data(), yy, j, pdj);
  };
  if (cfg->jac_mode == 0)
    s.jac_csc = [&](double, const double* yy, double* pd) { ode_jac_csc(n, *cfg, par, rates.data(), yy, pd); };

  int n_record = raco_n_record(o->t0, o->t_max, o->dt_first_step, o->ratio_tstep);
  if (n_record > o->n_record) n_record = o->n_record;
  // src/chemistry.f90:428-438: timer, runtime_laststep = huge, max_time_per_step
  double coef[5];
  raco_model_runtime_coefs(n.R, NEQ, n.NNZ, coef);
  const bool budget = o->max_runtime_allowed > 0.0;
  const double max_time_per_step = 5.0 / (double)n_record * o->max_runtime_allowed;
  double time_laststep = 0.0, runtime_laststep = 1.7976931348623157e308;
  bool premature = false;
  double t = o->t0;
  double t_step = o->dt_first_step;
  double tout = t + t_step;
  touts[0] = t;
  if (record) for (int k = 0; k < NEQ; ++k) record[k] = y[k];
  int NERR = 0, nerr_c = 0, quality = 0, ISTATE = 1, n_record_real = 1;
  long NST = 0, NFE = 0, NJE = 0, NLU = 0, nrestart = 0;
  int lastNST = 0, lastNFE = 0, lastNJE = 0, lastNLU = 0;
  auto harvest = [&]() {  // counters are reset by ISTATE = 1 calls: accumulate deltas
    NST += s.NST - lastNST; NFE += s.NFE - lastNFE; NJE += s.NJE - lastNJE; NLU += s.NLU - lastNLU;
    lastNST = s.NST; lastNFE = s.NFE; lastNJE = s.NJE; lastNLU = s.NLU;
  };
  const int iH = n.special[S_HI], iE = n.special[S_E], igH = n.special[S_gH],
            igH2 = n.special[S_gH2], igH2O = n.special[S_gH2O];
  int i;
  for (i = 2; i <= n_record; ++i) {
    if (tout >= o->t_max) tout = o->t_max;
    if (ISTATE == 1) { harvest(); lastNST = lastNFE = lastNJE = lastNLU = 0; ++nrestart; }
    ISTATE = s.call(y, &t, tout, rtols, atols, 4, ISTATE, 5, o->mxstep_per_interval, o->t_max, o->t_max);
    if (ISTATE == 1) ISTATE = 2;  // TOUT == T immediate return keeps ISTATE; cannot happen here
    touts[i - 1] = t;
    if (record) for (int k = 0; k < NEQ; ++k) record[(size_t)(i - 1) * NEQ + k] = y[k];
    n_record_real = i;
    if (budget) {   // src/chemistry.f90:480-494
      harvest();
      const double time_thisstep = coef[0] * (double)NFE + coef[1] * (double)NJE + coef[2] * (double)NLU +
                                   coef[3] * (double)s.n_solve + coef[4] * (double)NST;
      const double runtime_thisstep = time_thisstep - time_laststep;
      if (runtime_thisstep > std::max(10.0 * runtime_laststep, 0.5 * o->max_runtime_allowed) ||
          time_thisstep > o->max_runtime_allowed) { premature = true; break; }   // 'Premature finish'
      if (runtime_thisstep > max_time_per_step) ISTATE = 1;
      time_laststep = time_thisstep;
      runtime_laststep = runtime_thisstep;
    }
    if (t >= o->t_max) break;
    if (ISTATE < 0) {
      NERR += 1;
      nerr_c += 1;
      // ode_solver_error_handling: -4 / -5 loosen the offending component
      if (ISTATE == -4 || ISTATE == -5) {
        int idx = s.IMXER;  // IWORK(16)
        if (idx <= N) {
          rtols[idx - 1] = std::min(rtols[idx - 1] * 10.0, 1e-3);
          atols[idx - 1] = std::min(atols[idx - 1] * 100.0, 1e-20);
        } else {
          rtols[idx - 1] = std::min(rtols[idx - 1] * 10.0, 1e-2);
          atols[idx - 1] = std::min(atols[idx - 1] * 100.0, 1.0);
        }
      }
      if (ISTATE == -7) { quality += 1024; break; }  // error_stop in the reference
      if (ISTATE == -3) { quality += 256; break; }
      if (nerr_c < 3) ISTATE = 3;
      else { ISTATE = 1; nerr_c = 0; }
    }
    auto big = [&](int idx, double lim) { return idx > 0 && std::fabs(y[idx - 1]) > lim; };
    if (std::isnan(y[NEQ - 1]) || big(igH2, 1.0) || big(igH2O, 1.0) || big(igH, 1.0) || big(iH, 2.0) ||
        big(iE, 1.0) || y[NEQ - 1] <= 0.0) {
      quality += 512;
      break;
    }
    if (i % o->steps_reset_solver == 0) ISTATE = 1;
    t_step = t_step * o->ratio_tstep;
    tout = t + t_step;
  }
  harvest();
  for (int r = n_record_real + 1; r <= o->n_record; ++r) {
    touts[r - 1] = t;
    if (record) for (int k = 0; k < NEQ; ++k) record[(size_t)(r - 1) * NEQ + k] = y[k];
  }
  if (NERR > (int)(0.1f * (float)n_record)) quality += 1;
  if (t <= 0.5 * o->t_max) quality += 2;
  *t_final = t;
  *n_record_real_out = n_record_real;
  *istate_last = ISTATE;
  *quality_out = quality;
  if (stats) {
    for (int k = 0; k < 16; ++k) stats[k] = 0.0;
    stats[0] = (double)NST; stats[1] = (double)NFE; stats[2] = (double)NJE; stats[3] = (double)NLU;
    stats[4] = s.NQU; stats[5] = (double)s.n_solve; stats[6] = NERR; stats[7] = (double)nrestart;
    stats[8] = (double)s.n_cfail; stats[9] = (double)s.n_efail; stats[10] = n_record_real;
    stats[11] = ISTATE; stats[12] = s.HU;
    stats[13] = coef[0] * (double)NFE + coef[1] * (double)NJE + coef[2] * (double)NLU +
                coef[3] * (double)s.n_solve + coef[4] * (double)NST;
    stats[14] = premature ? 1.0 : 0.0;
  }
  return 0;
}

int raco_evol_solve_batch(const raco_net* h, const raco_cfg* cfg, int ncell, const double* par,
                          const double* y0, int tol_j, double RTOL, double ATOL,
                          const raco_solve_opts* o, int nthreads, double* y_final, double* t_final,
                          int* istate, int* quality, double* stats) {
  const int NEQ = h->net.NEQ;
  std::atomic<int> next(0), err(0);
  auto work = [&]() {
    std::vector<double> y(NEQ), rt(NEQ), at(NEQ), touts(o->n_record);
    for (;;) {
      int c = next.fetch_add(1);
      if (c >= ncell) break;
      const double* p = par + (size_t)c * RACO_NPAR;
      for (int k = 0; k < NEQ; ++k) y[k] = y0[(size_t)c * NEQ + k];
      raco_set_solver_flags_alt(h, tol_j, RTOL, ATOL, p[RACO_P_ratioDust2HnucNum], rt.data(), at.data());
      int nrec, ist, q;
      double tf;
      int rc = raco_evol_solve(h, cfg, p, o, y.data(), rt.data(), at.data(), touts.data(), nullptr, &tf,
                               &nrec, &ist, &q, stats ? stats + (size_t)c * 16 : nullptr);
      if (rc != 0) err = rc;
      for (int k = 0; k < NEQ; ++k) y_final[(size_t)c * NEQ + k] = y[k];
      t_final[c] = tf; istate[c] = ist; quality[c] = q;
    }
  };
  if (nthreads <= 1) work();
  else {
    std::vector<std::thread> th;
    for (int k = 0; k < nthreads; ++k) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
  return err.load();
}

}  // extern "C"
