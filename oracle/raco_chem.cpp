#include <algorithm>
#include <vector>
#include <cstdlib>
// raco_chem.cpp -- ORACLE (test infrastructure only, see raco.h): rate coefficients,
// RHS and Jacobian for one cell, restating src/chemistry.f90:591-966,1007-1086,
// 1532-1590 and the fixed-T branches of src/disk.f90:4569-4659, 4746-4903.
#include "raco_internal.hpp"

namespace raco {

// getStickingCoeff, src/chemistry.f90:1068-1086
static double sticking(double mass_num, double T) {
  const double beta = 2.5, S0_H = 1.0, T0_H = 0.5 * (52.0 + 25.0);
  double T0 = mass_num * T0_H;
  double r = T / T0;
  double tmp = (1.0 + r) * (1.0 + r) * std::sqrt(1.0 + r);
  return S0_H * (1.0 + beta * r) / tmp;
}

// getMobility, src/chemistry.f90:1542-1568
static double mobility(const raco_cfg& c, double vibfreq, double massnum, double Edesorb, double Tdust) {
  const double w = 1e-8;
  double m = vibfreq * std::exp(std::max(
      -Edesorb * c.Diff2DesorRatio / Tdust,
      -2.0 * w / phy_hbarPlanck_CGS *
          std::sqrt(2.0 * massnum * (phy_mProton_CGS * phy_kBoltzmann_CGS * c.Diff2DesorRatio) * Edesorb)));
  if (std::fabs(massnum - 1.0) <= 1e-4 && c.use_special_gH_mobi && !c.update_gH_params_realtime) {
    double E = c.special_gH_E_diff;
    m = vibfreq * std::exp(std::max(
        -E / Tdust,
        -2.0 * w / phy_hbarPlanck_CGS * std::sqrt(2.0 * massnum * (phy_mProton_CGS * phy_kBoltzmann_CGS * E))));
  }
  if (std::isnan(m)) m = 0.0;
  return m;
}

// getBranchingRatio, src/chemistry.f90:1571-1590
static double branching(const Net& n, int i, double Tdust) {
  if (n.itype[i] < 63) return 1.0;
  double A = n.ABC[3 * i], B = n.ABC[3 * i + 1], C = n.ABC[3 * i + 2];
  double b;
  if (C != 0.0) {
    b = A * std::exp(std::max(
            -C / Tdust,
            -2.0 * B * 1e-8 / phy_hbarPlanck_CGS *
                std::sqrt(2.0 * n.T_range[2 * i] * phy_mProton_CGS * phy_kBoltzmann_CGS * C)));
  } else {
    b = A;
  }
  if (std::isnan(b)) b = 0.0;
  return b;
}

static double fss(const Net& n, int i, const double* p, bool toStar) {
  int k = n.fss_kind[i];
  if (k == 0) return 1.0;
  return p[(toStar ? RACO_P_fss_toStar_H2 : RACO_P_fss_toISM_H2) + (k - 1)];
}

// chem_cal_rates, src/chemistry.f90:591-966
int cal_rates(const Net& n, const raco_cfg& c, const double* p, double* rates) {
  const double Tgas = p[RACO_P_Tgas], Tdust = p[RACO_P_Tdust];
  const double T300 = Tgas / 300.0;
  const double TemperatureReduced = phy_kBoltzmann_SI * Tgas /
      (phy_elementaryCharge_SI * phy_elementaryCharge_SI * phy_CoulombConst_SI /
       (p[RACO_P_GrainRadius_CGS] * 1e-2));
  double JNegaPosi, JChargeNeut;
  if (TemperatureReduced > 0.0) {
    JNegaPosi = (1.0 + 1.0 / TemperatureReduced) * (1.0 + std::sqrt(2.0 / (2.0 + TemperatureReduced)));
    JChargeNeut = 1.0 + std::sqrt(phy_Pi / 2.0 / TemperatureReduced);
  } else {
    JNegaPosi = 0.0; JChargeNeut = 0.0;
  }
  const double sig_dust = p[RACO_P_sigdust_ave];  // evol_dust_size = .false.
  const double cosmicray_rela = p[RACO_P_zeta_cosmicray_H2] / const_cosmicRay_intensity_0 *
                                std::exp(-p[RACO_P_Ncol_toISM] / const_cosmicray_attenuate_N);
  const double Xray_rela = p[RACO_P_zeta_Xray_H2] / const_cosmicRay_intensity_0;
  const double f_H2_cov_modi = 1.0;
  const double D = p[RACO_P_ratioDust2HnucNum], S = p[RACO_P_SitesPerGrain];
  std::vector<double> adsorb_coeff(n.N, std::nan("")), desorb_coeff(n.N, std::nan(""));

  for (int i = 0; i < n.R; ++i) {
    const double A = n.ABC[3 * i], B = n.ABC[3 * i + 1], C = n.ABC[3 * i + 2];
    const double Tlo = n.T_range[2 * i], Thi = n.T_range[2 * i + 1];
    double k = 0.0;
    switch (n.itype[i]) {
      case 5:
        if (Tgas <= 0.0) k = 0.0;
        else if (C < 0.0) {
          if (Tlo > Tgas) k = A * std::pow(Tlo / 300.0, B) * std::exp(-C / Tlo);
          else if (Thi < Tgas) k = A * std::pow(Thi / 300.0, B) * std::exp(-C / Thi);
          else k = A * std::pow(T300, B) * std::exp(-C / Tgas);
        } else {
          k = A * std::pow(T300, B) * std::exp(-C / Tgas);
        }
        break;
      case 6:
        if (Tlo > Tgas) k = 0.0;
        else if (Thi < Tgas) k = 0.0;
        else k = A * std::pow(T300, B) * std::exp(-C / Tgas);
        break;
      case 1:
        k = A * (cosmicray_rela + Xray_rela);
        break;
      case 2: case 20:
        k = A * (C / (1.0 - p[RACO_P_omega_albedo]) * cosmicray_rela + Xray_rela);
        break;
      case 3:
        if (!n.first_is_H2[i]) {
          k = A * (p[RACO_P_G0_UV_toISM] * std::exp(-C * p[RACO_P_Av_toISM]) * fss(n, i, p, false) +
                   p[RACO_P_G0_UV_toStar] * std::exp(-C * p[RACO_P_Av_toStar]) * fss(n, i, p, true));
        } else {
          k = A * (p[RACO_P_G0_UV_toISM] * std::exp(-C * p[RACO_P_Av_toISM]) * fss(n, i, p, false) +
                   p[RACO_P_G0_UV_H2phd] * fss(n, i, p, true));
        }
        break;
      case 21:
        if (Tgas <= 0.0) k = 0.0;
        else {
          int id1 = n.reac[3 * i], id2 = n.reac[3 * i + 1], id3;
          if (id1 < 1 || id2 < 1) { set_error("type 21 needs two reactants"); return -1; }
          if (n.elements[RACO_NELEM * (id1 - 1) + 2] == 0) id3 = id1;
          else if (n.elements[RACO_NELEM * (id2 - 1) + 2] == 0) id3 = id2;
          else { set_error("Species name problem with type 21."); return -1; }
          int charge3 = n.elements[RACO_NELEM * (id1 - 1)] * n.elements[RACO_NELEM * (id2 - 1)];
          double m = n.mass_num[id3 - 1] * phy_mProton_CGS;
          if (charge3 == -1) k = std::sqrt(8.0 * phy_kBoltzmann_CGS / phy_Pi * Tgas / m) * sig_dust * JNegaPosi;
          else if (charge3 == 0) k = std::sqrt(8.0 * phy_kBoltzmann_CGS / phy_Pi * Tgas / m) * sig_dust * JChargeNeut;
          else { set_error("Charge problem with type 21."); return -1; }
          if (sig_dust <= 1e-30) k = 0.0;
        }
        break;
      case 13:
        k = p[RACO_P_phflux_Lya] * A * fss(n, i, p, true);
        break;
      case 0:
        if (Tgas <= 0.0) k = 0.0;
        else {
          double st = sticking(n.mass_num[n.reac[3 * i] - 1], Tgas);
          double tmp = std::sqrt(8.0 / phy_Pi * phy_kBoltzmann_CGS * Tgas / phy_mProton_CGS);
          k = 0.5 * st * sig_dust * tmp * D;
          if (sig_dust <= 1e-30) k = 0.0;
        }
        break;
      case 61:
        if (Tgas <= 0.0) k = 0.0;
        else {
          int r1 = n.reac[3 * i];
          double st = sticking(n.mass_num[r1 - 1], Tgas);
          double m = n.mass_num[r1 - 1] * phy_mProton_CGS;
          k = st * A * sig_dust * p[RACO_P_ndust_tot] *
              std::sqrt(8.0 / phy_Pi * phy_kBoltzmann_CGS * Tgas / m);
          if (sig_dust <= 1e-30) k = 0.0;
        }
        adsorb_coeff[n.reac[3 * i] - 1] = k;
        break;
      case 62: {
        int r1 = n.reac[3 * i];
        double Edesorb_eff = C * f_H2_cov_modi;
        k = n.vib_freq[r1 - 1] * (std::exp(-Edesorb_eff / Tdust) +
                                  CosmicDesorpPreFactor * cosmicray_rela * std::exp(-Edesorb_eff / CosmicDesorpGrainT));
        if (sig_dust <= 1e-30) k = 0.0;
        desorb_coeff[r1 - 1] = k;
        k = k * (S * D);
        break;
      }
      case 63: {
        int i1 = n.reac[3 * i];
        double tmp = mobility(c, n.vib_freq[i1 - 1], n.mass_num[i1 - 1], n.Edesorb[i1 - 1] * f_H2_cov_modi, Tdust) / S;
        double br = branching(n, i, Tdust);
        if (n.first_is_gH[i]) {
          if (c.H2_form_use_moeq) {
            int ig = n.counterpart[i1 - 1];
            k = tmp / (tmp + desorb_coeff[i1 - 1]) * adsorb_coeff[ig - 1] / D;
          } else {
            k = tmp / D * br;
          }
          if (sig_dust <= 1e-30) k = 0.0;
        } else {
          k = tmp / D * br;
        }
        break;
      }
      case 64: {
        int i1 = n.reac[3 * i], i2 = n.reac[3 * i + 1];
        double br = branching(n, i, Tdust);
        k = (mobility(c, n.vib_freq[i1 - 1], n.mass_num[i1 - 1], n.Edesorb[i1 - 1] * f_H2_cov_modi, Tdust) +
             mobility(c, n.vib_freq[i2 - 1], n.mass_num[i2 - 1], n.Edesorb[i2 - 1] * f_H2_cov_modi, Tdust)) /
            (S * D) * br;
        if (sig_dust <= 1e-30) k = 0.0;
        break;
      }
      case 75: {
        double photoyield = A + B * Tdust;
        k = (p[RACO_P_G0_UV_toStar_photoDesorb] * phy_Habing_photon_flux_CGS +
             p[RACO_P_G0_UV_toISM] * phy_Habing_photon_flux_CGS * std::exp(-phy_UVext2Av * p[RACO_P_Av_toISM])) *
            sig_dust * D * photoyield;
        if (sig_dust <= 1e-30) k = 0.0;
        break;
      }
      default:
        k = 0.0;
    }
    k = k * phy_SecondsPerYear;
    if (n.n_reac[i] == 2 && n.itype[i] < 60) k = k * p[RACO_P_n_gas];
    rates[i] = k;
    // duplicate-set resolution, src/chemistry.f90:948-964 (order dependent)
    for (int kk1 : n.dupli[i]) {
      int kk = kk1 - 1;
      double v[4] = {std::fabs(n.T_range[2 * kk] - Tgas), std::fabs(n.T_range[2 * kk + 1] - Tgas),
                     std::fabs(Tlo - Tgas), std::fabs(Thi - Tgas)};
      int i1 = 0;  // MINLOC: first minimum
      for (int q = 1; q < 4; ++q) if (v[q] < v[i1]) i1 = q;
      if (i1 == 0 || i1 == 1) { rates[i] = 0.0; break; }
      rates[kk] = 0.0;
    }
  }
  return 0;
}

// flux of reaction i and its classification; returns false when the reaction is
// skipped (case default: cycle)
static inline bool flux(const Net& n, const raco_cfg& c, const double* p, const double* rates,
                        const double* y, int i, double& rtmp) {
  const int r1 = n.reac[3 * i], r2 = n.reac[3 * i + 1];
  switch (n.itype[i]) {
    case 5: case 6: case 21: case 64:
      rtmp = rates[i] * y[r1 - 1] * y[r2 - 1];
      if (y[r1 - 1] < 0.0 && y[r2 - 1] < 0.0) rtmp = -rtmp;
      return true;
    case 1: case 2: case 3: case 13: case 61: case 20: case 0:
      rtmp = rates[i] * y[r1 - 1];
      return true;
    case 62: case 75: {
      double tmp1 = p[RACO_P_ratioDust2HnucNum] * p[RACO_P_SitesPerGrain];
      if (n.itype[i] == 75) tmp1 *= n.ABC[3 * i + 2];
      if (tmp1 <= 0.0) rtmp = rates[i];
      else {
        double tmp = y[r1 - 1] / tmp1;
        rtmp = (tmp <= 1e-4) ? rates[i] * tmp : rates[i] * (1.0 - std::exp(-tmp));
      }
      return true;
    }
    case 63:
      rtmp = rates[i] * y[r1 - 1] * y[r1 - 1];
      if (y[r1 - 1] < 0.0) rtmp = -rtmp;
      return true;
    default:
      return false;
  }
}

// chem_ode_f (evolT = .false.), src/disk.f90:4569-4659
void ode_f(const Net& n, const raco_cfg& c, const double* p, const double* rates, const double* y,
           double* ydot) {
  for (int i = 0; i < n.NEQ; ++i) ydot[i] = 0.0;
  static const int fsum_rev = getenv("ORACLE_FSUM_REVERSE") ? 1 : 0;   // diagnostics: summation order
  static const int fsum_gpu = getenv("ORACLE_FSUM_GPU") ? atoi(getenv("ORACLE_FSUM_GPU")) : 0;
  std::vector<std::vector<double>> terms;
  if (fsum_gpu) terms.resize(n.NEQ);
  for (int ii = 0; ii < n.R; ++ii) {
    const int i = fsum_rev ? n.R - 1 - ii : ii;
    double rtmp;
    if (n.itype[i] == 63 && n.first_is_gH[i] && c.H2_form_use_moeq) {
      int r1 = n.reac[3 * i];
      int i1 = n.counterpart[r1 - 1];
      rtmp = rates[i] * y[i1 - 1] * y[r1 - 1];
      ydot[i1 - 1] -= rtmp;
      ydot[r1 - 1] += rtmp;
      if (y[r1 - 1] < 0.0) rtmp = -rtmp;
    } else if (!flux(n, c, p, rates, y, i, rtmp)) {
      continue;
    }
    if (fsum_gpu) {
      // diagnostics: collect the terms per species (net stoichiometry per reaction, as the GPU does)
      int sp[7], cf[7], m = 0;
      auto add = [&](int s_, int c_) { for (int q = 0; q < m; ++q) if (sp[q] == s_) { cf[q] += c_; return; } sp[m] = s_; cf[m] = c_; ++m; };
      for (int j = 0; j < n.n_reac[i]; ++j) add(n.reac[3 * i + j] - 1, -1);
      for (int j = 0; j < n.n_prod[i]; ++j) add(n.prod[4 * i + j] - 1, +1);
      for (int q = 0; q < m; ++q) if (cf[q] != 0) terms[sp[q]].push_back((double)cf[q] * rtmp);
      continue;
    }
    for (int j = 0; j < n.n_reac[i]; ++j) ydot[n.reac[3 * i + j] - 1] -= rtmp;
    for (int j = 0; j < n.n_prod[i]; ++j) ydot[n.prod[4 * i + j] - 1] += rtmp;
  }
  if (fsum_gpu) {
    for (int s_ = 0; s_ < n.NEQ; ++s_) {
      const std::vector<double>& t = terms[s_];
      double tot = 0.0;
      for (size_t c0 = 0; c0 < t.size(); c0 += 32) {
        double acc[4] = {0, 0, 0, 0};
        const size_t c1 = std::min(t.size(), c0 + 32);
        if (fsum_gpu == 1) { for (size_t q = c0; q < c1; ++q) acc[(q - c0) & 3] += t[q]; }
        else { for (size_t q = c0; q < c1; ++q) acc[0] += t[q]; }
        tot += (acc[0] + acc[1]) + (acc[2] + acc[3]);
      }
      ydot[s_] = tot;
    }
  }
  ydot[n.NEQ - 1] = 0.0;
}

// d(flux_i)/d(y_j) for column j (1-based); returns false when reaction skipped
static inline bool dflux(const Net& n, const raco_cfg& c, const double* p, const double* rates,
                         const double* y, int i, int j, double& rtmp) {
  const int r1 = n.reac[3 * i], r2 = n.reac[3 * i + 1];
  switch (n.itype[i]) {
    case 5: case 6: case 21: case 64:
      if (j == r1) rtmp = (r1 != r2) ? rates[i] * y[r2 - 1] : 2.0 * rates[i] * y[r2 - 1];
      else if (j == r2) rtmp = (r1 != r2) ? rates[i] * y[r1 - 1] : 2.0 * rates[i] * y[r1 - 1];
      else rtmp = 0.0;
      if (y[r1 - 1] < 0.0 && y[r2 - 1] < 0.0) rtmp = -rtmp;
      return true;
    case 1: case 2: case 3: case 13: case 61: case 20: case 0:
      rtmp = (j != r1) ? 0.0 : rates[i];
      return true;
    case 62: case 75:
      if (j != r1) rtmp = 0.0;
      else {
        double tmp2 = p[RACO_P_ratioDust2HnucNum] * p[RACO_P_SitesPerGrain];
        if (n.itype[i] == 75) tmp2 *= n.ABC[3 * i + 2];
        if (tmp2 <= 0.0) rtmp = 0.0;
        else {
          double tmp1 = 1.0 / tmp2;
          double tmp = y[r1 - 1] * tmp1;
          rtmp = (tmp <= 1e-4) ? rates[i] * tmp1 : rates[i] * tmp1 * std::exp(-tmp);
        }
      }
      return true;
    case 63:
      rtmp = (j == r1) ? 2.0 * rates[i] * y[r1 - 1] : 0.0;
      if (y[r1 - 1] < 0.0) rtmp = -rtmp;
      return true;
    default:
      return false;
  }
}

// chem_ode_jac (evolT = .false.), src/disk.f90:4746-4903: one column, FULL reaction scan
void ode_jac_col(const Net& n, const raco_cfg& c, const double* p, const double* rates,
                 const double* y, int j, double* pdj) {
  for (int i = 0; i < n.NEQ; ++i) pdj[i] = 0.0;
  for (int i = 0; i < n.R; ++i) {
    double rtmp;
    if (n.itype[i] == 63 && n.first_is_gH[i] && c.H2_form_use_moeq) {
      int r1 = n.reac[3 * i];
      int i1 = n.counterpart[r1 - 1];
      if (j == r1) { rtmp = rates[i] * y[i1 - 1]; pdj[i1 - 1] -= rtmp; pdj[r1 - 1] += rtmp; }
      else if (j == i1) { rtmp = rates[i] * y[r1 - 1]; pdj[i1 - 1] -= rtmp; pdj[r1 - 1] += rtmp; }
      else rtmp = 0.0;
      if (y[r1 - 1] < 0.0) rtmp = -rtmp;
    } else if (!dflux(n, c, p, rates, y, i, j, rtmp)) {
      continue;
    }
    if (rtmp != 0.0) {
      for (int k = 0; k < n.n_reac[i]; ++k) pdj[n.reac[3 * i + k] - 1] -= rtmp;
      for (int k = 0; k < n.n_prod[i]; ++k) pdj[n.prod[4 * i + k] - 1] += rtmp;
    }
  }
  pdj[n.NEQ - 1] = 0.0;
}

// The same matrix built with one O(R) sweep: for each reaction only the columns
// of its own reactants can be non-zero.  Accumulation order inside a
// slot is by ascending reaction index, exactly as in the column form.
void ode_jac_csc(const Net& n, const raco_cfg& c, const double* p, const double* rates,
                 const double* y, double* pd) {
  for (int k = 0; k < n.NNZ; ++k) pd[k] = 0.0;
  const int NEQ = n.NEQ;
  for (int i = 0; i < n.R; ++i) {
    if (n.itype[i] == 63 && n.first_is_gH[i] && c.H2_form_use_moeq) {
      // rare branch: fall back to the column form for the two columns involved
      continue;
    }
    int cols[2] = {n.reac[3 * i], n.n_reac[i] >= 2 ? n.reac[3 * i + 1] : 0};
    int ncol = (cols[1] > 0 && cols[1] != cols[0]) ? 2 : 1;
    for (int q = 0; q < ncol; ++q) {
      int j = cols[q];
      if (j <= 0) continue;
      double rtmp;
      if (!dflux(n, c, p, rates, y, i, j, rtmp)) break;
      if (rtmp == 0.0) continue;
      for (int k = 0; k < n.n_reac[i]; ++k)
        pd[n.slot_of[(size_t)(n.reac[3 * i + k] - 1) + (size_t)(j - 1) * NEQ]] -= rtmp;
      for (int k = 0; k < n.n_prod[i]; ++k)
        pd[n.slot_of[(size_t)(n.prod[4 * i + k] - 1) + (size_t)(j - 1) * NEQ]] += rtmp;
    }
  }
  if (c.H2_form_use_moeq) {
    // patch the columns touched by the moment-equation branch with the column form
    std::vector<double> pdj(NEQ);
    std::vector<char> done(NEQ + 1, 0);
    for (int i = 0; i < n.R; ++i) {
      if (!(n.itype[i] == 63 && n.first_is_gH[i])) continue;
      int r1 = n.reac[3 * i];
      int cc[2] = {r1, n.counterpart[r1 - 1]};
      for (int j : cc) {
        if (j <= 0 || done[j]) continue;
        done[j] = 1;
        ode_jac_col(n, c, p, rates, y, j, pdj.data());
        for (int k = n.ia[j - 1] - 1; k < n.ia[j] - 1; ++k) pd[k] = pdj[n.ja[k] - 1];
      }
    }
  }
}

}  // namespace raco

using namespace raco;
extern "C" {

int raco_cal_rates(const raco_net* h, const raco_cfg* c, const double* par, double* rates) {
  return cal_rates(h->net, *c, par, rates);
}
void raco_ode_f(const raco_net* h, const raco_cfg* c, const double* par, const double* rates,
                const double* y, double* ydot) {
  ode_f(h->net, *c, par, rates, y, ydot);
}
void raco_ode_jac_col(const raco_net* h, const raco_cfg* c, const double* par, const double* rates,
                      const double* y, int j, double* pdj) {
  ode_jac_col(h->net, *c, par, rates, y, j, pdj);
}
void raco_ode_jac_csc(const raco_net* h, const raco_cfg* c, const double* par, const double* rates,
                      const double* y, double* pd) {
  ode_jac_csc(h->net, *c, par, rates, y, pd);
}

// Rounding scales for the parity tests: the same sums with every term taken in
// absolute value (sum_r |flux_r| per participating species / per Jacobian slot).
void raco_ode_f_abs(const raco_net* h, const raco_cfg* c, const double* par, const double* rates,
                    const double* y, double* out) {
  const Net& n = h->net;
  for (int i = 0; i < n.NEQ; ++i) out[i] = 0.0;
  for (int i = 0; i < n.R; ++i) {
    double r;
    if (!flux(n, *c, par, rates, y, i, r)) continue;
    r = std::fabs(r);
    for (int j = 0; j < n.n_reac[i]; ++j) out[n.reac[3 * i + j] - 1] += r;
    for (int j = 0; j < n.n_prod[i]; ++j) out[n.prod[4 * i + j] - 1] += r;
  }
}
void raco_ode_jac_csc_abs(const raco_net* h, const raco_cfg* c, const double* par, const double* rates,
                          const double* y, double* pd) {
  const Net& n = h->net;
  const int NEQ = n.NEQ;
  for (int k = 0; k < n.NNZ; ++k) pd[k] = 0.0;
  for (int i = 0; i < n.R; ++i) {
    int cols[2] = {n.reac[3 * i], n.n_reac[i] >= 2 ? n.reac[3 * i + 1] : 0};
    int ncol = (cols[1] > 0 && cols[1] != cols[0]) ? 2 : 1;
    for (int q = 0; q < ncol; ++q) {
      int j = cols[q];
      double r;
      if (!dflux(n, *c, par, rates, y, i, j, r)) break;
      r = std::fabs(r);
      for (int k = 0; k < n.n_reac[i]; ++k) pd[n.slot_of[(size_t)(n.reac[3 * i + k] - 1) + (size_t)(j - 1) * NEQ]] += r;
      for (int k = 0; k < n.n_prod[i]; ++k) pd[n.slot_of[(size_t)(n.prod[4 * i + k] - 1) + (size_t)(j - 1) * NEQ]] += r;
    }
  }
}

// chem_set_solver_flags_alt(j), src/chemistry.f90:205-268
void raco_set_solver_flags_alt(const raco_net* h, int j, double RTOL, double ATOL, double D,
                               double* rt, double* at) {
  const Net& n = h->net;
  const int N = n.N, NEQ = n.NEQ;
  double r, a, rT, aT;
  switch (j) {
    case 1: r = RTOL; a = ATOL; rT = 1e-3; aT = 1e-1; break;
    case 2: r = std::min(RTOL * 1e1, 1e-4); a = std::min(ATOL * 1e5, 1e-25); rT = 1e-2; aT = 1e-1; break;
    case 3: r = std::min(RTOL * 1e2, 1e-4); a = std::min(ATOL * 1e10, 1e-20); rT = 1e-3; aT = 1.0; break;
    case 4: r = std::min(RTOL * 1e2, 1e-4); a = std::min(ATOL * 1e10, 1e-18); rT = 1e-3; aT = 1.0; break;
    default:
      r = std::min(RTOL * std::pow(2.0, j), 1e-3); a = std::min(ATOL * std::pow(1e2, j), 1e-15);
      rT = 1e-2; aT = 1.0;
  }
  for (int i = 0; i < NEQ; ++i) { rt[i] = r; at[i] = a; }
  rt[N] = rT; at[N] = aT;
  for (int i = 0; i < 10; ++i) {
    int s = n.special[i];
    if (s > 0) { rt[s - 1] = std::max(RTOL, 1e-4); at[s - 1] = std::max(ATOL, 1e-30); }
  }
  if (n.special[S_Grain0] > 0) {
    for (int s : {n.special[S_Grain0], n.special[S_GrainM], n.special[S_GrainP]}) {
      if (s > 0) { rt[s - 1] = 1e-4; at[s - 1] = std::max(D * 1e-6, 1e-30); }
    }
  }
  for (int s : n.grain_idx) { rt[s - 1] = std::max(RTOL, 1e-3); at[s - 1] = std::max(ATOL, D * 1e-8); }
}

}  // extern "C"
