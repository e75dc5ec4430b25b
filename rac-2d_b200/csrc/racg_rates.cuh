// racg_rates.cuh -- rate coefficient of one reaction for one cell (device).
// Restates the branches of chem_cal_rates (reference src/chemistry.f90:591-966) and
// its helpers getStickingCoeff (1068-1086), getMobility (1542-1568),
// getBranchingRatio (1571-1590), f_selfshielding_toISM/toStar (1007-1063) with the
// species-name predicates pre-resolved on the host into `rcode`.
#pragma once
#include "racg_dev.cuh"
#include "racg_host.hpp"

namespace racg {

// per-cell quantities shared by all reactions (src/chemistry.f90:603-649)
struct CellCommon {
  double Tgas, Tdust, T300, JNegaPosi, JChargeNeut, sig_dust, cr, xr, D, S, n_gas, ndust_tot;
  double omega, G0ism, G0star, G0H2phd, G0phdes, AvISM, AvStar, Lya;
  double fssISM[5], fssStar[5];
};

template <typename ParGet>
__device__ __forceinline__ void cell_common(const racg_cfg& c, ParGet par, CellCommon& cc) {
  cc.Tgas = par(RACG_P_Tgas); cc.Tdust = par(RACG_P_Tdust);
  cc.T300 = cc.Tgas / 300.0;
  double Tred = c.phy_kBoltzmann_SI * cc.Tgas /
                (c.phy_elementaryCharge_SI * c.phy_elementaryCharge_SI * c.phy_CoulombConst_SI /
                 (par(RACG_P_GrainRadius_CGS) * 1e-2));
  if (Tred > 0.0) {
    cc.JNegaPosi = (1.0 + 1.0 / Tred) * (1.0 + sqrt(2.0 / (2.0 + Tred)));
    cc.JChargeNeut = 1.0 + sqrt(c.phy_Pi / 2.0 / Tred);
  } else { cc.JNegaPosi = 0.0; cc.JChargeNeut = 0.0; }
  cc.sig_dust = par(RACG_P_sigdust_ave);
  cc.cr = par(RACG_P_zeta_cosmicray_H2) / c.const_cosmicRay_intensity_0 *
          exp(-par(RACG_P_Ncol_toISM) / c.const_cosmicray_attenuate_N);
  cc.xr = par(RACG_P_zeta_Xray_H2) / c.const_cosmicRay_intensity_0;
  cc.D = par(RACG_P_ratioDust2HnucNum); cc.S = par(RACG_P_SitesPerGrain);
  cc.n_gas = par(RACG_P_n_gas); cc.ndust_tot = par(RACG_P_ndust_tot);
  cc.omega = par(RACG_P_omega_albedo);
  cc.G0ism = par(RACG_P_G0_UV_toISM); cc.G0star = par(RACG_P_G0_UV_toStar);
  cc.G0H2phd = par(RACG_P_G0_UV_H2phd); cc.G0phdes = par(RACG_P_G0_UV_toStar_photoDesorb);
  cc.AvISM = par(RACG_P_Av_toISM); cc.AvStar = par(RACG_P_Av_toStar); cc.Lya = par(RACG_P_phflux_Lya);
  cc.fssISM[0] = 1.0; cc.fssStar[0] = 1.0;
  for (int k = 0; k < 4; ++k) {
    cc.fssISM[k + 1] = par(RACG_P_fss_toISM_H2 + k);
    cc.fssStar[k + 1] = par(RACG_P_fss_toStar_H2 + k);
  }
}

__device__ __forceinline__ double dev_sticking(double mass_num, double T) {
  const double T0 = mass_num * (0.5 * (52.0 + 25.0));
  const double r = T / T0;
  const double tmp = (1.0 + r) * (1.0 + r) * sqrt(1.0 + r);
  return 1.0 * (1.0 + 2.5 * r) / tmp;
}

__device__ __forceinline__ double dev_mobility(const racg_cfg& c, double vib, double mass, double Ed,
                                               double Tdust) {
  double m = vib * exp(fmax(-Ed * c.Diff2DesorRatio / Tdust,
                            -2.0 * 1e-8 / c.phy_hbarPlanck_CGS *
                                sqrt(2.0 * mass * (c.phy_mProton_CGS * c.phy_kBoltzmann_CGS * c.Diff2DesorRatio) * Ed)));
  if (fabs(mass - 1.0) <= 1e-4 && c.use_special_gH_mobi) {
    const double E = c.special_gH_E_diff;
    m = vib * exp(fmax(-E / Tdust, -2.0 * 1e-8 / c.phy_hbarPlanck_CGS *
                                       sqrt(2.0 * mass * (c.phy_mProton_CGS * c.phy_kBoltzmann_CGS * E))));
  }
  // NOTE: Fortran MAX/fmax differ for NaN arguments; the reference then maps NaN -> 0
  if (isnan(m)) m = 0.0;
  return m;
}

__device__ __forceinline__ double dev_branching(const racg_cfg& c, double A, double B, double C,
                                                double Tlo, double Tdust) {
  double b;
  if (C != 0.0) {
    b = A * exp(fmax(-C / Tdust, -2.0 * B * 1e-8 / c.phy_hbarPlanck_CGS *
                                     sqrt(2.0 * Tlo * c.phy_mProton_CGS * c.phy_kBoltzmann_CGS * C)));
  } else b = A;
  if (isnan(b)) b = 0.0;
  return b;
}

// rate coefficient in the reference's native unit (s^-1 based), before the conversion to yr^-1,
// the n_gas factor of two-body gas reactions and the duplicate-set resolution
__device__ __forceinline__ double rate_coeff_raw(const DevNet& net, const CellCommon& cc, int i) {
  const racg_cfg& c = net.cfg;
  const int code = net.rcode[i];
  const int cls = code & 0xff, fk = (code >> 8) & 0xf;
  const double A = net.rA[i], B = net.rB[i], C = net.rC[i], Tlo = net.rTlo[i], Thi = net.rThi[i];
  const double* X = net.rX + (size_t)6 * i;
  const double Tgas = cc.Tgas, Tdust = cc.Tdust, sig = cc.sig_dust;
  double k = 0.0;
  switch (cls) {
    case RC_ARRH:
      if (Tgas <= 0.0) k = 0.0;
      else if (C < 0.0) {
        if (Tlo > Tgas) k = A * pow(Tlo / 300.0, B) * exp(-C / Tlo);
        else if (Thi < Tgas) k = A * pow(Thi / 300.0, B) * exp(-C / Thi);
        else k = A * pow(cc.T300, B) * exp(-C / Tgas);
      } else k = A * pow(cc.T300, B) * exp(-C / Tgas);
      break;
    case RC_ARRH_STRICT:
      if (Tlo > Tgas || Thi < Tgas) k = 0.0;
      else k = A * pow(cc.T300, B) * exp(-C / Tgas);
      break;
    case RC_CR: k = A * (cc.cr + cc.xr); break;
    case RC_CRPHOT: k = A * (C / (1.0 - cc.omega) * cc.cr + cc.xr); break;
    case RC_PHOTO:
      k = A * (cc.G0ism * exp(-C * cc.AvISM) * cc.fssISM[fk] + cc.G0star * exp(-C * cc.AvStar) * cc.fssStar[fk]);
      break;
    case RC_PHOTO_H2:
      k = A * (cc.G0ism * exp(-C * cc.AvISM) * cc.fssISM[fk] + cc.G0H2phd * cc.fssStar[fk]);
      break;
    case RC_GRAIN_NP: case RC_GRAIN_N0:
      if (Tgas <= 0.0) k = 0.0;
      else {
        const double m = X[0] * c.phy_mProton_CGS;
        k = sqrt(8.0 * c.phy_kBoltzmann_CGS / c.phy_Pi * Tgas / m) * sig *
            (cls == RC_GRAIN_NP ? cc.JNegaPosi : cc.JChargeNeut);
        if (sig <= 1e-30) k = 0.0;
      }
      break;
    case RC_LYA: k = cc.Lya * A * cc.fssStar[fk]; break;
    case RC_H2FORM0:
      if (Tgas <= 0.0) k = 0.0;
      else {
        const double st = dev_sticking(X[0], Tgas);
        const double tmp = sqrt(8.0 / c.phy_Pi * c.phy_kBoltzmann_CGS * Tgas / c.phy_mProton_CGS);
        k = 0.5 * st * sig * tmp * cc.D;
        if (sig <= 1e-30) k = 0.0;
      }
      break;
    case RC_ADSORB:
      if (Tgas <= 0.0) k = 0.0;
      else {
        const double st = dev_sticking(X[0], Tgas);
        const double m = X[0] * c.phy_mProton_CGS;
        k = st * A * sig * cc.ndust_tot * sqrt(8.0 / c.phy_Pi * c.phy_kBoltzmann_CGS * Tgas / m);
        if (sig <= 1e-30) k = 0.0;
      }
      break;
    case RC_DESORB:
      k = X[0] * (exp(-C / Tdust) + c.CosmicDesorpPreFactor * cc.cr * exp(-C / c.CosmicDesorpGrainT));
      if (sig <= 1e-30) k = 0.0;
      k = k * (cc.S * cc.D);
      break;
    case RC_SURF_AA: {
      const double tmp = dev_mobility(c, X[0], X[1], X[2], Tdust) / cc.S;
      const double br = dev_branching(c, A, B, C, Tlo, Tdust);
      k = tmp / cc.D * br;
      if (((code >> 13) & 1) && sig <= 1e-30) k = 0.0;
      break;
    }
    case RC_SURF_AB: {
      const double br = dev_branching(c, A, B, C, Tlo, Tdust);
      k = (dev_mobility(c, X[0], X[1], X[2], Tdust) + dev_mobility(c, X[3], X[4], X[5], Tdust)) /
          (cc.S * cc.D) * br;
      if (sig <= 1e-30) k = 0.0;
      break;
    }
    case RC_PHOTODES: {
      const double photoyield = A + B * Tdust;
      k = (cc.G0phdes * c.phy_Habing_photon_flux_CGS +
           cc.G0ism * c.phy_Habing_photon_flux_CGS * exp(-c.phy_UVext2Av * cc.AvISM)) *
          sig * cc.D * photoyield;
      if (sig <= 1e-30) k = 0.0;
      break;
    }
    default: k = 0.0;
  }
  return k;
}

// rate coefficient in yr^-1, before duplicate-set resolution (src/chemistry.f90:936-942)
__device__ __forceinline__ double rate_coeff(const DevNet& net, const CellCommon& cc, int i) {
  double k = rate_coeff_raw(net, cc, i) * net.cfg.phy_SecondsPerYear;
  if ((net.rcode[i] >> 12) & 1) k = k * cc.n_gas;
  return k;
}

// duplicate-set resolution of reaction i = dup_reac[d] (src/chemistry.f90:948-964):
// returns through the callback which reactions are zeroed.
template <typename Zero>
__device__ __forceinline__ void resolve_dupli(const DevNet& net, double Tgas, int d, Zero zero) {
  const int i = net.dup_reac[d];
  const double Tlo = net.rTlo[i], Thi = net.rThi[i];
  for (int q = net.dup_ptr[d]; q < net.dup_ptr[d + 1]; ++q) {
    const int kk = net.dup_list[q];
    const double v0 = fabs(net.rTlo[kk] - Tgas), v1 = fabs(net.rThi[kk] - Tgas);
    const double v2 = fabs(Tlo - Tgas), v3 = fabs(Thi - Tgas);
    int i1 = 0; double vm = v0;                     // MINLOC: first minimum
    if (v1 < vm) { vm = v1; i1 = 1; }
    if (v2 < vm) { vm = v2; i1 = 2; }
    if (v3 < vm) { vm = v3; i1 = 3; }
    if (i1 <= 1) { zero(i); break; }
    zero(kk);
  }
}

}  // namespace racg
