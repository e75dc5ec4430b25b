"""Synthetic cell batches for the chemistry hot path.

Follows the reference's own single-point convention (src/test_cases_bak.f90:47-56)
and the per-cell derivations of src/vertical_structure.f90:186-225.  Pure numpy,
host only.  PRNG: numpy Philox keyed with seed 20240613; cell id = row index of a
single stream, so cell i is the same whatever ncell is.
"""
import numpy as np

NPAR = 32
PAR_NAMES = [
    "Tgas", "Tdust", "n_gas", "GrainRadius_CGS", "sigdust_ave", "ndust_tot",
    "ratioDust2HnucNum", "SitesPerGrain", "zeta_cosmicray_H2", "zeta_Xray_H2",
    "Ncol_toISM", "omega_albedo", "G0_UV_toISM", "G0_UV_toStar", "G0_UV_H2phd",
    "G0_UV_toStar_photoDesorb", "Av_toISM", "Av_toStar", "phflux_Lya",
    "fss_toISM_H2", "fss_toISM_CO", "fss_toISM_H2O", "fss_toISM_OH",
    "fss_toStar_H2", "fss_toStar_CO", "fss_toStar_H2O", "fss_toStar_OH",
]
P = {n: i for i, n in enumerate(PAR_NAMES)}
SEED = 20240613
NU = 16  # uniforms drawn per cell

phy_Pi = 3.1415926535897932384626433
phy_mProton_CGS = 1.67262158e-24
phy_colDen2Av_coeff = 5.3e-22  # src/sub_global_variables.f90:90


def _uniforms(ncell, first_cell=0, seed=SEED):
    bg = np.random.Philox(key=seed)
    # 16 doubles per cell = 4 Philox blocks of 4x64 bit; advance() counts 256-bit blocks
    bg.advance(first_cell * (NU // 4))
    return np.random.Generator(bg).random((ncell, NU))


def _logu(u, lo, hi):
    return 10.0 ** (np.log10(lo) + u * (np.log10(hi) - np.log10(lo)))


def grain_constants(a=1e-5, d2g_mass=0.01, mmw=1.4, rho=2.0):
    """GrainRadius_CGS, sigdust_ave, SitesPerGrain, ratioDust2HnucNum."""
    sig = phy_Pi * a * a
    sites = 4.0 * sig * 1e15
    D = d2g_mass * (phy_mProton_CGS * mmw) / (4.0 * phy_Pi / 3.0 * a ** 3 * rho)
    return a, sig, sites, D


def cell_params(ncell, first_cell=0, stratum=None, seed=SEED):
    """par[ncell, NPAR] for config 2/3 (stratum=None) or one stiffness stratum 1..4 of
    config 5."""
    u = _uniforms(ncell, first_cell, seed)
    par = np.zeros((ncell, NPAR))
    n_gas = _logu(u[:, 0], 1e3, 1e13)
    Tgas = _logu(u[:, 1], 8.0, 3000.0)
    Tdust = np.clip(Tgas * 10.0 ** (-0.5 * u[:, 2]), 5.0, 1500.0)
    G0s = _logu(u[:, 3], 1e-2, 1e8)
    Avs = np.where(u[:, 4] < 0.1, 0.0, 20.0 * u[:, 5])
    Avi = np.where(u[:, 6] < 0.1, 0.0, 20.0 * u[:, 7])
    if stratum == 1:    # cold dense midplane
        n_gas = _logu(u[:, 0], 1e9, 1e13); Tgas = _logu(u[:, 1], 8.0, 30.0)
        Avs = 10.0 + 10.0 * u[:, 5]; Avi = 10.0 + 10.0 * u[:, 7]
        Tdust = np.clip(Tgas * 10.0 ** (-0.5 * u[:, 2]), 5.0, 1500.0)
    elif stratum == 2:  # warm molecular
        n_gas = _logu(u[:, 0], 1e6, 1e9); Tgas = _logu(u[:, 1], 30.0, 300.0)
        Avs = 1.0 + 9.0 * u[:, 5]; Avi = 1.0 + 9.0 * u[:, 7]
        Tdust = np.clip(Tgas * 10.0 ** (-0.5 * u[:, 2]), 5.0, 1500.0)
    elif stratum == 3:  # PDR surface
        n_gas = _logu(u[:, 0], 1e3, 1e7); Tgas = _logu(u[:, 1], 100.0, 3000.0)
        Avs = u[:, 5]; Avi = u[:, 7]; G0s = _logu(u[:, 3], 1e3, 1e8)
        Tdust = np.clip(Tgas * 10.0 ** (-0.5 * u[:, 2]), 5.0, 1500.0)
    elif stratum == 4:  # ice-line band
        Tdust = 80.0 + 100.0 * u[:, 2]
        Tgas = np.clip(Tdust * 10.0 ** (0.5 * u[:, 1]), 8.0, 3000.0)
    a, sig, sites, D = grain_constants()
    par[:, P["Tgas"]] = Tgas
    par[:, P["Tdust"]] = Tdust
    par[:, P["n_gas"]] = n_gas
    par[:, P["GrainRadius_CGS"]] = a
    par[:, P["sigdust_ave"]] = sig
    par[:, P["ndust_tot"]] = n_gas * D
    par[:, P["ratioDust2HnucNum"]] = D
    par[:, P["SitesPerGrain"]] = sites
    par[:, P["zeta_cosmicray_H2"]] = 1.36e-17
    par[:, P["zeta_Xray_H2"]] = _logu(u[:, 8], 1e-19, 1e-11)
    par[:, P["Ncol_toISM"]] = Avi / phy_colDen2Av_coeff
    par[:, P["omega_albedo"]] = 0.5
    par[:, P["G0_UV_toISM"]] = 1.0
    par[:, P["G0_UV_toStar"]] = G0s
    att = np.exp(-2.6 * Avs / 1.086)
    par[:, P["G0_UV_H2phd"]] = G0s * att * 0.3
    par[:, P["G0_UV_toStar_photoDesorb"]] = G0s * att
    par[:, P["Av_toISM"]] = Avi
    par[:, P["Av_toStar"]] = Avs
    par[:, P["phflux_Lya"]] = np.where(u[:, 9] < 0.5, 0.0, _logu(u[:, 10], 1e4, 1e12))
    par[:, P["fss_toISM_H2"]] = _logu(u[:, 11], 1e-8, 1.0)
    par[:, P["fss_toISM_CO"]] = _logu(u[:, 12], 1e-8, 1.0)
    par[:, P["fss_toISM_H2O"]] = 1.0
    par[:, P["fss_toISM_OH"]] = 1.0
    par[:, P["fss_toStar_H2"]] = _logu(u[:, 13], 1e-8, 1.0)
    par[:, P["fss_toStar_CO"]] = _logu(u[:, 14], 1e-8, 1.0)
    par[:, P["fss_toStar_H2O"]] = 1.0
    par[:, P["fss_toStar_OH"]] = 1.0
    return par


def stratified_params(ncell, first_cell=0, seed=SEED):
    """Config 5: four equal stiffness strata interleaved cell by cell (cell id mod 4)."""
    par = np.zeros((ncell, NPAR))
    ids = np.arange(first_cell, first_cell + ncell)
    for s in range(4):
        m = (ids % 4) == s
        if m.any():
            full = cell_params(ncell, first_cell, stratum=s + 1, seed=seed)
            par[m] = full[m]
    return par


def initial_state(y0_species, par, i_Grain0):
    """y[ncell, NEQ]: template IC + Grain0 = ratioDust2HnucNum + T slot = Tgas
    (set_initial_condition_4solver, src/disk.f90:2057-2067). i_Grain0 is 1-based, 0 = absent."""
    ncell = par.shape[0]
    N = y0_species.shape[0]
    y = np.zeros((ncell, N + 1))
    y[:, :N] = y0_species[None, :]
    if i_Grain0 > 0:
        y[:, i_Grain0 - 1] = par[:, P["ratioDust2HnucNum"]]
    y[:, N] = par[:, P["Tgas"]]
    return y
