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


# ---------------------------------------------------------------------------
# configs[3] ("full ncol=200 Andrews-disk grid, all cells of one structure iteration batched"):
# HOST-SIDE EMULATION.  The real grid and its per-cell fields come from the Fortran host
# (grid refinement src/grid.f90, Monte-Carlo radiative transfer, heating/cooling, shielding
# pre-pass), none of which exists here.  What is kept from the reference: the
# analytic gas density (Andrews 2009 as coded in src/grid.f90:1741-1818) with the gas parameters
# of inp/template_configure.dat:134-144, the column-wise grid with ncol = 200 between rmin = 0.1
# and rmax = 200 AU starting from dr0 = 0.02 AU (template_configure.dat:4-10), the vertical
# refinement rule "density varies by at most max_ratio_to_be_uniform = 1.5 within a cell" with
# cell sizes in [0.02 AU, min(10 AU, 0.1 x distance to the star)] and the density floor 1e3 cm-3
# (template_configure.dat:14-23), column densities to the ISM (vertical) and to the star (along the
# ray), A_V = 5.3e-22 N (src/sub_global_variables.f90:90), the Draine (1996) eq. 37 H2 shielding
# (src/disk.f90:1888-1897).  Analytic stand-ins, NOT the reference's computed fields: dust and gas
# temperatures (two-layer prescription), the stellar UV / Ly-alpha / X-ray fields (inverse-square
# from TW-Hya-like values at 1 AU), CO shielding (power law in N_CO; the Visser table is not read).
phy_AU2cm = 1.49597871e13
phy_Msun_CGS = 1.98892e33


def andrews_dens(r, z, Md=2e-2, rin=0.1, rout=200.0, rc=80.0, hc=10.0, gam=1.5, psi=1.0,
                 r0_in_exp=3.5, rs_in_exp=1e2, p_in_exp=1.0, f_in_exp=1e-5, particlemass=1.4 * phy_mProton_CGS):
    """n_H [cm-3] at (r, z) [AU]; restates Andrews_dens (src/grid.f90:1741-1818), useNumDens = T."""
    r = np.asarray(r, dtype=float); z = np.asarray(z, dtype=float)
    t3 = np.exp(-(rin / rc) ** (2.0 - gam)); t4 = np.exp(-(rout / rc) ** (2.0 - gam))
    sigma_c = (2.0 - gam) * Md / (2.0 * phy_Pi * rc ** 2) / (t3 - t4)
    rrc = np.maximum(r, 1e-30) / rc
    rlog = np.log(rrc)
    t1 = np.exp(-gam * rlog); t2 = rrc * rrc * t1
    taper = np.where(r < r0_in_exp, np.exp(-(np.maximum(r0_in_exp - r, 0.0) / rs_in_exp) ** p_in_exp) * f_in_exp, 1.0)
    sigma = sigma_c * t1 * np.exp(-t2) * taper
    h = hc * np.exp(psi * rlog)
    e = 0.5 * (z / h) ** 2
    n = sigma / (np.sqrt(2.0 * phy_Pi) * h) * np.exp(-np.minimum(e, 700.0)) * phy_Msun_CGS / (phy_AU2cm ** 3 * particlemass)
    return np.where((r < rin) | (r > rout) | (e >= 700.0), 0.0, n)


def andrews_disk_cells(ncol=200, rmin=0.1, rmax=200.0, dr0=0.02, max_ratio=1.5, nmin=1e3,
                       smallest=0.02, largest=10.0, largest_frac=0.1, zmax=200.0):
    """(par[ncell, NPAR], geom[ncell, 4] = rmin, rmax, zmin, zmax in AU) of the emulated grid, cells
    ordered column by column from the inner edge outward and from the surface down to the midplane
    (the order in which the reference's do_chemical_stuff sweeps them)."""
    # geometric column widths: dr0 * q**i, sum = rmax - rmin
    lo, hi = 1.0 + 1e-9, 2.0
    for _ in range(200):
        q = 0.5 * (lo + hi)
        if dr0 * (q ** ncol - 1.0) / (q - 1.0) > rmax - rmin: hi = q
        else: lo = q
    edges = rmin + dr0 * (q ** np.arange(ncol + 1) - 1.0) / (q - 1.0)
    edges[-1] = rmax
    cells = []
    for c in range(ncol):
        r0, r1 = edges[c], edges[c + 1]
        rc_ = 0.5 * (r0 + r1)
        h = 10.0 * (rc_ / 80.0)
        zs = [0.0]
        while zs[-1] < zmax:
            z = zs[-1]
            if float(andrews_dens(rc_, z)) < nmin: break
            dz = -z + np.sqrt(z * z + 2.0 * h * h * np.log(max_ratio))
            dz = min(max(dz, smallest), largest, max(smallest, largest_frac * np.hypot(rc_, z)))
            zs.append(min(z + dz, zmax))
        col = [(r0, r1, zs[k], zs[k + 1]) for k in range(len(zs) - 1)]
        cells.extend(col[::-1])                  # top of the column first
    geom = np.array(cells)
    rcen = 0.5 * (geom[:, 0] + geom[:, 1]); zcen = 0.5 * (geom[:, 2] + geom[:, 3])
    ncell = geom.shape[0]
    n_gas = np.maximum(andrews_dens(rcen, zcen), nmin)
    # column density to the ISM: the cells above in the same column (+ half of the cell itself)
    Ncol_ism = np.zeros(ncell)
    i = 0
    while i < ncell:
        j = i
        while j < ncell and geom[j, 0] == geom[i, 0]: j += 1
        dN = n_gas[i:j] * (geom[i:j, 3] - geom[i:j, 2]) * phy_AU2cm
        Ncol_ism[i:j] = np.cumsum(dN) - 0.5 * dN
        i = j
    # column density to the star: along the straight ray from the origin, 256 log-spaced samples
    s = np.concatenate([[0.0], np.logspace(-4, 0, 256)])
    sm = 0.5 * (s[1:] + s[:-1]); ds = np.diff(s)
    dist = np.hypot(rcen, zcen)
    Ncol_star = (andrews_dens(rcen[:, None] * sm[None, :], zcen[:, None] * sm[None, :]) * ds[None, :]).sum(axis=1) * dist * phy_AU2cm
    Av_ism = phy_colDen2Av_coeff * Ncol_ism
    Av_star = phy_colDen2Av_coeff * Ncol_star
    # analytic stand-ins for the radiative-transfer / thermal-balance outputs
    T_atm = 550.0 * dist ** -0.5
    T_mid = 120.0 * rcen ** -0.5
    w = np.exp(-np.minimum(Av_star, Av_ism * 4.0))          # 1 in the irradiated surface, 0 in the shielded interior
    Tdust = np.clip(T_mid + (T_atm - T_mid) * w, 5.0, 1500.0)
    Tgas = np.clip(Tdust * (1.0 + 4.0 * np.exp(-np.minimum(Av_star, Av_ism))), 8.0, 3000.0)
    G0s = 3e6 / dist ** 2
    att = np.exp(-2.6 * Av_star / 1.086)
    def h2_shield(N_H2, dv=1e5):                            # Draine 1996 eq. 37 (src/disk.f90:1888-1897)
        x = N_H2 / 5e14; b5 = dv / 1e5; t = np.sqrt(1.0 + x)
        return np.minimum(1.0, 0.965 / (1.0 + x / b5) ** 2 + 0.035 / t * np.exp(-8.5e-4 * t))
    def co_shield(N_CO):
        return np.minimum(1.0, (np.maximum(N_CO, 1e-30) / 1e15) ** -0.75)      # stand-in for the Visser et al. table
    a, sig, sites, D = grain_constants()
    par = np.zeros((ncell, NPAR))
    par[:, P["Tgas"]] = Tgas
    par[:, P["Tdust"]] = Tdust
    par[:, P["n_gas"]] = n_gas
    par[:, P["GrainRadius_CGS"]] = a
    par[:, P["sigdust_ave"]] = sig
    par[:, P["ndust_tot"]] = n_gas * D
    par[:, P["ratioDust2HnucNum"]] = D
    par[:, P["SitesPerGrain"]] = sites
    par[:, P["zeta_cosmicray_H2"]] = 1.36e-17
    par[:, P["zeta_Xray_H2"]] = np.clip(1e-10 / dist ** 2 * np.exp(-Ncol_star / 1e23), 1e-19, 1e-11)
    par[:, P["Ncol_toISM"]] = Ncol_ism
    par[:, P["omega_albedo"]] = 0.5
    par[:, P["G0_UV_toISM"]] = 1.0
    par[:, P["G0_UV_toStar"]] = G0s
    par[:, P["G0_UV_H2phd"]] = G0s * att * 0.3
    par[:, P["G0_UV_toStar_photoDesorb"]] = G0s * att
    par[:, P["Av_toISM"]] = Av_ism
    par[:, P["Av_toStar"]] = Av_star
    par[:, P["phflux_Lya"]] = 1e14 / dist ** 2 * att
    par[:, P["fss_toISM_H2"]] = h2_shield(0.5 * Ncol_ism)
    par[:, P["fss_toStar_H2"]] = h2_shield(0.5 * Ncol_star)
    par[:, P["fss_toISM_CO"]] = co_shield(1e-4 * Ncol_ism)
    par[:, P["fss_toStar_CO"]] = co_shield(1e-4 * Ncol_star)
    for nm in ("fss_toISM_H2O", "fss_toISM_OH", "fss_toStar_H2O", "fss_toStar_OH"):
        par[:, P[nm]] = 1.0
    return par, geom
