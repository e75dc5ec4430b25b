"""Host-side mirror of the reference's chemistry interface over the C-ABI of libracg.so.

Names follow the reference (src/chemistry.f90): ``chem_read_reactions``,
``chem_load_initial_abundances``, ``chem_cal_rates``, ``chem_ode_f`` / ``chem_ode_jac``,
``chem_set_solver_flags_alt``, ``chem_evol_solve``.  Arrays are passed exactly as a Fortran
host would pass them: ``a(ncell, item)`` column-major, i.e. numpy arrays of shape
``(ncell, item)`` in Fortran order, which is the ``[item][cell]`` device layout.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NPAR = 32
NSTAT = 16
NAME_LEN = 12
NELEM = 20


class RacgError(RuntimeError):
    pass


class Cfg(C.Structure):
    _fields_ = [("Diff2DesorRatio", C.c_double), ("special_gH_E_diff", C.c_double),
                ("H2_form_use_moeq", C.c_int), ("use_special_gH_mobi", C.c_int),
                ("update_gH_params_realtime", C.c_int), ("evol_dust_size", C.c_int)] + \
               [(n, C.c_double) for n in (
                   "phy_Pi", "phy_elementaryCharge_SI", "phy_CoulombConst_SI", "phy_mProton_CGS",
                   "phy_kBoltzmann_SI", "phy_kBoltzmann_CGS", "phy_hbarPlanck_CGS",
                   "phy_SecondsPerYear", "phy_Habing_photon_flux_CGS", "phy_UVext2Av",
                   "const_cosmicray_attenuate_N", "const_cosmicRay_intensity_0",
                   "CosmicDesorpPreFactor", "CosmicDesorpGrainT")]


class SolveParams(C.Structure):
    _fields_ = [("ratio_tstep", C.c_double), ("mxstep_per_interval", C.c_int),
                ("steps_reset_solver", C.c_int), ("nrec_max", C.c_int), ("tol_policy_j", C.c_int),
                ("RTOL", C.c_double), ("ATOL", C.c_double), ("max_runtime_allowed", C.c_double)]


def lib_path():
    # RACG_LIB: A/B measurements of build variants (e.g. libracg_nt256.so) from the Python
    # harness; the C library itself reads no environment variables
    return os.environ.get("RACG_LIB") or os.path.join(_HERE, "libracg.so")


def build(force=False):
    """Compile csrc/ into libracg.so for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc")
    so = lib_path()
    deps = [os.path.join(src, f) for f in os.listdir(src)
            if f.endswith((".cu", ".cuh", ".cpp", ".hpp"))] + \
           [os.path.join(_HERE, "..", "include", "racg.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["make", "-s", "-j4", "-C", src, "../libracg.so"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = lib_path()
        if not os.path.exists(so):
            raise RacgError(f"{so} is missing: run rac2d_b200.build() (or `make -C rac-2d_b200/csrc`); "
                            "there is no CPU fallback")
        L = C.CDLL(so)
        L.racg_last_error.restype = C.c_char_p
        L.chem_read_reactions.restype = C.c_void_p
        L.chem_read_reactions.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.racg_launch_count.restype = C.c_long
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _check(rc):
    if rc != 0:
        raise RacgError(f"libracg error {rc}: {lib().racg_last_error().decode()}")


def default_cfg():
    c = Cfg()
    lib().racg_default_cfg(C.byref(c))
    return c


def _f(a, dtype=np.float64):
    """Fortran-ordered contiguous view a(ncell, item)."""
    return np.asfortranarray(a, dtype=dtype)


class ChemNetwork:
    """chem_read_reactions + chem_load_reactions + chem_parse_reactions +
    chem_get_dupli_reactions (host side; stays Fortran in production)."""

    def __init__(self, filename_chemical_network):
        L = lib()
        err = C.create_string_buffer(256)
        self.h = C.c_void_p(L.chem_read_reactions(filename_chemical_network.encode(), err, 256))
        if not self.h:
            raise RacgError(err.value.decode())
        R, N, nd = C.c_int(), C.c_int(), C.c_int()
        L.chem_host_sizes(self.h, C.byref(R), C.byref(N), C.byref(nd))
        self.R, self.N, self.NEQ = R.value, N.value, N.value + 1
        R, N = self.R, self.N
        self.reac = np.zeros((R, 3), np.int32)       # memory = reac(3,R) column-major
        self.prod = np.zeros((R, 4), np.int32)
        self.n_reac = np.zeros(R, np.int32)
        self.n_prod = np.zeros(R, np.int32)
        self.itype = np.zeros(R, np.int32)
        self.ABC = np.zeros((R, 3))
        self.T_range = np.zeros((R, 2))
        self.ctype_raw = np.zeros(2 * R, np.uint8)
        self.names_raw = np.zeros(NAME_LEN * N, np.uint8)
        self.elements = np.zeros((N, NELEM), np.int32)
        self.mass_num = np.zeros(N)
        self.vib_freq = np.zeros(N)
        self.Edesorb = np.zeros(N)
        self.dupli_ptr = np.zeros(R + 1, np.int32)
        self.dupli_list = np.zeros(max(nd.value, 1), np.int32)
        L.chem_host_tables(self.h, _p(self.reac), _p(self.prod), _p(self.n_reac), _p(self.n_prod),
                           _p(self.itype), _p(self.ABC), _p(self.T_range), _p(self.ctype_raw),
                           _p(self.names_raw), _p(self.elements), _p(self.mass_num), _p(self.vib_freq),
                           _p(self.Edesorb), _p(self.dupli_ptr), _p(self.dupli_list))
        self.names = [bytes(self.names_raw[NAME_LEN * i:NAME_LEN * (i + 1)]).decode().strip()
                      for i in range(N)]
        self.ctype = [bytes(self.ctype_raw[2 * i:2 * i + 2]).decode() for i in range(R)]

    def index(self, name):
        """1-based species index, 0 if absent (chem_idx_some_spe convention)."""
        return self.names.index(name) + 1 if name in self.names else 0

    def chem_load_initial_abundances(self, filename_initial_abundances):
        y0 = np.zeros(self.N)
        rc = lib().chem_load_initial_abundances(self.h, filename_initial_abundances.encode(), _p(y0))
        if rc != 0:
            raise RacgError(f"chem_load_initial_abundances failed ({rc})")
        return y0

    def create_solver(self, cfg=None, device=None):
        return ChemSolver(self, cfg, device)

    def __del__(self):
        try:
            if self.h:
                lib().chem_host_free(self.h)
        except Exception:
            pass


def write_chemical_data(dirname, iiter, abundances, col_den_toStar=None, col_den_toISM=None, nspecies=None):
    """chemical_data_iter_NNNN.bin of the reference (src/data_dump.f90:88-162) from batch arrays:
    abundances (ncell, >= nspecies), col_den_* (ncell, ncd)."""
    ab = _f(abundances)
    ncell = ab.shape[0]
    ns = nspecies or ab.shape[1]
    ncd = 0 if col_den_toStar is None else np.asarray(col_den_toStar).shape[1]
    cs = np.ascontiguousarray(col_den_toStar, np.float64) if ncd else None
    ci = np.ascontiguousarray(col_den_toISM, np.float64) if ncd else None
    _check(lib().racg_write_chemical_data(dirname.encode(), C.c_int(iiter), C.c_int(ncell), C.c_int(ns), _p(ab),
                                          C.c_int(ncd), _p(cs), _p(ci)))


def read_chemical_data(dirname, iiter, ncell, nspecies, ncd):
    ab = np.zeros((ncell, nspecies), order="F")
    cs = np.zeros((ncell, ncd))
    ci = np.zeros((ncell, ncd))
    _check(lib().racg_read_chemical_data(dirname.encode(), C.c_int(iiter), C.c_int(ncell), C.c_int(nspecies), _p(ab),
                                         C.c_int(ncd), _p(cs) if ncd else None, _p(ci) if ncd else None))
    return ab, cs, ci


class ChemSolver:
    """racg_handle: the GPU chemistry solver for one network."""

    def __init__(self, net, cfg=None, device=None):
        L = lib()
        self.net = net
        self.cfg = cfg or default_cfg()
        if device is not None:
            _check(L.racg_set_device(C.c_int(device)))
        self.h = C.c_void_p()
        _check(L.racg_network_create(
            C.byref(self.h), C.c_int(net.R), C.c_int(net.N), _p(net.reac), _p(net.prod), _p(net.n_reac),
            _p(net.n_prod), _p(net.itype), _p(net.ABC), _p(net.T_range), _p(net.ctype_raw),
            _p(net.names_raw), _p(net.elements), _p(net.mass_num), _p(net.vib_freq), _p(net.Edesorb),
            _p(net.dupli_ptr), _p(net.dupli_list), C.byref(self.cfg)))
        s = (C.c_int * 8)()
        _check(L.racg_network_sizes(self.h, s))
        (self.R, self.N, self.NEQ, self.NNZ, self.NNZ_diag, self.nnz_lu, self.ntail,
         self.nlevels) = list(s)

    def close(self):
        if self.h:
            lib().racg_destroy(self.h)
            self.h = C.c_void_p()

    # ---- devices and options --------------------------------------------
    def use_devices(self, devices=None):
        """Replicate the network on more GPUs of the node (None: every visible one); the
        host-pointer chem_evol_solve then shards every batch over them."""
        if devices is None:
            _check(lib().racg_use_devices(self.h, C.c_int(0), None))
        else:
            arr = (C.c_int * len(devices))(*devices)
            _check(lib().racg_use_devices(self.h, C.c_int(len(devices)), arr))
        return int(lib().racg_device_count(self.h))

    def set_option(self, name, value):
        _check(lib().racg_set_option(self.h, name.encode(), C.c_double(value)))

    def describe(self):
        buf = C.create_string_buffer(8192)
        _check(lib().racg_network_describe(self.h, buf, C.c_int(8192)))
        return buf.value.decode()

    def model_runtime_coefs(self):
        c = np.zeros(5)
        _check(lib().racg_model_runtime_coefs(self.h, _p(c)))
        return c

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup queries -------------------------------------------------
    def pattern(self):
        ia = np.zeros(self.NEQ + 1, np.int32)
        ja = np.zeros(self.NNZ, np.int32)
        _check(lib().racg_network_pattern(self.h, _p(ia), _p(ja)))
        return ia, ja

    def ordering(self):
        perm = np.zeros(self.N, np.int32)
        _check(lib().racg_network_ordering(self.h, _p(perm)))
        return perm

    # ---- host-pointer entry points (what the Fortran host calls) -------
    def chem_cal_rates(self, cellpar):
        par = _f(cellpar)
        ncell = par.shape[0]
        rates = np.zeros((ncell, self.R), order="F")
        _check(lib().racg_rates(self.h, C.c_int(ncell), _p(par), _p(rates)))
        return rates

    def chem_ode_f_jac(self, cellpar, y, rates, want_f=True, want_jac=True):
        par, y, rates = _f(cellpar), _f(y), _f(rates)
        ncell = par.shape[0]
        ydot = np.zeros((ncell, self.NEQ), order="F") if want_f else None
        pd = np.zeros((ncell, self.NNZ), order="F") if want_jac else None
        _check(lib().racg_rhs_jac(self.h, C.c_int(ncell), _p(par), _p(y), _p(rates), _p(ydot), _p(pd)))
        return ydot, pd

    def chem_set_solver_flags_alt(self, j, RTOL, ATOL, cellpar):
        par = _f(cellpar)
        ncell = par.shape[0]
        rt = np.zeros((ncell, self.NEQ), order="F")
        at = np.zeros((ncell, self.NEQ), order="F")
        _check(lib().racg_solver_flags_alt(self.h, C.c_int(j), C.c_double(RTOL), C.c_double(ATOL),
                                           C.c_int(ncell), _p(par), _p(rt), _p(at)))
        return rt, at

    @staticmethod
    def n_record(t0, t_max, dt_first_step, ratio):
        """chem_evol_solve_prepare_run_once, src/chemistry.f90:1894-1899."""
        return int(np.ceil(np.log((t_max - t0) / dt_first_step * (ratio - 1.0) + 1.0) / np.log(ratio))) + 1

    def chem_evol_solve(self, cellpar, y0, rtol=None, atol=None, t0=0.0, t_max=1e6,
                        dt_first_step=1e-8, ratio_tstep=1.1, mxstep_per_interval=6000,
                        steps_reset_solver=50, tol_policy_j=1, RTOL=1e-4, ATOL=1e-30,
                        want_record=False, want_touts=True, max_runtime_allowed=0.0, nrec_max=None):
        """The batch replacement of the loop body around `call chem_evol_solve`
        (src/disk.f90:1686).  Scalars t0/t_max/dt_first_step may be per-cell arrays."""
        par, y0 = _f(cellpar), _f(y0)
        ncell = par.shape[0]
        t0a = np.ascontiguousarray(np.broadcast_to(np.asarray(t0, np.float64), (ncell,)))
        tma = np.ascontiguousarray(np.broadcast_to(np.asarray(t_max, np.float64), (ncell,)))
        dta = np.ascontiguousarray(np.broadcast_to(np.asarray(dt_first_step, np.float64), (ncell,)))
        nrec = max([self.n_record(a, b, c, ratio_tstep) for a, b, c in
                    set(zip(t0a.tolist(), tma.tolist(), dta.tolist()))] or [2])
        if nrec_max is not None:
            nrec = nrec_max
        sp = SolveParams(ratio_tstep, mxstep_per_interval, steps_reset_solver, nrec, tol_policy_j, RTOL, ATOL,
                         max_runtime_allowed)
        rt = _f(rtol) if rtol is not None else None
        at = _f(atol) if atol is not None else None
        yf = np.zeros((ncell, self.NEQ), order="F")
        tf = np.zeros(ncell)
        touts = np.zeros((ncell, max(nrec, 1)), order="F") if want_touts else None
        rec = np.zeros((ncell, self.NEQ, max(nrec, 1)), order="F") if want_record else None
        nrr = np.zeros(ncell, np.int32)
        ist = np.zeros(ncell, np.int32)
        q = np.zeros(ncell, np.int32)
        st = np.zeros((ncell, NSTAT), order="F")
        _check(lib().racg_solve_batch(self.h, C.c_int(ncell), _p(par), _p(y0), _p(rt), _p(at), _p(t0a),
                                      _p(tma), _p(dta), C.byref(sp), _p(yf), _p(tf), _p(touts), _p(rec),
                                      _p(nrr), _p(ist), _p(q), _p(st)))
        return dict(y=yf, t_final=tf, touts=touts, record=rec, n_record_real=nrr, istate=ist,
                    quality=q, stats=st, nrec_max=nrec)

    def calc_this_cell(self, cellpar, y0, t_max=1e6, dt_first_step0=1e-8, nlocal_iter=4, ratio_tstep=1.1,
                       mxstep_per_interval=6000, steps_reset_solver=50, RTOL=1e-4, ATOL=1e-30,
                       max_runtime_allowed=0.0):
        """The batch form of calc_this_cell's local-iteration loop (src/disk.f90:1651-1791):
        tolerance ladder, continuation from t_final, rectify_abundances, last-good-record harvest."""
        par, y0 = _f(cellpar), _f(y0)
        ncell = par.shape[0]
        tma = np.ascontiguousarray(np.broadcast_to(np.asarray(t_max, np.float64), (ncell,)))
        sp = SolveParams(ratio_tstep, mxstep_per_interval, steps_reset_solver, 0, 1, RTOL, ATOL, max_runtime_allowed)
        ab = np.zeros((ncell, self.NEQ), order="F")
        tf = np.zeros(ncell)
        q = np.zeros(ncell, np.int32)
        ist = np.zeros(ncell, np.int32)
        nit = np.zeros(ncell, np.int32)
        rh2 = np.zeros(ncell)
        nmol = np.zeros(ncell)
        st = np.zeros((ncell, NSTAT), order="F")
        _check(lib().racg_calc_batch(self.h, C.c_int(ncell), _p(par), _p(y0), _p(tma), C.c_double(dt_first_step0),
                                     C.byref(sp), C.c_int(nlocal_iter), _p(ab), _p(tf), _p(q), _p(ist), _p(nit),
                                     _p(rh2), _p(nmol), _p(st)))
        return dict(abundances=ab, t_final=tf, quality=q, istate=ist, n_iter_used=nit,
                    R_H2_form_rate_coeff=rh2, n_mol_on_grain=nmol, stats=st)

    # ---- device-pointer entry points (buffers already in HBM) ----------
    def rates_dev(self, ncell, d_par, d_rates, stream=0):
        _check(lib().racg_rates_dev(self.h, C.c_int(ncell), C.c_void_p(d_par), C.c_void_p(d_rates),
                                    C.c_void_p(stream)))

    def rhs_jac_dev(self, ncell, d_par, d_y, d_rates, d_ydot, d_pd, stream=0):
        _check(lib().racg_rhs_jac_dev(self.h, C.c_int(ncell), C.c_void_p(d_par), C.c_void_p(d_y),
                                      C.c_void_p(d_rates), C.c_void_p(d_ydot or None),
                                      C.c_void_p(d_pd or None), C.c_void_p(stream)))

    def solve_batch_dev(self, ncell, sp, d_par, d_y0, d_t0, d_tmax, d_dt, d_yf, d_tf, d_nrec, d_ist,
                        d_q, d_stats, d_rtol=0, d_atol=0, d_touts=0, d_record=0, stream=0):
        v = lambda x: C.c_void_p(x or None)
        _check(lib().racg_solve_batch_dev(self.h, C.c_int(ncell), v(d_par), v(d_y0), v(d_rtol), v(d_atol),
                                          v(d_t0), v(d_tmax), v(d_dt), C.byref(sp), v(d_yf), v(d_tf),
                                          v(d_touts), v(d_record), v(d_nrec), v(d_ist), v(d_q), v(d_stats),
                                          v(stream)))

    def debug_fjac(self, cellpar, y, con=0.0):
        """f and J (as a CSC value array in the slot order of pattern()) from the integrator's
        in-kernel routines"""
        cellpar = _f(cellpar); y = _f(y)
        ncell = cellpar.shape[0]
        nstore = C.c_int(0)
        c2s = np.zeros(self.NNZ, dtype=np.int32)
        _check(lib().racg_debug_fjac(self.h, 0, None, None, None, None, _p(c2s), C.byref(nstore), C.c_double(0.0)))
        par_t = np.ascontiguousarray(cellpar.T); y_t = np.ascontiguousarray(y.T)
        f_t = np.zeros((self.NEQ, ncell)); j_t = np.zeros((nstore.value, ncell))
        _check(lib().racg_debug_fjac(self.h, ncell, _p(par_t), _p(y_t), _p(f_t), _p(j_t), _p(c2s), C.byref(nstore),
                                     C.c_double(con)))
        pd = np.zeros((ncell, self.NNZ))
        m = c2s >= 0
        pd[:, m] = j_t[c2s[m], :].T
        return f_t.T.copy(), pd

    def selfcheck(self):
        """host-side consistency check of the factorisation / solve schedules (no GPU needed)"""
        _check(lib().racg_selfcheck(self.h))

    def selfcheck_damaged(self, mode):
        """the same check on a deliberately damaged copy of the schedules: must raise"""
        _check(lib().racg_selfcheck_damaged(self.h, C.c_int(mode)))

    def launch_count(self):
        return int(lib().racg_launch_count(self.h))

    def phase_cycles(self):
        out = np.zeros(32)
        _check(lib().racg_phase_cycles(self.h, _p(out)))
        names = ["rates", "f", "jac", "fact_head", "fact_schur", "fact_tail", "solve", "vec", "glu_loop",
                 "total", "ncell", "pbuild", "tail_inv", "solve_fwd", "solve_tail", "solve_bwd",
                 "f_flux", "f_gather", "glu_pivmul", "glu_flat", "glu_narrow", "glu_wide", "solve_spmv", "glu_copy",
                 "tail_L0", "tail_Lall", "tail_w0mid", "tail_U0", "tail_Utop", "blk_diag", "blk_panel", "blk_update"]
        return {n: out[i] for i, n in enumerate(names)}
