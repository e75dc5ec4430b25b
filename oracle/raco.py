"""ctypes binding of the CPU oracle (oracle/libraco.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
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


class Cfg(C.Structure):
    _fields_ = [("Diff2DesorRatio", C.c_double), ("special_gH_E_diff", C.c_double),
                ("H2_form_use_moeq", C.c_int), ("use_special_gH_mobi", C.c_int),
                ("update_gH_params_realtime", C.c_int), ("jac_mode", C.c_int)]


class SolveOpts(C.Structure):
    _fields_ = [("t0", C.c_double), ("t_max", C.c_double), ("dt_first_step", C.c_double),
                ("ratio_tstep", C.c_double), ("mxstep_per_interval", C.c_int),
                ("steps_reset_solver", C.c_int), ("n_record", C.c_int),
                ("max_runtime_allowed", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "libraco.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h", ".hpp"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libraco.so"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "libraco.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.raco_net_load.restype = C.c_void_p
        L.raco_net_load.argtypes = [C.c_char_p]
        L.raco_last_error.restype = C.c_char_p
        L.raco_lsodes_create.restype = C.c_void_p
        L.raco_n_record.argtypes = [C.c_double] * 4
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def default_cfg(jac_mode=0):
    return Cfg(0.5, 225.0, 0, 0, 0, jac_mode)


class Network:
    """The reference's setup chain on one network file (chemistry.f90:1427-1454 ...)."""

    def __init__(self, path):
        L = lib()
        self.h = C.c_void_p(L.raco_net_load(path.encode()))
        if not self.h:
            raise RuntimeError(L.raco_last_error().decode())
        s = (C.c_int * 8)()
        L.raco_net_sizes(self.h, s)
        (self.R, self.N, self.NEQ, self.NNZ, self.nGrain, self.n_dupli, self.nnz_diag,
         self.nnz_ldu) = list(s)
        R, N = self.R, self.N
        self.reac = np.zeros((R, 3), np.int32)
        self.prod = np.zeros((R, 4), np.int32)
        self.n_reac = np.zeros(R, np.int32)
        self.n_prod = np.zeros(R, np.int32)
        self.itype = np.zeros(R, np.int32)
        self.ABC = np.zeros((R, 3))
        self.T_range = np.zeros((R, 2))
        ct = np.zeros(2 * R, np.uint8)
        L.raco_net_tables(self.h, _p(self.reac), _p(self.prod), _p(self.n_reac), _p(self.n_prod),
                          _p(self.itype), _p(self.ABC), _p(self.T_range), _p(ct))
        self.ctype = [bytes(ct[2 * i:2 * i + 2]).decode() for i in range(R)]
        nm = np.zeros(12 * N, np.uint8)
        self.elements = np.zeros((N, 20), np.int32)
        self.mass_num = np.zeros(N)
        self.vib_freq = np.zeros(N)
        self.Edesorb = np.zeros(N)
        self.counterpart = np.zeros(N, np.int32)
        L.raco_net_species(self.h, _p(nm), _p(self.elements), _p(self.mass_num), _p(self.vib_freq),
                           _p(self.Edesorb), _p(self.counterpart))
        self.names = [bytes(nm[12 * i:12 * i + 12]).decode().strip() for i in range(N)]
        self.dupli_ptr = np.zeros(R + 1, np.int32)
        self.dupli_list = np.zeros(max(self.n_dupli, 1), np.int32)
        L.raco_net_dupli(self.h, _p(self.dupli_ptr), _p(self.dupli_list))
        self.special = np.zeros(32, np.int32)
        L.raco_net_special(self.h, _p(self.special))
        self.grain_idx = np.zeros(max(self.nGrain, 1), np.int32)
        L.raco_net_grain_species(self.h, _p(self.grain_idx))
        self.grain_idx = self.grain_idx[:self.nGrain]
        self.ia = np.zeros(self.NEQ + 1, np.int32)
        self.ja = np.zeros(self.NNZ, np.int32)
        L.raco_net_pattern(self.h, _p(self.ia), _p(self.ja))

    def load_initial_abundances(self, path):
        y0 = np.zeros(self.N)
        rc = lib().raco_load_initial_abundances(self.h, path.encode(), _p(y0))
        if rc != 0:
            raise RuntimeError(lib().raco_last_error().decode())
        return y0

    def cal_rates(self, par, cfg=None):
        cfg = cfg or default_cfg()
        par = np.ascontiguousarray(par, np.float64)
        k = np.zeros(self.R)
        rc = lib().raco_cal_rates(self.h, C.byref(cfg), _p(par), _p(k))
        if rc != 0:
            raise RuntimeError(lib().raco_last_error().decode())
        return k

    def ode_f(self, par, rates, y, cfg=None):
        cfg = cfg or default_cfg()
        yd = np.zeros(self.NEQ)
        lib().raco_ode_f(self.h, C.byref(cfg), _p(np.ascontiguousarray(par)), _p(rates),
                         _p(np.ascontiguousarray(y)), _p(yd))
        return yd

    def ode_jac_col(self, par, rates, y, j, cfg=None):
        cfg = cfg or default_cfg()
        pd = np.zeros(self.NEQ)
        lib().raco_ode_jac_col(self.h, C.byref(cfg), _p(np.ascontiguousarray(par)), _p(rates),
                               _p(np.ascontiguousarray(y)), C.c_int(j), _p(pd))
        return pd

    def ode_jac_csc(self, par, rates, y, cfg=None):
        cfg = cfg or default_cfg()
        pd = np.zeros(self.NNZ)
        lib().raco_ode_jac_csc(self.h, C.byref(cfg), _p(np.ascontiguousarray(par)), _p(rates),
                               _p(np.ascontiguousarray(y)), _p(pd))
        return pd

    def ode_f_abs(self, par, rates, y, cfg=None):
        cfg = cfg or default_cfg()
        out = np.zeros(self.NEQ)
        lib().raco_ode_f_abs(self.h, C.byref(cfg), _p(np.ascontiguousarray(par)), _p(rates),
                             _p(np.ascontiguousarray(y)), _p(out))
        return out

    def ode_jac_csc_abs(self, par, rates, y, cfg=None):
        cfg = cfg or default_cfg()
        pd = np.zeros(self.NNZ)
        lib().raco_ode_jac_csc_abs(self.h, C.byref(cfg), _p(np.ascontiguousarray(par)), _p(rates),
                                   _p(np.ascontiguousarray(y)), _p(pd))
        return pd

    def solver_flags_alt(self, j, RTOL, ATOL, D):
        rt = np.zeros(self.NEQ)
        at = np.zeros(self.NEQ)
        lib().raco_set_solver_flags_alt(self.h, C.c_int(j), C.c_double(RTOL), C.c_double(ATOL),
                                        C.c_double(D), _p(rt), _p(at))
        return rt, at

    def evol_solve(self, par, y, rtols, atols, t0=0.0, t_max=1e6, dt_first_step=1e-8, ratio=1.1,
                   mxstep=6000, steps_reset=50, cfg=None, want_record=True, max_runtime_allowed=0.0):
        """chem_evol_solve for one cell. Returns dict(y, touts, record, t_final, ...)."""
        cfg = cfg or default_cfg()
        nrec = lib().raco_n_record(t0, t_max, dt_first_step, ratio)
        o = SolveOpts(t0, t_max, dt_first_step, ratio, mxstep, steps_reset, nrec, max_runtime_allowed)
        y = np.array(y, np.float64)
        rt = np.array(rtols, np.float64)
        at = np.array(atols, np.float64)
        touts = np.zeros(nrec)
        rec = np.zeros((nrec, self.NEQ)) if want_record else None
        tf = C.c_double()
        nr = C.c_int()
        ist = C.c_int()
        q = C.c_int()
        st = np.zeros(16)
        rc = lib().raco_evol_solve(self.h, C.byref(cfg), _p(np.ascontiguousarray(par, np.float64)),
                                   C.byref(o), _p(y), _p(rt), _p(at), _p(touts), _p(rec),
                                   C.byref(tf), C.byref(nr), C.byref(ist), C.byref(q), _p(st))
        if rc != 0:
            raise RuntimeError(lib().raco_last_error().decode())
        return dict(y=y, touts=touts, record=rec, t_final=tf.value, n_record_real=nr.value,
                    istate=ist.value, quality=q.value, stats=st, rtols=rt, atols=at)

    def evol_solve_batch(self, par, y0, tol_j=1, RTOL=1e-4, ATOL=1e-30, t0=0.0, t_max=1e6,
                         dt_first_step=1e-8, ratio=1.1, mxstep=6000, steps_reset=50, nthreads=1,
                         cfg=None, max_runtime_allowed=0.0):
        """par[ncell,NPAR], y0[ncell,NEQ] (cell-major). CPU baseline."""
        cfg = cfg or default_cfg()
        par = np.ascontiguousarray(par, np.float64)
        y0 = np.ascontiguousarray(y0, np.float64)
        ncell = par.shape[0]
        nrec = lib().raco_n_record(t0, t_max, dt_first_step, ratio)
        o = SolveOpts(t0, t_max, dt_first_step, ratio, mxstep, steps_reset, nrec, max_runtime_allowed)
        yf = np.zeros((ncell, self.NEQ))
        tf = np.zeros(ncell)
        ist = np.zeros(ncell, np.int32)
        q = np.zeros(ncell, np.int32)
        st = np.zeros((ncell, 16))
        rc = lib().raco_evol_solve_batch(self.h, C.byref(cfg), C.c_int(ncell), _p(par), _p(y0),
                                         C.c_int(tol_j), C.c_double(RTOL), C.c_double(ATOL),
                                         C.byref(o), C.c_int(nthreads), _p(yf), _p(tf), _p(ist),
                                         _p(q), _p(st))
        if rc != 0:
            raise RuntimeError(lib().raco_last_error().decode())
        return dict(y=yf, t_final=tf, istate=ist, quality=q, stats=st)
