"""CPU tests that PIN the oracle: against the reference's only in-tree known-answer data
(the DLSODES documentation example, src/opkdmain.f:1919-2133), against the parser golden
values of SURVEY App. B/E, and against scipy's independent stiff integrators."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import IC_GARROD, NET_A, NET_B, NET_C, ROOT

GOLD = os.path.join(ROOT, "tests", "golden")
RK = [0, 0.1, 10, 50, 2.5, 0.1, 10, 50, 2.5, 50, 5, 50, 50, 50, 30, 100, 2.5, 100, 2.5, 50, 50]


def fex(y):
    """FEX of the DLSODES example (src/opkdmain.f:2002-2030)."""
    Y = lambda i: y[i - 1]
    rk = RK
    yd = np.zeros(12)
    yd[0] = -rk[1] * Y(1)
    yd[1] = rk[1] * Y(1) + rk[11] * rk[14] * Y(4) + rk[19] * rk[14] * Y(5) - rk[3] * Y(2) * Y(3) - rk[15] * Y(2) * Y(12) - rk[2] * Y(2)
    yd[2] = rk[2] * Y(2) - rk[5] * Y(3) - rk[3] * Y(2) * Y(3) - rk[7] * Y(10) * Y(3) + rk[11] * rk[14] * Y(4) + rk[12] * rk[14] * Y(6)
    yd[3] = rk[3] * Y(2) * Y(3) - rk[11] * rk[14] * Y(4) - rk[4] * Y(4)
    yd[4] = rk[15] * Y(2) * Y(12) - rk[19] * rk[14] * Y(5) - rk[16] * Y(5)
    yd[5] = rk[7] * Y(10) * Y(3) - rk[12] * rk[14] * Y(6) - rk[8] * Y(6)
    yd[6] = rk[17] * Y(10) * Y(12) - rk[20] * rk[14] * Y(7) - rk[18] * Y(7)
    yd[7] = rk[9] * Y(10) - rk[13] * rk[14] * Y(8) - rk[10] * Y(8)
    yd[8] = rk[4] * Y(4) + rk[16] * Y(5) + rk[8] * Y(6) + rk[18] * Y(7)
    yd[9] = (rk[5] * Y(3) + rk[12] * rk[14] * Y(6) + rk[20] * rk[14] * Y(7) + rk[13] * rk[14] * Y(8)
             - rk[7] * Y(10) * Y(3) - rk[17] * Y(10) * Y(12) - rk[6] * Y(10) - rk[9] * Y(10))
    yd[10] = rk[10] * Y(8)
    yd[11] = rk[6] * Y(10) + rk[19] * rk[14] * Y(5) + rk[20] * rk[14] * Y(7) - rk[15] * Y(2) * Y(12) - rk[17] * Y(10) * Y(12)
    return yd


def jex(y, j):
    """JEX of the DLSODES example (src/opkdmain.f:2032-2095), column j."""
    rk = RK
    Y = lambda i: y[i - 1]
    p = np.zeros(13)
    if j == 1: p[1] = -rk[1]; p[2] = rk[1]
    elif j == 2: p[2] = -rk[3] * Y(3) - rk[15] * Y(12) - rk[2]; p[3] = rk[2] - rk[3] * Y(3); p[4] = rk[3] * Y(3); p[5] = rk[15] * Y(12); p[12] = -rk[15] * Y(12)
    elif j == 3: p[2] = -rk[3] * Y(2); p[3] = -rk[5] - rk[3] * Y(2) - rk[7] * Y(10); p[4] = rk[3] * Y(2); p[6] = rk[7] * Y(10); p[10] = rk[5] - rk[7] * Y(10)
    elif j == 4: p[2] = rk[11] * rk[14]; p[3] = rk[11] * rk[14]; p[4] = -rk[11] * rk[14] - rk[4]; p[9] = rk[4]
    elif j == 5: p[2] = rk[19] * rk[14]; p[5] = -rk[19] * rk[14] - rk[16]; p[9] = rk[16]; p[12] = rk[19] * rk[14]
    elif j == 6: p[3] = rk[12] * rk[14]; p[6] = -rk[12] * rk[14] - rk[8]; p[9] = rk[8]; p[10] = rk[12] * rk[14]
    elif j == 7: p[7] = -rk[20] * rk[14] - rk[18]; p[9] = rk[18]; p[10] = rk[20] * rk[14]; p[12] = rk[20] * rk[14]
    elif j == 8: p[8] = -rk[13] * rk[14] - rk[10]; p[10] = rk[13] * rk[14]; p[11] = rk[10]
    elif j == 10: p[3] = -rk[7] * Y(3); p[6] = rk[7] * Y(3); p[7] = rk[17] * Y(12); p[8] = rk[9]; p[10] = -rk[7] * Y(3) - rk[17] * Y(12) - rk[6] - rk[9]; p[12] = rk[6] - rk[17] * Y(12)
    elif j == 12: p[2] = -rk[15] * Y(2); p[5] = rk[15] * Y(2); p[7] = rk[17] * Y(10); p[10] = -rk[17] * Y(10); p[12] = -rk[15] * Y(2) - rk[17] * Y(10)
    return p[1:]


def test_dlsodes_documentation_example(oracle):
    """The oracle's DLSODES restatement reproduces the reference's printed table: every
    component to the printed 6 digits, the step counts and NST/NFE/NJE/NLU/NNZ/NNZLU."""
    gold = json.load(open(os.path.join(GOLD, "dlsodes_example.json")))
    L = oracle.lib()
    # MF=121: structure from JAC at a perturbed y = pattern of JEX plus the diagonal
    yp = np.full(12, 0.3)
    ia, ja = [1], []
    for j in range(1, 13):
        col = jex(yp, j)
        ja += sorted(set([j] + [i + 1 for i in range(12) if col[i] != 0]))
        ia.append(len(ja) + 1)
    assert len(ja) == gold["counters"]["nnz"]
    FCB = C.CFUNCTYPE(None, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)
    JCB = C.CFUNCTYPE(None, C.c_int, C.c_double, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_void_p)

    def f_(n, t, y, yd, ctx):
        r = fex(np.ctypeslib.as_array(y, (12,)))
        for i in range(12):
            yd[i] = r[i]

    def j_(n, t, y, j, pd, ctx):
        r = jex(np.ctypeslib.as_array(y, (12,)), j)
        for i in range(12):
            pd[i] = r[i]
    fcb, jcb = FCB(f_), JCB(j_)
    h = C.c_void_p(L.raco_lsodes_create(12, (C.c_int * 13)(*ia), (C.c_int * len(ja))(*ja), fcb, jcb, None))
    L.raco_lsodes_call.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_double, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    y = np.zeros(12); y[0] = 1.0
    t = C.c_double(0.0)
    rt = np.full(12, 1e-4); at = np.full(12, 1e-6)
    ist = 1
    st = (C.c_int * 10)()
    hu = C.c_double()
    for out in gold["outputs"]:
        ist = L.raco_lsodes_call(h, y.ctypes.data, C.byref(t), out["t"], rt.ctypes.data, at.ctypes.data, 1, ist, 0, 500, 0.0, 0.0)
        assert ist == 2
        L.raco_lsodes_stats(h, st, C.byref(hu))
        assert st[0] == out["nst"]
        assert abs(hu.value - out["hu"]) <= 5.1e-4 * out["hu"]
        for a, b in zip(y, out["y"]):
            # printed with 6 significant digits
            assert abs(a - b) <= 5.1e-6 * abs(b), (out["t"], a, b)
    c = gold["counters"]
    assert (st[0], st[1], st[2], st[7], st[6], st[8] + st[9] + 12) == (c["nst"], c["nfe"], c["nje"], c["nlu"], c["nnz"], c["nnzlu"])
    L.raco_lsodes_free(h)


@pytest.mark.parametrize("path", [NET_A, NET_B, NET_C])
def test_parser_golden_values(oracle, path):
    gold = json.load(open(os.path.join(GOLD, "network_golden.json")))[os.path.basename(path)]
    net = oracle.Network(path)
    assert (net.R, net.N, net.NEQ, net.NNZ, net.nnz_diag, net.nGrain) == \
        (gold["R"], gold["N"], gold["NEQ"], gold["NNZ"], gold["NNZ_diag"], gold["n_grain_species"])
    for name, idx in gold["species_index"].items():
        assert net.names.index(name) + 1 == idx, name
    from collections import Counter
    it = Counter(net.itype.tolist())
    for k, v in gold["itype_count"].items():
        assert it[int(k)] == v, k
    assert {str(k): v for k, v in Counter(net.n_reac.tolist()).items()} == gold["n_reac_count"]
    assert {str(k): v for k, v in Counter(net.n_prod.tolist()).items()} == gold["n_prod_count"]
    # duplicate sets: a set of size s contributes 1 + 2 + ... + (s-1) list entries
    sizes = Counter()
    members = {}
    for i in range(net.R):
        tw = net.dupli_list[net.dupli_ptr[i]:net.dupli_ptr[i + 1]]
        if len(tw):
            members[min(tw)] = max(members.get(min(tw), 0), len(tw) + 1)
    sizes = Counter(members.values())
    assert sizes.get(2, 0) == gold["dupli_sets_size2"] and sizes.get(3, 0) == gold["dupli_sets_size3"]
    # pattern invariants: sorted rows, T column full, T row at the 10 heating/cooling species
    assert net.ia[0] == 1 and net.ia[-1] == net.NNZ + 1
    for c in range(net.NEQ):
        rows = net.ja[net.ia[c] - 1:net.ia[c + 1] - 1]
        assert np.all(np.diff(rows) > 0)
    assert net.ia[net.NEQ] - net.ia[net.NEQ - 1] == net.NEQ


def test_initial_abundances(oracle):
    net = oracle.Network(NET_A)
    y0 = net.load_initial_abundances(IC_GARROD)
    # HD is in the IC file but not in the network: dropped; total H = 2*0.5 + 2*1.8e-4
    assert abs(y0[net.names.index("H2")] - 0.5 / 1.00036) < 1e-15
    assert y0[net.names.index("E-")] == 0.0
    el = net.elements.astype(float)
    assert abs(y0 @ el[:, 3] - 1.0) < 1e-14
    assert y0 @ el[:, 0] == 0.0


def test_jacobian_forms_and_finite_differences(oracle):
    """O(R) scatter Jacobian == the reference's column-by-column form (bit-exact), and both
    match finite differences of chem_ode_f on a populated state."""
    import rac2d_b200.synth as synth
    net = oracle.Network(NET_A)
    y0s = net.load_initial_abundances(IC_GARROD)
    par = synth.cell_params(2)
    rng = np.random.default_rng(0)
    for c in range(2):
        k = net.cal_rates(par[c])
        y = synth.initial_state(y0s, par, int(net.special[14]))[c]
        y[:net.N] = 10.0 ** rng.uniform(-14, -4, net.N)
        pd = net.ode_jac_csc(par[c], k, y)
        fa = net.ode_f_abs(par[c], k, y)
        f0 = net.ode_f(par[c], k, y)
        for j in list(rng.integers(1, net.N + 1, 25)) + [net.names.index(s) + 1 for s in ("H", "H2", "E-", "gH", "gH2O", "Grain0")]:
            col = net.ode_jac_col(par[c], k, y, int(j))
            sl = slice(net.ia[j - 1] - 1, net.ia[j] - 1)
            dense = np.zeros(net.NEQ)
            dense[net.ja[sl] - 1] = pd[sl]
            assert np.array_equal(dense, col)
            h = 0.25 * abs(y[j - 1])
            y2 = y.copy(); y2[j - 1] += h
            y3 = y.copy(); y3[j - 1] -= h
            fd = (net.ode_f(par[c], k, y2) - net.ode_f(par[c], k, y3)) / (2 * h)
            # f is at most quadratic in y_j except the saturating desorption terms, so the
            # central difference is exact up to rounding ~ eps * sum|terms| / h
            sat = np.zeros(net.NEQ, bool)
            tol = 1e-9 * np.abs(col).max() + 64 * 2.2e-16 * fa / h
            ok = np.abs(fd - col) <= tol
            bad = np.nonzero(~ok)[0]
            for b in bad:   # rows touched by itype 62/75 of species j: compare loosely
                assert abs(fd[b] - col[b]) <= 0.05 * abs(col[b]) + tol[b], (j, b)


def test_oracle_trajectory_vs_scipy(oracle):
    """Independent cross-check (trajectories are otherwise unpinned by the reference):
    scipy BDF at rtol 1e-9 driven by the oracle's f/J on one cell, compared at 8 output times."""
    from scipy.integrate import solve_ivp
    import rac2d_b200.synth as synth
    net = oracle.Network(NET_A)
    y0s = net.load_initial_abundances(IC_GARROD)
    par = synth.cell_params(2)
    c = 1
    y0 = synth.initial_state(y0s, par, int(net.special[14]))[c]
    k = net.cal_rates(par[c])
    rt, at = net.solver_flags_alt(1, 1e-4, 1e-30, par[c, 6])
    r = net.evol_solve(par[c], y0, rt, at)
    assert r["quality"] == 0 and r["istate"] == 2 and r["t_final"] == 1e6 and r["n_record_real"] == 316
    NEQ = net.NEQ
    rows = net.ja - 1
    cols = np.repeat(np.arange(NEQ), np.diff(net.ia))

    def jac(t, y):
        J = np.zeros((NEQ, NEQ))
        J[rows, cols] = net.ode_jac_csc(par[c], k, y)
        return J
    sel = r["touts"][[100, 200, 280, 315]]
    sol = solve_ivp(lambda t, y: net.ode_f(par[c], k, y), (0, 1e6), y0, method="BDF", jac=jac,
                    rtol=1e-8, atol=1e-40, t_eval=sel, first_step=1e-12)
    assert sol.status == 0
    for i, tt in enumerate(sel):
        idx = int(np.where(r["touts"] == tt)[0][0])
        a, b = r["record"][idx][:net.N], sol.y[:net.N, i]
        m = np.abs(b) > 1e-12
        assert np.max(np.abs(a[m] - b[m]) / np.abs(b[m])) < 1e-3


def test_solver_flags_and_nrecord(oracle):
    net = oracle.Network(NET_A)
    assert oracle.lib().raco_n_record(0.0, 1e6, 1e-8, 1.1) == 316
    D = 2.8e-12
    rt, at = net.solver_flags_alt(1, 1e-4, 1e-30, D)
    N = net.N
    assert rt[N] == 1e-3 and at[N] == 1e-1
    g = net.grain_idx - 1
    assert np.all(rt[g] == 1e-3) and np.all(at[g] == D * 1e-8)
    for s in ("Grain0", "Grain-", "Grain+"):
        i = net.names.index(s)
        assert rt[i] == 1e-4 and at[i] == D * 1e-6
    assert rt[net.names.index("CO")] == 1e-4 and at[net.names.index("CO")] == 1e-30
    rt2, at2 = net.solver_flags_alt(2, 1e-4, 1e-30, D)
    assert rt2[net.names.index("CH")] == 1e-4 and at2[net.names.index("CH")] == 1e-25
