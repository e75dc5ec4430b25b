"""Ad-hoc: in-kernel f / J of the integrator vs the oracle at late-time states of a cell."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rac2d_b200 as rb, raco
inp = os.path.join(ROOT, "tests", "golden", "inp")
fn = os.path.join(inp, "rate06_dipole_reformated_again_withgrain.dat")
net = rb.ChemNetwork(fn); onet = raco.Network(fn)
sol = net.create_solver()
y0s = net.chem_load_initial_abundances(os.path.join(inp, "initial_condition_Garrod08_mod_waterice.dat"))
c = int(sys.argv[1])
par = rb.synth.cell_params(1, first_cell=c)
y0 = rb.synth.initial_state(y0s, par, net.index("Grain0"))
rt, at = onet.solver_flags_alt(1, 1e-4, 1e-30, par[0, 6])
rates = onet.cal_rates(par[0])
for tm in [1e-2, 1e3, 1e5]:
    o = onet.evol_solve(par[0], y0[0], rt, at, t_max=tm, want_record=False)
    y = o["y"].copy()
    fo = onet.ode_f(par[0], rates, y); jo = onet.ode_jac_csc(par[0], rates, y)
    fa = onet.ode_f_abs(par[0], rates, y); ja = onet.ode_jac_csc_abs(par[0], rates, y)
    fg, jg = sol.debug_fjac(par, y[None, :])
    N = net.N
    ef = np.abs(fg[0, :N] - fo[:N]) / np.maximum(fa[:N], 1e-300)
    m = ja > 0
    ej = np.abs(jg[0][m] - jo[m]) / ja[m]
    bad = np.nonzero(m)[0][np.argsort(-ej)[:5]]
    print("t=%.0e  f: max err/scale %.2e   J: max err/scale %.2e  (nnz %d, nonzero oracle %d, nonzero gpu %d, neg y %d)" % (
        tm, ef.max(), ej.max(), m.sum(), (jo != 0).sum(), (jg[0] != 0).sum(), (y[:N] < 0).sum()))
    for k in bad[:3]:
        print("    slot", k, "oracle", jo[k], "gpu", jg[0][k], "scale", ja[k])

    # linear solve accuracy: (I + con*J) x = f with con = -h*el0, reference = dense LU with partial
    # pivoting in extended precision
    import scipy.sparse as sp
    ia = np.asarray(onet.ia) - 1; ja = np.asarray(onet.ja) - 1
    NEQ = net.NEQ
    Jd = sp.csc_matrix((jo, ja, ia), shape=(NEQ, NEQ))[:N, :N].toarray()
    def solve_ld(A, b):
        A = A.astype(np.longdouble).copy(); b = b.astype(np.longdouble).copy(); n = len(b)
        for k in range(n):
            p = k + int(np.argmax(np.abs(A[k:, k])))
            if p != k: A[[k, p]] = A[[p, k]]; b[[k, p]] = b[[p, k]]
            l = A[k + 1:, k] / A[k, k]
            A[k + 1:, k:] -= np.outer(l, A[k, k:]); b[k + 1:] -= l * b[k]
        x = np.zeros(n, dtype=np.longdouble)
        for k in range(n - 1, -1, -1): x[k] = (b[k] - A[k, k + 1:] @ x[k + 1:]) / A[k, k]
        return x
    ew = 1.0 / (rt[:N] * np.abs(y[:N]) + at[:N])
    for h in [1e-2 * tm, tm]:
        con = -h * 0.5
        xg, _ = sol.debug_fjac(par, y[None, :], con=con)
        P = np.eye(N) + con * Jd
        xr = solve_ld(P, fo[:N]).astype(np.float64)
        x64 = np.linalg.solve(P, fo[:N])
        wr = lambda v: np.sqrt(np.mean((v * ew) ** 2))
        print("   h=%.1e: wrms(x_ref) %.3e  wrms(x_gpu - x_ref) %.3e  wrms(x_lapack64 - x_ref) %.3e   max|x_gpu-x_ref|/max|x_ref| %.2e" % (
            h, wr(xr), wr(xg[0, :N] - xr), wr(x64 - xr), np.max(np.abs(xg[0, :N] - xr)) / np.max(np.abs(xr))))
