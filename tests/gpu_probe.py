"""Ad-hoc GPU probe (not a pytest): timing + phase breakdown of the integrator."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rac2d_b200 as rb
inp = os.path.join(ROOT, "tests", "golden", "inp")
ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 296
netname = sys.argv[2] if len(sys.argv) > 2 else "rate06_dipole_reformated_again_withgrain.dat"
net = rb.ChemNetwork(os.path.join(inp, netname))
sol = net.create_solver()
print("sizes", sol.R, sol.N, sol.NNZ, "nnzLU", sol.nnz_lu, "tail", sol.ntail, "levels", sol.nlevels, flush=True)
y0s = net.chem_load_initial_abundances(os.path.join(inp, "initial_condition_Garrod08_mod_waterice.dat"))
par = rb.synth.cell_params(ncell)
y0 = rb.synth.initial_state(y0s, par, net.index("Grain0"))
for rep in range(2):
    t = time.time()
    res = sol.chem_evol_solve(par, y0, want_touts=False)
    dt = time.time() - t
    st = res["stats"]
    print(f"rep {rep}: {ncell} cells in {dt:.3f}s -> {ncell/dt:.1f} cells/s; istate", np.unique(res["istate"], return_counts=True),
          "quality", np.unique(res["quality"], return_counts=True), flush=True)
print("mean NST %.0f NFE %.0f NJE %.0f NLU %.0f nsolve %.0f ; max NST %.0f" % (st[:,0].mean(), st[:,1].mean(), st[:,2].mean(), st[:,3].mean(), st[:,5].mean(), st[:,0].max()))
ph = sol.phase_cycles()
tot = ph["total"]
print("phase share of CTA cycles:", {k: round(v / tot, 4) for k, v in ph.items() if k in ("rates","f","jac","fact_head","fact_schur","fact_tail","solve","vec","glu_loop")})
print("cycles per cell (sum over CTAs / ncell): %.3e" % (tot / max(ph["ncell"], 1)))
nlu, nsolve, nfe, nje, nst = st[:,3].sum(), st[:,5].sum(), st[:,1].sum(), st[:,2].sum(), st[:,0].sum()
print("pbuild/LU %.0f tail_inv/LU %.0f ; solve fwd %.0f tail %.0f bwd %.0f" % (ph["pbuild"]/nlu, ph["tail_inv"]/nlu, ph["solve_fwd"]/nsolve, ph["solve_tail"]/nsolve, ph["solve_bwd"]/nsolve))
print("per LU: glu_pivmul %.0f flat %.0f narrow %.0f wide %.0f copy %.0f | per solve: spmv %.0f | per f: flux %.0f gather %.0f" % (
    ph["glu_pivmul"]/nlu, ph["glu_flat"]/nlu, ph["glu_narrow"]/nlu, ph["glu_wide"]/nlu, ph["glu_copy"]/nlu, ph["solve_spmv"]/nsolve, ph["f_flux"]/nfe, ph["f_gather"]/nfe))
print("glu loop total per LU (io slot): %.0f" % (ph["glu_loop"]/nlu))
print("cycles per op: LU %.0f (head %.0f schur %.0f tail %.0f) solve %.0f f %.0f jac %.0f vec/step %.0f" % (
    (ph["fact_head"]+ph["fact_schur"]+ph["fact_tail"])/nlu, ph["fact_head"]/nlu, ph["fact_schur"]/nlu, ph["fact_tail"]/nlu,
    ph["solve"]/nsolve, ph["f"]/nfe, ph["jac"]/nje, ph["vec"]/nst))

if len(sys.argv) > 3:   # compare the first cells with the CPU oracle
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import raco
    onet = raco.Network(os.path.join(inp, netname))
    tol = np.array([1e-2 if nm.startswith("g") else 1e-3 for nm in net.names])
    worst = 0.0
    for c in range(int(sys.argv[3])):
        rt, at = onet.solver_flags_alt(1, 1e-4, 1e-30, par[c, 6])
        o = onet.evol_solve(par[c], y0[c], rt, at, want_record=False)
        m = np.abs(o["y"][:net.N]) > 1e-12
        d = float(np.max(np.abs(res["y"][c, :net.N][m] - o["y"][:net.N][m]) / (np.abs(o["y"][:net.N][m]) * tol[m])))
        worst = max(worst, d)
        print("cell", c, "max diff/(10 RTOL_i |X|) = %.3e" % d, "steps gpu/oracle", int(st[c, 0]), int(o["stats"][0]) if "stats" in o else -1, flush=True)
    print("worst", worst)
