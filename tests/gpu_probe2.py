"""Ad-hoc GPU probe (not a pytest): timing + phase breakdown + parity of a few cells."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rac2d_b200 as rb
inp = os.path.join(ROOT, "tests", "golden", "inp")
ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 592
ncheck = int(sys.argv[2]) if len(sys.argv) > 2 else 8
netname = sys.argv[3] if len(sys.argv) > 3 else "rate06_dipole_reformated_again_withgrain.dat"
net = rb.ChemNetwork(os.path.join(inp, netname))
sol = net.create_solver()
print(sol.describe(), flush=True)
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
print("phase share:", {k: round(v / tot, 4) for k, v in ph.items() if k in ("rates","f","jac","fact_head","fact_schur","fact_tail","solve","vec")})
nlu, nsolve, nfe, nje, nst = st[:,3].sum(), st[:,5].sum(), st[:,1].sum(), st[:,2].sum(), st[:,0].sum()
print("cycles/cell %.3e ; per step %.0f" % (tot / max(ph["ncell"], 1), tot / nst))
print("per LU: total %.0f pbuild %.0f glu_loop %.0f (pivmul %.0f flat %.0f narrow %.0f wide %.0f) copy %.0f tail %.0f (inv %.0f)" % (
    (ph["fact_head"]+ph["fact_schur"]+ph["fact_tail"])/nlu, ph["pbuild"]/nlu, ph["glu_loop"]/nlu, ph["glu_pivmul"]/nlu, ph["glu_flat"]/nlu,
    ph["glu_narrow"]/nlu, ph["glu_wide"]/nlu, ph["glu_copy"]/nlu, ph["fact_tail"]/nlu, ph["tail_inv"]/nlu))
print("per solve: total %.0f fwd %.0f tail %.0f bwd %.0f spmv %.0f | per f: %.0f (flux %.0f gather %.0f) | jac %.0f | vec/step %.0f" % (
    ph["solve"]/nsolve, ph["solve_fwd"]/nsolve, ph["solve_tail"]/nsolve, ph["solve_bwd"]/nsolve, ph["solve_spmv"]/nsolve,
    ph["f"]/nfe, ph["f_flux"]/nfe, ph["f_gather"]/nfe, ph["jac"]/max(nje,1), ph["vec"]/nst))
print("tail sweeps per solve: warp0 L chain %.0f | last warp until its L chain ends %.0f | warp0 between its chains %.0f | warp0 U chain %.0f | top U chain %.0f" % (
    ph["tail_L0"]/nsolve, ph["tail_Lall"]/nsolve, ph["tail_w0mid"]/nsolve, ph["tail_U0"]/nsolve, ph["tail_Utop"]/nsolve))
print("blocked tail LU per factorisation: diagonal blocks %.0f | panel solves %.0f | trailing updates %.0f" % (
    ph["blk_diag"]/nlu, ph["blk_panel"]/nlu, ph["blk_update"]/nlu))
if ncheck:
    import raco
    onet = raco.Network(os.path.join(inp, netname))
    tol = np.array([1e-2 if nm.startswith("g") else 1e-3 for nm in net.names])
    worst = 0.0
    for c in range(ncheck):
        rt, at = onet.solver_flags_alt(1, 1e-4, 1e-30, par[c, 6])
        o = onet.evol_solve(par[c], y0[c], rt, at, want_record=False)
        m = np.abs(o["y"][:net.N]) > 1e-12
        d = float(np.max(np.abs(res["y"][c, :net.N][m] - o["y"][:net.N][m]) / (np.abs(o["y"][:net.N][m]) * tol[m])))
        worst = max(worst, d)
        print("cell", c, "diff/(10 RTOL_i |X|) = %.3e" % d, "steps gpu/oracle", int(st[c, 0]), int(o["stats"][0]), "istate", res["istate"][c], o["istate"], flush=True)
    print("worst", worst)
