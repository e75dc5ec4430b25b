import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rac2d_b200 as rb
inp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inp")
net = rb.ChemNetwork(os.path.join(inp, "rate06_dipole_reformated_again_withgrain.dat"))
sol = net.create_solver()
y0s = net.chem_load_initial_abundances(os.path.join(inp, "initial_condition_Garrod08_mod_waterice.dat"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
par = rb.synth.cell_params(n, first_cell=first)
y0 = rb.synth.initial_state(y0s, par, net.index("Grain0"))
res = sol.chem_evol_solve(par, y0, want_touts=False)
st = res["stats"]
o = np.argsort(-st[:, 0])[:8]
for c in o:
    print("cell", c, "NST", int(st[c,0]), "NFE", int(st[c,1]), "NJE", int(st[c,2]), "NLU", int(st[c,3]), "nsolve", int(st[c,5]), "NERR", int(st[c,6]), "nrestart", int(st[c,7]), "ncfail", int(st[c,8]), "nefail", int(st[c,9]), "nrec", int(st[c,10]), "istate", res["istate"][c], "q", res["quality"][c], "t_final", res["t_final"][c])
    P = rb.synth.P
    print("    n_gas %.3e Tgas %.1f Tdust %.1f G0 %.2e Av %.2f zetaX %.2e" % (par[c,P["n_gas"]], par[c,P["Tgas"]], par[c,P["Tdust"]], par[c,P["G0_UV_toStar"]], par[c,P["Av_toStar"]], par[c,P["zeta_Xray_H2"]]))
print("sum NST", st[:,0].sum(), "median", np.median(st[:,0]))
