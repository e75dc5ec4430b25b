"""Ad-hoc: step counters of one synthetic cell as a function of t_max (GPU)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rac2d_b200 as rb
inp = os.path.join(ROOT, "tests", "golden", "inp")
net = rb.ChemNetwork(os.path.join(inp, "rate06_dipole_reformated_again_withgrain.dat"))
sol = net.create_solver()
y0s = net.chem_load_initial_abundances(os.path.join(inp, "initial_condition_Garrod08_mod_waterice.dat"))
c = int(sys.argv[1])
par1 = rb.synth.cell_params(1, first_cell=c)
tm = np.array([1e2, 1e4, 1e5, 1e6])
par = np.repeat(par1, len(tm), axis=0)
y0 = rb.synth.initial_state(y0s, par, net.index("Grain0"))
res = sol.chem_evol_solve(par, y0, t_max=tm, want_touts=False)
for k in range(len(tm)):
    st = res["stats"][k]
    print("t_max %.0e NST %d NFE %d NJE %d NLU %d ncfail %d nefail %d NERR %d NQU %d HU %.3e nrestart %d" % (tm[k], st[0], st[1], st[2], st[3], st[8], st[9], st[6], st[4], st[12], st[7]))
