"""Ad-hoc: which of K2 / K3 misbehaves at a given batch size (run one case per process)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rac2d_b200 as rb
inp = os.path.join(ROOT, "tests", "golden", "inp")
which, ncell = sys.argv[1], int(sys.argv[2])
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 2
net = rb.ChemNetwork(os.path.join(inp, "rate06_dipole_reformated_again_withgrain.dat"))
sol = net.create_solver()
sol.set_option("k3_variant", variant)
y0s = net.chem_load_initial_abundances(os.path.join(inp, "initial_condition_Garrod08_mod_waterice.dat"))
par = rb.synth.cell_params(ncell)
y = rb.synth.initial_state(y0s, par, net.index("Grain0"))
y[:, :net.N] += 1e-12
k = sol.chem_cal_rates(par)
f, pd = sol.chem_ode_f_jac(par, y, k, want_f=(which == "k2"), want_jac=(which == "k3"))
out = f if which == "k2" else pd
print(which, ncell, variant, "ok, checksum", float(np.abs(out).sum()), flush=True)
