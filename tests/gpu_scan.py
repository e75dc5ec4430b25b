"""Ad-hoc GPU scan (not a pytest): integrate cells [first, first+n) plus extra cells of the
config-2 stream with a given work budget and save the per-cell return codes and counters.
usage: gpu_scan.py out.npz nrep budget first n [extra cell ids...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rac2d_b200 as rb
inp = os.path.join(ROOT, "tests", "golden", "inp")
out, mode, budget, first, n = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
extra = [int(a) for a in sys.argv[6:]]
net = rb.ChemNetwork(os.path.join(inp, "rate06_dipole_reformated_again_withgrain.dat"))
sol = net.create_solver()
y0s = net.chem_load_initial_abundances(os.path.join(inp, "initial_condition_Garrod08_mod_waterice.dat"))
ids = list(range(first, first + n)) + extra
par = np.vstack([rb.synth.cell_params(n, first_cell=first)] + [rb.synth.cell_params(1, first_cell=c) for c in extra])
y0 = rb.synth.initial_state(y0s, par, net.index("Grain0"))
for rep in range(max(mode, 1)):   # rep > 0: queue served heaviest-first from the previous batch
    t = time.time()
    res = sol.chem_evol_solve(par, y0, want_touts=False, max_runtime_allowed=budget)
    dt = time.time() - t
    print(f"rep {rep}: {len(ids)} cells in {dt:.2f}s = {len(ids)/dt:.1f} cells/s", flush=True)
st = res["stats"]
print(f"lib {os.environ.get('RACG_LIB','default')} nrep {mode} budget {budget}: {len(ids)} cells in {dt:.2f}s = {len(ids)/dt:.1f} cells/s; "
      f"istate!=2: {int((res['istate']!=2).sum())} quality!=0: {int((res['quality']!=0).sum())} premature: {int(st[:,14].sum())} "
      f"sum NST {st[:,0].sum():.0f} max NST {st[:,0].max():.0f} max cfail {st[:,8].max():.0f}", flush=True)
np.savez_compressed(out, ids=np.array(ids), t_final=res["t_final"], istate=res["istate"],
                    quality=res["quality"], nrec=res["n_record_real"], stats=st, seconds=dt)
