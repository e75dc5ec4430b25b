"""Generates tests/golden/sweep_r02.npz: the CPU oracle's results on the benchmarked cell
stream (BASELINE.json configs[1]: Philox seed 20240613, rate06-withgrain, Garrod08 IC,
RTOL 1e-4 / ATOL 1e-30 policy j=1, max_runtime_allowed = 60 model seconds) for cells
0..1999 plus the cells the round-1 review singled out, once nominal and once with the
initial abundances perturbed by 1e-13 relative (round-off level for a run at RTOL 1e-4):
cells whose return codes differ between the two oracle runs are integrated at the edge of
what round-off allows, and no two implementations of the reference algorithm (the Fortran
binary on another compiler included) can be expected to agree on them.

    python tests/golden/make_sweep_fixture.py [nthreads]      (about 10 minutes on 8 cores)

Test infrastructure only (uses oracle/); reads nothing outside the repository.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import raco  # noqa: E402
import rac2d_b200.synth as synth  # noqa: E402

INP = os.path.join(ROOT, "tests", "golden", "inp")
NCELL = 2000
SPECIAL = [3628, 4926, 6116, 6321, 7542, 7944, 8231, 9670, 13114, 18564, 26688, 32998, 36370,
           2091, 11980, 15520, 1650, 14392, 24682, 25791]
BUDGET = 60.0
PERT = 1e-13


def cell_ids():
    return list(range(NCELL)) + [c for c in SPECIAL if c >= NCELL]


def inputs(net_names_index_grain0, y0s):
    ids = cell_ids()
    par = np.vstack([synth.cell_params(NCELL)] + [synth.cell_params(1, first_cell=c) for c in ids[NCELL:]])
    y0 = synth.initial_state(y0s, par, net_names_index_grain0)
    return ids, par, y0


def perturb(y0, N, ids):
    out = y0.copy()
    for k, c in enumerate(ids):
        rng = np.random.default_rng(977 + c)
        out[k, :N] *= 1.0 + PERT * (2.0 * rng.random(N) - 1.0)
    return out


def main():
    nth = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    net = raco.Network(os.path.join(INP, "rate06_dipole_reformated_again_withgrain.dat"))
    y0s = net.load_initial_abundances(os.path.join(INP, "initial_condition_Garrod08_mod_waterice.dat"))
    ids, par, y0 = inputs(net.names.index("Grain0") + 1, y0s)
    a = net.evol_solve_batch(par, y0, nthreads=nth, max_runtime_allowed=BUDGET)
    b = net.evol_solve_batch(par, perturb(y0, net.N, ids), nthreads=nth, max_runtime_allowed=BUDGET)
    # |y_perturbed - y| / (10 RTOL_i |y|) over species above 1e-12: the oracle's own sensitivity
    tol = np.array([1e-2 if nm.startswith("g") else 1e-3 for nm in net.names])
    p_viol = np.zeros(len(ids), np.float32)
    for k in range(len(ids)):
        ya, yb = a["y"][k, :net.N].astype(np.float32).astype(np.float64), b["y"][k, :net.N]
        m = np.abs(ya) > 1e-12
        if m.any():
            p_viol[k] = np.max(np.abs(yb[m] - ya[m]) / (np.abs(ya[m]) * tol[m]))
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "sweep_r02.npz"), ids=np.array(ids, np.int32),
        istate=a["istate"], quality=a["quality"], t_final=a["t_final"],
        nrec=a["stats"][:, 10].astype(np.int32), nst=a["stats"][:, 0].astype(np.int32),
        ncfail=a["stats"][:, 8].astype(np.int32), premature=a["stats"][:, 14].astype(np.int8),
        runtime=a["stats"][:, 13].astype(np.float32), y=a["y"][:, :net.N].astype(np.float32),
        p_istate=b["istate"], p_quality=b["quality"], p_t_final=b["t_final"],
        p_nrec=b["stats"][:, 10].astype(np.int32), p_viol=p_viol, budget=BUDGET, pert=PERT)
    stable = (a["istate"] == b["istate"]) & (a["quality"] == b["quality"]) & (a["t_final"] == b["t_final"])
    print(f"{len(ids)} cells; oracle vs perturbed oracle: {int((~stable).sum())} cells change their return codes")


if __name__ == "__main__":
    main()
