"""GPU parity of the integrator's own in-kernel f / Jacobian routines (not the stand-alone
K2/K3 kernels) against the oracle, at the initial state and at evolved states."""
import numpy as np
import pytest

from conftest import IC_GARROD, NET_A

pytestmark = pytest.mark.gpu


def test_in_kernel_f_and_jacobian_match_oracle(rb, oracle):
    net = rb.ChemNetwork(NET_A)
    onet = oracle.Network(NET_A)
    sol = net.create_solver(device=0)
    y0s = net.chem_load_initial_abundances(IC_GARROD)
    ncell = 6
    par = rb.synth.stratified_params(ncell)
    y0 = rb.synth.initial_state(y0s, par, net.index("Grain0"))
    N = net.N
    states = []
    for c in range(ncell):
        rt, at = onet.solver_flags_alt(1, 1e-4, 1e-30, par[c, 6])
        states.append(y0[c])
        for tm in (1e-2, 1e3):       # evolved states: fluxes span ~30 decades there
            states.append(onet.evol_solve(par[c], y0[c], rt, at, t_max=tm, want_record=False)["y"])
    pars = np.repeat(par, 3, axis=0)
    ys = np.array(states)
    fg, jg = sol.debug_fjac(pars, ys)
    for k in range(len(states)):
        rates = onet.cal_rates(pars[k])
        fo, jo = onet.ode_f(pars[k], rates, ys[k]), onet.ode_jac_csc(pars[k], rates, ys[k])
        fa, ja = onet.ode_f_abs(pars[k], rates, ys[k]), onet.ode_jac_csc_abs(pars[k], rates, ys[k])
        # tolerance: 1e-13 of the sum of absolute terms (the rounding scale of each sum)
        assert np.all(np.abs(fg[k, :N] - fo[:N]) <= 1e-13 * fa[:N] + 1e-300), k
        assert np.all(np.abs(jg[k] - jo) <= 1e-13 * ja + 1e-300), k
