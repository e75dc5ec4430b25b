"""Ad-hoc: achieved bandwidth of the stand-alone K1/K2/K3 kernels."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rac2d_b200 as rb
inp = os.path.join(ROOT, "tests", "golden", "inp")
net = rb.ChemNetwork(os.path.join(inp, "rate06_dipole_reformated_again_withgrain.dat"))
sol = net.create_solver(device=0)
y0s = net.chem_load_initial_abundances(os.path.join(inp, "initial_condition_Garrod08_mod_waterice.dat"))
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
dev = torch.device("cuda", 0); f64 = dict(dtype=torch.float64, device=dev)
kpar = rb.synth.cell_params(nk)
d_kpar = torch.from_numpy(np.ascontiguousarray(kpar.T)).to(dev)
d_ky = torch.from_numpy(np.ascontiguousarray(rb.synth.initial_state(y0s, kpar, net.index("Grain0")).T)).to(dev)
d_ky[:net.N] += 1e-12
R, NEQ = sol.R, sol.NEQ
d_k = torch.empty((R, nk), **f64); d_yd = torch.empty((NEQ, nk), **f64); d_pd = torch.empty((sol.NNZ, nk), **f64)
s_ = torch.cuda.current_stream().cuda_stream
for name, fn, nbytes in (
    ("K1 rates", lambda: sol.rates_dev(nk, d_kpar.data_ptr(), d_k.data_ptr(), s_), 8.0 * (rb.NPAR + R)),
    ("K2 rhs", lambda: sol.rhs_jac_dev(nk, d_kpar.data_ptr(), d_ky.data_ptr(), d_k.data_ptr(), d_yd.data_ptr(), 0, s_), 8.0 * (R + 2 * NEQ)),
    ("K3 jac (pipelined: cp.async rate rows, prefetched slot words; 2 CTAs/SM, 64-cell tiles)", lambda: sol.rhs_jac_dev(nk, d_kpar.data_ptr(), d_ky.data_ptr(), d_k.data_ptr(), 0, d_pd.data_ptr(), s_), 8.0 * (R + NEQ + sol.NNZ)),
    ("K3 jac (pipelined, 1 CTA of 1024 threads per SM, two buffers, one barrier per group)", lambda: (sol.set_option("k3_variant", 5), sol.rhs_jac_dev(nk, d_kpar.data_ptr(), d_ky.data_ptr(), d_k.data_ptr(), 0, d_pd.data_ptr(), s_), sol.set_option("k3_variant", 3)), 8.0 * (R + NEQ + sol.NNZ)),
    ("K3 jac (jac_kernel_wide<2,512>)", lambda: (sol.set_option("k3_variant", 2), sol.rhs_jac_dev(nk, d_kpar.data_ptr(), d_ky.data_ptr(), d_k.data_ptr(), 0, d_pd.data_ptr(), s_), sol.set_option("k3_variant", 3)), 8.0 * (R + NEQ + sol.NNZ)),
    ("K3 jac (jac_kernel_wide<4,1024>)", lambda: (sol.set_option("k3_variant", 4), sol.rhs_jac_dev(nk, d_kpar.data_ptr(), d_ky.data_ptr(), d_k.data_ptr(), 0, d_pd.data_ptr(), s_), sol.set_option("k3_variant", 3)), 8.0 * (R + NEQ + sol.NNZ))):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); tt = 0.0
    for _ in range(5):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tt += e0.elapsed_time(e1)
    print("%s: %.3f ms  %.0f GB/s (%.1f%% of 6542.7)" % (name, tt / 5, nbytes * nk / (tt / 5 * 1e-3) / 1e9, 100 * nbytes * nk / (tt / 5 * 1e-3) / 1e9 / 6542.7))

# the K3 variants add the same terms in the same order: bitwise equal
outs = {}
for v in (3, 5, 2):
    sol.set_option("k3_variant", v); d_pd.zero_()
    sol.rhs_jac_dev(nk, d_kpar.data_ptr(), d_ky.data_ptr(), d_k.data_ptr(), 0, d_pd.data_ptr(), s_); torch.cuda.synchronize()
    outs[v] = d_pd.clone()
sol.set_option("k3_variant", 3)
print("K3 variants bitwise equal:", bool(torch.equal(outs[3], outs[5])), bool(torch.equal(outs[3], outs[2])))
