"""CPU tests of the product's host side: the C++ loader mirror and the symbolic setup
against the (independently written) oracle, the C-ABI surface, the synthetic generator
and the multi-GPU sharding logic (gloo, world_size 2).  No compute call needs a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import IC_GARROD, IC_LOMETAL, NET_A, NET_B, NET_C, ROOT


@pytest.mark.parametrize("path", [NET_A, NET_B, NET_C])
def test_loader_and_pattern_bit_exact_vs_oracle(rb, oracle, path):
    """Reaction indexing and the sparsity pattern must be bit-exact (north_star)."""
    net = rb.ChemNetwork(path)
    o = oracle.Network(path)
    assert (net.R, net.N) == (o.R, o.N)
    assert net.names == o.names and net.ctype == o.ctype
    for a, b in ((net.reac, o.reac), (net.prod, o.prod), (net.n_reac, o.n_reac), (net.n_prod, o.n_prod),
                 (net.itype, o.itype), (net.ABC, o.ABC), (net.T_range, o.T_range), (net.elements, o.elements),
                 (net.mass_num, o.mass_num), (net.dupli_ptr, o.dupli_ptr),
                 (net.dupli_list[:o.n_dupli], o.dupli_list[:o.n_dupli])):
        assert np.array_equal(a, b)
    assert np.array_equal(net.vib_freq, o.vib_freq, equal_nan=True)
    assert np.array_equal(net.Edesorb, o.Edesorb, equal_nan=True)
    sol = net.create_solver()          # host-only handle when no GPU is visible
    ia, ja = sol.pattern()
    assert np.array_equal(ia, o.ia) and np.array_equal(ja, o.ja)
    assert (sol.NNZ, sol.NNZ_diag) == (o.NNZ, o.nnz_diag)
    perm = sol.ordering()
    assert sorted(perm.tolist()) == list(range(1, net.N + 1))
    # fill of the product's ordering is within 10 % of the oracle's minimum-degree ordering
    # (oracle counts the T slot's full row and column as well)
    assert sol.nnz_lu <= 1.10 * o.nnz_ldu
    for ic in (IC_GARROD, IC_LOMETAL):
        assert np.array_equal(net.chem_load_initial_abundances(ic), o.load_initial_abundances(ic))


@pytest.mark.parametrize("path", [NET_A, NET_B, NET_C])
def test_factorisation_and_solve_schedules_are_consistent(rb, path):
    """Every update of the symbolic LU appears exactly once and in the level of its pivot, the
    targets of a level are distinct and never operands of that level, the staged solves cover
    every L/U entry of the head block once (racg_selfcheck, host only)."""
    net = rb.ChemNetwork(path)
    sol = net.create_solver()
    sol.selfcheck()


@pytest.mark.parametrize("mode", [1, 2])
def test_schedule_selfcheck_detects_damage(rb, mode):
    """The self-check is not vacuous: a copy of the schedule with one damaged entry (a wrong
    operand position / a target written twice) is rejected (racg_selfcheck_damaged)."""
    net = rb.ChemNetwork(NET_A)
    sol = net.create_solver()
    sol.selfcheck()
    with pytest.raises(rb.RacgError, match="self-check"):
        sol.selfcheck_damaged(mode)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU arm) prints one JSON line with the contract's keys."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-cells", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "cells/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"


def test_solver_flags_alt_matches_oracle(rb, oracle):
    net = rb.ChemNetwork(NET_A)
    o = oracle.Network(NET_A)
    sol = net.create_solver()
    par = rb.synth.cell_params(5)
    for j in (1, 2, 3, 4, 6):
        rt, at = sol.chem_set_solver_flags_alt(j, 1e-4, 1e-30, par)
        for c in range(5):
            ro, ao = o.solver_flags_alt(j, 1e-4, 1e-30, par[c, 6])
            assert np.array_equal(rt[c], ro) and np.array_equal(at[c], ao)
    assert sol.n_record(0.0, 1e6, 1e-8, 1.1) == 316


def test_c_abi_exports_every_declared_symbol(rb):
    """libracg.so loads and exports every function include/racg.h declares."""
    hdr = open(os.path.join(ROOT, "include", "racg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(racg_[a-z_]+)\s*\(", hdr))
    assert len(names) >= 15
    L = C.CDLL(rb.lib_path())
    for n in sorted(names):
        assert hasattr(L, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", rb.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (racg_\w+)", out))
    assert names <= exported


def test_no_gpu_means_loud_failure_not_fallback(rb):
    """Without a CUDA device every compute entry point fails with RACG_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    net = rb.ChemNetwork(NET_A)
    sol = net.create_solver()
    par = rb.synth.cell_params(2)
    with pytest.raises(rb.RacgError, match="no CPU fallback"):
        sol.chem_cal_rates(par)
    y0 = rb.synth.initial_state(net.chem_load_initial_abundances(IC_GARROD), par, net.index("Grain0"))
    with pytest.raises(rb.RacgError, match="no CPU fallback"):
        sol.chem_evol_solve(par, y0)


def test_product_never_touches_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "rac-2d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "raco" not in src and "libraco" not in src, os.path.join(dirpath, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libracg.so")], capture_output=True, text=True).stdout
    assert "raco" not in out


def test_bad_network_inputs_are_rejected(rb):
    net = rb.ChemNetwork(NET_A)
    bad = net.itype.copy()
    # a type-21 (ion + grain) reaction whose reactants are both grains -> the reference's error_stop
    i21 = int(np.nonzero(net.itype == 21)[0][0])
    saved = net.reac[i21].copy()
    g0 = net.index("Grain0")
    net.reac[i21, 0] = g0
    net.reac[i21, 1] = net.index("Grain-")
    with pytest.raises(rb.RacgError, match="type 21"):
        net.create_solver()
    net.reac[i21] = saved
    cfg = rb.default_cfg()
    cfg.H2_form_use_moeq = 1
    with pytest.raises(rb.RacgError, match="not supported"):
        net.create_solver(cfg)
    del bad


def test_synthetic_cells_deterministic_and_prefix_stable(rb):
    a = rb.synth.cell_params(100)
    b = rb.synth.cell_params(40, first_cell=30)
    assert np.array_equal(a[30:70], b)
    assert np.array_equal(a, rb.synth.cell_params(100))
    P = rb.synth.P
    assert a[:, P["n_gas"]].min() >= 1e3 and a[:, P["n_gas"]].max() <= 1e13
    assert a[:, P["Tgas"]].min() >= 8 and a[:, P["Tdust"]].min() >= 5
    assert np.all(a[:, P["ndust_tot"]] == a[:, P["n_gas"]] * a[:, P["ratioDust2HnucNum"]])
    s = rb.synth.stratified_params(16)
    assert np.all(s[0::4, P["n_gas"]] >= 1e9) and np.all(s[2::4, P["Av_toStar"]] < 1.0)


def test_world_size_2_sharding_with_gloo(tmp_path):
    """N>1 path on CPU: the cell stream is dealt round-robin to the ranks (as bench.py does),
    results are gathered with one all_gather; the union equals the single-rank batch."""
    script = tmp_path / "shard.py"
    script.write_text(f'''
import os, sys
sys.path.insert(0, {ROOT!r})
import numpy as np, torch, torch.distributed as dist
import rac2d_b200 as rb
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ncell = 24
par = np.ascontiguousarray(rb.synth.cell_params(world * ncell)[rank::world])
# stand-in for the solve: a deterministic per-cell function of the inputs
yf = torch.from_numpy(np.ascontiguousarray(par[:, :8].T.copy()))
out = torch.empty((world,) + tuple(yf.shape), dtype=yf.dtype)
dist.all_gather_into_tensor(out.view(-1), yf.view(-1))
full = rb.synth.cell_params(world * ncell)
for r in range(world):
    assert np.array_equal(out[r].numpy().T, full[r::world, :8])
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == world
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
''')
    env = dict(os.environ)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_chemical_data_file_has_the_reference_record_layout(rb, tmp_path):
    """f4: chemical_data_iter_NNNN.bin (back_cells_chemical_data, src/data_dump.f90:88-162) is a
    direct-access file of fixed-length records [abundances(nSpecies), col_den_toStar(ncd),
    col_den_toISM(ncd)] in native doubles, one per leaf cell, no record markers."""
    ncell, ns, ncd = 7, 467, 10
    rng = np.random.default_rng(3)
    ab = np.asfortranarray(rng.random((ncell, ns + 1)))      # batch layout incl. the T slot
    cs, ci = rng.random((ncell, ncd)), rng.random((ncell, ncd))
    rb.write_chemical_data(str(tmp_path), 12, ab, cs, ci, nspecies=ns)
    fn = tmp_path / "chemical_data_iter_0012.bin"
    raw = np.fromfile(fn, dtype=np.float64)
    assert raw.size == ncell * (ns + 2 * ncd)                # record_len = 8 * (nSpecies + 2 ncd)
    rec = raw.reshape(ncell, ns + 2 * ncd)
    assert np.array_equal(rec[:, :ns], ab[:, :ns]) and np.array_equal(rec[:, ns:ns + ncd], cs)
    assert np.array_equal(rec[:, ns + ncd:], ci)
    ab2, cs2, ci2 = rb.read_chemical_data(str(tmp_path), 12, ncell, ns, ncd)
    assert np.array_equal(ab2, ab[:, :ns]) and np.array_equal(cs2, cs) and np.array_equal(ci2, ci)
    rb.write_chemical_data(str(tmp_path), -1, ab, nspecies=ns)
    assert (tmp_path / "chemical_data.bin").stat().st_size == 8 * ncell * ns


def test_andrews_disk_emulation_grid():
    """configs[3] stand-in (rac-2d_b200/synth.py:andrews_disk_cells): 200 columns, the density of
    src/grid.f90:1741-1818 integrates to the configured disk mass, the refinement rule holds, the
    fields are finite and inside the ranges the rate routines accept; deterministic."""
    import rac2d_b200.synth as synth
    par, geom = synth.andrews_disk_cells()
    par2, _ = synth.andrews_disk_cells()
    assert np.array_equal(par, par2)
    P = synth.P
    assert len(np.unique(geom[:, 0])) == 200 and 3000 < par.shape[0] < 20000
    assert abs(geom[:, 0].min() - 0.1) < 1e-12 and abs(geom[:, 1].max() - 200.0) < 1e-9
    assert np.isfinite(par).all()
    # mass of the gridded part: both sides of the midplane, mean molecular weight 1.4 per H nucleus
    r0, r1, z0, z1 = geom.T
    vol = np.pi * (r1 ** 2 - r0 ** 2) * (z1 - z0) * 2.0 * synth.phy_AU2cm ** 3
    rc, zc = 0.5 * (r0 + r1), 0.5 * (z0 + z1)
    mass = (synth.andrews_dens(rc, zc) * vol).sum() * 1.4 * synth.phy_mProton_CGS / synth.phy_Msun_CGS
    # Md normalises the untapered profile between rin and rout; inside r0_in_exp = 3.5 AU the taper
    # removes all but 1e-5 of it
    e = lambda r: np.exp(-(r / 80.0) ** 0.5)
    expect = 2e-2 * (e(3.5) - e(200.0)) / (e(0.1) - e(200.0))
    assert abs(mass / expect - 1.0) < 0.03, (mass, expect)
    # refinement: the analytic density varies by <= 1.5 (+ the size clamps) across a cell's height
    ratio = synth.andrews_dens(rc, z0) / np.maximum(synth.andrews_dens(rc, z1), 1e-300)
    free = ((z1 - z0) > 0.0201) & (z1 < 199.9)
    assert np.all(ratio[free] <= 1.5 * (1 + 1e-9))
    # columns inner -> outer, each from the surface down to the midplane
    assert np.all(np.diff(geom[:, 0]) >= 0)
    same = np.diff(geom[:, 0]) == 0
    assert np.all(np.diff(geom[:, 2])[same] < 0)
    n, Tg, Td = par[:, P["n_gas"]], par[:, P["Tgas"]], par[:, P["Tdust"]]
    assert n.min() >= 1e3 and 8 <= Tg.min() and Tg.max() <= 3000 and 5 <= Td.min() and Td.max() <= 1500
    for k in ("fss_toISM_H2", "fss_toStar_H2", "fss_toISM_CO", "fss_toStar_CO"):
        assert np.all((par[:, P[k]] > 0) & (par[:, P[k]] <= 1))
    # Av grows downward inside a column
    assert np.all(np.diff(par[:, P["Av_toISM"]])[same] > 0)
