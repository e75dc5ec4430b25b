#!/usr/bin/env python
"""bench.py -- cells integrated to 1 Myr per second on N B200s (BASELINE.json metric).

A "step" = one pass of the hot path over one batch of synthetic cells: the batched
replacement of RAC-2D's `do_chemical_stuff` cell loop (rate coefficients + stiff BDF
integration 1e-8 -> 1e6 yr of every cell).  Workload at N=1 = BASELINE.json configs[1]:
1e4 synthetic cells, rate06-withgrain network, on one B200; for N>1 every rank gets its
own 1e4 cells (weak scaling; cells are independent, nothing runs between GPUs on the
path, the final abundance gather is one NCCL all_gather).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port)
    python bench.py --network rate12-withGrain               # configs[2] network
    python bench.py --stratified                             # configs[4] stiffness strata
    python bench.py --andrews                                # configs[3]: emulated ncol=200 Andrews-disk grid
    python bench.py --inlib --gpus N                         # N GPUs through ONE racg_solve_batch
                                                             # call (racg_use_devices), no torchrun

PyTorch is used only for device buffers, streams/events and torch.distributed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
INP = os.path.join(ROOT, "tests", "golden", "inp")
NETWORKS = {
    "rate06-withgrain": "rate06_dipole_reformated_again_withgrain.dat",
    "rate12-withGrain": "rate12_withGrain_lowH2Bind_hiObind.dat",
    "rate06-template": "rate06_withgrain_lowH2Bind_hiOBind_lowCObind.dat",
}
IC = "initial_condition_Garrod08_mod_waterice.dat"
METRIC = "cells integrated to 1 Myr per second"
UNIT = "cells/s"
# chemsol_params%max_runtime_allowed of the reference's template configuration
# (inp/template_configure.dat:31), in model seconds (include/racg.h)
MAX_RUNTIME = 60.0


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(network, ncell):
    """dram__bytes_read.sum + dram__bytes_write.sum of one integrate_kernel launch of this
    workload from the committed ncu capture (profiles/r02/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "r02", "traffic.json")
    try:
        with open(p) as f:
            t = json.load(f)
        e = t.get(f"{network}:{ncell}")
        return (float(e["dram_bytes"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(stats, R, NEQ, NNZ, nnz_lu):
    """SURVEY.md 8(d) whole-integration formula from the measured per-cell counters:
    n_f*B(K2) + n_J*B(K3) + n_LU*B(K4) + n_solve*B(K5) + n_step*(q+3)*2*8*NEQ (q ~ 5 -> 8)."""
    nst, nfe, nje, nlu, nsolve = (stats[k].sum() for k in (0, 1, 2, 3, 5))
    b_k2 = 8.0 * (R + 2 * NEQ)
    b_k3 = 8.0 * (R + NEQ + NNZ)
    b_k4 = 8.0 * (NNZ + nnz_lu)
    b_k5 = 8.0 * (nnz_lu + 2 * NEQ)
    b_step = 8.0 * 2 * 8 * NEQ
    return nfe * b_k2 + nje * b_k3 + nlu * b_k4 + nsolve * b_k5 + nst * b_step


def config_dict(args, R, NEQ, NNZ, synth):
    """identical in both arms (the driver compares them)"""
    which = "configs[4] (stiffness-stratified cells, 4 strata interleaved)" if args.stratified else \
        ("configs[2] network" if args.network == "rate12-withGrain" else "configs[1]")
    if args.andrews:
        which = ("configs[3] (ncol=200 Andrews-2009 disk grid of inp/template_configure.dat, all cells of one structure "
                 "iteration in one batch; HOST-SIDE EMULATION of the Fortran host's grid and per-cell fields, "
                 "rac-2d_b200/synth.py:andrews_disk_cells)")
    return {"workload": f"{which}: {args.ncell} synthetic cells per GPU, {args.network} (R={R}, NEQ={NEQ}, NNZ={NNZ}), "
                        f"Garrod08 waterice IC, t=1e-8..1e6 yr, RTOL 1e-4 ATOL 1e-30 (policy j=1), mxstep 6000, "
                        f"reset every 50 outputs, evolT=F, max_runtime_allowed={MAX_RUNTIME:g} s "
                        f"(template_configure.dat; deterministic work model)",
            "cells_per_gpu": args.ncell, "seed": synth.SEED,
            "sharding": "first n_gpus*cells_per_gpu cells of the stream, dealt round-robin to the ranks",
            "scheduling": "work queue served heaviest-first from the previous step's per-cell cost "
                          "(value_cold / e2e_cold: first call, queue in cell order)",
            "l2": "256 MB buffer written between timed iterations"}


def make_cells(synth, args, n, first=0):
    if args.andrews:   # the emulated disk grid is one fixed list of cells (column by column)
        return np.ascontiguousarray(synth.andrews_disk_cells()[0][first:first + n])
    return synth.stratified_params(n, first_cell=first) if args.stratified else synth.cell_params(n, first_cell=first)


def cpu_sample_ids(ncell, nsample):
    """a fixed-stride sample of the step's batch (not its first cells)"""
    nsample = min(nsample, ncell)
    return (np.arange(nsample) * (ncell // nsample)).astype(np.int64)


def run_cpu(onet, raco, par, y0, ids, cores, jac_mode):
    t = time.perf_counter()
    o = onet.evol_solve_batch(par[ids], y0[ids], nthreads=cores, cfg=raco.default_cfg(jac_mode=jac_mode),
                              max_runtime_allowed=MAX_RUNTIME)
    return len(ids) / (time.perf_counter() - t), o


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference algorithm on the host cores (the Fortran
    `rac` cannot be built: no Fortran compiler in the image).  Rank 0 only.  Each step is a
    bounded, fixed-stride sample of the workload."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import raco
    raco.build()
    import rac2d_b200.synth as synth
    netf = os.path.join(INP, NETWORKS[args.network])
    onet = raco.Network(netf)
    y0s = onet.load_initial_abundances(os.path.join(INP, IC))
    cores = os.cpu_count() or 1
    nsample = args.cpu_cells if args.cpu_cells > 0 else 4 * cores
    par = make_cells(synth, args, args.ncell)
    y0 = synth.initial_state(y0s, par, int(onet.special[14]))
    ids = cpu_sample_ids(args.ncell, nsample)
    times = []
    for it in range(args.warmup + args.steps):
        t = time.perf_counter()
        run_cpu(onet, raco, par, y0, ids, cores, args.cpu_jac_mode)
        dt = time.perf_counter() - t
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = len(ids) * len(times) / total
    best, _ = run_cpu(onet, raco, par, y0, ids, cores, 0) if args.cpu_jac_mode == 1 else (value, None)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, onet.R, onet.NEQ, onet.NNZ, synth),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "value_best_effort_O(R)_jacobian": best,
                         "sample": f"{len(ids)} cells per step, every {args.ncell // len(ids)}-th cell of the step's "
                                   f"{args.ncell}-cell batch; oracle C++ port of chem_evol_solve+DLSODES (not the "
                                   f"Fortran rac binary), jac_mode={args.cpu_jac_mode} "
                                   f"({'O(R) Jacobian' if args.cpu_jac_mode == 0 else 'reference-faithful O(NEQ*R) Jacobian'}), "
                                   f"{cores} threads, dynamic queue"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def code_histogram(istate, quality, premature):
    h = {}
    for i, q, p in zip(istate.tolist(), quality.tolist(), premature.tolist()):
        k = f"istate={i},quality={q}" + (",premature" if p else "")
        h[k] = h.get(k, 0) + 1
    return dict(sorted(h.items(), key=lambda kv: -kv[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="racg", choices=["racg", "reference"])
    ap.add_argument("--ncell", type=int, default=10000, help="cells per GPU per step")
    ap.add_argument("--network", default="rate06-withgrain", choices=list(NETWORKS))
    ap.add_argument("--stratified", action="store_true", help="configs[4]: stiffness-stratified cells")
    ap.add_argument("--andrews", action="store_true",
                    help="configs[3]: every cell of an emulated ncol=200 Andrews-disk grid in one batch (cells per GPU = grid cells // gpus)")
    ap.add_argument("--inlib", action="store_true",
                    help="drive --gpus N GPUs from this one process through racg_use_devices + racg_solve_batch")
    ap.add_argument("--cpu-cells", type=int, default=0, help="cells in the CPU sample (0 = 8 x cores; 4 x cores per step of --impl reference)")
    ap.add_argument("--cpu-jac-mode", type=int, default=1,
                    help="1 = reference-faithful O(NEQ*R) Jacobian (default), 0 = O(R) Jacobian")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-ncell", type=int, default=75776,
                    help="cells for the K1-K3 roofline kernels (148 SMs x 128-cell tiles x 4 waves)")
    args = ap.parse_args()
    if args.andrews:
        sys.path.insert(0, ROOT)
        import rac2d_b200.synth as _synth
        args.ncell = _synth.andrews_disk_cells()[0].shape[0] // max(args.gpus, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import rac2d_b200 as rb
    from rac2d_b200.chem import SolveParams

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libracg has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    inlib = args.inlib and world == 1 and args.gpus > 1
    ngpu = args.gpus if inlib else world

    net = rb.ChemNetwork(os.path.join(INP, NETWORKS[args.network]))
    sol = net.create_solver(device=local_rank)
    if inlib:
        ngpu = sol.use_devices(list(range(args.gpus)))
    y0s = net.chem_load_initial_abundances(os.path.join(INP, IC))
    ncell = args.ncell
    # the first ngpu*ncell cells of the synthetic stream, dealt round-robin to the ranks (every
    # shard sees the same mix of stiffness; no data-path exchange between shards)
    if inlib:
        par = make_cells(rb.synth, args, ngpu * ncell)
    else:
        par = np.ascontiguousarray(make_cells(rb.synth, args, world * ncell)[rank::world])
    y0 = rb.synth.initial_state(y0s, par, net.index("Grain0"))
    NEQ, R = sol.NEQ, sol.R
    sp = SolveParams(1.1, 6000, 50, 0, 1, 1e-4, 1e-30, MAX_RUNTIME)
    f64 = dict(dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, **f64)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world > 1:
            t = torch.tensor([x], **f64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # ---- e2e through the host-pointer C-ABI call (what the Fortran host calls): host buffers in,
    # H2D + solve + D2H inside the timed region; every rank runs its shard, the slowest counts.
    # First call = cold (queue in cell order), then warm calls.
    h_par = np.asfortranarray(par)
    h_y0 = np.asfortranarray(y0)
    nloc = h_par.shape[0]
    e2e_t = []
    res = None
    for it in range(1 + max(1, min(args.steps, 2))):
        flush.fill_(0.0)
        barrier()
        t = time.perf_counter()
        res = sol.chem_evol_solve(h_par, h_y0, want_touts=False, max_runtime_allowed=MAX_RUNTIME, nrec_max=0)
        torch.cuda.synchronize()
        e2e_t.append(allmax(time.perf_counter() - t))
    ntot = ngpu * ncell
    h2d = ntot * 8 * (rb.NPAR + NEQ + 3)
    d2h = ntot * (8 * (NEQ + 1 + rb.NSTAT) + 4 * 3)
    e2e_cold = ntot / e2e_t[0]
    e2e_val = ntot / (sum(e2e_t[1:]) / len(e2e_t[1:]))

    if inlib:
        # one process, N GPUs behind the C-ABI: there is no device-pointer variant of the sharded
        # call, so the device-timed value does not exist; report the host-call numbers
        stats = res["stats"].T
        istate, quality = res["istate"], res["quality"]
        ms_per_step = 1e3 * sum(e2e_t[1:]) / len(e2e_t[1:])
        value, value_cold, launches = e2e_val, e2e_cold, sol.launch_count()
        clocks = None
        gy = res["y"]
        tfin = res["t_final"]
    else:
        # ---- device-resident arm: inputs already in HBM, CUDA events on the launch stream
        d_par = torch.from_numpy(np.ascontiguousarray(par.T)).to(dev)   # [item][cell] == Fortran a(ncell,item)
        d_y0 = torch.from_numpy(np.ascontiguousarray(y0.T)).to(dev)
        d_t0 = torch.zeros(ncell, **f64)
        d_tmax = torch.full((ncell,), 1e6, **f64)
        d_dt = torch.full((ncell,), 1e-8, **f64)
        d_yf = torch.empty((NEQ, ncell), **f64)
        d_tf = torch.empty(ncell, **f64)
        d_nrec = torch.empty(ncell, dtype=torch.int32, device=dev)
        d_ist = torch.empty(ncell, dtype=torch.int32, device=dev)
        d_q = torch.empty(ncell, dtype=torch.int32, device=dev)
        d_st = torch.empty((rb.NSTAT, ncell), **f64)
        gathered = torch.empty((world, NEQ, ncell), **f64) if world > 1 else None

        def step():
            stream = torch.cuda.current_stream().cuda_stream
            sol.solve_batch_dev(ncell, sp, d_par.data_ptr(), d_y0.data_ptr(), d_t0.data_ptr(), d_tmax.data_ptr(),
                                d_dt.data_ptr(), d_yf.data_ptr(), d_tf.data_ptr(), d_nrec.data_ptr(),
                                d_ist.data_ptr(), d_q.data_ptr(), d_st.data_ptr(), stream=stream)
            if world > 1:   # the only collective: final abundance gather over NVLink
                dist.all_gather_into_tensor(gathered.view(-1), d_yf.view(-1))

        # cold step: the handle has no per-cell cost for this batch size on the device path yet
        sol.set_option("warm_order", 0)
        flush.fill_(0.0)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(); step(); c1.record()
        barrier()
        value_cold = world * ncell / (allmax(c0.elapsed_time(c1)) * 1e-3)
        sol.set_option("warm_order", 1)
        for _ in range(max(args.warmup - 1, 1)):
            flush.fill_(0.0)
            step()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = sol.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for k in range(args.steps):
            flush.fill_(0.0)               # L2 flush between timed iterations (outside the events)
            ev[k][0].record()
            step()
            ev[k][1].record()
        barrier()
        launches = sol.launch_count() - l0
        sampler.stop_flag = True
        ms = allmax(sum(a.elapsed_time(b) for a, b in ev))
        ms_per_step = ms / args.steps
        value = world * ncell * args.steps / (ms * 1e-3)
        stats = d_st.cpu().numpy()
        istate = d_ist.cpu().numpy()
        quality = d_q.cpu().numpy()
        gy = d_yf.cpu().numpy().T
        tfin = d_tf.cpu().numpy()
        clocks = sampler.summary()

    if rank == 0:
        peak, peak_src = measured_peaks()
        # ---- roofline of the dominant kernel (integrate_kernel: one launch per step)
        abytes = algorithmic_bytes(stats, R, NEQ, sol.NNZ, sol.nnz_lu)
        ach = abytes / (ms_per_step * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic(args.network, ncell)
        roof = {"kernel": "integrate_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": abytes,
                "note": "per-cell state is L2/shared-memory resident by design (traffic << algorithmic bytes); "
                        "achieved = SURVEY 8(d) whole-integration bytes from measured counters / launch time"}
        kern = {}
        if not inlib:
            # ---- RHS / Jacobian stand-alone kernels (the 'RHS+Jac HBM GB/s vs peak' half of the metric)
            nk = args.kernel_ncell
            kpar = rb.synth.cell_params(nk)
            d_kpar = torch.from_numpy(np.ascontiguousarray(kpar.T)).to(dev)
            d_ky = torch.from_numpy(np.ascontiguousarray(rb.synth.initial_state(y0s, kpar, net.index("Grain0")).T)).to(dev)
            d_ky[:net.N] += 1e-12
            d_k = torch.empty((R, nk), **f64)
            d_yd = torch.empty((NEQ, nk), **f64)
            d_pd = torch.empty((sol.NNZ, nk), **f64)
            s_ = torch.cuda.current_stream().cuda_stream
            for name, fn, nbytes in (
                ("rates_kernel(K1)", lambda: sol.rates_dev(nk, d_kpar.data_ptr(), d_k.data_ptr(), s_), 8.0 * (rb.NPAR + R)),
                ("rhs_kernel(K2)", lambda: sol.rhs_jac_dev(nk, d_kpar.data_ptr(), d_ky.data_ptr(), d_k.data_ptr(), d_yd.data_ptr(), 0, s_), 8.0 * (R + 2 * NEQ)),
                ("jac_kernel_pipe(K3)", lambda: sol.rhs_jac_dev(nk, d_kpar.data_ptr(), d_ky.data_ptr(), d_k.data_ptr(), 0, d_pd.data_ptr(), s_), 8.0 * (R + NEQ + sol.NNZ)),
            ):
                for _ in range(3):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                tt = 0.0
                for _ in range(5):
                    flush.fill_(0.0)
                    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                    tt += e0.elapsed_time(e1)
                gbs = nbytes * nk / (tt / 5 * 1e-3) / 1e9
                kern[name] = {"ms": tt / 5, "cells": nk, "algorithmic_bytes_per_cell": nbytes, "achieved_GBs": gbs,
                              "frac_of_hbm_peak": gbs / peak}
            # K1 is FP64/SFU bound: report it against the measured DFMA peak as well
            k1 = kern["rates_kernel(K1)"]
            k1["bound"] = "fp64"
            k1["note"] = ("one pow + one exp per reaction dominate; measured DFMA peak 1.99 warp-DFMA/cycle/SM "
                          "(profiles/r01/fp64_micro.log) = 37.1 TFLOP/s at 1.965 GHz")
        # ---- CPU baseline on this box's host cores (bounded fixed-stride sample of the step's batch)
        cpu = None
        if not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import raco
            raco.build()
            onet = raco.Network(os.path.join(INP, NETWORKS[args.network]))
            cores = os.cpu_count() or 1
            nsample = args.cpu_cells if args.cpu_cells > 0 else 8 * cores
            ids = cpu_sample_ids(nloc, nsample)
            v1, o = run_cpu(onet, raco, par, y0, ids, cores, 1)
            v0, _ = run_cpu(onet, raco, par, y0, ids, cores, 0)
            oy = o["y"][:, :net.N]
            g = gy[ids][:, :net.N]
            ok_both = (o["istate"] == 2) & (o["quality"] == 0) & (o["t_final"] == 1e6) & \
                      (istate[ids] == 2) & (quality[ids] == 0)
            tol = np.array([1e-2 if nm.startswith("g") else 1e-3 for nm in net.names])   # 10 x RTOL_i
            viol = []
            for k in np.where(ok_both)[0]:
                m = np.abs(oy[k]) > 1e-12
                viol.append(float(np.max(np.abs(g[k][m] - oy[k][m]) / (np.abs(oy[k][m]) * tol[m]))))
            viol = np.array(viol) if viol else np.zeros(1)
            cpu = {"value": v1, "unit": UNIT, "cores": cores, "kind": "port",
                   "value_best_effort_O(R)_jacobian": v0,
                   "sample": f"{len(ids)} cells, every {nloc // len(ids)}-th cell of the step's batch; oracle C++ port of "
                             f"the reference algorithm (not the Fortran rac binary); value: reference-faithful "
                             f"O(NEQ*R) Jacobian (jac_mode=1), value_best_effort: O(R) Jacobian; {cores} threads",
                   "parity_on_sample": {
                       "cells": int(len(ids)),
                       "return_code_mismatches(istate,quality,t_final)": int(np.sum(
                           (o["istate"] != istate[ids]) | (o["quality"] != quality[ids]) |
                           (np.abs(o["t_final"] - tfin[ids]) > 1e-9 * np.abs(o["t_final"])))),
                       "abundance_diff_over_10xRTOL_i(X>1e-12)": {
                           "median": float(np.median(viol)), "p99": float(np.percentile(viol, 99)),
                           "max": float(viol.max()), "cells_compared": int(ok_both.sum())}}}
        prem = stats[14] != 0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpu, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, R, NEQ, sol.NNZ, rb.synth),
            "value_cold": value_cold,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "n_gpus": ngpu, "value_cold": e2e_cold,
                    "note": ("one racg_solve_batch call sharding the batch over the GPUs (racg_use_devices)" if inlib else
                             "racg_solve_batch with host buffers on every rank, slowest rank counts") +
                            "; pinned staging and copies inside the timed region"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "kernels": kern,
            "cpu_baseline": cpu,
            "solver": {"return_codes": code_histogram(istate, quality, prem),
                       "mean_steps": float(stats[0].mean()), "max_steps": float(stats[0].max()),
                       "mean_f": float(stats[1].mean()), "mean_jac": float(stats[2].mean()),
                       "mean_lu": float(stats[3].mean()), "mean_solves": float(stats[5].mean()),
                       "max_model_runtime_s": float(stats[13].max())},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
