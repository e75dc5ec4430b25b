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


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def workload_text(ncell, network, R, NEQ, NNZ):
    return (f"configs[1]: {ncell} synthetic cells per GPU, {network} (R={R}, NEQ={NEQ}, NNZ={NNZ}), "
            f"Garrod08 waterice IC, t=1e-8..1e6 yr, RTOL 1e-4 ATOL 1e-30 (policy j=1), mxstep 6000, "
            f"reset every 50 outputs, evolT=F")


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference algorithm on the host cores (the Fortran
    `rac` cannot be built: no Fortran compiler in the image).  Rank 0 only."""
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
    nsample = args.cpu_cells if args.cpu_cells > 0 else max(2 * cores, 16)
    par = synth.cell_params(nsample)
    y0 = synth.initial_state(y0s, par, int(onet.special[14]))
    cfg = raco.default_cfg(jac_mode=args.cpu_jac_mode)
    times = []
    for it in range(args.warmup + args.steps):
        t = time.perf_counter()
        onet.evol_solve_batch(par, y0, nthreads=cores, cfg=cfg)
        dt = time.perf_counter() - t
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = nsample * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(args.ncell, args.network, onet.R, onet.NEQ, onet.NNZ),
                   "cells_per_gpu": args.ncell, "seed": synth.SEED,
                   "cells_per_step": nsample,
                   "note": "each step integrates a bounded sample (the first cells_per_step cells) of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {nsample} cells of the same synthetic stream per step, "
                                   f"oracle C++ port of chem_evol_solve+DLSODES, jac_mode={args.cpu_jac_mode} "
                                   f"({'O(R) Jacobian' if args.cpu_jac_mode == 0 else 'reference-faithful O(NEQ*R) Jacobian'}), "
                                   f"{cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="racg", choices=["racg", "reference"])
    ap.add_argument("--ncell", type=int, default=10000, help="cells per GPU per step")
    ap.add_argument("--network", default="rate06-withgrain", choices=list(NETWORKS))
    ap.add_argument("--cpu-cells", type=int, default=0, help="cells in the CPU-baseline sample (0 = 2 x cores)")
    ap.add_argument("--cpu-jac-mode", type=int, default=1,
                    help="1 = reference-faithful O(NEQ*R) Jacobian (default), 0 = O(R) Jacobian")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-ncell", type=int, default=75776,
                    help="cells for the K1-K3 roofline kernels (148 SMs x 128-cell tiles x 4 waves)")
    args = ap.parse_args()

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

    net = rb.ChemNetwork(os.path.join(INP, NETWORKS[args.network]))
    sol = net.create_solver(device=local_rank)
    y0s = net.chem_load_initial_abundances(os.path.join(INP, IC))
    ncell = args.ncell
    # the first world*ncell cells of the synthetic stream, dealt round-robin to the ranks (every
    # shard sees the same mix of stiffness; no data-path exchange between shards)
    par = np.ascontiguousarray(rb.synth.cell_params(world * ncell)[rank::world])
    y0 = rb.synth.initial_state(y0s, par, net.index("Grain0"))
    NEQ, R = sol.NEQ, sol.R
    nrec = sol.n_record(0.0, 1e6, 1e-8, 1.1)
    sp = SolveParams(1.1, 6000, 50, nrec, 1, 1e-4, 1e-30)

    f64 = dict(dtype=torch.float64, device=dev)
    # [item][cell] device layout == Fortran a(ncell,item)
    d_par = torch.from_numpy(np.ascontiguousarray(par.T)).to(dev)
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
    flush = torch.empty(256 * 1024 * 1024 // 8, **f64)     # > 126 MB L2

    def step():
        stream = torch.cuda.current_stream().cuda_stream
        sol.solve_batch_dev(ncell, sp, d_par.data_ptr(), d_y0.data_ptr(), d_t0.data_ptr(), d_tmax.data_ptr(),
                            d_dt.data_ptr(), d_yf.data_ptr(), d_tf.data_ptr(), d_nrec.data_ptr(),
                            d_ist.data_ptr(), d_q.data_ptr(), d_st.data_ptr(), stream=stream)
        if world > 1:   # the only collective: final abundance gather over NVLink
            dist.all_gather_into_tensor(gathered.view(-1), d_yf.view(-1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
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
    ms = sum(a.elapsed_time(b) for a, b in ev)
    if world > 1:
        t = torch.tensor([ms], **f64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * ncell * args.steps / (ms * 1e-3)

    stats = d_st.cpu().numpy()
    istate = d_ist.cpu().numpy()
    quality = d_q.cpu().numpy()
    phases = sol.phase_cycles()

    # ---- e2e through the host-pointer C-ABI call (what the Fortran host calls): host buffers in,
    # H2D + solve + D2H inside the timed region; every rank runs its shard, the slowest counts
    h_par = np.asfortranarray(par)
    h_y0 = np.asfortranarray(y0)
    e2e_t = []
    for it in range(1 + max(1, min(args.steps, 2))):
        barrier()
        t = time.perf_counter()
        res = sol.chem_evol_solve(h_par, h_y0, want_touts=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        if world > 1:
            tt = torch.tensor([dt], **f64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        if it > 0:
            e2e_t.append(dt)
    h2d = world * 8 * (rb.NPAR * ncell + NEQ * ncell + 3 * ncell)
    d2h = world * (8 * (NEQ * ncell + ncell + rb.NSTAT * ncell) + 4 * 3 * ncell)
    e2e_val = world * ncell / (sum(e2e_t) / len(e2e_t))

    if rank == 0:
        peak, peak_src = measured_peaks()
        # ---- roofline of the dominant kernel (integrate_kernel: one launch per step)
        abytes = algorithmic_bytes(stats, R, NEQ, sol.NNZ, sol.nnz_lu)
        ach = abytes / (ms_per_step * 1e-3) / 1e9
        roof = {"kernel": "integrate_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": abytes,
                "note": "per-cell state is L2/shared-memory resident by design; achieved = SURVEY 8(d) "
                        "whole-integration bytes from measured counters / launch time"}
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
        kern = {}
        for name, fn, nbytes in (
            ("rates_kernel(K1)", lambda: sol.rates_dev(nk, d_kpar.data_ptr(), d_k.data_ptr(), s_), 8.0 * (rb.NPAR + R)),
            ("rhs_kernel(K2)", lambda: sol.rhs_jac_dev(nk, d_kpar.data_ptr(), d_ky.data_ptr(), d_k.data_ptr(), d_yd.data_ptr(), 0, s_), 8.0 * (R + 2 * NEQ)),
            ("jac_kernel(K3)", lambda: sol.rhs_jac_dev(nk, d_kpar.data_ptr(), d_ky.data_ptr(), d_k.data_ptr(), 0, d_pd.data_ptr(), s_), 8.0 * (R + NEQ + sol.NNZ)),
        ):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            tt = 0.0
            for _ in range(5):
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                tt += e0.elapsed_time(e1)
            gbs = nbytes * nk / (tt / 5 * 1e-3) / 1e9
            kern[name] = {"ms": tt / 5, "cells": nk, "algorithmic_bytes_per_cell": nbytes, "achieved_GBs": gbs,
                          "frac_of_hbm_peak": gbs / peak}
        # ---- CPU baseline on this box's host cores (bounded sample)
        cpu = None
        if not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import raco
            raco.build()
            onet = raco.Network(os.path.join(INP, NETWORKS[args.network]))
            cores = os.cpu_count() or 1
            nsample = args.cpu_cells if args.cpu_cells > 0 else max(2 * cores, 16)
            t = time.perf_counter()
            o = onet.evol_solve_batch(par[:nsample], y0[:nsample], nthreads=cores,
                                      cfg=raco.default_cfg(jac_mode=args.cpu_jac_mode))
            dt = time.perf_counter() - t
            m = np.abs(o["y"][:, :net.N]) > 1e-12
            gy = d_yf.cpu().numpy().T[:nsample, :net.N]
            rel = float(np.max(np.abs(gy[m] - o["y"][:, :net.N][m]) / np.abs(o["y"][:, :net.N][m])))
            cpu = {"value": nsample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {nsample} cells of the step's batch, oracle C++ port of the reference "
                             f"algorithm (not the Fortran rac binary), jac_mode={args.cpu_jac_mode}, {cores} threads",
                   "max_rel_diff_vs_gpu_X>1e-12": rel}
        clocks = sampler.summary()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(ncell, args.network, R, NEQ, sol.NNZ),
                       "cells_per_gpu": ncell, "seed": rb.synth.SEED,
                       "sharding": "first n_gpus*cells_per_gpu cells of the stream, dealt round-robin to the ranks",
                       "scheduling": "work queue served heaviest-first from the previous step's per-cell cost",
                       "l2": "256 MB buffer written between timed iterations"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "n_gpus": world, "note": "racg_solve_batch with host buffers on every rank, slowest rank counts"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "kernels": kern,
            "cpu_baseline": cpu,
            "solver": {"istate_ok_frac": float(np.mean(istate == 2)), "quality0_frac": float(np.mean(quality == 0)),
                       "mean_steps": float(stats[0].mean()), "mean_f": float(stats[1].mean()),
                       "mean_jac": float(stats[2].mean()), "mean_lu": float(stats[3].mean()),
                       "mean_solves": float(stats[5].mean()),
                       "phase_share": {k: (v / phases["total"] if phases["total"] else None)
                                       for k, v in phases.items() if k not in ("total", "ncell")}},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
