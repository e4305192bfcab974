#!/usr/bin/env python
"""bench.py -- BASELINE.json metric on B200: model evaluations / s of flux_density_grid (100 t x 3 nu)
over a batch of seeded synthetic parameter draws, plus the MCMC log-likelihood rate (config 5).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU via torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores

A "step" is one pass of the hot path over one batch (per GPU) of `--batch` parameter sets
(SURVEY.md section 8d draw, seed = 1000 + rank).  `value` is timed with CUDA events on the launching
stream with inputs and outputs resident in HBM; `e2e` goes through the host-buffer C-ABI entry point
with pinned host memory (H2D of the parameters, D2H of the fluxes inside the timed region).
L2 is flushed between timed iterations (a 256 MiB device memset outside the event brackets).
"""
from __future__ import annotations

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

from vegasafterglow_b200 import abi, configs  # noqa: E402

METRIC = "model evals/sec (flux_density_grid 100t x 3nu)"
UNIT = "evals/s"


def workload(batch, rank=0):
    t, nu = configs.C1()[1:]
    P = configs.random_draw(batch, seed=1000 + rank)
    return P, t, nu


def loglike_workload(batch, rank=0):
    """BASELINE.json config 5: FS+RS tophat on-axis walkers, 5 bands x 20 epochs series + data."""
    P = configs.random_draw(batch, seed=2000 + rank, rvs=True)
    ts = np.sort(np.tile(np.logspace(2.5, 6.5, 20), 5))
    nus = np.tile([1e9, 5e9, 4.84e14, 1e17, 1e18], 20)
    return P, ts, nus


def config_dict(args, n_gpus):
    return {
        "workload": f"C1-shape batch: {args.batch} seeded random-draw TophatJet/ISM on-axis forward-shock "
                    f"synchrotron parameter sets per GPU per step, flux_density_grid 100 t x 3 nu "
                    f"(SURVEY.md 8d draw); config 5 log-likelihood reported under 'loglike'",
        "batch_per_gpu": args.batch,
        "global_batch": args.batch * n_gpus,
        "n_t": 100,
        "n_nu": 3,
        "parallelism": f"walker-partition x{n_gpus} (no data-path collective)",
        "l2": "flushed between timed iterations (256 MiB memset outside the event brackets)",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """--impl reference: the unmodified reference (oracle/_ref, built from /root/reference by
    oracle/Makefile) on all host cores, same workload/metric; each step = one full batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref  # the CPU-baseline leg is one of the places allowed to execute oracle/

    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libvagref.so not built"}))
        return
    cores = ref.hardware_threads()
    n = min(args.batch, 4096)
    P, t, nu = workload(n)
    for _ in range(max(args.warmup, 1)):
        ref.flux_density_grid(P[: max(64, n // 8)], t, nu, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.flux_density_grid(P, t, nu, n_threads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(args, 1) | {"batch_per_gpu": n, "global_batch": n},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{n} parameter sets per step x {args.steps} steps, std::thread over the unmodified "
                                   f"reference (oracle/ref_driver.cpp), {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="parameter sets per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from vegasafterglow_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    eng = Engine(local)
    P, t, nu = workload(args.batch, rank)
    n = P.size
    stream = torch.cuda.current_stream().cuda_stream

    # ---- device-resident inputs/outputs ("value") -------------------------------------------------
    d_p = torch.from_numpy(P.view(np.uint8).copy()).to(dev)
    d_t, d_nu = torch.from_numpy(t).to(dev), torch.from_numpy(nu).to(dev)
    d_out = torch.empty((n, abi.NCOMP, nu.size, t.size), dtype=torch.float64, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    eng.set_capacity(256, 128)

    def step_dev():
        eng.flux_density_grid_dev(d_p.data_ptr(), n, d_t.data_ptr(), t.size, d_nu.data_ptr(), nu.size,
                                  d_out.data_ptr(), d_st.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.set_profiling(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage_acc = {}
    launches = 0
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        step_dev()
        ev[i][1].record()
        torch.cuda.synchronize()
        for k, v in eng.last_stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
        launches += eng.last_launch_count()
    barrier()
    wall = time.perf_counter() - wall0
    eng.set_profiling(False)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    assert int(d_st.abs().sum()) == 0, "a model reported a status bit"
    assert bool(torch.isfinite(d_out).all())

    # ---- end to end through the host-buffer C-ABI call (pinned host memory) -----------------------
    h_p = torch.from_numpy(P.view(np.uint8).copy()).pin_memory()
    h_t, h_nu = torch.from_numpy(t.copy()).pin_memory(), torch.from_numpy(nu.copy()).pin_memory()
    h_out = torch.empty((n, abi.NCOMP, nu.size, t.size), dtype=torch.float64).pin_memory()
    h_st = torch.zeros(n, dtype=torch.int32).pin_memory()
    lib = eng._lib

    def step_e2e():
        rc = lib.vag_flux_density_grid(eng._h, h_p.data_ptr(), n, h_t.data_ptr(), t.size, h_nu.data_ptr(), nu.size,
                                       h_out.data_ptr(), h_st.data_ptr())
        assert rc == 0, lib.vag_last_error()

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - e0
    clocks = sampler.stop() if rank == 0 else None

    # ---- config 5: batched log-likelihood (series chi2), device-resident ---------------------------
    Pl, ts, nus = loglike_workload(args.batch, rank)
    rng = np.random.default_rng(42)
    lnF = np.log(1e-26 * (1 + 0.05 * rng.standard_normal(ts.size)) * (ts / 1e3) ** -1.0)
    sig = np.full(ts.size, 0.1)
    wgt = np.ones(ts.size)
    d_pl = torch.from_numpy(Pl.view(np.uint8).copy()).to(dev)
    d_arr = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (ts, nus, lnF, sig, wgt)]
    d_chi2 = torch.empty(n, dtype=torch.float64, device=dev)
    gathered = torch.empty(n * world, dtype=torch.float64, device=dev) if world > 1 else None

    def step_ll():
        eng.chi2_series_dev(d_pl.data_ptr(), n, *[a.data_ptr() for a in d_arr], ts.size, d_chi2.data_ptr(),
                            d_st.data_ptr(), stream)
        if world > 1:  # the only inter-GPU traffic of the path: gather of float64[n] log-likelihoods
            dist.all_gather_into_tensor(gathered, d_chi2)

    for _ in range(args.warmup):
        step_ll()
    barrier()
    la, lb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    la.record()
    for _ in range(args.steps):
        step_ll()
    lb.record()
    barrier()
    ll_ms = la.elapsed_time(lb)

    # ---- max over ranks ------------------------------------------------------------------------------
    times = torch.tensor([dev_ms, e2e_s * 1e3, ll_ms, wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, ll_ms, wall_ms = (float(x) for x in times.tolist())

    if rank == 0:
        total_models = n * world * args.steps
        value = total_models / (dev_ms * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        fp64_peak = eng.measure_fp64_peak()
        per = {k: v / args.steps for k, v in stage_acc.items()}  # ms per step of rank 0
        # algorithmic work per model evaluation (SURVEY.md 8d / DESIGN.md section 5), C1 shape
        flops = {"dynamics": 4.6e4, "radiation": 2.1e4, "eats": 9.9e5}
        bytes_eats = 23e3
        dominant = max(("grid", "dynamics", "eats"), key=lambda k: per.get(k, 0.0))
        eats_s = per["eats"] * 1e-3
        roofline = {"bound": "hbm", "kernel": "k_eats", "achieved": n * bytes_eats / eats_s / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": n * bytes_eats / eats_s / 1e9 / hbm_peak, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                    "note": "the path is FP64-pipe / dependent-latency bound, not HBM bound (SURVEY.md 8d): see "
                            "roofline_fp64 for the per-kernel FP64 fractions"}
        roofline_fp64 = {"peak_tflops": fp64_peak, "peak_source": "DFMA probe kernel run in this process",
                         "dominant_stage": dominant, "ms_per_step": per,
                         "kernels": {k: {"achieved_tflops": n * f / (per[k] * 1e-3) / 1e12,
                                         "frac": n * f / (per[k] * 1e-3) / 1e12 / fp64_peak} for k, f in flops.items()
                                     if per.get(k, 0) > 0}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(args, world),
            "e2e": {"value": total_models / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(P.nbytes + t.nbytes + nu.nbytes),
                    "d2h_bytes_per_step": int(h_out.numel() * 8 + h_st.numel() * 4)},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "roofline_fp64": roofline_fp64,
            "loglike": {"metric": "MCMC loglike evals/s (4096-walker FS+RS tophat, 100-point 5-band series)",
                        "value": n * world * args.steps / (ll_ms * 1e-3), "unit": UNIT, "ms_per_step": ll_ms / args.steps,
                        "collective": "all_gather float64[n] over NCCL" if world > 1 else "none (1 GPU)"},
            "wall_ms_per_step_incl_flush": wall_ms / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import ref  # CPU-baseline leg: allowed to execute oracle/

                if ref.available():
                    cores = ref.hardware_threads()
                    m = min(n, 2048)
                    ref.flux_density_grid(P[:64], t, nu, n_threads=cores)
                    c0 = time.perf_counter()
                    reps = 3
                    for _ in range(reps):
                        ref.flux_density_grid(P[:m], t, nu, n_threads=cores)
                    cdt = (time.perf_counter() - c0) / reps
                    line["cpu_baseline"] = {"value": m / cdt, "unit": UNIT, "cores": cores, "kind": "reference",
                                            "sample": f"first {m} parameter sets of the same batch x {reps} repeats, unmodified "
                                                      f"reference via oracle/_ref/libvagref.so with {cores} std::threads"}
            except Exception as exc:  # noqa: BLE001
                line["cpu_baseline"] = {"unavailable": repr(exc)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
