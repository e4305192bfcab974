#!/usr/bin/env python
"""bench.py -- BASELINE.json metric on B200: model evaluations / s of flux_density_grid (100 t x 3 nu)
over a batch of seeded synthetic parameter draws, plus the MCMC log-likelihood rate (config 5).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU via torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores

A "step" is one pass of the hot path over one batch (per GPU) of `--batch` parameter sets
(SURVEY.md section 8d draw, seed = 1000 + rank).  `value` is timed with CUDA events on the launching
stream with inputs and outputs resident in HBM; `e2e` goes through the host-buffer C-ABI entry point
with pinned host memory (H2D of the parameters, D2H of the fluxes inside the timed region).
L2 is flushed before every step (a 256 MiB device memset on the step's stream).  Up to `--inflight`
independent steps (batches) are in flight per GPU, each on its own context/stream; the timed region runs
from the first launch to the completion of the last kernel (CUDA events on the device timeline).
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
REF_BUILD = ("oracle/Makefile: g++ -O3 -march=x86-64-v3 -ffp-contract=fast -freciprocal-math, no LTO "
             "(the reference's CMake uses -march=native and LTO where available)")
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


def extra_workloads(rank=0):
    """BASELINE.json configs 2-5 as seeded batches (SURVEY.md section 8d request shapes): name -> (P, t, nu, series)."""
    ts = np.sort(np.tile(np.logspace(2.5, 6.5, 20), 5))
    nus = np.tile([1e9, 5e9, 4.84e14, 1e17, 1e18], 20)
    return {
        "C2_gaussian_offaxis_200x8": (configs.random_draw(512, seed=3000 + rank, jet="gaussian", theta_obs_max=0.4),
                                      np.logspace(2, 8, 200), np.logspace(9, 18, 8), False),
        "C3_fs_rs_wind_100x3": (configs.random_draw(4096, seed=4000 + rank, rvs=True, medium="wind"), np.logspace(1, 7, 100),
                                np.array([1e9, 4.84e14, 1e18]), False),
        "C4_powerlaw_ssc_kn_50x40": (configs.random_draw(128, seed=5000 + rank, jet="powerlaw", theta_obs_max=0.3, ssc=True, kn=True),
                                     np.logspace(2, 7, 50), np.logspace(9, 27, 40), False),
        "C5_gaussian_fs_rs_series100": (configs.random_draw(1024, seed=6000 + rank, rvs=True, jet="gaussian", theta_obs_max=0.4),
                                        ts, nus, True),
    }


# algorithmic work per stage from the batch's work counters (SURVEY.md section 8d formulas):
#   ODE        4.6e4 flop per forward-shock row (650 RHS x 70), 2.9e5 per forward+reverse row (950 x 300)
#   radiation  400 flop per shock-table cell
#   EATS       120 flop per (EATS cell x distinct frequency) + 12 per (EATS row x output element)
#   grid (K0)  6 pdf evaluations per quadrature attempt: 70 flop each for theta, 25 per theta node for phi, plus the
#              ~1.8 k profile evaluations of the 512-point scans / 101-point pre-scans at 30 flop
#   HBM bytes  8 [cells (7 shock + 17 photon doubles) x 2 (write + read) + outputs]
def algorithmic_work(wk, n_models, n_t_obs, n_nu_distinct, n_out_per_model):
    return {
        "grid": wk["quad_attempts_theta"] * 6 * 70 + wk["quad_phi_evals"] * 6 * 25 + n_models * 1800 * 30,
        "dynamics": wk["rows_fwd"] * 4.6e4 + wk["rows_pair"] * 2.9e5,
        "radiation": wk["cells"] * 400,
        "eats": wk["eats_cells"] * n_nu_distinct * 120 + wk["eats_rows"] * n_out_per_model * 12,
        "bytes": 8 * (wk["cells"] * 24 * 2 + n_models * n_out_per_model * 2),
    }


def bind_to_gpu_numa(index):
    """Pin this process to the CPUs `nvidia-smi topo -m` lists for its GPU, so that the pinned host buffers of the
    end-to-end arm (first touched right after) and the threads that drive the copies are NUMA-local to the GPU."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout.splitlines()
        hdr = next(l for l in out if "CPU Affinity" in l)
        col = [c.strip() for c in hdr.split("\t") if c.strip()].index("CPU Affinity")
        row = next(l for l in out if l.startswith(f"GPU{index}\t") or l.startswith(f"GPU{index} "))
        cells = [c.strip() for c in row.split("\t") if c.strip()]
        spec = cells[col + 1]  # the row carries its own label in column 0
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return spec
    except Exception:  # noqa: BLE001 -- topology not available: leave the affinity alone
        pass
    return None


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
        "l2": "flushed before every step (256 MiB memset on the step's stream, inside the timed region)",
        "inflight": f"{args.inflight} independent steps in flight per GPU (one vag_context + stream each)",
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


def _time_ref(fn, reps):
    fn()  # warm (page-in, thread pool)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def reference_side_lines(ref, cores, rank=0, reps=2):
    """The reference's own CPU path on the other metrics / configs of BASELINE.json, each on a BOUNDED sample."""
    out = {}
    Pl, ts, nus = loglike_workload(2048, rank)
    rng = np.random.default_rng(42)
    lnF = np.log(1e-26 * (1 + 0.05 * rng.standard_normal(ts.size)) * (ts / 1e3) ** -1.0)
    dt = _time_ref(lambda: ref.chi2_series(Pl, ts, nus, lnF, np.full(ts.size, 0.1), np.ones(ts.size), n_threads=cores), reps)
    out["loglike"] = {"value": Pl.size / dt, "unit": UNIT, "cores": cores, "kind": "reference",
                      "sample": f"{Pl.size} walkers of the config-5 workload x {reps} repeats (vagref_chi2_series: Model + "
                                f"flux_density + chi2 per walker, {cores} std::threads)"}
    sizes = {"C2_gaussian_offaxis_200x8": 64, "C3_fs_rs_wind_100x3": 512, "C4_powerlaw_ssc_kn_50x40": 32,
             "C5_gaussian_fs_rs_series100": 128}
    for name, (P, t, nu, series) in extra_workloads(rank).items():
        m = sizes[name]
        fn = (lambda: ref.flux_density_series(P[:m], t, nu, n_threads=cores)) if series else \
             (lambda: ref.flux_density_grid(P[:m], t, nu, n_threads=cores))
        dt = _time_ref(fn, reps)
        out[name] = {"value": m / dt, "unit": UNIT, "cores": cores, "kind": "reference",
                     "sample": f"first {m} parameter sets of the config's batch x {reps} repeats, {cores} std::threads"}
    return out


def run_reference(args):
    """--impl reference: the unmodified reference (oracle/_ref, built from /root/reference by
    oracle/Makefile) on all host cores, same workload/metric; each step = one bounded sample of the batch."""
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
    cfg = config_dict(args, args.gpus)
    cfg["reference_batch_per_step"] = n  # what one reference step actually processed (a rate, so the metric is comparable)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference",
                         "build": REF_BUILD,
                         "sample": f"first {n} parameter sets of the {args.batch}-set batch per step x {args.steps} steps, "
                                   f"std::thread over the unmodified reference (oracle/ref_driver.cpp), {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    try:
        side = reference_side_lines(ref, cores)
        line["loglike"] = side.pop("loglike")
        line["configs"] = side
    except Exception as exc:  # noqa: BLE001
        line["configs"] = {"unavailable": repr(exc)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16384, help="parameter sets per GPU per step")
    ap.add_argument("--ll-batch", type=int, default=4096, help="walkers per GPU per log-likelihood step (config 5)")
    ap.add_argument("--inflight", type=int, default=4, help="independent batches (steps) in flight per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-configs", action="store_true", help="skip the C2 / C3 / C4 / C5-Gaussian side measurements")
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
    numa = bind_to_gpu_numa(local)  # before any pinned allocation / worker thread exists
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from concurrent.futures import ThreadPoolExecutor

    P, t, nu = workload(args.batch, rank)
    n = P.size
    S = max(1, args.inflight)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # One slot = one vag_context (own CUDA stream + HBM workspaces) + its own I/O buffers.  Steps are
    # independent batches, so up to `--inflight` of them are in flight at once: while one batch sits in
    # the dependent-latency-bound ODE kernel (1 warp/SM at 4096 rows) the EATS kernel of another fills
    # the SMs.  Every step still runs the full pipeline on its full batch.
    class Slot:
        def __init__(self):
            self.eng = Engine(local)
            self.eng.set_capacity(256, 128)
            # host-buffer calls transfer only the component planes the batch has (total + forward
            # synchrotron here), as the reference's FluxDict leaves absent components empty
            # ... and, with a single emission component in the batch, `total` (identical to it bit for bit) is not
            # shipped a second time (VAG_OUT_PRESENT_ALIAS_TOTAL, include/vag.h)
            self.eng.set_output_mode(True, alias_total=True)
            self.stream = torch.cuda.Stream(device=dev)
            self.d_p = torch.from_numpy(P.view(np.uint8).copy()).to(dev)
            self.d_t, self.d_nu = torch.from_numpy(t).to(dev), torch.from_numpy(nu).to(dev)
            self.d_out = torch.empty((n, abi.NCOMP, nu.size, t.size), dtype=torch.float64, device=dev)
            self.d_st = torch.zeros(n, dtype=torch.int32, device=dev)
            self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            self.launches = 0
            # pinned host buffers of the end-to-end arm
            self.h_p = torch.from_numpy(P.view(np.uint8).copy()).pin_memory()
            self.h_t, self.h_nu = torch.from_numpy(t.copy()).pin_memory(), torch.from_numpy(nu.copy()).pin_memory()
            self.h_out = torch.empty((n, abi.NCOMP, nu.size, t.size), dtype=torch.float64).pin_memory()
            self.h_st = torch.zeros(n, dtype=torch.int32).pin_memory()

        def step_dev(self):
            with torch.cuda.stream(self.stream):
                self.flush.zero_()  # L2 flush between iterations (256 MiB > 126 MB L2)
            self.eng.flux_density_grid_dev(self.d_p.data_ptr(), n, self.d_t.data_ptr(), t.size, self.d_nu.data_ptr(),
                                           nu.size, self.d_out.data_ptr(), self.d_st.data_ptr(), self.stream.cuda_stream)
            self.launches += self.eng.last_launch_count()

        def step_e2e(self):
            lib = self.eng._lib
            rc = lib.vag_flux_density_grid(self.eng._h, self.h_p.data_ptr(), n, self.h_t.data_ptr(), t.size,
                                           self.h_nu.data_ptr(), nu.size, self.h_out.data_ptr(), self.h_st.data_ptr())
            assert rc == 0, lib.vag_last_error()

    slots = [Slot() for _ in range(S)]
    eng = slots[0].eng
    pool = ThreadPoolExecutor(max_workers=S)

    def run_steps(method, k):
        """k steps, round-robin over the slots; each slot's steps run in order on its own thread."""
        def worker(si):
            for _ in range(si, k, S):
                getattr(slots[si], method)()
        list(pool.map(worker, range(min(S, k))))

    def timed(method, k):
        """Device-timeline duration [ms] from the first launch to the completion of the last kernel."""
        main = torch.cuda.current_stream()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(main)
        for sl in slots:
            sl.stream.wait_event(ev0)
        w0 = time.perf_counter()
        run_steps(method, k)
        for sl in slots:
            e = torch.cuda.Event()
            e.record(sl.stream)
            main.wait_event(e)
        ev1.record(main)
        barrier()
        return ev0.elapsed_time(ev1), (time.perf_counter() - w0) * 1e3

    run_steps("step_dev", args.warmup * S)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for sl in slots:
        sl.launches = 0
    dev_ms, wall_ms_local = timed("step_dev", args.steps)
    launches = sum(sl.launches for sl in slots)
    for sl in slots:
        assert int(sl.d_st.abs().sum()) == 0, "a model reported a status bit"
        assert bool(torch.isfinite(sl.d_out).all())
    wall = wall_ms_local * 1e-3

    # per-stage device times (CUDA events inside the library), sequential, one slot, not overlapped
    eng.set_profiling(True)
    stage_acc = {}
    n_prof = 5
    for _ in range(n_prof):
        slots[0].step_dev()
        torch.cuda.synchronize()
        for k_, v in eng.last_stage_ms().items():
            stage_acc[k_] = stage_acc.get(k_, 0.0) + v
    work_main = eng.last_work()
    eng.set_profiling(False)

    def stage_report(per_ms, wk, n_models, n_t_obs, n_nu_distinct, n_out, fp64_peak, hbm_peak):
        """Per-stage achieved FP64 rate (algorithmic flop of the batch / un-overlapped CUDA-event stage time) and the HBM
        figure of the EATS stage, from the batch's own work counters."""
        aw = algorithmic_work(wk, n_models, n_t_obs, n_nu_distinct, n_out)
        kern = {}
        for k_ in ("grid", "dynamics", "radiation", "eats"):
            if per_ms.get(k_, 0) > 0:
                tf = aw[k_] / (per_ms[k_] * 1e-3) / 1e12
                kern[k_] = {"ms": per_ms[k_], "algorithmic_gflop": aw[k_] / 1e9, "achieved_tflops": tf, "frac": tf / fp64_peak}
        gbs = aw["bytes"] / (per_ms["eats"] * 1e-3) / 1e9 if per_ms.get("eats", 0) > 0 else None
        return kern, aw, gbs

    def measure_config(P_, t_, nu_, series, steps, warm):
        """Device-resident throughput of one more BASELINE.json config on slot 0 (L2 flushed before every step, CUDA
        events on the launching stream), then profiled passes for the per-stage split."""
        sl = slots[0]
        m = P_.size
        d_p = torch.from_numpy(P_.view(np.uint8).copy()).to(dev)
        d_t, d_nu = torch.from_numpy(np.ascontiguousarray(t_)).to(dev), torch.from_numpy(np.ascontiguousarray(nu_)).to(dev)
        shape = (m, abi.NCOMP, t_.size) if series else (m, abi.NCOMP, nu_.size, t_.size)
        d_o = torch.empty(shape, dtype=torch.float64, device=dev)
        d_s = torch.zeros(m, dtype=torch.int32, device=dev)

        def step():
            with torch.cuda.stream(sl.stream):
                sl.flush.zero_()
            if series:
                sl.eng.flux_density_series_dev(d_p.data_ptr(), m, d_t.data_ptr(), d_nu.data_ptr(), t_.size, d_o.data_ptr(),
                                               d_s.data_ptr(), sl.stream.cuda_stream)
            else:
                sl.eng.flux_density_grid_dev(d_p.data_ptr(), m, d_t.data_ptr(), t_.size, d_nu.data_ptr(), nu_.size,
                                             d_o.data_ptr(), d_s.data_ptr(), sl.stream.cuda_stream)
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(sl.stream)
        for _ in range(steps):
            step()
        b.record(sl.stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        st_or = int(torch.bitwise_or(d_s, torch.zeros_like(d_s)).max()) if m else 0
        assert bool(torch.isfinite(d_o).all()), "non-finite flux"
        sl.eng.set_profiling(True)
        acc = {}
        for _ in range(3):
            step()
            torch.cuda.synchronize()
            for k_, v in sl.eng.last_stage_ms().items():
                acc[k_] = acc.get(k_, 0.0) + v / 3
        wk = sl.eng.last_work()
        sl.eng.set_profiling(False)
        return ms, acc, wk, st_or

    # ---- end to end through the host-buffer C-ABI call (pinned host memory) -----------------------
    run_steps("step_e2e", args.warmup * S)
    barrier()
    e0 = time.perf_counter()
    run_steps("step_e2e", args.steps)
    barrier()
    e2e_s = time.perf_counter() - e0
    clocks = sampler.stop() if rank == 0 else None
    h_out, h_st = slots[0].h_out, slots[0].h_st
    alias = eng.last_total_alias()
    assert alias == abi.COMPONENTS.index("fwd_sync"), "forward-shock batch: total aliases the fwd_sync plane"
    assert bool(torch.isfinite(h_out[:, alias]).all()) and float(h_out[:, alias].min()) > 0 and int(h_st.abs().sum()) == 0
    d2h_planes = 1

    # ---- config 5: batched log-likelihood (series chi2), device-resident ---------------------------
    Pl, ts, nus = loglike_workload(args.ll_batch, rank)
    n_ll = Pl.size
    rng = np.random.default_rng(42)
    lnF = np.log(1e-26 * (1 + 0.05 * rng.standard_normal(ts.size)) * (ts / 1e3) ** -1.0)
    sig = np.full(ts.size, 0.1)
    wgt = np.ones(ts.size)
    gathered = torch.empty(n_ll * world, dtype=torch.float64, device=dev) if world > 1 else None
    for sl in slots:
        sl.d_pl = torch.from_numpy(Pl.view(np.uint8).copy()).to(dev)
        sl.d_arr = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (ts, nus, lnF, sig, wgt)]
        sl.d_chi2 = torch.empty(n_ll, dtype=torch.float64, device=dev)
        sl.d_st_ll = torch.zeros(n_ll, dtype=torch.int32, device=dev)

        def step_ll(self=sl):
            self.eng.chi2_series_dev(self.d_pl.data_ptr(), n_ll, *[a.data_ptr() for a in self.d_arr], ts.size,
                                     self.d_chi2.data_ptr(), self.d_st_ll.data_ptr(), self.stream.cuda_stream)
        sl.step_ll = step_ll

    # The only inter-GPU traffic of the path is the gather of float64[n] log-likelihoods after every step.  NCCL
    # operations must be issued in the same order on every rank, so the worker threads (one per in-flight batch)
    # never call the collective themselves: they record an event when their step is enqueued, and THIS thread
    # issues the all-gathers in step order on a separate communication stream that waits on those events.
    S_ll = S
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None

    def run_ll(k):
        if world == 1:
            def worker(si):
                for _ in range(si, k, S_ll):
                    slots[si].step_ll()
            list(pool.map(worker, range(min(S_ll, k))))
            return
        import threading
        enqueued = [threading.Event() for _ in range(k)]
        gather_issued = [threading.Event() for _ in range(k)]
        ev_done = [torch.cuda.Event() for _ in range(k)]
        ev_gathered = [torch.cuda.Event() for _ in range(k)]

        def worker(si):
            for i in range(si, k, S_ll):
                if i >= S_ll:  # this slot's chi2 vector is free once the gather of its previous step has run
                    gather_issued[i - S_ll].wait()
                    slots[si].stream.wait_event(ev_gathered[i - S_ll])
                slots[si].step_ll()
                ev_done[i].record(slots[si].stream)
                enqueued[i].set()

        futs = [pool.submit(worker, si) for si in range(min(S_ll, k))]
        for i in range(k):
            enqueued[i].wait()
            comm_stream.wait_event(ev_done[i])
            with torch.cuda.stream(comm_stream):
                dist.all_gather_into_tensor(gathered, slots[i % S_ll].d_chi2)
            ev_gathered[i].record(comm_stream)
            gather_issued[i].set()
        for f in futs:
            f.result()

    run_ll(args.warmup * S_ll)
    barrier()
    main = torch.cuda.current_stream()
    la, lb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    la.record(main)
    for sl in slots:
        sl.stream.wait_event(la)
    if comm_stream is not None:
        comm_stream.wait_event(la)
    run_ll(args.steps)
    for st_ in [sl.stream for sl in slots] + ([comm_stream] if comm_stream is not None else []):
        e = torch.cuda.Event()
        e.record(st_)
        main.wait_event(e)
    lb.record(main)
    barrier()
    ll_ms = la.elapsed_time(lb)

    # ---- config 5 as stated (north_star): ONE 4096-walker ensemble partitioned over the N GPUs by walker,
    # one evaluation at a time (an MCMC step needs the whole ensemble's log-likelihoods before it can move):
    # per rank a contiguous block of ceil(4096 / N) walkers, then the NCCL all-gather of the chi2 vector.
    from vegasafterglow_b200 import parallel

    n_ens = 4096
    Pe = loglike_workload(n_ens, 0)[0]
    lo, hi = parallel.partition(n_ens, world, rank)
    per_rank = -(-n_ens // world)
    sl0 = slots[0]
    d_pe = torch.from_numpy(Pe[lo:hi].view(np.uint8).copy()).to(dev)
    d_blk = torch.full((per_rank,), float("inf"), dtype=torch.float64, device=dev)
    d_ens = torch.empty(per_rank * world, dtype=torch.float64, device=dev)
    d_st_e = torch.zeros(per_rank, dtype=torch.int32, device=dev)

    def step_ensemble():
        sl0.eng.chi2_series_dev(d_pe.data_ptr(), hi - lo, *[a.data_ptr() for a in sl0.d_arr], ts.size,
                                d_blk.data_ptr(), d_st_e.data_ptr(), sl0.stream.cuda_stream)
        if world > 1:
            with torch.cuda.stream(sl0.stream):
                dist.all_gather_into_tensor(d_ens, d_blk)

    for _ in range(args.warmup):
        step_ensemble()
    barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record(main)
    sl0.stream.wait_event(ea)
    for _ in range(args.steps):
        step_ensemble()
    e = torch.cuda.Event()
    e.record(sl0.stream)
    main.wait_event(e)
    eb.record(main)
    barrier()
    ens_ms = ea.elapsed_time(eb)

    # ---- per-stage split of the config-5 batch, and the other BASELINE.json configs (single-GPU measurements) ------
    eng.set_profiling(True)
    acc_ll = {}
    for _ in range(3):
        slots[0].step_ll()
        torch.cuda.synchronize()
        for k_, v in eng.last_stage_ms().items():
            acc_ll[k_] = acc_ll.get(k_, 0.0) + v / 3
    work_ll = eng.last_work()
    eng.set_profiling(False)
    extra = {}
    if world == 1 and not args.skip_configs:
        for name, (P_, t_, nu_, series) in extra_workloads(rank).items():
            ms_, acc_, wk_, st_ = measure_config(P_, t_, nu_, series, steps=max(3, min(args.steps, 5)), warm=3)
            extra[name] = (P_.size, t_.size, nu_.size, int(np.unique(nu_).size), series, ms_, acc_, wk_, st_)

    # ---- max over ranks ------------------------------------------------------------------------------
    times = torch.tensor([dev_ms, e2e_s * 1e3, ll_ms, wall * 1e3, ens_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, ll_ms, wall_ms, ens_ms = (float(x) for x in times.tolist())

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
        per = {k: v / n_prof for k, v in stage_acc.items()}  # ms per (un-overlapped) step of rank 0
        dominant = max(("grid", "dynamics", "eats"), key=lambda k: per.get(k, 0.0))
        kern, aw, eats_gbs = stage_report(per, work_main, n, t.size, nu.size, nu.size * t.size, fp64_peak, hbm_peak)
        # DRAM bytes of one k_eats launch from the committed `ncu --set full` capture (profiles/), when it was
        # taken at this batch size
        traffic = None
        for tf_name in ("r02_traffic.json", "r01_traffic.json"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", tf_name)))["k_eats_fs_grid"]
                if int(tj["batch"]) == n:
                    traffic = float(tj["dram_bytes_per_launch"])
                    break
            except (OSError, KeyError, ValueError):
                pass
        probe_clock = None
        try:
            q = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=10).stdout.strip()
            probe_clock = float(q.splitlines()[0])
        except Exception:  # noqa: BLE001
            pass
        roofline = {"bound": "hbm", "kernel": "k_eats", "achieved": eats_gbs, "peak": hbm_peak,
                    "unit": "GB/s", "frac": eats_gbs / hbm_peak, "traffic": traffic,
                    "algorithmic_bytes_per_launch": aw["bytes"],
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                    "note": "the path is FP64-pipe / dependent-latency bound, not HBM bound (SURVEY.md 8d): see "
                            "roofline_fp64 for the per-kernel FP64 fractions"}
        roofline_fp64 = {"peak_tflops": fp64_peak,
                         "peak_source": "DFMA probe kernel (k_fp64_peak: 8 independent FMA chains / thread, 8 x 148 CTAs x 256) run "
                                        "in this process right before this line was assembled; MEASURED_PEAKS.json has no FP64 entry",
                         "sm_mhz_after_probe": probe_clock,
                         "times": "per-stage CUDA-event times of ONE batch running alone (un-overlapped, 5 passes averaged); the "
                                  "headline ms_per_step is the overlapped pipeline with --inflight batches and is smaller than their sum",
                         "work": "algorithmic flop of THIS batch from its work counters (vag_last_work) x the SURVEY.md 8d per-unit "
                                 "figures (bench.py algorithmic_work)",
                         "dominant_stage": dominant, "ms_per_step": per, "kernels": kern}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(args, world),
            "e2e": {"value": total_models / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(P.nbytes + t.nbytes + nu.nbytes),
                    "d2h_bytes_per_step": int(n * d2h_planes * nu.size * t.size * 8 + h_st.numel() * 4),
                    "cpu_affinity": numa,
                    "note": "VAG_OUT_PRESENT_ALIAS_TOTAL: of out[n][5][n_nu][n_t] only the fwd_sync plane crosses PCIe -- the three "
                            "absent components (no SSC, no reverse shock in this workload) are not materialised, and `total`, "
                            "identical to the only present component, is returned as an alias (vag_last_total_alias); pinned "
                            "host buffers allocated after binding the process to the GPU's NUMA-local CPUs"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "roofline_fp64": roofline_fp64,
            "loglike": {"metric": f"MCMC loglike evals/s ({n_ll}-walker batch per GPU per step, FS+RS tophat, 100-point "
                                  f"5-band series, {S_ll} steps in flight)",
                        "value": n_ll * world * args.steps / (ll_ms * 1e-3), "unit": UNIT, "ms_per_step": ll_ms / args.steps,
                        "collective": "all_gather float64[n] over NCCL" if world > 1 else "none (1 GPU)",
                        "ensemble_4096": {
                            "what": f"one 4096-walker ensemble split over {world} GPU(s) by walker ({per_rank} per GPU), one "
                                    f"evaluation in flight, chi2 all-gathered over NCCL (strong scaling of config 5)",
                            "ms_per_ensemble": ens_ms / args.steps,
                            "value": n_ens * args.steps / (ens_ms * 1e-3), "unit": UNIT}},
            "wall_ms_per_step": wall_ms / args.steps,
        }
        kern_ll, _, _ = stage_report(acc_ll, work_ll, n_ll, ts.size, 5, ts.size, fp64_peak, hbm_peak)
        line["loglike"]["stages_one_batch_alone"] = kern_ll
        if extra:
            cfgs = {}
            for name, (m_, nt_, nnu_, nnu_d, series, ms_, acc_, wk_, st_) in extra.items():
                kern_, aw_, gbs_ = stage_report(acc_, wk_, m_, nt_, nnu_d if series else nnu_, nt_ if series else nnu_ * nt_,
                                                fp64_peak, hbm_peak)
                cfgs[name] = {"batch": m_, "request": f"{nt_} points series" if series else f"{nt_} t x {nnu_} nu grid",
                              "value": m_ / (ms_ * 1e-3), "unit": UNIT, "ms_per_step": ms_, "status_or": st_,
                              "timing": "device-resident, one batch at a time on one stream, L2 flushed before every step, CUDA events",
                              "rows_per_model": (wk_["rows_fwd"] + wk_["rows_pair"]) / m_,
                              "eats_cells_per_model": wk_["eats_cells"] / m_,
                              "stages": kern_, "eats_hbm": {"achieved_gbs": gbs_, "frac": (gbs_ or 0) / hbm_peak}}
            line["configs"] = cfgs
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import ref  # CPU-baseline leg: allowed to execute oracle/

                if ref.available():
                    cores = ref.hardware_threads()
                    side = reference_side_lines(ref, cores, rank)
                    line["loglike"]["cpu_baseline"] = side.pop("loglike")
                    for name, cb in side.items():
                        if name in line.get("configs", {}):
                            line["configs"][name]["cpu_baseline"] = cb
                    m = min(n, 16384)  # x 3 repeats: ~1.6 s on 16 threads, ~26 core-seconds
                    ref.flux_density_grid(P[:64], t, nu, n_threads=cores)
                    c0 = time.perf_counter()
                    reps = 3
                    for _ in range(reps):
                        ref.flux_density_grid(P[:m], t, nu, n_threads=cores)
                    cdt = (time.perf_counter() - c0) / reps
                    line["cpu_baseline"] = {"value": m / cdt, "unit": UNIT, "cores": cores, "kind": "reference",
                                            "build": REF_BUILD,
                                            "sample": f"first {m} parameter sets of the same batch x {reps} repeats, unmodified "
                                                      f"reference via oracle/_ref/libvagref.so with {cores} std::threads"}
            except Exception as exc:  # noqa: BLE001
                line["cpu_baseline"] = {"unavailable": repr(exc)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
