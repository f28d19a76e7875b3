#!/usr/bin/env python
"""bench.py — MPC QP solves/sec of the batched B200 tracking step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libcudampc.so)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, restated (oracle/), all host cores

A "step" is one pass of the hot path over one batch: `batch` independent horizon-50 tracking QPs
(BASELINE.json configs[2]: 65,536 problems, steering-rate limit +-0.02 rad/step) per GPU.  Under torchrun
each rank solves its own shard of one seeded global batch (weak scaling, no collective on the data path);
torch.distributed is used only to gather the per-rank times and counters.  One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MPC QP solves/sec (horizon 50, batched)"
UNIT = "solves/s"
HORIZON = 50
BATCH = 65536
SEED = 3
DU_DELTA = 0.02            # rad/step, configs[2] "tight steering-rate limits"
EPS = 1e-6                 # parity setting of BASELINE.json (u0 within 1e-5 at eps_abs = eps_rel = 1e-6)
POLISH_PASSES = 5
POLISH_RETRY = 4           # with 4, every problem of the bench batch ends on a polished KKT point (2 leaves 1 of 65,536)
EARLY_POLISH = 1          # finish as soon as a polish certifies a KKT point of a settled active set

# canonical flop model of BASELINE.md §2 / SURVEY.md §8d (n = 11N+5, m = 19N+7, nnz(A) = 43N+5)
_NNZ_L = {15: 706, 20: 941, 50: 2349}
# DRAM traffic per solve of K_solve from the ncu --set full capture profiles/ncu_full_r01_solve_summary.txt
# (dram__bytes_read.sum + dram__bytes_write.sum = 8.49 + 56.18 MB for a 4,096-problem launch): dominated by the
# warm-start / polish back-up of the ADMM iterate (12 KB per save), not by the 4.1 KB of algorithmic I/O.
TRAFFIC_BYTES_PER_SOLVE = 64.67e6 / 4096


def flop_model(N: int):
    n, m, nnzA = 11 * N + 5, 19 * N + 7, 43 * N + 5
    nnzL = _NNZ_L.get(N, int(round(47 * N + 1)))
    f_iter = 4 * nnzL + 4 * nnzA + 8 * n + 10 * m
    f_chol = 121 * n
    return dict(f_iter=f_iter, f_chol=f_chol, f_solve=4 * nnzL, f_lin=40 * N)


def flops_of_batch(N, iters, n_fac, n_solves):
    """flops = n_fac*F_chol + iters*F_iter + (polish triangular solves)*4 nnz(L) + F_lin, from the kernel's counters."""
    fm = flop_model(N)
    it, nf, ns = float(np.sum(iters)), float(np.sum(n_fac)), float(np.sum(n_solves))
    return nf * fm["f_chol"] + it * fm["f_iter"] + max(ns - it, 0.0) * fm["f_solve"] + len(iters) * fm["f_lin"]


def io_bytes(N, B):
    """algorithmic bytes per solve: in 32N+80, out 48N+56."""
    return B * (32 * N + 80), B * (48 * N + 56)


def shard_range(rank: int, world: int, per_rank: int):
    """Static contiguous split of the global batch index."""
    return rank * per_rank, per_rank


def product_params(N=HORIZON, du=DU_DELTA):
    from rrt_mpc_b200 import MPCConfig
    p = MPCConfig(horizon=N).to_parameters(0.8)
    return dataclasses.replace(p, du_bounds=((-12.0, 12.0), (-du, du)))


def gather_metrics(local: dict, world: int):
    """All-gather a flat dict of floats (summary metrics only)."""
    if world == 1:
        return [local]
    import torch
    import torch.distributed as dist
    keys = sorted(local)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(local[k]) for k in keys], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [dict(zip(keys, o.cpu().tolist())) for o in out]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        # under load = upper half of the samples (the sampler also sees the idle gaps between steps)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's restatement of the reference path, one process per host core
# ----------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    lo, hi, N, du, seed, batch, eps, scaling = args
    from oracle import c_oracle as CO
    from oracle import mpc_numpy as O
    from rrt_mpc_b200.synthetic import make_batch
    p = dataclasses.replace(O.Params(horizon=N), du_bounds=((-12.0, 12.0), (-du, du)))
    x0, ref, up = make_batch(batch, N, seed, start=lo, count=hi - lo)
    CO.lib()
    t = time.perf_counter()
    r = CO.solve_batch(p, x0, ref, up, eps_abs=eps, eps_rel=eps, scaling=scaling, polish_passes=1)
    dt = time.perf_counter() - t
    return dt, int((r["status"] == 1).sum()), float(r["iters"].mean())


def cpu_baseline(sample: int, N=HORIZON, eps=EPS, scaling=10, cores=None):
    """Cold solves (the reference rebuilds its problem every call, mpc_controller.py:119) of the first `sample`
    problems of the bench batch with the restated OSQP path at OSQP's default scaling; input generation excluded."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    per = max(1, sample // cores)
    jobs = [(i * per, (i + 1) * per, N, DU_DELTA, SEED, BATCH, eps, scaling) for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    tmax = max(r[0] for r in res)
    n = per * cores
    return {"value": n / tmax, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n} of the {BATCH} bench problems, cold start each, eps {eps:g}, OSQP-equivalent C restatement "
                      f"(Ruiz scaling {scaling}, polish), one process per core, {tmax:.1f} s",
            "solved": int(sum(r[1] for r in res)), "mean_iters": float(np.mean([r[2] for r in res]))}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (restated; cvxpy/osqp are absent offline)."""
    if rank != 0:
        return
    vals = []
    sample = args.cpu_sample
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(sample)
        if i >= args.warmup:
            vals.append(cb)
    v = float(np.mean([c["value"] for c in vals]))
    ms = 1e3 * sample / v
    cb = dict(vals[-1]); cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{BATCH} independent tracking QPs, horizon {HORIZON}, steering-rate +-{DU_DELTA} (BASELINE configs[2]); "
                                   f"each step = a bounded sample of {sample} problems on all host cores", "eps": EPS},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="problems per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=2048, help="problems in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--latency-launches", type=int, default=100, help="launches of one resident wave for the latency percentiles")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the MPC hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from rrt_mpc_b200 import MPCController, SolverSettings
    from rrt_mpc_b200.synthetic import make_batch

    N, B = HORIZON, args.batch
    warmup = max(3, args.warmup)
    params = product_params()
    settings = SolverSettings(eps_abs=EPS, eps_rel=EPS, polish_passes=POLISH_PASSES, polish_retry=POLISH_RETRY, early_polish=bool(EARLY_POLISH))
    start, count = shard_range(rank, world, B)
    x0, ref, up = make_batch(B * world, N, SEED, start=start, count=count)
    ctl = MPCController(params, settings, device=local_rank, max_batch=B)
    dev = torch.device("cuda", local_rank)
    d_x0, d_ref, d_up = (torch.as_tensor(a).to(dev) for a in (x0, ref, up))
    pin = lambda a: torch.as_tensor(a).pin_memory().numpy()
    h_x0, h_ref, h_up = pin(x0), pin(ref), pin(up)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)
    peak_tf = ctl.fp64_peak_tflops()

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    for _ in range(warmup):
        res = ctl.solve_batch(d_x0, d_ref, u_prev=d_up)
    sync()

    # ---- device-resident timing (value) ----------------------------------------------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = ctl.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync()
    t_wall = time.perf_counter()
    for e0, e1 in ev:
        flush.zero_()                       # L2 flush between timed iterations (untimed)
        e0.record(stream)
        res = ctl.solve_batch(d_x0, d_ref, u_prev=d_up)
        e1.record(stream)
    sync()
    t_wall = time.perf_counter() - t_wall
    launches = ctl.launch_count() - launches0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    clk = clocks.stop() if rank == 0 else None

    iters = res.iters.cpu().numpy(); info = res.info.cpu().numpy(); status = res.status.cpu().numpy()
    flops = flops_of_batch(N, iters, info[:, 1], info[:, 3])

    # ---- same batch, OSQP-literal termination (polish only after the ADMM residual test; no early polish) -----------
    lit = SolverSettings(eps_abs=EPS, eps_rel=EPS, polish_passes=POLISH_PASSES, polish_retry=POLISH_RETRY, early_polish=False)
    ctl.solve_batch(d_x0, d_ref, u_prev=d_up, settings=lit)
    sync()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_()
    l0.record(stream); lres = ctl.solve_batch(d_x0, d_ref, u_prev=d_up, settings=lit); l1.record(stream)
    sync()
    lit_ms = l0.elapsed_time(l1)
    lit_iters = lres.iters.cpu().numpy(); lit_info = lres.info.cpu().numpy()
    lit_flops = flops_of_batch(N, lit_iters, lit_info[:, 1], lit_info[:, 3])
    u0_gap = float((lres.u0 - res.u0).abs().max().item())

    # ---- per-launch latency of one resident wave (BASELINE metric: p99 per-step latency) -----------------------------
    wave = min(B, 148 * ctl.problems_per_sm())
    lat = []
    for i in range(args.latency_launches):
        o = (i * wave) % max(1, B - wave + 1)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream); ctl.solve_batch(d_x0[o:o + wave], d_ref[o:o + wave], u_prev=d_up[o:o + wave]); a1.record(stream)
        torch.cuda.synchronize(dev)
        lat.append(a0.elapsed_time(a1))
    lat = np.array(lat[5:]) if len(lat) > 10 else np.array(lat)

    # ---- end-to-end through the public API with host buffers ---------------------------------------
    ctl.pinned_outputs = True            # page-locked result arrays, reused per call (inputs are pinned above)
    for _ in range(1):
        ctl.solve_batch(h_x0, h_ref, u_prev=h_up)
    sync()
    e2e_t = []
    for _ in range(max(2, min(args.steps, 3))):
        t = time.perf_counter()
        hres = ctl.solve_batch(h_x0, h_ref, u_prev=h_up)
        e2e_t.append(time.perf_counter() - t)
    sync()
    h2d = B * (4 + 4 * (N + 1) + 2) * 8
    d2h = B * (2 + 4 * (N + 1) + 2 * N + 2) * 8 + B * 6 * 4
    assert np.array_equal(hres.status, status)

    local = {"ms_total": float(np.sum(step_ms)), "ms_max_step": float(np.max(step_ms)), "solves": float(B * args.steps), "flops": flops,
             "solved": float((status == 1).sum()), "iters_mean": float(iters.mean()), "iters_max": float(iters.max()),
             "e2e_s": float(np.mean(e2e_t)), "launches": float(launches), "polished": float((info[:, 2] > 0).sum()),
             "n_fac_mean": float(info[:, 1].mean()), "wall_s": t_wall, "lit_ms": lit_ms, "lit_solves": float(B)}
    allm = gather_metrics(local, world)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    t_ms = max(m["ms_total"] for m in allm)                     # max over ranks
    total_solves = sum(m["solves"] for m in allm)
    value = total_solves / (t_ms * 1e-3)
    ms_per_step = t_ms / args.steps
    e2e_value = (B * world) / max(m["e2e_s"] for m in allm)
    # roofline of the dominant (only) kernel, per launch on rank 0: canonical flops / CUDA-event duration
    ach_tf = flops / (float(np.mean(step_ms)) * 1e-3) / 1e12
    inb, outb = io_bytes(N, B)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_ach = (inb + outb) / (float(np.mean(step_ms)) * 1e-3) / 1e9
    cb = None
    if not args.no_cpu_baseline:
        cb = cpu_baseline(args.cpu_sample)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "p99_ms_per_step": float(np.percentile(step_ms, 99)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{B} independent tracking QPs per GPU, horizon {N}, 4 states / 2 controls, steering-rate +-{DU_DELTA} rad/step "
                               f"(BASELINE.json configs[2]), seed {SEED}", "batch_per_gpu": B, "horizon": N, "eps_abs": EPS, "eps_rel": EPS,
                   "polish_passes": POLISH_PASSES, "polish_retry": POLISH_RETRY, "early_polish": EARLY_POLISH, "parallelism": f"{world} x independent shards, no data-path collective",
                   "l2": "working set per step (inputs 112 MB + outputs 161 MB + 803 MB warm-start state) exceeds the 126 MB L2; "
                         "a 256 MB write flushes L2 between timed steps"},
        "solve_stats": {"solved_frac": sum(m["solved"] for m in allm) / (B * world), "iters_mean": float(np.mean([m["iters_mean"] for m in allm])),
                        "iters_max": max(m["iters_max"] for m in allm), "polished_frac": sum(m["polished"] for m in allm) / (B * world),
                        "factorisations_mean": float(np.mean([m["n_fac_mean"] for m in allm])),
                        "problems_per_sm": ctl.problems_per_sm(), "smem_doubles_per_problem": ctl.workspace_doubles()},
        "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf if peak_tf > 0 else None,
                     "traffic": TRAFFIC_BYTES_PER_SOLVE * B, "traffic_unit": "bytes per launch (ncu r01, scaled per solve)",
                     "note": "binding roofline is the non-tensor fp64 pipe (SURVEY.md 8d); achieved = canonical flops (BASELINE.md model, from the "
                             "kernel's own iteration/factorisation counters) / CUDA-event time of one launch; peak = DFMA throughput measured "
                             "live by cudampc_fp64_peak_tflops (MEASURED_PEAKS.json has no fp64 figure)",
                     "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                             "algorithmic_bytes_per_launch": inb + outb, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        "osqp_literal": {"value": sum(m["lit_solves"] for m in allm) / (max(m["lit_ms"] for m in allm) * 1e-3), "unit": UNIT,
                         "iters_mean": float(lit_iters.mean()), "achieved_tflops": lit_flops / (lit_ms * 1e-3) / 1e12,
                         "frac": lit_flops / (lit_ms * 1e-3) / 1e12 / peak_tf if peak_tf > 0 else None, "max_abs_u0_gap_vs_early_polish": u0_gap,
                         "note": "same kernel with early_polish off: ADMM runs until the eps 1e-6 residual test passes, then polishes"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "latency": {"batch": int(wave), "launches": int(len(lat)), "p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)),
                    "note": "one launch of exactly one resident wave of problems (SMs x problems/SM), device-resident, rank 0"},
        "gpu_launches": int(sum(m["launches"] for m in allm)),
        "clocks": clk,
        "cpu_baseline": cb,
        "wall_s_timed_region": max(m["wall_s"] for m in allm),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
