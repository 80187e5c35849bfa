#!/usr/bin/env python
"""bench.py — agent-steps/sec of the Flockers hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one Schedule::step of the Flockers model: the fused neighbour-gather + boids kernel
over every agent (K4) followed by Field2D::lazy_update (cell-list rebuild: scan + scatter).
N=1 workload: BASELINE config 2 — 1,000,000 agents, 4000x4000 toroidal world, disc 10/1.5,
radius 10, relaxed query, Philox seed 42, synthetic uniform init.
  value     whole-job agent-steps/s, state resident in HBM, CUDA-event time per step, L2 flushed
            between timed steps
  e2e       same metric through kg_field2d_step_boids_host with pinned HOST buffers: every step
            uploads all agents, rebuilds, steps, rebuilds and downloads the result
  roofline  the dominant kernel (K4 step) against the measured HBM peak, algorithmic bytes
            40 B/agent + 8 B/cell per launch
  cpu_baseline  the oracle (C++ restatement of the reference, 1 thread — the reference's
            single-world step is sequential) on a bounded sample of the same workload
--impl reference times that oracle alone (the Rust reference cannot be built in this image).
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
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

DISC = float(np.float32(10.0) / np.float32(1.5))
DENSITY = 10000.0 / (400.0 * 400.0)  # BASELINE config 1: 10k agents in 400x400
SEED = 42
L2_FLUSH_BYTES = 256 << 20
METRIC = "agent-steps/sec (Flockers 1M/64M agents) at 1/2/4/8 B200; % HBM roofline"


def world_for(n_agents):
    return float(np.sqrt(n_agents / DENSITY))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [v.strip() for v in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------- reference arm / CPU
def oracle_rate(n_agents, steps, warmup, budget_s=150.0):
    """agent-steps/s of the oracle (1 thread) on the Flockers workload; shrinks the world at
    constant density when the requested run would not fit the time budget."""
    import oracle_binding as ob
    per_agent_step = 3.8e-6  # seconds, measured on this image's host class at 1M agents
    n = n_agents
    while n > 20000 and n * per_agent_step * (steps + warmup) > budget_s:
        n //= 2
    w = world_for(n)
    m = ob.Flockers(w, w, n, DISC, True, ob.boids_params(radius=10.0, exact=0, seed=SEED))
    m.init()
    if warmup:
        m.step(warmup)
    sec = m.time_steps(steps)
    sample = (f"{n} agents ({w:.0f}x{w:.0f}, same density/geometry) x {steps} Schedule::step after "
              f"{warmup} warm-up, single thread")
    return n * steps / sec, sec, n, sample


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    n_agents = args.agents
    rate, sec, n, sample = oracle_rate(n_agents, args.steps, args.warmup)
    import oracle_binding as ob
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Flockers {n_agents} agents, {world_for(n_agents):.0f}^2 toroidal, "
                               f"disc 10/1.5, radius 10, relax query, seed {SEED}",
                   "note": "restated reference (C++ oracle), Rust toolchain unavailable"},
        "cpu_baseline": {"value": rate, "unit": "agent-steps/s", "cores": 1, "kind": "port",
                         "sample": sample, "host_cores": int(ob.lib().okg_hardware_concurrency())},
        "e2e": {"value": rate, "unit": "agent-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import krabmaga_b200 as kb

    rank, world, local = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = max(world, 1)
    device = local
    torch.cuda.set_device(device)

    n_agents = args.agents
    w = world_for(n_agents)
    params = kb.boids_params(radius=10.0, exact=0, seed=SEED)
    field = kb.Field2D(w, w, DISC, True, capacity=n_agents, device=device)
    field.init_flockers(n_agents, SEED + rank)
    field.lazy_update()
    ncells = field.dw * field.dh

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-state throughput (value)
    params.step = 0
    field.run_boids(params, args.warmup)
    field.sync()
    launches0 = kb._abi.lib().kg_launch_count()
    sampler = ClockSampler(device)
    barrier()
    sampler.start()
    params.step = args.warmup
    ms = field.run_boids_timed(params, args.steps, L2_FLUSH_BYTES)
    barrier()
    clocks = sampler.stop()
    launches = kb._abi.lib().kg_launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = n_agents * n_gpus * args.steps / (ms_max * 1e-3)

    # ---- per-kernel device time (roofline of the dominant kernel), same loop under the profiler
    field.profile(True)
    field.profile_read(reset=True)
    params.step = args.warmup + args.steps
    field.run_boids_timed(params, args.steps, L2_FLUSH_BYTES)
    prof = field.profile_read(reset=True)
    field.profile(False)
    peak, peak_src = measured_peaks()
    step_ms, step_n = prof["step"]
    k4_bytes = 40.0 * n_agents + 8.0 * ncells
    k4_gbs = k4_bytes / (step_ms / step_n * 1e-3) / 1e9 if step_n else 0.0
    kern = {}
    alg = {"step": k4_bytes, "scan": 12.0 * ncells, "scatter": 40.0 * n_agents + 8.0 * ncells}
    total_ms = sum(v[0] for v in prof.values())
    for k, (kms, kn) in prof.items():
        if kn and kms > 0:
            per = kms / (kn if k != "scan" else kn / 3.0)
            kern[k] = {"ms_per_step": kms / args.steps, "share": kms / total_ms,
                       "gbs": alg[k] / (per * 1e-3) / 1e9 if k in alg else None}
    step_alg_bytes = 80.0 * n_agents + 16.0 * ncells
    roofline = {
        "bound": "hbm", "kernel": "step_boids_kernel<relax> (K4: neighbour gather + boids force + "
                                   "position update + histogram)",
        "achieved": k4_gbs, "peak": peak, "unit": "GB/s", "frac": k4_gbs / peak, "traffic": None,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": k4_bytes,
        "whole_step": {"algorithmic_bytes": step_alg_bytes,
                       "achieved": step_alg_bytes * args.steps / (ms_max * 1e-3) / 1e9,
                       "frac": step_alg_bytes * args.steps / (ms_max * 1e-3) / 1e9 / peak},
        "kernels": kern,
        "note": "K4 is FP32-issue bound (two IEEE divisions per candidate pair), see DESIGN.md",
    }

    # ---- e2e through the host-buffer entry point
    e2e = None
    if not args.no_e2e:
        inp = {k: kb._abi.pinned_empty(n_agents, np.uint32 if k == "id" else np.float32)
               for k in ("id", "x", "y", "ldx", "ldy")}
        out = {k: kb._abi.pinned_empty(n_agents, np.uint32 if k == "id" else np.float32)
               for k in ("id", "x", "y", "ldx", "ldy")}
        d = field.download(with_cells=False)
        for k in inp:
            inp[k][:] = d[k]
        f2 = kb.Field2D(w, w, DISC, True, capacity=n_agents, device=device)
        e2e_steps = max(3, min(args.steps, 30))
        e2e_ms = 0.0
        for i in range(3 + e2e_steps):
            params.step = 1000 + i
            f2.l2_flush(L2_FLUSH_BYTES)
            f2.timer_start()
            f2.step_boids_host(params, inp, out)
            dt = f2.timer_stop()
            if i >= 3:
                e2e_ms += dt
            inp, out = out, inp  # the next step consumes this step's host result
        t2 = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": n_agents * n_gpus * e2e_steps / (float(t2.item()) * 1e-3),
               "unit": "agent-steps/s", "h2d_bytes_per_step": 20 * n_agents,
               "d2h_bytes_per_step": 20 * n_agents, "steps": e2e_steps,
               "api": "kg_field2d_step_boids_host (pinned host SoA in, pinned host SoA out)"}
        f2.close()

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        import oracle_binding as ob
        rate, sec, n_cpu, sample = oracle_rate(n_agents, 4, 1, budget_s=30.0)
        cpu = {"value": rate, "unit": "agent-steps/s", "cores": 1, "kind": "port", "sample": sample,
               "host_cores": int(ob.lib().okg_hardware_concurrency())}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"Flockers {n_agents} agents per GPU, {w:.0f}^2 toroidal, disc 10/1.5 "
                                   f"(3x3-cell window), radius 10, relax query, Philox seed {SEED}",
                       "agents": n_agents * n_gpus, "cells": ncells,
                       "parallelism": "single GPU" if n_gpus == 1 else f"{n_gpus} independent replicas",
                       "l2": f"flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write, untimed)",
                       "order": "KG_ORDER_ANY"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    field.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents", type=int, default=1_000_000, help="agents per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
