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
K4_KERNEL = "step_boids_packed_kernel"   # the K4 that KG_K4_AUTO launches for the BASELINE geometry
METRIC = "agent-steps/sec (Flockers 1M/64M agents) at 1/2/4/8 B200; % HBM roofline"


def world_for(n_agents):
    return float(np.sqrt(n_agents / DENSITY))


def flockers_workload(n_agents, gpus):
    """config.workload of the Flockers line — ONE function for both arms, so that the driver's
    same_config comparison sees identical strings."""
    w = world_for(n_agents)
    if gpus <= 1:
        return (f"Flockers {n_agents} agents, {w:.0f}^2 toroidal world, disc 10/1.5 (3x3-cell window), "
                f"radius 10, relax query, Philox seed {SEED}")
    return (f"Flockers {n_agents} agents in ONE {w:.0f}^2 toroidal world, x-strip decomposed over {gpus} "
            f"GPUs, disc 10/1.5 (3x3-cell window), radius 10, relax query, Philox seed {SEED}")


def blocks_for(step_s, steps, want_s=0.1, cap=64):
    """how many K-step blocks make the timed window at least `want_s` long"""
    return int(max(1, min(cap, np.ceil(want_s / max(step_s * steps, 1e-9)))))


def recorded_traffic(kernel, agents):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    `ncu --set full` capture (profiles/r02_traffic.json), or None if that configuration was not
    captured."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        rec = json.load(f)
    for e in rec.get("captures", []):
        if e["kernel"] == kernel and e.get("agents") == agents:
            return e["dram_bytes_per_launch"]
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region.  NVML is polled from a
    thread every 2 ms (the timed region of a 1M-agent run lasts only milliseconds, far below
    nvidia-smi's own start-up time); nvidia-smi -lms is the fallback when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []
        self.samples, self.stop_flag, self.thread, self.nvml = [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id
            self.h = self._by_bus(pynvml, device)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    @staticmethod
    def _by_bus(pynvml, device):
        import torch
        p = torch.cuda.get_device_properties(device)
        want = f"{p.pci_domain_id:08X}:{p.pci_bus_id:02X}:{p.pci_device_id:02X}.0"
        try:
            return pynvml.nvmlDeviceGetHandleByPciBusId(want.encode())
        except Exception:
            return pynvml.nvmlDeviceGetHandleByIndex(device)

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) \
                    if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, mx, pw, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self.stop_flag = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
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
        if self.nvml is not None:
            self.stop_flag = True
            if self.thread:
                self.thread.join(timeout=1)
            nv = self.nvml
            names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            reasons = sorted({k for _, _, _, rs in self.samples for k, bit in names.items() if rs & bit})
            sm = [v[0] for v in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None,
                    "sm_max_mhz": float(max(v[1] for v in self.samples)) if sm else None,
                    "power_w_max": float(max(v[2] for v in self.samples)) if sm else None,
                    "samples": len(sm), "reasons": reasons, "source": "nvml, 2 ms period"}
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
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def pin_to_gpu_numa(local):
    """N > 1: run this rank on the CPUs next to its GPU, so that the pinned host buffers of the e2e leg are
    allocated on that socket (8 ranks moving 2.5 GB per step through one socket's memory otherwise).
    Best effort: returns what it did for the line's `host_affinity` field."""
    if os.environ.get("KG_BENCH_AFFINITY", "1") in ("0", ""):
        return "off"
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:00.0".encode())
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = sorted(near & allowed)
        if not use:
            return f"none of the GPU's {len(near)} local CPUs is available to this process"
        os.sched_setaffinity(0, use)
        return f"{len(use)} CPUs local to the GPU (of {len(allowed)} allowed)"
    except Exception as e:  # noqa: BLE001 - purely advisory
        return f"unavailable ({type(e).__name__})"


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
    if args.workload != "flockers":
        cpu = ff_cpu_baseline(steps=max(1, min(args.steps, 5))) if args.workload == "forest_fire" \
            else sweep_cpu_baseline(steps=max(1, min(args.steps, 5)))
        line = {"impl": "reference", "metric": FF_METRIC if args.workload == "forest_fire" else SWEEP_METRIC,
                "value": cpu["value"], "unit": cpu["unit"], "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u8" if args.workload == "forest_fire" else "f32",
                "data": "synthetic",
                "config": {"workload": cpu["sample"],
                           "note": "restated reference (C++ oracle), Rust toolchain unavailable"},
                "cpu_baseline": cpu,
                "e2e": {"value": cpu["value"], "unit": cpu["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0
    n_agents = args.agents or (1_000_000 if args.gpus <= 1 else 64_000_000)
    rate, sec, n, sample = oracle_rate(n_agents, args.steps, args.warmup)
    import oracle_binding as ob
    extrapolated = n != n_agents
    if extrapolated:
        sample += (f"; RATE MEASURED ON {n} AGENTS and extrapolated to the {n_agents}-agent workload at "
                   "constant density (per-agent cost of the restated reference does not depend on the world size)")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        # time of one step of the NAMED workload: measured when it was run at full size, else the
        # measured per-agent rate applied to the named population (extrapolated: true)
        "ms_per_step": 1e3 * n_agents / rate, "higher_is_better": True,
        "scaling": "weak" if args.gpus <= 1 else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": flockers_workload(n_agents, args.gpus),
                   "requested_agents": n_agents, "measured_agents": n, "extrapolated": extrapolated,
                   "measured_ms_per_step": 1e3 * sec / args.steps,
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
def clocks_line(sampler):
    return sampler.stop()


def run_ours(args):
    import copy

    import torch
    import torch.distributed as dist

    rank, world, local = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    affinity = pin_to_gpu_numa(local) if world > 1 else None
    rc = 0
    if args.workload == "forest_fire":
        line = run_forest_fire(args, torch, dist, rank, world, local)
    elif args.workload == "sweep":
        line = run_sweep(args, torch, dist, rank, world, local)
    else:
        # the headline: BASELINE config 2 (N=1) / config 3 (N>1)
        line = run_strips(args, torch, dist, rank, world, local) if world > 1 else run_single(args, torch, local)
        # parity leg, outside every timed region: the multi-GPU data plane (peer stores over NVLink,
        # flag protocol, IPC mappings) against one GPU, bit for bit
        parity = None if args.no_parity else run_parity(torch, dist, rank, world, local)
        extra = {}
        if not args.no_extra:
            # configs 4 and 5 folded into the same line, so that the driver's record holds them
            a4 = copy.copy(args)
            a4.steps = max(args.steps, 200)
            ff = run_forest_fire(a4, torch, dist, rank, world, local)
            sw = run_sweep(args, torch, dist, rank, world, local)
            bw = run_block_world(args, dist, rank, world) if world > 1 else None
            if rank == 0:
                keep = ("metric", "value", "unit", "steps", "ms_per_step", "blocks", "scaling", "dtype", "config",
                        "roofline", "cpu_baseline", "e2e", "gpu_launches")
                extra = {"forest_fire": {k: ff[k] for k in keep if k in ff},
                         "sweep": {k: sw[k] for k in keep if k in sw}}
                if bw:
                    extra["block_world"] = bw
        if rank == 0:
            line["parity"] = parity
            line["extra"] = extra
            if affinity:
                line["host_affinity"] = affinity
            line["gpu_launches"] = int(line["gpu_launches"]) + sum(int(v.get("gpu_launches", 0)) for v in extra.values())
            if parity and parity.get("mismatches"):
                rc = 1
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


def run_block_world(args, dist, rank, world):
    """N > 1, informational: the headline world over a 2-D block decomposition (csrc/block.cu) that rank 0 drives
    over every GPU of the box — wall clock, because the exchange goes through the host.  Never fails the line."""
    out = None
    # the other ranks must wait on the CPU: an NCCL barrier would park a spinning kernel on their GPUs, which rank 0
    # is about to use from its own context
    try:
        cpu_group = dist.new_group(backend="gloo")
    except Exception:  # noqa: BLE001 - fall back to the default group (slower leg, still correct)
        cpu_group = None
    dist.barrier(group=cpu_group)
    if rank == 0:
        try:
            import time
            import krabmaga_b200 as kb
            from krabmaga_b200 import blocks
            n = args.agents or 64_000_000
            w = world_for(n)
            nbx = {2: 2, 4: 2, 8: 4}.get(world, world)
            f0 = kb.Field2D(w, w, DISC, True, capacity=n, device=0)
            f0.init_flockers(n, SEED)
            init = f0.download(unbuffered=True, with_cells=False)
            f0.close()
            bw = blocks.BlockWorld(w, w, DISC, 10.0, nbx, world // nbx, list(range(world)), n, slack=1.5)
            bw.upload(init)
            del init
            params = kb.boids_params(radius=10.0, exact=0, seed=SEED)
            params.step = 0
            bw.run_boids(params, 5)
            steps = max(args.steps, 20)
            launches0 = kb._abi.lib().kg_launch_count()
            t0 = time.perf_counter()
            params.step = 5
            bw.run_boids(params, steps)
            dt = time.perf_counter() - t0
            launches = kb._abi.lib().kg_launch_count() - launches0
            bw.close()
            out = {"metric": METRIC, "value": n * steps / dt, "unit": "agent-steps/s", "steps": steps,
                   "gpu_launches": int(launches),
                   "ms_per_step": 1e3 * dt / steps, "timing": "wall clock around kg_blocks_run (host-orchestrated exchange)",
                   "config": {"workload": f"Flockers {n} agents over {nbx} x {world // nbx} blocks on {world} GPUs, "
                                          "one process driving every block"}}
        except Exception as e:  # noqa: BLE001 - informational leg
            out = {"error": f"{type(e).__name__}: {e}"[:300]}
    dist.barrier(group=cpu_group)
    return out


# --------------------------------------------------------------------------- parity legs
def run_parity(torch, dist, rank, world, local, n=1_000_000, steps=30, side=4096, ff_steps=200):
    """N > 1: a 1M-agent world in KG_ORDER_CANONICAL stepped `steps` times over the N real strips
    (one per GPU, migrants and ghosts stored into the neighbours' inboxes over NVLink) against the same
    run on rank 0's single field, compared bit for bit after an all-gather of (id, x, y, ldx, ldy);
    and a side^2 Forest-Fire grid over N row strips against rank 0's single grid.  Checks the seam
    semantics of field_2d.rs:472-516 (clamped window: no wrap halo) across real devices.
    N = 1: the default K4 against the generic, reference-shaped kernel on the same world."""
    import krabmaga_b200 as kb
    from krabmaga_b200 import gridstrips, strips

    w = world_for(n)
    params = kb.boids_params(radius=10.0, exact=0, seed=SEED)

    def single(variant=None):
        f = kb.Field2D(w, w, DISC, True, capacity=n, device=local)
        f.set_order(True)
        if variant is not None:
            f.set_kernel_variant(variant)
        f.init_flockers(n, SEED)
        f.lazy_update()
        params.step = 0
        f.run_boids(params, steps)
        d = f.download(with_cells=False)
        f.close()
        o = np.argsort(d["id"], kind="stable")
        return {k: v[o] for k, v in d.items()}

    def diff(got, want):
        if len(got["id"]) != len(want["id"]) or (got["id"] != want["id"]).any():
            return max(1, abs(len(got["id"]) - len(want["id"])))
        return int(sum((got[k].view(np.uint32) != want[k].view(np.uint32)).sum() for k in ("x", "y", "ldx", "ldy")))

    def blocks_run(devices):
        """the same world over a 2-D block decomposition (csrc/block.cu), driven by this process"""
        from krabmaga_b200 import blocks
        nb = len(devices) if len(devices) > 1 else 4
        nbx = {2: 2, 4: 2, 8: 4}.get(nb, nb)
        f0 = kb.Field2D(w, w, DISC, True, capacity=n, device=local)
        f0.init_flockers(n, SEED)
        init = f0.download(unbuffered=True, with_cells=False)
        f0.close()
        bw = blocks.BlockWorld(w, w, DISC, 10.0, nbx, nb // nbx, devices, n, canonical_order=True, slack=2.0)
        bw.upload(init)
        params.step = 0
        bw.run_boids(params, steps)
        d = bw.download()
        bw.close()
        o = np.argsort(d["id"], kind="stable")
        return {k: v[o] for k, v in d.items()}, f"{nbx} x {nb // nbx} blocks on {len(set(devices))} GPU(s)"

    if world == 1:
        want = single(kb._abi.KG_K4_GENERIC)
        bad = diff(single(), want)
        got_b, shape = blocks_run([local])
        bad_b = diff(got_b, want)
        return {"checked": True, "mismatches": bad + bad_b, "blocks_mismatches": bad_b,
                "what": f"Flockers {n} agents x {steps} steps, KG_ORDER_CANONICAL: default K4 (packed kernel) vs "
                        f"the generic reference-shaped kernel, every f32 bit of x, y, last_d: {bad} differing words; "
                        f"the same world over {shape}: {bad_b}; oracle parity is the -m gpu test suite and smoke()"}

    cap, hcap, mcap = strips.default_capacities(n, w, w, DISC, 10.0, world, slack=2.0)
    st = strips.StripField2D(w, w, DISC, 10.0, rank, world, cap, hcap, mcap, device=local)
    st.set_order(True)
    strips.connect_ipc(st, dist)
    st.init_flockers(n, SEED)
    dist.barrier()
    st.prepare()
    params.step = 0
    st.run_boids(params, steps)
    st.sync()
    mine = st.download()
    stats = st.stats()
    dist.barrier()
    st.close()
    parts = [None] * world
    dist.all_gather_object(parts, {k: np.ascontiguousarray(v) for k, v in mine.items()})
    mig = [None] * world
    dist.all_gather_object(mig, (stats["migrants_out"], stats["halo_left"] + stats["halo_right"]))

    g = gridstrips.StripDenseNumberGrid2D(side, side, rank, world, device=local)
    gridstrips.connect_ipc(g, dist)
    g.init_forest_fire(0.6, SEED)
    dist.barrier()
    g.prepare()
    dist.barrier()
    g.run_stencil(ff_steps)
    g.sync()
    rows = g.download()
    dist.barrier()
    g.close()
    grids = [None] * world
    dist.all_gather_object(grids, rows)

    out = None
    if rank == 0:
        got = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
        o = np.argsort(got["id"], kind="stable")
        got = {k: v[o] for k, v in got.items()}
        want = single()
        bad = diff(got, want)
        got_b, shape = blocks_run(list(range(world)))   # rank 0 drives every GPU of the box for this leg
        bad_b = diff(got_b, want)
        ref = kb.DenseNumberGrid2D(side, side, device=local)
        ref.init_forest_fire(0.6, SEED)
        ref.run_stencil(ff_steps)
        bad_ff = int((np.concatenate(grids, axis=0) != ref.download()).sum())
        ref.close()
        out = {"checked": True, "mismatches": bad + bad_ff + bad_b,
               "what": f"Flockers {n} agents x {steps} steps KG_ORDER_CANONICAL over {world} strips on {world} "
                       f"GPUs vs rank 0's single field, every f32 bit of (id, x, y, ldx, ldy): {bad} differing "
                       f"words; the same world over {shape}: {bad_b}; Forest Fire {side}^2 x {ff_steps} steps over "
                       f"{world} row strips vs one grid: {bad_ff} differing cells",
               "flockers_mismatches": bad, "blocks_mismatches": bad_b, "forest_fire_mismatches": bad_ff,
               "migrants_out_total": int(sum(m[0] for m in mig)), "halo_agents_last_step": int(sum(m[1] for m in mig))}
    dist.barrier()
    return out


def run_single(args, torch, device):
    import krabmaga_b200 as kb

    n_agents = args.agents or 1_000_000
    w = world_for(n_agents)
    params = kb.boids_params(radius=10.0, exact=0, seed=SEED)
    field = kb.Field2D(w, w, DISC, True, capacity=n_agents, device=device)
    field.init_flockers(n_agents, SEED)
    field.lazy_update()
    ncells = field.dw * field.dh
    flush = 0 if args.no_flush else L2_FLUSH_BYTES

    # ---- resident-state throughput (value)
    params.step = 0
    field.run_boids(params, args.warmup)
    field.sync()
    # EXACTLY args.steps steps per timed block; the block is repeated until the timed window is
    # >= 100 ms (a 1M-agent step lasts tens of microseconds) and the median block is reported
    params.step = args.warmup
    est = field.run_boids_timed(params, 3, flush) / 3 * 1e-3
    nblocks = blocks_for(est, args.steps)
    sampler = ClockSampler(device)
    torch.cuda.synchronize()
    sampler.start()
    block_ms = []
    launches0 = kb._abi.lib().kg_launch_count()
    for b in range(nblocks):
        params.step = args.warmup + 3 + b * args.steps
        block_ms.append(field.run_boids_timed(params, args.steps, flush))
        if b == 0:
            launches = kb._abi.lib().kg_launch_count() - launches0
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = float(np.median(block_ms))
    value = n_agents * args.steps / (ms * 1e-3)

    # ---- per-kernel device time (roofline of the dominant kernel), same loop under the profiler
    field.profile(True)
    field.profile_read(reset=True)
    params.step = args.warmup + 3 + nblocks * args.steps
    field.run_boids_timed(params, max(args.steps, 50), flush)
    prof = field.profile_read(reset=True)
    field.profile(False)
    peak, peak_src = measured_peaks()
    step_ms, step_n = prof["step"]
    k4_bytes = 40.0 * n_agents + 8.0 * ncells
    k4_gbs = k4_bytes / (step_ms / step_n * 1e-3) / 1e9 if step_n else 0.0
    alg = {"step": k4_bytes, "scan": 8.0 * ncells, "scatter": 40.0 * n_agents + 8.0 * ncells}
    total_ms = sum(v[0] for v in prof.values())
    kern = {}
    for k, (kms, kn) in prof.items():
        if kn and kms > 0:
            kern[k] = {"us_per_launch": 1e3 * kms / kn, "launches": int(kn), "share": kms / total_ms,
                       "gbs": alg[k] / (kms / kn * 1e-3) / 1e9 if k in alg else None}
    step_alg_bytes = 80.0 * n_agents + 16.0 * ncells
    whole = step_alg_bytes * args.steps / (ms * 1e-3) / 1e9
    # K4's real ceiling is the FP32 pipe: per candidate pair 21 f32 lane-operations
    # that the reference's arithmetic fixes (IEEE division = reciprocal + 2 refinements + 3 per
    # numerator), ~25.0 candidates per agent at this density, ~150 more per agent after the loop;
    # B200: 148 SMs x 128 lanes x 1.965 GHz.
    cand = 9.0 * n_agents / (field.max_x * field.max_y)
    lane_ops = n_agents * (21.0 * cand + 200.0)
    # 120 lane-ops/clk/SM is what scalar and two-lane FP32 instructions sustain on this part
    # (tools/probe/pipe_probe.cu, profiles/r02_fp32_pipe_probe.jsonl); nominal is 128
    fp32_peak = 148 * 120 * 1.965e9
    k4_s = step_ms / step_n * 1e-3 if step_n else float("inf")
    roofline = {
        "bound": "hbm", "kernel": "step_boids_packed_kernel (K4: neighbour gather + boids force + "
                                   "position update + histogram)",
        "achieved": k4_gbs, "peak": peak, "unit": "GB/s", "frac": k4_gbs / peak,
        "traffic": recorded_traffic(K4_KERNEL, n_agents),
        "peak_source": peak_src, "algorithmic_bytes_per_launch": k4_bytes,
        "whole_step": {"algorithmic_bytes": step_alg_bytes, "achieved": whole, "frac": whole / peak},
        "kernels": kern,
        "fp32_pipe": {"lane_ops_per_launch": lane_ops, "achieved_tlops": lane_ops / k4_s / 1e12,
                      "peak_tlops": fp32_peak / 1e12, "frac": lane_ops / k4_s / fp32_peak,
                      "candidates_per_agent": cand},
        "note": "K4 is bound by FP32 lane throughput and lane occupancy, not by HBM: the candidate loop runs "
                "on two-lane FADD2/FMUL2/FFMA2 instructions, which halve issue slots but not pipe cycles "
                "(DESIGN.md §3, profiles/r02_ncu_k4_packed.txt, r02_fp32_pipe_probe.jsonl); fp32_pipe.peak is the "
                "probe's measured 120 lane-ops/clk/SM; kernels[] are isolated per-launch times from a profiled "
                "pass, the step itself overlaps them (dependent launches)",
    }

    # ---- e2e through the host-buffer entry point
    e2e = None
    if not args.no_e2e:
        keys = ("id", "x", "y", "ldx", "ldy")
        inp = {k: kb._abi.pinned_empty(n_agents, np.uint32 if k == "id" else np.float32) for k in keys}
        out = {k: kb._abi.pinned_empty(n_agents, np.uint32 if k == "id" else np.float32) for k in keys}
        d = field.download(with_cells=False)
        order = np.argsort(d["id"], kind="stable")      # agent i at index i: init_flockers' ids are 0 .. n-1
        for k in inp:
            inp[k][:] = d[k][order]
        f2 = kb.Field2D(w, w, DISC, True, capacity=n_agents, device=device)
        e2e_steps = max(3, min(args.steps, 30))

        def timed(call, a, b):
            total = 0.0
            for i in range(3 + e2e_steps):
                params.step = 1000 + i
                f2.l2_flush(flush)
                f2.timer_start()
                call(params, a, b)
                dt = f2.timer_stop()
                if i >= 3:
                    total += dt
                a, b = b, a  # the next step consumes this step's host result
            return total

        # headline: the by-position entry — the host keeps its agents in arrays (agent i = id i, as Flockers'
        # State::init numbers them), sends x, y, last_d and gets them back in place: 16 B per agent each way
        # (the four arrays of a side are slices of ONE pinned block, which the entry moves as one copy)
        def block4(src):
            blk = kb._abi.pinned_empty(4 * n_agents, np.float32)
            views = {k: blk[j * n_agents:(j + 1) * n_agents] for j, k in enumerate(("x", "y", "ldx", "ldy"))}
            if src is not None:
                for k in views:
                    views[k][:] = src[k]
            views["_block"] = blk           # keeps the pinned allocation alive
            return views
        pos_in, pos_out = block4(inp), block4(None)
        ms_ordered = timed(f2.step_boids_host_ordered, pos_in, pos_out)
        # the keyed entry (ids travel both ways, result in cell order): 20 B per agent each way
        ms_keyed = timed(f2.step_boids_host, inp, out)
        e2e = {"value": n_agents * e2e_steps / (ms_ordered * 1e-3), "unit": "agent-steps/s",
               "h2d_bytes_per_step": 16 * n_agents, "d2h_bytes_per_step": 16 * n_agents,
               "steps": e2e_steps,
               "api": "kg_field2d_step_boids_host_ordered (pinned host x, y, last_d in; the same arrays out, "
                      "agent i at index i, ids implicit)",
               "keyed": {"value": n_agents * e2e_steps / (ms_keyed * 1e-3), "unit": "agent-steps/s",
                         "h2d_bytes_per_step": 20 * n_agents, "d2h_bytes_per_step": 20 * n_agents,
                         "api": "kg_field2d_step_boids_host (ids travel both ways, result in cell order)"}}
        f2.close()

    # ---- CPU baseline: bounded sample of the same workload
    cpu = None
    if not args.no_cpu_baseline:
        import oracle_binding as ob
        rate, sec, n_cpu, sample = oracle_rate(n_agents, 4, 1, budget_s=30.0)
        cpu = {"value": rate, "unit": "agent-steps/s", "cores": 1, "kind": "port", "sample": sample,
               "host_cores": int(ob.lib().okg_hardware_concurrency())}

    # the N>1 runs use the 64M-agent world (BASELINE config 3): time it on this one GPU too, so
    # that strong-scaling efficiency can be computed against the same workload
    scaling_ref = None
    if not args.no_scaling_ref and not args.agents:
        field.close()
        n64 = 64_000_000
        w64 = world_for(n64)
        f64 = kb.Field2D(w64, w64, DISC, True, capacity=n64, device=device)
        f64.init_flockers(n64, SEED)
        f64.lazy_update()
        params.step = 0
        f64.run_boids(params, 3)
        params.step = 3
        ms64 = f64.run_boids_timed(params, 10, 0)
        scaling_ref = {"workload": f"Flockers {n64} agents on ONE GPU (the world the N>1 runs cut into strips)",
                       "value": n64 * 10 / (ms64 * 1e-3), "unit": "agent-steps/s", "ms_per_step": ms64 / 10,
                       "steps": 10}
        f64.close()
        field = None

    line = {
        "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "blocks": {"count": nblocks, "ms": block_ms, "reported": "median block / steps"},
        "config": {"workload": flockers_workload(n_agents, 1),
                   "agents": n_agents, "cells": ncells, "parallelism": "single GPU",
                   "l2": "not flushed" if args.no_flush else
                         f"flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write, untimed)",
                   "order": "KG_ORDER_ANY"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "scaling_reference": scaling_ref,
    }
    if field is not None:
        field.close()
    return line


def run_strips(args, torch, dist, rank, world, local):
    """N > 1: BASELINE config 3 — one world of 64M agents cut into N x-strips, one rank per GPU,
    halo exchange and migration over NVLink peer memory (no collective on the data path)."""
    import krabmaga_b200 as kb
    from krabmaga_b200 import strips

    n_total = args.agents or 64_000_000
    w = world_for(n_total)
    params = kb.boids_params(radius=10.0, exact=0, seed=SEED)
    cap, hcap, mcap = strips.default_capacities(n_total, w, w, DISC, 10.0, world, slack=1.25)
    strip = strips.StripField2D(w, w, DISC, 10.0, rank, world, cap, hcap, mcap, device=local)
    strips.connect_ipc(strip, dist)
    strip.init_flockers(n_total, SEED)
    dist.barrier()
    strip.prepare()
    flush = 0 if args.no_flush else L2_FLUSH_BYTES
    max_x, _, dw, dh = strips.grid_dims(w, w, DISC)
    ncells = dw * dh

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    params.step = 0
    strip.run_boids(params, args.warmup)
    strip.sync()
    params.step = args.warmup
    barrier()
    est = reduce_max(strip.run_boids_timed(params, 3, flush)) / 3 * 1e-3
    nblocks = blocks_for(est, args.steps)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    block_ms = []
    launches0 = strip.stats()["launches"]
    for b in range(nblocks):
        params.step = args.warmup + 3 + b * args.steps
        barrier()
        mine = strip.run_boids_timed(params, args.steps, flush)
        barrier()
        block_ms.append(reduce_max(mine))      # device time of the block, max over ranks
        if b == 0:
            launches = strip.stats()["launches"] - launches0
    clocks = sampler.stop()
    st = strip.stats()
    ms_max = float(np.median(block_ms))
    value = n_total * args.steps / (ms_max * 1e-3)
    owned = torch.tensor([st["n_owned"], st["migrants_out"], st["halo_left"] + st["halo_right"]],
                         dtype=torch.float64, device="cuda")
    gathered = [torch.zeros_like(owned) for _ in range(world)]
    dist.all_gather(gathered, owned)

    # ---- e2e: every step each rank uploads the agents it owns from pinned host memory, the strips
    # rebuild + exchange halos, step once, and each rank downloads what it owns afterwards
    e2e = None
    if not args.no_e2e:
        keys = ("id", "x", "y", "ldx", "ldy")
        bufs = [{k: kb._abi.pinned_empty(cap, np.uint32 if k == "id" else np.float32) for k in keys}
                for _ in range(2)]
        cur = strip.download(out=bufs[0])
        e2e_steps = max(3, min(args.steps, 10))
        e2e_ms, h2d, d2h = 0.0, 0, 0
        for i in range(2 + e2e_steps):
            params.step = 1000 + i
            dist.barrier()
            strip.timer_start()
            strip.clear()
            strip.upload(cur["id"], cur["x"], cur["y"], cur["ldx"], cur["ldy"])
            n_up = len(cur["id"])
            strip.prepare()
            strip.step_boids(params)
            cur = strip.download(out=bufs[(i + 1) % 2])
            dt = strip.timer_stop()
            if i >= 2:
                e2e_ms += dt
                h2d += 20 * n_up
                d2h += 20 * len(cur["id"])
        e2e_max = reduce_max(e2e_ms)
        tot = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot)
        e2e = {"value": n_total * e2e_steps / (e2e_max * 1e-3), "unit": "agent-steps/s",
               "h2d_bytes_per_step": int(tot[0].item() / e2e_steps),
               "d2h_bytes_per_step": int(tot[1].item() / e2e_steps), "steps": e2e_steps,
               "api": "kg_strip_clear/upload/prepare/step_boids/download per rank (pinned host SoA)"}

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        step_alg_bytes = 80.0 * n_total + 16.0 * ncells
        whole = step_alg_bytes * args.steps / (ms_max * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "blocks": {"count": nblocks, "ms": block_ms, "reported": "median block (max over ranks) / steps"},
            "config": {"workload": flockers_workload(n_total, world),
                       "agents": n_total, "cells": ncells,
                       "parallelism": f"{world} x-strips, per-step halo (line) + migration (ring) "
                                      "by peer stores over NVLink, no collective",
                       "per_rank": [{"owned": int(g[0]), "migrants_out_total": int(g[1]),
                                     "halo_agents": int(g[2])} for g in gathered],
                       "l2": "not flushed" if args.no_flush else
                             f"flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write, untimed)",
                       "note": "N=1 runs BASELINE config 2 (1M agents); N>1 runs config 3 (64M total)"},
            "roofline": {"bound": "hbm", "kernel": "whole step (K4 + exchange + rebuild), all ranks",
                         "achieved": whole, "peak": peak * world, "unit": "GB/s",
                         "frac": whole / (peak * world), "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": step_alg_bytes},
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
    dist.barrier()
    strip.close()
    return line if rank == 0 else None


# --------------------------------------------------------------------------- config 4: Forest Fire
FF_METRIC = "cell-updates/sec (Forest Fire on DenseNumberGrid2D, u8 states); % HBM roofline"


def ff_cpu_baseline(steps=3, side=4096):
    import oracle_binding as ob
    o = ob.ForestFire(side, side)
    o.init(0.6, SEED)
    o.step(20)                       # let the front enter the grid
    sec = o.time_steps(steps)
    return {"value": side * side * steps / sec, "unit": "cell-updates/s", "cores": 1, "kind": "port",
            "sample": f"{side}x{side} grid (same rule, same init) x {steps} steps after 20 warm-up, single thread",
            "host_cores": int(ob.lib().okg_hardware_concurrency())}


def ff_roofline(gbs, peak, peak_src, world, cells, steps, launches):
    """2 algorithmic bytes per cell-step.  run_stencil fuses up to 8 steps into one pass over
    the grid (forest_fire_u8_multi_kernel: step t in, step t+T out, the levels between in registers), so a
    launch processes T x cells units while HBM sees each cell once in and once out: `achieved` (algorithmic
    bytes / time, the contract's definition) exceeds the HBM peak by design; `hbm_frac_of_pass_bytes` is the
    share of the peak the bytes actually moved amount to."""
    per_launch = max(1.0, steps / max(1, launches))
    fused = per_launch > 1.5
    kernel = (f"forest_fire_u8_multi_kernel (K5 on bit planes, {per_launch:.1f} steps per launch per GPU)" if fused
              else "forest_fire_u8_kernel (K5), one launch per step per GPU")
    out = {"bound": "hbm", "kernel": kernel, "achieved": gbs, "peak": peak * world, "unit": "GB/s",
           "frac": gbs / (peak * world),
           "traffic": recorded_traffic("forest_fire_u8_multi_kernel" if fused else "forest_fire_u8_kernel", cells)
           if world == 1 else None,
           "peak_source": peak_src, "algorithmic_bytes_per_launch": 2.0 * cells / world * per_launch,
           "steps_per_launch": per_launch}
    if fused:
        out["hbm_frac_of_pass_bytes"] = gbs / per_launch / (peak * world)
        out["note"] = ("temporal blocking: one pass reads step t and writes step t+T (T = 8, 4 or 2), so frac > 1 "
                       "against the per-step algorithmic bytes; hbm_frac_of_pass_bytes counts 2 B per cell per PASS")
    return out


def run_forest_fire(args, torch, dist, rank, world, local):
    """BASELINE config 4: 32768 x 32768 u8 grid, Moore-8 rule, strong scaling over row strips."""
    import krabmaga_b200 as kb
    from krabmaga_b200 import gridstrips

    side = args.side or 32768
    cells = side * side

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    strip = gridstrips.StripDenseNumberGrid2D(side, side, rank, world, device=local)
    gridstrips.connect_ipc(strip, dist if world > 1 else None)
    strip.init_forest_fire(0.6, SEED)
    barrier()
    strip.prepare()
    barrier()
    strip.run_stencil(args.warmup)
    strip.sync()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = kb._abi.lib().kg_launch_count()
    ms = strip.run_stencil_timed(args.steps)
    launches = kb._abi.lib().kg_launch_count() - launches0
    barrier()
    clocks = sampler.stop()
    ms_max = reduce_max(ms)
    value = cells * args.steps / (ms_max * 1e-3)

    # e2e: host grid in, one step, host grid out (each rank moves its own rows), per step
    e2e = None
    if not args.no_e2e:
        own = (strip.x1 - strip.x0) * side
        hin = kb._abi.pinned_empty(own, np.uint8)
        hout = kb._abi.pinned_empty(own, np.uint8)
        strip.download(out=hin)
        e2e_steps = max(3, min(args.steps, 5))
        tot = 0.0
        for i in range(2 + e2e_steps):
            barrier()
            t0 = time.perf_counter()
            strip.upload(hin)
            if world > 1:
                dist.barrier()
            strip.prepare()
            if world > 1:
                dist.barrier()
            strip.run_stencil(1)
            strip.download(out=hout)          # straight into page-locked memory
            dt = time.perf_counter() - t0
            if i >= 2:
                tot += dt
            hin, hout = hout, hin             # the next step consumes this step's host result
        tot = reduce_max(tot)
        e2e = {"value": cells * e2e_steps / tot, "unit": "cell-updates/s",
               "h2d_bytes_per_step": cells, "d2h_bytes_per_step": cells, "steps": e2e_steps,
               "api": "kg_gridstrip_upload/prepare/run_stencil(1)/download per rank, page-locked host buffers both "
                      "ways, host wall clock (includes the synchronous copies)"}

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        gbs = 2.0 * cells * args.steps / (ms_max * 1e-3) / 1e9
        line = {
            "metric": FF_METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": f"Forest Fire {side}x{side} u8 grid, Moore-8 rule, density 0.6, "
                                   f"column 0 burning, Philox seed {SEED}",
                       "cells": cells,
                       "parallelism": "single GPU" if world == 1 else
                                      f"{world} row strips, halo rows pushed by the stencil kernel over NVLink",
                       "l2": f"working set {2 * cells / world / 2**20:.0f} MiB per GPU per step "
                             "(inputs larger than the 126 MB L2, no flush needed)"},
            "roofline": ff_roofline(gbs, peak, peak_src, world, cells, args.steps, launches),
            "cpu_baseline": None if (args.no_cpu_baseline or world > 1) else ff_cpu_baseline(),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
    barrier()
    strip.close()
    return line


# --------------------------------------------------------------------------- config 5: explore sweep
SWEEP_METRIC = "agent-steps/sec (explore sweep: independent Flockers replicas of 16k agents); % HBM roofline"


def sweep_cpu_baseline(n=16384, steps=3):
    import oracle_binding as ob
    cores = int(ob.lib().okg_hardware_concurrency())
    reps = max(cores, 8)
    sec, work = ob.flockers_sweep(world_for(n), world_for(n), DISC, True, n,
                                  ob.boids_params(radius=10.0, exact=0, seed=SEED), reps, steps, 0)
    return {"value": work / sec, "unit": "agent-steps/s", "cores": cores, "kind": "port",
            "sample": f"{reps} replicas x {n} agents x {steps} steps, one replica per host thread "
                      f"(explore_parallel!'s rayon fan-out), {cores} threads",
            "host_cores": cores}


def run_sweep(args, torch, dist, rank, world, local):
    """BASELINE config 5: 4096 independent replicas of 16,384 agents (512x512 world each), dealt to
    the GPUs round-robin; replicas only, no communication."""
    import krabmaga_b200 as kb

    total_reps = args.replicas or 4096
    n = 16384
    w = world_for(n)
    mine = len(range(rank, total_reps, world))
    params = [kb.boids_params(radius=10.0, exact=0, seed=SEED + r) for r in range(rank, total_reps, world)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    b = kb.FlockerBatch((w, w), n, mine, DISC, True, params, device=local)
    b.init()
    b.run(args.warmup)
    b.sync()
    flush = 0 if args.no_flush else L2_FLUSH_BYTES
    est = reduce_max(b.run_timed(3, flush)) / 3 * 1e-3
    nblocks = blocks_for(est, args.steps, cap=8)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    block_ms = []
    launches0 = kb._abi.lib().kg_launch_count()
    for blk in range(nblocks):
        barrier()
        mine_ms = b.run_timed(args.steps, flush)
        barrier()
        block_ms.append(reduce_max(mine_ms))
        if blk == 0:
            launches = kb._abi.lib().kg_launch_count() - launches0
    clocks = sampler.stop()
    ms_max = float(np.median(block_ms))
    agents = total_reps * n
    value = agents * args.steps / (ms_max * 1e-3)

    e2e = None
    if not args.no_e2e:
        # one whole sweep through the public entry point: configurations in, result rows out
        e2e_reps = min(total_reps, 1024 * world)   # enough work to amortise allocation and init
        seeds = [SEED + r for r in range(rank, e2e_reps, world)]
        dts = []
        for _ in range(3):                         # median of three whole sweeps (the first pays allocator warm-up)
            barrier()
            t0 = time.perf_counter()
            rows = kb.explore_parallel(args.steps, 1, (w, w), n, DISC, {"seed": seeds},
                                       mode=kb.ExploreMode.Matched, devices=(local,))
            dts.append(reduce_max(time.perf_counter() - t0))
        dt = float(np.median(dts))
        e2e = {"value": e2e_reps * n * args.steps / dt, "unit": "agent-steps/s",
               "h2d_bytes_per_step": 48 * e2e_reps // max(args.steps, 1),
               "d2h_bytes_per_step": 64 * e2e_reps // max(args.steps, 1), "steps": args.steps,
               "api": f"explore_parallel({e2e_reps} replicas x {args.steps} steps): create + init + run + "
                      "device-side reductions + output rows, host wall clock, median of 3 sweeps",
               "sweep_seconds": dts, "rows": len(rows)}

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        ncells = 78 * 78 * total_reps
        step_alg_bytes = 80.0 * agents + 16.0 * ncells
        whole = step_alg_bytes * args.steps / (ms_max * 1e-3) / 1e9
        line = {
            "metric": SWEEP_METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"explore sweep: {total_reps} independent Flockers replicas of {n} agents "
                                   f"({w:.0f}^2 toroidal each), disc 10/1.5, radius 10, relax query, "
                                   f"seed {SEED}+replica",
                       "agents": agents, "replicas_per_gpu": mine,
                       "parallelism": "replicas only: replica i -> GPU i % G, no communication",
                       "l2": "not flushed" if args.no_flush else
                             f"flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write, untimed)"},
            "roofline": {"bound": "hbm", "kernel": "whole step (batch K4 + scan + scatter), all ranks",
                         "achieved": whole, "peak": peak * world, "unit": "GB/s",
                         "frac": whole / (peak * world), "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": step_alg_bytes},
            "cpu_baseline": None if (args.no_cpu_baseline or world > 1) else sweep_cpu_baseline(),
            "blocks": {"count": nblocks, "ms": block_ms, "reported": "median block (max over ranks) / steps"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
    barrier()
    b.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents", type=int, default=0,
                    help="total agents (default: 1M at N=1, 64M at N>1 as in BASELINE.json)")
    ap.add_argument("--workload", default="flockers", choices=["flockers", "forest_fire", "sweep"],
                    help="flockers = BASELINE configs 2/3 (the headline metric; default), "
                         "forest_fire = config 4, sweep = config 5")
    ap.add_argument("--side", type=int, default=0, help="forest_fire: grid side (default 32768)")
    ap.add_argument("--replicas", type=int, default=0, help="sweep: total replicas (default 4096)")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true",
                    help="flockers: do not fold the forest_fire / sweep results (configs 4, 5) into the line")
    ap.add_argument("--no-parity", action="store_true", help="flockers: skip the parity leg")
    ap.add_argument("--no-scaling-ref", action="store_true",
                    help="N=1: skip the extra 64M-agent single-GPU timing")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
