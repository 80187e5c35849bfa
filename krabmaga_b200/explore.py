"""explore mirror: the reference's replica fan-out for parameter sweeps, with the replicas batched
on the device instead of spread over rayon tasks.

explore_parallel   <- explore_parallel!(nstep, rep_conf, State, input{..}, output[..], mode)
                      src/explore/model_exploration.rs:354-423
explore_sequential <- explore_sequential!   :232-312 (same rows, one replica per launch)
explore_distributed <- explore_distributed_mpi!  src/explore/mpi/model_exploration.rs:121-290
                      (configurations dealt to ranks, rows gathered on the root)
ExploreMode        <- src/lib.rs:481-487 (Exaustive [sic] = cartesian product, Matched = zip)
shard              <- replica i -> device i % G; mirrors explore/mpi/model_exploration.rs:206, 217

Rows follow build_dataframe!'s FrameRow (:451-540): (conf_num, conf_rep, *inputs, *outputs,
run_duration, step_per_sec).  Inputs are the Flockers fixture's swept quantities — any field of
KgBoidsParams (cohesion, avoidance, randomness, consistency, momentum, jump, radius, seed) and the
state's own constructor arguments `dim` ((w, h) tuples) and `initial_flockers`: explore_parallel!
builds an arbitrary state per configuration (model_exploration.rs:387-410).  The replicas of one
device batch share world and population, so runs are grouped by (dim, initial_flockers) and every
group is advanced as its own batches; the rows come back in run order regardless.

field_names / write_csv <- the DataFrame trait and write_csv, src/lib.rs:1781-1800: header =
the row's field names, one record per row, every value through its string form, "<name>.csv".
"""
import csv
import enum
import itertools
import time

import numpy as np

from . import _abi as abi
from .batch import FlockerBatch

PARAM_FIELDS = ("cohesion", "avoidance", "randomness", "consistency", "momentum", "jump", "radius",
                "seed")
STATE_FIELDS = ("dim", "initial_flockers")     # Flocker::new(dim, initial_flockers), state.rs:24-32
INPUT_FIELDS = PARAM_FIELDS + STATE_FIELDS


class ExploreMode(enum.Enum):
    Exaustive = 0   # spelled as in the reference
    Matched = 1


def build_configurations(inputs, mode):
    """List of dicts, one per configuration.  Exaustive: cartesian product with the FIRST input
    varying slowest, the order build_configurations! produces (src/lib.rs:1724-1748)."""
    names = list(inputs)
    for n in names:
        if n not in INPUT_FIELDS:
            raise ValueError(f"unknown input {n!r}; the batched Flockers state takes {INPUT_FIELDS}")
    if mode == ExploreMode.Exaustive:
        combos = itertools.product(*[inputs[n] for n in names])
    else:
        lens = {len(inputs[n]) for n in names}
        if len(lens) > 1:
            raise ValueError("Matched mode needs inputs of equal length")
        combos = zip(*[inputs[n] for n in names])
    return [dict(zip(names, c)) for c in combos]


def deal_runs(n_runs, n_devices, max_per_batch):
    """Which runs go where: run i -> device i % G (round-robin, as the reference's MPI variant deals
    configurations to ranks, explore/mpi/model_exploration.rs:206), each device's share cut into
    batches of at most `max_per_batch`.  Returns [(device_slot, [run indices])...]."""
    out = []
    for g in range(n_devices):
        mine = list(range(g, n_runs, n_devices))
        for lo in range(0, len(mine), max_per_batch):
            out.append((g, mine[lo:lo + max_per_batch]))
    return out


def run_seed(seed, conf_num, rep):
    """Philox key of one run: SplitMix64 of (seed, configuration, repetition).  Every run is its own
    stream, as with the reference's rand::rng() per run (model_exploration.rs:387-410): no two
    configurations share placement or noise, and (seed s, rep 1) is not (seed s + 1, rep 0)."""
    z = (int(seed) * 0x9E3779B97F4A7C15 + int(conf_num) * 0xBF58476D1CE4E5B9 + int(rep) * 0x94D049BB133111EB
         + 0x2545F4914F6CDD1D) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def _params_for(conf, conf_num, rep, base_seed):
    kw = dict(radius=10.0, exact=0, seed=base_seed)
    kw.update({k: v for k, v in conf.items() if k in PARAM_FIELDS})
    p = abi.boids_params(**kw)
    p.seed = run_seed(kw["seed"], conf_num, rep)
    return p


def default_outputs(red):
    """Example output column from the device-side reductions (FlockerBatch.reduce): the polarisation
    |mean last_d| / JUMP of each replica (order parameter).  An output function marked
    `from_reduce = True` receives the per-replica sums (replicas x 64 bytes cross the bus); any other
    receives the downloaded population, dict of [replicas, n] arrays, as before."""
    vx, vy = red["sum_ldx"] / red["n"], red["sum_ldy"] / red["n"]
    return {"polarisation": np.sqrt(vx * vx + vy * vy) / 0.7}


default_outputs.from_reduce = True


def _run_runs(runs, confs, nstep, dim, initial_flockers, discretization, outputs, devices, toroidal,
              max_replicas_per_batch, base_seed, canonical_order):
    """Rows for `runs` = [(conf_num, conf_rep)...], in that order: the runs are dealt to `devices`
    round-robin and advanced as batches of at most `max_replicas_per_batch` replicas."""
    rows = [None] * len(runs)
    # runs whose configuration names its own world / population form their own batches
    groups = {}
    for k, (i, _) in enumerate(runs):
        d = confs[i].get("dim", dim)
        key = ((float(d[0]), float(d[1])), int(confs[i].get("initial_flockers", initial_flockers)))
        groups.setdefault(key, []).append(k)
    for (gdim, gn), members in groups.items():
        for g, sub in deal_runs(len(members), len(devices), max_replicas_per_batch):
            chunk = [members[j] for j in sub]
            params = [_params_for(confs[runs[k][0]], runs[k][0], runs[k][1], base_seed) for k in chunk]
            b = FlockerBatch(gdim, gn, len(chunk), discretization, toroidal, params,
                             device=devices[g], canonical_order=canonical_order)
            b.init()
            b.sync()
            t0 = time.perf_counter()
            b.run(nstep)
            b.sync()
            dt = time.perf_counter() - t0
            if not outputs:
                out = {}
            elif getattr(outputs, "from_reduce", False):
                out = outputs(b.reduce())       # sums computed on the device: no population download
            else:
                out = outputs(b.download())
            b.close()
            for j, k in enumerate(chunk):
                i, r = runs[k]
                # the replicas of a batch run concurrently: each row reports the batch's wall time
                rows[k] = dict(conf_num=i, conf_rep=r, **confs[i], effective_seed=int(params[j].seed),
                               **{name: float(col[j]) for name, col in out.items()},
                               run_duration=dt, step_per_sec=nstep / dt)
    return rows


def explore_parallel(nstep, rep_conf, dim, initial_flockers, discretization, inputs,
                     mode=ExploreMode.Matched, outputs=default_outputs, devices=(0,), toroidal=True,
                     max_replicas_per_batch=4096, base_seed=42, canonical_order=False):
    """Runs n_conf * rep_conf independent simulations of `nstep` steps and returns the rows.

    The runs are dealt to `devices` round-robin (run i -> devices[i % G]); each device advances its
    share as batches of at most `max_replicas_per_batch` replicas.  `canonical_order` sorts every
    bag by id (KG_ORDER_CANONICAL): results then do not depend on how the runs were batched."""
    confs = build_configurations(inputs, mode)
    runs = [(i, r) for i in range(len(confs)) for r in range(rep_conf)]   # run / rep_conf, run % rep_conf
    return _run_runs(runs, confs, nstep, dim, initial_flockers, discretization, outputs, devices, toroidal,
                     max_replicas_per_batch, base_seed, canonical_order)


def explore_distributed(nstep, rep_conf, dim, initial_flockers, discretization, inputs,
                        mode=ExploreMode.Matched, outputs=default_outputs, device=0, toroidal=True,
                        max_replicas_per_batch=4096, base_seed=42, canonical_order=False, group=None,
                        root=0):
    """explore_distributed_mpi! (src/explore/mpi/model_exploration.rs:121-290) over
    torch.distributed, one process per GPU: every rank builds the same configuration table,
    takes configurations rank, rank + world, ... with all their repetitions (`i*num_procs +
    my_rank`, :206, :217), runs them as batches on its own device, and `root` gathers the rows
    (gather_varcount_into_root, :247-283).  Replicas only — no data-path collective.  Returns the rows ordered by (conf_num,
    conf_rep) on `root`, None elsewhere."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    confs = build_configurations(inputs, mode)
    runs = [(i, r) for i in range(rank, len(confs), world) for r in range(rep_conf)]
    rows = _run_runs(runs, confs, nstep, dim, initial_flockers, discretization, outputs, (device,), toroidal,
                     max_replicas_per_batch, base_seed, canonical_order)
    parts = [None] * world if rank == root else None
    dist.gather_object(rows, parts, dst=root, group=group)
    if rank != root:
        return None
    return sorted((row for part in parts for row in part), key=lambda r: (r["conf_num"], r["conf_rep"]))


def explore_sequential(nstep, rep_conf, dim, initial_flockers, discretization, inputs,
                       mode=ExploreMode.Matched, outputs=default_outputs, device=0, toroidal=True,
                       base_seed=42, canonical_order=False):
    """Same rows with one replica per launch (explore_sequential!, :232-312)."""
    return explore_parallel(nstep, rep_conf, dim, initial_flockers, discretization, inputs, mode,
                            outputs, (device,), toroidal, 1, base_seed, canonical_order)


def field_names(rows):
    """DataFrame::field_names (src/lib.rs:1797): column names in FrameRow declaration order —
    conf_num, conf_rep, inputs, outputs, run_duration, step_per_sec."""
    return list(rows[0].keys()) if rows else []


def write_csv(name, rows):
    """write_csv(name, &dataframe) (src/lib.rs:1781-1792): writes `<name>.csv`, header first, then
    one record per row with every field formatted as a string.  Returns the path.  An empty
    dataframe still creates the file (the reference opens the Writer before looking at the rows)."""
    path = f"{name}.csv"
    names = field_names(rows)
    with open(path, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(names)
        for row in rows:
            if list(row.keys()) != names:
                raise ValueError("rows of one dataframe must share their fields")
            w.writerow([_field_str(row[k]) for k in names])
    return path


def _field_str(v):
    # Rust's Display for floats prints the shortest string that round-trips; repr() does the same
    if isinstance(v, (float, np.floating)):
        return repr(float(v))
    return str(int(v)) if isinstance(v, (int, np.integer)) else str(v)
