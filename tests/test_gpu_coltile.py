"""The column-chunk K4 (KG_K4_COLTILE) against the per-agent packed kernel and the oracle.

It walks an agent's 3x3 window y-outer instead of the reference's x-outer order (field_2d.rs:502-512),
so its f32 sums may differ from the other kernels' in the last bits.  Two kinds of test:
  * inputs for which the order cannot matter (positions and last_d on a coarse binary grid, avoidance
    weight 0): every sum is exact, so the result must equal the packed kernel's BIT FOR BIT — this pins
    the candidate set, the self exclusion, the staging and the epilogue;
  * general inputs: within the north-star's 1e-5 of the packed kernel and of the oracle.
"""
import numpy as np
import pytest

import krabmaga_b200 as kb
from krabmaga_b200 import _abi as abi
from parity_util import NORTH_STAR_DISC, both_params, by_id, random_agents

pytestmark = pytest.mark.gpu


def clustered_agents(n, w, seed, blobs=6, sigma=12.0):
    rng = np.random.default_rng(seed)
    a = random_agents(n, w, w, seed)
    c = rng.random((blobs, 2)) * w
    k = rng.integers(0, blobs, n)
    x = (c[k, 0] + rng.normal(0, sigma, n)) % w
    y = (c[k, 1] + rng.normal(0, sigma, n)) % w
    a["x"] = np.minimum(x.astype(np.float32), np.nextafter(np.float32(w), np.float32(0)))
    a["y"] = np.minimum(y.astype(np.float32), np.nextafter(np.float32(w), np.float32(0)))
    return a


def quantize(a):
    """positions on multiples of 1/64, last_d on multiples of 2^-10: dx, dy and the consistency sums
    are then exact in f32 whatever the order of the additions"""
    for k, q in (("x", 64.0), ("y", 64.0), ("ldx", 1024.0), ("ldy", 1024.0)):
        a[k] = (np.round(a[k].astype(np.float64) * q) / q).astype(np.float32)
    return a


CASES = [  # (name, n, w, maker)
    ("uniform", 20000, 500.0, lambda n, w: random_agents(n, w, w, seed=5)),
    ("north-star density, several chunks per column", 160000, 1600.0, lambda n, w: random_agents(n, w, w, seed=6)),
    ("clustered (crowded chunks take the global path)", 30000, 500.0, lambda n, w: clustered_agents(n, w, 7)),
    ("sparse (chunks spanning more rows than the tables hold)", 1500, 3000.0,
     lambda n, w: random_agents(n, w, w, seed=8)),
    ("tiny world", 300, 30.0, lambda n, w: random_agents(n, w, w, seed=9)),
]


def edge_agents(a, w, tiny=1e-7):
    """a few agents on the world's edges: x == w / y == w sit in the padding column / row; x < 2^-20
    takes the full-division lane path"""
    a["x"][:6] = [w, w, 0.0, tiny, w / 2, np.nextafter(np.float32(w), np.float32(0))]
    a["y"][:6] = [w, 3.0, w, 5.0, w, 0.0]
    return a


def one_step(variant, a, w, gp, cap_mult=1):
    f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=len(a["id"]) * cap_mult)
    f.set_kernel_variant(variant)
    f.set_object_locations(a["id"], a["x"], a["y"], a["ldx"], a["ldy"])
    f.lazy_update()
    gp.step = 5
    f.step_boids(gp)
    f.lazy_update()
    out = by_id(f.download())
    f.close()
    return out


@pytest.mark.parametrize("name,n,w,maker", CASES, ids=[c[0] for c in CASES])
def test_coltile_is_bit_exact_when_the_sums_are_exact(name, n, w, maker):
    a = edge_agents(quantize(maker(n, w)), w, tiny=0.0)
    _, gp = both_params(exact=0, seed=21, avoidance=0.0, cohesion=1.3, consistency=0.7, randomness=1.7,
                        momentum=1.1, jump=0.65)
    want = one_step(abi.KG_K4_AUTO, a, w, gp)
    got = one_step(abi.KG_K4_COLTILE, a, w, gp)
    for k in want:
        bad = np.flatnonzero(got[k].view(np.uint32) != want[k].view(np.uint32))
        assert len(bad) == 0, f"{k}: {len(bad)} of {n} differ, first ids {bad[:5]}"


@pytest.mark.parametrize("name,n,w,maker", CASES, ids=[c[0] for c in CASES])
def test_coltile_within_1e5_of_the_packed_kernel(name, n, w, maker):
    a = edge_agents(maker(n, w), w)
    _, gp = both_params(exact=0, seed=22)
    want = one_step(abi.KG_K4_AUTO, a, w, gp)
    got = one_step(abi.KG_K4_COLTILE, a, w, gp)
    for k in ("x", "y"):
        diff = np.abs(got[k].astype(np.float64) - want[k].astype(np.float64))
        diff = np.minimum(diff, w - diff)
        assert (diff <= 1e-5 * np.maximum(np.abs(want[k]), 1.0)).all(), float(diff.max())
    # last_d = jump * d/|d| amplifies summation-order noise when the five terms nearly cancel (as in
    # test_gpu_field2d.py's any-order test); dense clusters sum hundreds of candidates per agent
    for k in ("ldx", "ldy"):
        diff = np.abs(got[k].astype(np.float64) - want[k].astype(np.float64))
        assert (diff <= (1e-3 if "clustered" in name else 1e-4)).all(), float(diff.max())
        assert (diff <= 1e-5).mean() >= 0.999, float((diff > 1e-5).mean())


def test_coltile_trajectory_conserves_agents_and_stays_close():
    """40 steps: ids conserved, every agent inside the world, |last_d| = jump, and the flock stays within
    rounding-noise distance of the packed kernel's for the first steps"""
    n, w = 40000, 800.0
    a = random_agents(n, w, w, seed=31)
    _, gp = both_params(exact=0, seed=23)
    fs = {}
    for v in (abi.KG_K4_AUTO, abi.KG_K4_COLTILE):
        f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
        f.set_kernel_variant(v)
        f.set_object_locations(a["id"], a["x"], a["y"], a["ldx"], a["ldy"])
        f.lazy_update()
        gp.step = 0
        f.run_boids(gp, 2)
        fs[v] = (f, by_id(f.download()))
    for k in ("x", "y"):
        diff = np.abs(fs[abi.KG_K4_AUTO][1][k].astype(np.float64) - fs[abi.KG_K4_COLTILE][1][k].astype(np.float64))
        diff = np.minimum(diff, w - diff)
        assert np.quantile(diff, 0.999) <= 1e-4, float(np.quantile(diff, 0.999))
    f = fs[abi.KG_K4_COLTILE][0]
    gp.step = 2
    f.run_boids(gp, 38)
    d = f.download()
    assert sorted(d["id"].tolist()) == list(range(n))
    assert ((d["x"] >= 0) & (d["x"] <= w) & (d["y"] >= 0) & (d["y"] <= w)).all()
    sp = np.hypot(d["ldx"].astype(np.float64), d["ldy"].astype(np.float64))
    assert np.allclose(sp, 0.7, atol=1e-5)
    for v in fs:
        fs[v][0].close()


def test_coltile_in_canonical_order_runs_the_per_agent_kernel():
    """KG_ORDER_CANONICAL promises the reference's summation order: the variant must not apply"""
    n, w = 10000, 400.0
    a = random_agents(n, w, w, seed=41)
    _, gp = both_params(exact=0, seed=24)
    outs = []
    for v in (abi.KG_K4_AUTO, abi.KG_K4_COLTILE):
        f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
        f.set_order(True)
        f.set_kernel_variant(v)
        f.set_object_locations(a["id"], a["x"], a["y"], a["ldx"], a["ldy"])
        f.lazy_update()
        gp.step = 0
        f.run_boids(gp, 5)
        outs.append(by_id(f.download()))
        f.close()
    for k in outs[0]:
        assert (outs[0][k].view(np.uint32) == outs[1][k].view(np.uint32)).all(), k
