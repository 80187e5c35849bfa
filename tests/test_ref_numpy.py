"""The C++ oracle against an independent numpy.float32 restatement (tests/ref_numpy.py).

Both were written from the reference sources; they share no code.  Agreement bit for bit breaks
the oracle <-> golden-file circle for SURVEY §8 row G (the reference holds no numeric vectors and
seeds its RNG from the OS, so nothing of the reference's own pins Bird::step's output)."""
import os

import numpy as np
import pytest

import oracle_binding as ob
import ref_numpy as rn

DISC = float(np.float32(10.0) / np.float32(1.5))
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refnumpy_config1_20steps.npz")


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def oracle_run(w, h, disc, tor, agents, nsteps, **kw):
    m = ob.Flockers(w, h, len(agents["id"]), disc, tor, ob.boids_params(**kw), canonical_order=True)
    m.preset(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    m.init()
    m.step(nsteps)
    return dict(zip(("x", "y", "ldx", "ldy"), m.agents()))


def numpy_run(w, h, disc, tor, agents, nsteps, radius=10.0, exact=0, seed=42, **kw):
    wd = rn.World(w, h, disc, tor, seed=seed, radius=radius, exact=bool(exact), **kw)
    wd.preset(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    wd.step(nsteps)
    return wd.by_id()


def assert_same(a, b, what):
    for k in ("x", "y", "ldx", "ldy"):
        bad = np.flatnonzero(bits(a[k]) != bits(b[k]))
        assert len(bad) == 0, f"{what}: {k} differs for ids {bad[:8]} ({a[k][bad[:3]]} vs {b[k][bad[:3]]})"


def agents(ids, x, y, ldx=None, ldy=None):
    n = len(ids)
    return dict(id=np.asarray(ids, np.uint32), x=np.asarray(x, np.float32), y=np.asarray(y, np.float32),
                ldx=np.zeros(n, np.float32) if ldx is None else np.asarray(ldx, np.float32),
                ldy=np.zeros(n, np.float32) if ldy is None else np.asarray(ldy, np.float32))


def test_philox_restatements_agree():
    for ctr, key in [((0, 0, 0, 0), (0, 0)), ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2),
                     ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0))]:
        assert tuple(int(v) for v in ob.philox(ctr, key)) == rn.philox4x32_10(ctr, key)
    # Random123 known-answer vectors
    assert rn.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert rn.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6,
                                                                    0x6D5451FD)


@pytest.mark.parametrize("case", ["wrap_seam", "alone", "at_width", "pair_same_spot", "three_non_unit"])
def test_hand_built_cases(case):
    w = h = 100.0
    kw = {}
    if case == "wrap_seam":       # neighbours on both sides of x = 0 / x = w: the clamped window does NOT wrap
        a = agents([0, 1, 2], [0.5, 99.5, 3.0], [50.0, 50.5, 49.0], [0.7, -0.7, 0.0], [0.0, 0.0, 0.7])
    elif case == "alone":         # count == 0: vec holds only the agent itself
        a = agents([0, 1], [10.0, 80.0], [10.0, 80.0], [0.3, 0.0], [0.4, -0.7])
    elif case == "at_width":      # x == w exactly lands in the padding column; nobody sees it
        a = agents([0, 1, 2], [100.0, 99.0, 98.5], [20.0, 20.0, 21.0], [0.7, 0.0, 0.1], [0.0, 0.7, 0.2])
    elif case == "pair_same_spot":  # dx = dy = 0: zero sums, dis may be 0
        a = agents([0, 1], [40.0, 40.0], [40.0, 40.0])
    else:
        a = agents([5, 9, 2], [30.0, 33.0, 28.5], [30.0, 31.0, 35.0], [0.1, -0.2, 0.3], [0.5, 0.6, -0.7])
        kw = dict(cohesion=1.2, avoidance=0.85, consistency=0.9, randomness=1.4, momentum=0.95)
    for exact in (0, 1):
        if case == "three_non_unit":
            # oracle.agents() indexes by id 0..n-1: relabel to a dense id range, same order
            a = dict(a, id=np.array([1, 2, 0], np.uint32))
        want = oracle_run(w, h, DISC, True, a, 6, radius=10.0, exact=exact, seed=7, **kw)
        got = numpy_run(w, h, DISC, True, a, 6, radius=10.0, exact=exact, seed=7, **kw)
        assert_same(got, want, f"{case} exact={exact}")


def test_non_toroidal_and_wide_window():
    rng = np.random.default_rng(3)
    n, w = 300, 60.0
    a = agents(np.arange(n), rng.random(n) * w, rng.random(n) * w, rng.normal(size=n) * 0.4,
               rng.normal(size=n) * 0.4)
    a["x"] = np.minimum(a["x"], np.float32(59.99))
    a["y"] = np.minimum(a["y"], np.float32(59.99))
    for tor, disc, radius, exact in [(False, 4.0, 10.0, 0), (False, 4.0, 10.0, 1), (True, 2.5, 10.0, 1),
                                     (True, 7.0, 5.0, 0)]:
        want = oracle_run(w, w, disc, tor, a, 3, radius=radius, exact=exact, seed=11)
        got = numpy_run(w, w, disc, tor, a, 3, radius=radius, exact=exact, seed=11)
        assert_same(got, want, f"tor={tor} disc={disc} r={radius} exact={exact}")


def test_config1_20_steps_bit_exact_and_golden():
    """BASELINE config 1 (10,000 agents, 400 x 400, disc 10/1.5, radius 10, relaxed query, seed 42):
    the first 20 steps.  The numpy restatement's result is committed (tests/golden/, generated by
    tests/golden/make_refnumpy_golden.py); the oracle must reproduce it in every bit."""
    gold = np.load(GOLD)
    m = ob.Flockers(400.0, 400.0, 10000, DISC, True, ob.boids_params(radius=10.0, exact=0, seed=42),
                    canonical_order=True)
    m.init()
    x0, y0, _, _ = m.agents()
    assert (bits(x0) == bits(gold["x0"])).all() and (bits(y0) == bits(gold["y0"])).all()   # State::init
    m.step(20)
    want = dict(zip(("x", "y", "ldx", "ldy"), m.agents()))
    assert_same({k: gold[k] for k in want}, want, "config 1, 20 steps")


@pytest.mark.skipif(not os.environ.get("KG_SLOW"), reason="re-derives the golden file (about a minute); KG_SLOW=1")
def test_golden_file_is_what_ref_numpy_computes():
    gold = np.load(GOLD)
    wd = rn.World(400.0, 400.0, DISC, True, seed=42)
    wd.init(10000)
    wd.step(20)
    got = wd.by_id()
    assert_same(got, {k: gold[k] for k in ("x", "y", "ldx", "ldy")}, "golden regeneration")
