"""kg_field2d_step_custom — a model's own Agent::step as run-time compiled snippets (csrc/jit_agent.cuh):
  * Bird::step written as snippets must reproduce the built-in generic kernel bit for bit (pins the restated
    window walk, toroidal helpers and random stream of the generated source);
  * a model the library does not ship (drift to the neighbours' centroid, crowded agents stop) against a numpy
    restatement built on tests/ref_numpy.py's independent Field2D."""
import numpy as np
import pytest

import krabmaga_b200 as kb
import ref_numpy as rn
from custom_models import BIRD_FINISH, BIRD_PAIR, CENTROID_FINISH, CENTROID_PAIR
from krabmaga_b200 import _abi as abi
from parity_util import NORTH_STAR_DISC, both_params, by_id, random_agents

pytestmark = pytest.mark.gpu
F = np.float32

CASES = [  # (w, disc, toroidal, n, radius, exact, steps)
    (400.0, NORTH_STAR_DISC, True, 6000, 10.0, 0, 12),     # north-star geometry, relaxed
    (400.0, NORTH_STAR_DISC, True, 6000, 10.0, 1, 12),     # exact query (bird.rs:41)
    (10.0, 0.5, True, 64, 10.0, 1, 20),                    # the fixture: 41 x 41 cell window
    (300.0, NORTH_STAR_DISC, False, 2500, 10.0, 0, 10),    # non-toroidal field: the window wraps
    (90.0, 4.5, True, 2000, 10.0, 1, 10),                  # 5 x 5 window
]


@pytest.mark.parametrize("w,d,tor,n,radius,exact,steps", CASES)
def test_bird_step_as_snippets_equals_the_generic_kernel(w, d, tor, n, radius, exact, steps):
    agents = random_agents(n, w, w, seed=n + exact)
    agents["x"][:4] = [0.0, 1e-7, w - 1e-4, w / 2]
    agents["y"][:4] = [1e-8, 0.0, 3.0, w - 1e-4]
    weights = dict(cohesion=1.3, avoidance=0.9, randomness=1.7, consistency=0.7, momentum=1.1, jump=0.65)
    _, gp = both_params(radius=radius, exact=exact, seed=19, **weights)
    consts = [weights[k] for k in ("cohesion", "avoidance", "randomness", "consistency", "momentum", "jump")]
    out = []
    for custom in (False, True):
        f = kb.Field2D(w, w, d, tor, capacity=n)
        f.set_order(True)
        f.set_kernel_variant(abi.KG_K4_GENERIC)
        f.set_object_locations(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
        f.lazy_update()
        for s in range(steps):
            if custom:
                f.step_custom(BIRD_PAIR, BIRD_FINISH, consts, radius=radius, exact=exact, seed=19, step=s)
            else:
                gp.step = s
                f.step_boids(gp)
            f.lazy_update()
        out.append(by_id(f.download()))
        f.close()
    for k in out[0]:
        bad = np.flatnonzero(out[0][k].view(np.uint32) != out[1][k].view(np.uint32))
        assert len(bad) == 0, f"{k}: {len(bad)} of {n} differ, first ids {bad[:5]}"


def centroid_step(world, speed, limit):
    """the CENTROID snippets in numpy.float32, one agent after the other, on ref_numpy's field"""
    nx, ny, na, nb, keep = world.x.copy(), world.y.copy(), world.ldx.copy(), world.ldy.copy(), np.ones(len(world.x), bool)
    for k in range(len(world.x)):
        me, px, py = int(world.ids[k]), world.x[k], world.y[k]
        ax = ay = F(0)
        cnt = 0
        for e in world.neighbors(px, py, world.radius, world.exact):
            if int(world.ids[e]) != me:
                cnt += 1
                ax = ax + rn.toroidal_distance(px, world.x[e], world.w)
                ay = ay + rn.toroidal_distance(py, world.y[e], world.h)
        mx = my = F(0)
        if cnt > 0:
            mx, my = -ax / F(cnt), -ay / F(cnt)
        ln = np.sqrt(mx * mx + my * my)
        if ln > 0:
            mx, my = mx / ln * F(speed), my / ln * F(speed)
        nx[k], ny[k] = rn.toroidal_transform(px + mx, world.w), rn.toroidal_transform(py + my, world.h)
        na[k], nb[k] = mx, my
        keep[k] = not (F(cnt) > F(limit))
    return nx, ny, na, nb, keep


@pytest.mark.parametrize("exact", [0, 1])
def test_a_model_of_its_own_against_numpy(exact):
    n, w, d, radius, speed, limit = 700, 120.0, 5.0, 7.5, 0.6, 14
    a = random_agents(n, w, w, seed=3 + exact)
    world = rn.World(w, w, d, True, radius=radius, exact=bool(exact))
    world.preset(a["id"], a["x"], a["y"], a["ldx"], a["ldy"])
    f = kb.Field2D(w, w, d, True, capacity=n)
    f.set_order(True)
    f.set_object_locations(a["id"], a["x"], a["y"], a["ldx"], a["ldy"])
    f.lazy_update()
    for step in range(4):
        nx, ny, na, nb, keep = centroid_step(world, speed, limit)
        f.step_custom(CENTROID_PAIR, CENTROID_FINISH, [speed, limit], radius=radius, exact=exact, step=step, may_stop=True)
        f.lazy_update()
        got = f.download()
        ids = world.ids[keep]
        assert sorted(got["id"].tolist()) == sorted(ids.tolist()), step          # the crowded ones stopped
        o = np.argsort(got["id"])
        want = np.argsort(ids)
        for key, arr in (("x", nx), ("y", ny), ("ldx", na), ("ldy", nb)):
            assert (got[key][o].view(np.uint32) == arr[keep][want].view(np.uint32)).all(), (step, key)
        world.preset(ids, nx[keep], ny[keep], na[keep], nb[keep])
    assert f.num_objects() < n                                                   # somebody did stop
    f.close()


def test_a_snippet_that_does_not_compile():
    f = kb.Field2D(50.0, 50.0, 5.0, True, capacity=8)
    f.set_object_locations([1, 2], [1.0, 2.0], [1.0, 2.0])
    f.lazy_update()
    with pytest.raises(kb.KgError) as e:
        f.step_custom("acc[0] +* 1;", "nx = sx;", [])
    assert e.value.code == abi.KG_E_INVALID and "compile" in str(e.value)
    f.step_custom("", "nx = sx; ny = sy; na = 1.0f; nb = (float)nvec;", [], radius=100.0)   # an empty pair body is fine
    f.lazy_update()
    d = f.download()
    assert (d["ldx"] == 1.0).all() and (d["ldy"] == 2.0).all()
    f.close()


def test_field_model_runs_through_schedule_and_simulate():
    """the same custom model driven the reference's way: State + Schedule + simulate (lib.rs:1158-1175)"""
    n, w, d = 500, 100.0, 5.0
    a = random_agents(n, w, w, seed=8)
    make = lambda: kb.FieldModel((w, w), a, CENTROID_PAIR, CENTROID_FINISH, [0.5, 1e9], radius=7.5, discretization=d,
                                 canonical_order=True)
    st = make()
    sch = kb.Schedule()
    st.init(sch)
    for _ in range(6):
        sch.step_once(st)
    want = by_id(st.field1.download())
    st2 = make()
    kb.simulate(st2, 6, 1)
    got = by_id(st2.field1.download())
    for k in want:
        assert (got[k].view(np.uint32) == want[k].view(np.uint32)).all(), k
    sp = np.hypot(got["ldx"].astype(np.float64), got["ldy"].astype(np.float64))
    assert ((np.abs(sp - 0.5) < 1e-5) | (sp == 0)).all()
