"""Dynamic population on the device: Agent::is_stopped (agent.rs:18, schedule.rs:401-407)
and the births of State::after_step, against the oracle's LifeRule model — same ids, same agents,
every f32 bit, step after step."""
import numpy as np
import pytest

import krabmaga_b200 as kb
import oracle_binding as ob
from krabmaga_b200 import _abi as abi
from parity_util import NORTH_STAR_DISC, both_params, random_agents

pytestmark = pytest.mark.gpu


def gpu_population(f):
    d = f.download(with_cells=False)
    o = np.argsort(d["id"], kind="stable")
    return {k: v[o] for k, v in d.items()}


def assert_same_population(got, want, what):
    assert len(got["id"]) == len(want["id"]), (what, len(got["id"]), len(want["id"]))
    assert (got["id"] == want["id"]).all(), what
    for k in ("x", "y", "ldx", "ldy"):
        bad = np.flatnonzero(got[k].view(np.uint32) != want[k].view(np.uint32))
        assert len(bad) == 0, (what, k, got["id"][bad[:5]])


@pytest.mark.parametrize("exact,variant", [(0, abi.KG_K4_AUTO), (1, abi.KG_K4_AUTO), (0, abi.KG_K4_GENERIC)])
def test_births_and_deaths_match_the_oracle_step_by_step(exact, variant):
    n, w, cap = 3000, 240.0, 12000
    op, gp = both_params(exact=exact, seed=11, cohesion=1.1, avoidance=0.9, consistency=0.8, randomness=1.3,
                         momentum=0.97)
    death, birth, crowd = 0.03, 0.05, 45
    m = ob.Flockers(w, w, n, NORTH_STAR_DISC, True, op, canonical_order=True)
    m.set_life(death, birth, crowd, n)
    m.init()
    st = kb.Flocker((w, w), n, discretization=NORTH_STAR_DISC, params=gp, canonical_order=True,
                    life=kb.life_rule(death, birth, crowd), capacity=cap)
    st.field1.set_kernel_variant(variant)
    sch = kb.Schedule()
    st.init(sch)
    sizes = []
    for step in range(1, 41):
        m.step(1)
        sch.step_once(st)
        if step in (1, 2, 5, 10, 20, 40):
            want = m.population()
            assert_same_population(gpu_population(st.field1), want, f"step {step}")
            assert (st.born, st.stopped) == (want["born"], want["died"])
            sizes.append(len(want["id"]))
    assert st.born > 100 and st.stopped > 100 and len(set(sizes)) > 3     # the population really moved


def test_everybody_dies_and_the_field_empties():
    n, w = 500, 120.0
    _, gp = both_params(exact=0, seed=3)
    f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
    f.init_flockers(n, 3)
    f.lazy_update()
    gp.step = 0
    stopped, born = f.step_boids_life(gp, kb.life_rule(death_prob=1.0))
    assert (stopped, born) == (n, 0)
    f.lazy_update()
    assert f.num_objects() == 0
    assert f.step_boids_life(gp, kb.life_rule(death_prob=1.0)) == (0, 0)     # an empty field steps fine
    f.close()


def test_crowding_rule_uses_the_querys_neighbour_count():
    """crowd_limit stops exactly the agents whose query returned >= limit others (bird.rs:80 `count`)"""
    n, w = 4000, 200.0
    agents = random_agents(n, w, w, seed=21)
    op, gp = both_params(exact=0, seed=5)
    f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n)
    f.set_order(True)
    f.set_object_locations(agents["id"], agents["x"], agents["y"], agents["ldx"], agents["ldy"])
    f.lazy_update()
    offs, ids = f.neighbors_batch(np.stack([agents["x"], agents["y"]], axis=1), 10.0, exact=False)
    others = np.diff(offs) - 1
    limit = int(np.median(others))
    gp.step = 0
    stopped, born = f.step_boids_life(gp, kb.life_rule(crowd_limit=limit))
    assert born == 0 and stopped == int((others >= limit).sum())
    f.lazy_update()
    left = f.download(with_cells=False)["id"]
    assert sorted(left) == sorted(agents["id"][others < limit])
    f.close()


def test_births_that_do_not_fit_and_the_reserved_id_are_reported():
    n, w = 1000, 160.0
    _, gp = both_params(exact=0, seed=8)
    f = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=n + 5)
    f.init_flockers(n, 8)
    f.lazy_update()
    gp.step = 0
    with pytest.raises(kb.KgError) as e:
        f.step_boids_life(gp, kb.life_rule(birth_prob=0.5))
    assert e.value.code == abi.KG_E_CAPACITY
    f.close()
    g = kb.Field2D(w, w, NORTH_STAR_DISC, True, capacity=8)
    with pytest.raises(kb.KgError) as e:
        g.set_object_locations([1, 0xFFFFFFFF], [1.0, 2.0], [1.0, 2.0])
    assert e.value.code == abi.KG_E_INVALID
    g.close()
