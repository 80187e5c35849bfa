"""The oracle's dynamic-population model (LifeRule) checked against an independent restatement:
ref_numpy's step arithmetic plus the life rule written out in Python from the same definition
(include/krabgpu.h KgLifeRule).  Pins ids, survivors and children on the CPU side."""
import numpy as np

import oracle_binding as ob
import ref_numpy as rn

DISC = float(np.float32(10.0) / np.float32(1.5))


def numpy_life_run(w, n, seed, death, birth, crowd, nsteps):
    wd = rn.World(w, w, DISC, True, seed=seed)
    wd.init(n)
    next_id = n
    for _ in range(nsteps):
        k_n = len(wd.ids)
        keep, children = [], []
        new = []
        # parents in iter_objects order: bags by index, ascending id inside a bag
        order = [k for c in sorted(wd.bags) for k in wd.bags[c]]
        for k in range(k_n):
            me = int(wd.ids[k])
            others = sum(1 for e in wd.neighbors(wd.x[k], wd.y[k], wd.radius, False) if int(wd.ids[e]) != me)
            nx, ny, dx, dy = wd.step_agent(k)
            v = rn.philox4x32_10((me, wd.step_no & 0xFFFFFFFF, wd.step_no >> 32, 3), (seed & 0xFFFFFFFF, seed >> 32))
            dies = rn.uniform01(v[0]) < np.float32(death) or (crowd and others >= crowd)
            if not dies:
                new.append((me, nx, ny, dx, dy))
        for k in order:
            me = int(wd.ids[k])
            v = rn.philox4x32_10((me, wd.step_no & 0xFFFFFFFF, wd.step_no >> 32, 3), (seed & 0xFFFFFFFF, seed >> 32))
            if rn.uniform01(v[1]) < np.float32(birth):
                new.append((next_id, wd.x[k], wd.y[k], np.float32(0), np.float32(0)))
                next_id += 1
        step_no = wd.step_no + 1
        a = np.array(new, dtype=object).reshape(-1, 5)
        wd.preset(a[:, 0].astype(np.uint32), a[:, 1].astype(np.float32), a[:, 2].astype(np.float32),
                  a[:, 3].astype(np.float32), a[:, 4].astype(np.float32))
        wd.step_no = step_no
    return wd.by_id()


def test_oracle_life_model_equals_the_numpy_restatement():
    w, n, seed, death, birth, crowd, nsteps = 90.0, 400, 13, 0.04, 0.06, 30, 8
    m = ob.Flockers(w, w, n, DISC, True, ob.boids_params(radius=10.0, exact=0, seed=seed), canonical_order=True)
    m.set_life(death, birth, crowd, n)
    m.init()
    m.step(nsteps)
    want = m.population()
    got = numpy_life_run(w, n, seed, death, birth, crowd, nsteps)
    assert want["born"] > 20 and want["died"] > 20
    assert (got["id"] == want["id"]).all()
    for k in ("x", "y", "ldx", "ldy"):
        assert (got[k].view(np.uint32) == want[k].view(np.uint32)).all(), k
