"""A SECOND, independent restatement of the Flockers step in numpy.float32 scalar operations.

Test infrastructure only.  It shares no code with oracle/ (the C++ restatement): it was written
from the reference sources alone, so that the oracle's numeric output for SURVEY §8 row G is checked
by something other than itself and its own golden files (tests/test_ref_numpy.py compares the two
bit for bit).  It is still not the reference binary — the crate cannot be built in this image.

What it follows (paths relative to the krABMaga crate root):
  tests/model/flockers/bird.rs:39-155        Bird::step
  tests/model/flockers/state.rs:41-56        Flocker::init
  src/engine/fields/field_2d.rs:328-339      discretize
  src/engine/fields/field_2d.rs:386-440      get_neighbors_within_distance
  src/engine/fields/field_2d.rs:472-516      get_neighbors_within_relax_distance
  src/engine/fields/field_2d.rs:838-846      set_object_location
  src/engine/fields/field_2d.rs:926-1014     t_transform, check_circle, distance,
                                             toroidal_distance, toroidal_transform
Randomness: the shared Philox4x32-10 stream of DESIGN.md §4 (key = seed, counter = (agent id, step,
domain)), restated here from the published algorithm (Salmon et al., SC'11), u32 -> f32 as rand 0.9's
StandardUniform (24 high bits * 2^-24).

Bag order: the reference's bags hold agents in push order, which depends on the scheduler's pop
order.  Like the oracle's and the device's KG_ORDER_CANONICAL mode, a bag is walked in ascending id.
"""
import numpy as np

F = np.float32
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def uniform01(u):
    return F(u >> 8) * F(2.0 ** -24)


def draws(seed, agent_id, step, domain):
    v = philox4x32_10((agent_id, step & _MASK, step >> 32, domain), (seed & _MASK, seed >> 32))
    return uniform01(v[0]), uniform01(v[1])


# ------------------------------------------------------------------ field_2d.rs free functions
def t_transform(n, size):
    # Rust's % truncates toward zero
    r = int(np.fmod(n, size))
    return r if n >= 0 else r + size


def toroidal_transform(val, dim):
    if val >= F(0) and val < dim:
        return val
    val = np.fmod(val, dim)           # f32 `%` in Rust is C fmodf
    if val < F(0):
        val = val + dim
    return F(val)


def toroidal_distance(a, b, dim):
    if abs(a - b) <= dim / F(2):
        return a - b
    d = toroidal_transform(a, dim) - toroidal_transform(b, dim)
    if d * F(2) > dim:
        return d - dim
    if d * F(2) < -dim:
        return d + dim
    return d


def distance(ax, ay, bx, by, w, h, tor):
    if tor:
        dx, dy = toroidal_distance(ax, bx, w), toroidal_distance(ay, by, h)
    else:
        dx, dy = ax - bx, ay - by
    return np.sqrt(dx * dx + dy * dy)


def check_circle(bx, by, disc, w, h, lx, ly, dis, tor):
    nwx, nwy = F(bx) * disc, F(by) * disc
    ney = min(nwy + disc, h)
    swx = min(nwx + disc, w)
    d = [distance(nwx, nwy, lx, ly, w, h, tor), distance(nwx, ney, lx, ly, w, h, tor),
         distance(swx, nwy, lx, ly, w, h, tor), distance(swx, ney, lx, ly, w, h, tor)]
    if all(v <= dis for v in d):
        return 1
    if all(v > dis for v in d):
        return -1
    return 0


class World:
    """Field2D (default variant) + the Flockers population, double buffered."""

    def __init__(self, w, h, disc, toroidal, seed=42, radius=10.0, exact=False, jump=0.7, cohesion=1.0,
                 avoidance=1.0, randomness=1.0, consistency=1.0, momentum=1.0):
        self.w, self.h, self.disc, self.tor = F(w), F(h), F(disc), bool(toroidal)
        self.seed, self.radius, self.exact = int(seed), F(radius), bool(exact)
        self.jump, self.k_coh, self.k_avo = F(jump), F(cohesion), F(avoidance)
        self.k_rnd, self.k_con, self.k_mom = F(randomness), F(consistency), F(momentum)
        self.max_x = int(np.ceil(self.w / self.disc))
        self.max_y = int(np.ceil(self.h / self.disc))
        self.dw, self.dh = self.max_x + 1, self.max_y + 1
        self.step_no = 0
        self.ids = np.zeros(0, np.uint32)
        self.x = self.y = self.ldx = self.ldy = np.zeros(0, F)
        self.bags = {}

    # state.rs:41-56 with Philox draws (domain 0)
    def init(self, n):
        ids = np.arange(n, dtype=np.uint32)
        x, y = np.zeros(n, F), np.zeros(n, F)
        for i in range(n):
            r1, r2 = draws(self.seed, i, 0, 0)
            x[i], y[i] = self.w * r1, self.h * r2
        self.preset(ids, x, y, np.zeros(n, F), np.zeros(n, F))

    def preset(self, ids, x, y, ldx, ldy):
        self.ids = np.asarray(ids, np.uint32).copy()
        self.x, self.y = np.asarray(x, F).copy(), np.asarray(y, F).copy()
        self.ldx, self.ldy = np.asarray(ldx, F).copy(), np.asarray(ldy, F).copy()
        self._rebuild()

    def discretize(self, x, y):
        return int(np.floor(x / self.disc)), int(np.floor(y / self.disc))

    def _rebuild(self):
        # set_object_location for everyone, then lazy_update; bags in ascending id (canonical order)
        bags = {}
        for k in np.argsort(self.ids, kind="stable"):
            cx, cy = self.discretize(self.x[k], self.y[k])
            index = cx * self.dh + cy
            assert 0 <= index < self.dw * self.dh, "the reference would panic: bag index out of bounds"
            bags.setdefault(index, []).append(int(k))
        self.bags = bags

    def neighbors(self, lx, ly, dist, exact):
        """indices (into the state arrays) in the reference's visiting order"""
        out = []
        if dist <= F(0):
            return out
        dd = int(np.floor(dist / self.disc))
        cx, cy = self.discretize(lx, ly)
        min_i, max_i, min_j, max_j = cx - dd, cx + dd, cy - dd, cy + dd
        if self.tor:
            min_i, max_i = max(0, min_i), min(max_i, self.max_x - 1)
            min_j, max_j = max(0, min_j), min(max_j, self.max_y - 1)
        for i in range(min_i, max_i + 1):
            for j in range(min_j, max_j + 1):
                bx, by = t_transform(i, self.max_x), t_transform(j, self.max_y)
                check = check_circle(bx, by, self.disc, self.w, self.h, lx, ly, dist, self.tor) if exact else 1
                for k in self.bags.get(bx * self.dh + by, ()):
                    if check == 1 or (check == 0 and distance(lx, ly, self.x[k], self.y[k], self.w, self.h,
                                                              self.tor) <= dist):
                        out.append(k)
        return out

    # bird.rs:39-155 for the agent stored at index k
    def step_agent(self, k):
        me, px, py = int(self.ids[k]), self.x[k], self.y[k]
        vec = self.neighbors(px, py, self.radius, self.exact)
        zero = F(0)
        avo_x = avo_y = coh_x = coh_y = rnd_x = rnd_y = con_x = con_y = zero
        if vec:
            xa = ya = xc = yc = xs = ys = zero
            count = 0
            for e in vec:
                if me != int(self.ids[e]):
                    dx = toroidal_distance(px, self.x[e], self.w)
                    dy = toroidal_distance(py, self.y[e], self.h)
                    count += 1
                    square = dx * dx + dy * dy
                    xa = xa + dx / (square * square + F(1))
                    ya = ya + dy / (square * square + F(1))
                    xc, yc = xc + dx, yc + dy
                    xs, ys = xs + self.ldx[e], ys + self.ldy[e]
            if count > 0:
                c = F(count)
                xa, ya, xc, yc, xs, ys = xa / c, ya / c, xc / c, yc / c, xs / c, ys / c
                con_x, con_y = xs / c, ys / c          # divided twice, bird.rs:85-91
            else:
                con_x, con_y = xs, ys
            avo_x, avo_y = F(400) * xa, F(400) * ya
            coh_x, coh_y = -xc / F(10), -yc / F(10)
            r1, r2 = draws(self.seed, me, self.step_no, 1)
            xr, yr = r1 * F(2) - F(1), r2 * F(2) - F(1)
            length = np.sqrt(xr * xr + yr * yr)
            rnd_x, rnd_y = F(0.05) * xr / length, F(0.05) * yr / length
        dx = self.k_coh * coh_x + self.k_avo * avo_x + self.k_con * con_x + self.k_rnd * rnd_x \
            + self.k_mom * self.ldx[k]
        dy = self.k_coh * coh_y + self.k_avo * avo_y + self.k_con * con_y + self.k_rnd * rnd_y \
            + self.k_mom * self.ldy[k]
        dis = np.sqrt(dx * dx + dy * dy)
        if dis > zero:
            dx, dy = dx / dis * self.jump, dy / dis * self.jump
        nx = toroidal_transform(px + dx, self.w)
        ny = toroidal_transform(py + dy, self.w)          # `width` for both, bird.rs:146-147
        return F(nx), F(ny), F(dx), F(dy)

    def step(self, nsteps=1):
        with np.errstate(all="ignore"):
            for _ in range(nsteps):
                n = len(self.ids)
                nx, ny, ndx, ndy = np.zeros(n, F), np.zeros(n, F), np.zeros(n, F), np.zeros(n, F)
                for k in range(n):
                    nx[k], ny[k], ndx[k], ndy[k] = self.step_agent(k)
                self.x, self.y, self.ldx, self.ldy = nx, ny, ndx, ndy
                self._rebuild()
                self.step_no += 1

    def by_id(self):
        o = np.argsort(self.ids, kind="stable")
        return {"id": self.ids[o], "x": self.x[o], "y": self.y[o], "ldx": self.ldx[o], "ldy": self.ldy[o]}
