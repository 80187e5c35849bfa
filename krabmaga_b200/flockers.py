"""The Flockers fixture (tests/model/flockers/{bird,state}.rs) opted into the GPU field.

`Flocker` keeps the reference's State shape (step, field1, initial_flockers, dim); the only change a
model author makes is the field type.  The population lives in `field1` on the device; the
schedule holds one `Flock` proxy agent whose `step` is every Bird's `step` (bird.rs:39-155) in a
single fused launch, and `State::update` is `field1.lazy_update()` exactly as in state.rs:58-60.
"""
from . import _abi as abi
from .engine.agent import Agent
from .engine.fields.field_2d import Field2D
from .engine.state import State

# tests/model/flockers/state.rs:11-14
WIDTH, HEIGHT, DISCRETIZATION, TOROIDAL = 10.0, 10.0, 0.5, True


class Flock(Agent):
    """Proxy for all Birds of one field."""

    def step(self, state):
        p = state.params
        p.step = state.step
        if state.life is None:
            state.field1.step_boids(p)
        else:
            # dynamic population: is_stopped for every bird + the births of State::after_step
            stopped, born = state.field1.step_boids_life(p, state.life)
            state.stopped += stopped
            state.born += born

    def is_stopped(self, state):
        """the proxy stops when its population died out (agent.rs:18)"""
        return state.life is not None and state.field1.num_objects(unbuffered=True) == 0


class Flocker(State):
    def __init__(self, dim, initial_flockers, discretization=DISCRETIZATION, toroidal=TOROIDAL,
                 params=None, device=0, canonical_order=False, preset=None, life=None, capacity=None):
        """Flocker::new(dim, initial_flockers)  state.rs:24-32.  `params` defaults to the fixture's
        constants with the exact query (bird.rs:12-17, :41)."""
        self.step = 0
        self.dim = (float(dim[0]), float(dim[1]))
        self.initial_flockers = int(initial_flockers)
        self.discretization, self.toroidal, self.device = discretization, toroidal, device
        self.params = params or abi.boids_params(radius=10.0, exact=1)
        self.canonical_order = canonical_order
        self.preset = preset  # optional dict(id,x,y,ldx,ldy) instead of the Philox init
        self.life = life      # abi.life_rule(...): births and deaths on the device
        self.capacity = capacity
        self.stopped = self.born = 0
        self.field1 = None
        self._new_field()

    def _new_field(self):
        if self.field1 is not None:
            self.field1.close()
        cap = max(self.initial_flockers, len(self.preset["id"]) if self.preset else 0, 1, self.capacity or 0)
        self.field1 = Field2D(self.dim[0], self.dim[1], self.discretization, self.toroidal,
                              capacity=cap, device=self.device)
        self.field1.set_order(self.canonical_order)

    def reset(self):
        """state.rs:35-38"""
        self.step = 0
        self._new_field()

    def init(self, schedule):
        """state.rs:41-56: place the birds, schedule them (here: one proxy)."""
        if self.preset is not None:
            p = self.preset
            self.field1.set_object_locations(p["id"], p["x"], p["y"], p["ldx"], p["ldy"])
        else:
            self.field1.init_flockers(self.initial_flockers, self.params.seed)
        schedule.schedule_repeating(Flock(), 0.0, 0)

    def update(self, step):
        """state.rs:58-60"""
        self.step = step
        self.field1.lazy_update()
