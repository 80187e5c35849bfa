"""A Field2D model of the user's own on the GPU field: `State` + one proxy `Agent` whose `step` is every agent's
step, given as two CUDA C snippets (include/krabgpu.h KgCustomStep, csrc/jit_agent.cuh).  The shape is the
Flockers fixture's (flockers.py): `init` places the agents, the schedule holds the proxy, `update` is
`field1.lazy_update()`."""
from .engine.agent import Agent
from .engine.fields.field_2d import Field2D
from .engine.state import State


class FieldAgents(Agent):
    """Proxy for all agents of the model's field (Schedule::step calls it once per step)."""

    def step(self, state):
        state.field1.step_custom(state.pair, state.finish, state.consts, radius=state.radius, exact=state.exact,
                                 seed=state.seed, step=state.step, may_stop=state.may_stop)

    def is_stopped(self, state):
        """agent.rs:18 — the proxy stops once its population died out"""
        return state.may_stop and state.field1.num_objects(unbuffered=True) == 0


class FieldModel(State):
    def __init__(self, dim, agents, pair, finish, consts=(), radius=10.0, exact=False, discretization=1.0, toroidal=True,
                 seed=0, may_stop=False, device=0, canonical_order=False, capacity=None):
        """`agents` = dict(id, x, y, ldx, ldy): position and two floats of state per agent (a, b in the snippets)."""
        self.step = 0
        self.dim = (float(dim[0]), float(dim[1]))
        self.agents, self.pair, self.finish, self.consts = agents, pair, finish, list(consts)
        self.radius, self.exact, self.seed, self.may_stop = radius, exact, seed, may_stop
        self.discretization, self.toroidal, self.device = discretization, toroidal, device
        self.canonical_order, self.capacity = canonical_order, capacity
        self.field1 = None
        self._new_field()

    def _new_field(self):
        if self.field1 is not None:
            self.field1.close()
        self.field1 = Field2D(self.dim[0], self.dim[1], self.discretization, self.toroidal,
                              capacity=max(len(self.agents["id"]), self.capacity or 0, 1), device=self.device)
        self.field1.set_order(self.canonical_order)

    def reset(self):
        self.step = 0
        self._new_field()

    def init(self, schedule):
        a = self.agents
        self.field1.set_object_locations(a["id"], a["x"], a["y"], a["ldx"], a["ldy"])
        schedule.schedule_repeating(FieldAgents(), 0.0, 0)

    def update(self, step):
        self.step = step
        self.field1.lazy_update()
