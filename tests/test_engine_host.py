"""The host-side mirror of the reference's engine contracts (krabmaga_b200/engine/{schedule,state,
agent}.py, simulate.py) on plain CPU agents: the reference's own schedule tests
(tests/engine/schedule.rs:8-85) and the ordering rules of Schedule::step (schedule.rs:347-413).
No device is involved — with a GPU field the same Schedule carries one proxy agent."""
from krabmaga_b200.engine.agent import Agent
from krabmaga_b200.engine.schedule import Schedule
from krabmaga_b200.engine.state import State
from krabmaga_b200.simulate import simulate, simulate_explore, simulate_old


class MyNode(Agent):
    """tests/utils/mynode.rs"""

    def __init__(self, id, flag=False):
        self.id, self.flag = id, flag

    def __eq__(self, other):
        return self.id == other.id

    def step(self, state):
        state.trace.append(("step", self.id, state.schedule_step))


class Log(State):
    def __init__(self, agents=(), stop_at=None):
        self.trace, self.agents, self.stop_at, self.schedule_step = [], list(agents), stop_at, 0
        self.inits = 0

    def init(self, schedule):
        self.inits += 1
        for a, (t, o) in self.agents:
            schedule.schedule_repeating(a, t, o)

    def update(self, step):
        self.schedule_step = step
        self.trace.append(("update", step))

    def before_step(self, schedule):
        self.trace.append(("before", schedule.step))

    def after_step(self, schedule):
        self.trace.append(("after", schedule.step))

    def end_condition(self, schedule):
        return self.stop_at is not None and schedule.step >= self.stop_at


def test_schedule_operations():
    """tests/engine/schedule.rs:8-39"""
    schedule = Schedule()
    node1, node2 = MyNode(0), MyNode(1)
    schedule.schedule_repeating(node1, 0.0, 0)
    schedule.schedule_repeating(node2, 0.0, 0)
    agents = schedule.get_all_events()
    assert len(agents) == 2 and agents[0] == node1 and agents[1] == node2
    assert schedule.dequeue(node1, node1.id)
    assert schedule.get_all_events() == [node2]
    assert schedule.dequeue(node2, node2.id)
    assert schedule.get_all_events() == []
    assert not schedule.dequeue(node2, node2.id)


def test_distributed_schedule_operations():
    """tests/engine/schedule.rs:48-85"""
    schedule = Schedule()
    node1, node2 = MyNode(0), MyNode(1)
    assert schedule.distributed_schedule_repeating(node1, 0.0, 0) == (0, True)
    assert schedule.distributed_schedule_repeating(node2, 0.0, 0) == (1, True)
    assert schedule.get_all_events() == [node1, node2]
    assert schedule.dequeue(node1, node1.id)
    assert schedule.get_all_events() == [node2]
    assert schedule.dequeue(node2, node2.id)
    assert schedule.get_all_events() == []


def test_step_order_update_before_agents_after_update():
    """schedule.rs:347-413: the very first step starts with state.update(0); then before_step, every
    due agent in (time, ordering) order, after_step, step += 1, state.update(step)"""
    st = Log([(MyNode(0), (0.0, 1)), (MyNode(1), (0.0, 0)), (MyNode(2), (1.0, 0))])
    sch = Schedule()
    st.init(sch)
    sch.step_once(st)
    assert st.trace == [("update", 0), ("before", 0), ("step", 1, 0), ("step", 0, 0), ("after", 0), ("update", 1)]
    st.trace.clear()
    sch.step_once(st)        # time 1.0: the two repeating agents rescheduled at +1.0 and the late one
    assert st.trace[0] == ("before", 1) and st.trace[-2:] == [("after", 1), ("update", 2)]
    assert sorted(t[1] for t in st.trace if t[0] == "step") == [0, 1, 2]
    assert [t[1] for t in st.trace if t[0] == "step"][0] in (1, 2)    # ordering 0 before ordering 1
    assert sch.step == 2 and sch.time == 1.0


def test_empty_queue_branch_still_advances():
    """schedule.rs:357-365"""
    st, sch = Log(), Schedule()
    sch.step_once(st)
    assert st.trace == [("update", 0), ("before", 0), ("after", 0), ("update", 1)] and sch.step == 1


def test_stopped_agents_are_not_rescheduled():
    class Once(MyNode):
        def is_stopped(self, state):
            return True
    st = Log([(Once(7), (0.0, 0))])
    sch = Schedule()
    st.init(sch)
    sch.step_once(st)
    assert sch.get_all_events() == []


def test_simulate_entry_points():
    """tests/explore/simulate.rs:18-44 on a CPU state: one (duration, steps/s) pair per repetition;
    simulate! (lib.rs:1158-1175) re-inits per repetition and honours end_condition"""
    mk = lambda: Log([(MyNode(0), (0.0, 0))])
    assert simulate_old(mk(), 10, 0) == []
    for reps in (1, 2):
        res = simulate_old(mk(), 10, reps)
        assert len(res) == reps and all(d > 0 and sps > 0 for d, sps in res)
    assert len(simulate_explore(10, mk())) == 1
    st = simulate(Log([(MyNode(0), (0.0, 0))], stop_at=4), 10, 3)
    assert st.inits == 3
    assert [t for t in st.trace if t[0] == "step"] == [("step", 0, s) for _ in range(3) for s in range(4)]


def test_addplot_and_plot_mirrors(tmp_path):
    import pytest
    """addplot! / plot! (lib.rs:1163-1240): series are created at the first point, a missing plot is an error"""
    import krabmaga_b200 as kb
    from krabmaga_b200 import plots
    plots.DATA.clear()
    kb.addplot("Flock", "step", "polarisation")
    kb.plot("Flock", "s1", 1, 0.25)
    kb.plot("Flock", "s1", 2, 0.5)
    kb.plot("Flock", "s2", 1, -1.0)
    d = plots.DATA["Flock"]
    assert list(d.series) == ["s1", "s2"] and d.series["s1"] == [(1.0, 0.25), (2.0, 0.5)]
    assert (d.min_x, d.max_x, d.min_y, d.max_y) == (1, 2, -1.0, 0.5)
    with pytest.raises(KeyError):
        kb.plot("nope", "s", 0, 0)
    d.to_csv(tmp_path / "p.csv")
    assert (tmp_path / "p.csv").read_text().splitlines()[1] == "s1,1.0,0.25"
