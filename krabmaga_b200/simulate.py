"""Entry points mirroring the reference's macros for the hot path's callers.

simulate      <- simulate!(state, steps, reps, false)   src/lib.rs:1158-1175 (plain branch)
simulate_old  <- simulate_old!(state, steps, reps)      src/lib.rs:1473-1582
simulate_explore <- simulate_explore!(step, state)      src/explore/model_exploration.rs:160-190
"""
import time

from .engine.schedule import Schedule


def simulate(state, n_step, reps):
    s = state.as_state_mut()
    for _ in range(reps):
        schedule = Schedule()
        s.init(schedule)
        for _ in range(n_step):
            schedule.step_once(s)
            if s.end_condition(schedule):
                break
    return s


def _sync(state):
    for v in vars(state).values():
        if hasattr(v, "sync"):
            v.sync()


def simulate_old(state, n_step, reps):
    """Returns [(duration_seconds, steps_per_second)] per repetition, like the reference macro."""
    s = state.as_state_mut()
    results = []
    for _ in range(reps):
        schedule = Schedule()
        s.init(schedule)
        _sync(s)
        t0 = time.perf_counter()
        for _ in range(n_step):
            schedule.step_once(s)
            if s.end_condition(schedule):
                break
        _sync(s)  # device work is asynchronous: the run ends when the stream has drained
        dt = time.perf_counter() - t0
        results.append((dt, schedule.step / dt))
    return results


def simulate_explore(n_step, state):
    return simulate_old(state, n_step, 1)
