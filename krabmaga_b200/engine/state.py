"""`State` trait (src/engine/state.rs:45-63)."""


class State:
    def init(self, schedule):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def update(self, step):
        raise NotImplementedError

    def before_step(self, schedule):
        pass

    def after_step(self, schedule):
        pass

    def end_condition(self, schedule):
        return False

    def as_state_mut(self):
        return self

    def as_state(self):
        return self
