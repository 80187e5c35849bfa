"""`Agent` trait (src/engine/agent.rs:7-40)."""


class Agent:
    def step(self, state):
        raise NotImplementedError

    def is_stopped(self, state):
        return False

    def before_step(self, state):
        return None

    def after_step(self, state):
        return None
