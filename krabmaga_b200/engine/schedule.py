"""Sequential `Schedule` (src/engine/schedule.rs:227-413), kept on the host.

Row H of SURVEY §8 is *kept, not accelerated*: with a GPU field the queue holds one proxy agent
per device-resident population (so `step` never takes the empty-queue branch :357-365), and that
proxy's `step` launches the fused kernel for all real agents.
"""
import heapq
import itertools


class AgentImpl:
    """src/engine/agentimpl.rs:15-19"""
    __slots__ = ("id", "agent", "repeating")

    def __init__(self, agent, id):
        self.id, self.agent, self.repeating = id, agent, False


class Schedule:
    def __init__(self):
        """Schedule::new  schedule.rs:268-275"""
        self.step = 0
        self.time = 0.0
        self.events = []  # heap of (time, ordering, seq, AgentImpl); lower time/ordering first
        self.agent_ids_counting = 0
        self._seq = itertools.count()

    def schedule_once(self, agentimpl, the_time, the_ordering):
        """:284-286"""
        heapq.heappush(self.events, (float(the_time), int(the_ordering), next(self._seq), agentimpl))

    def schedule_repeating(self, agent, the_time, the_ordering):
        """:295-303"""
        a = AgentImpl(agent, self.agent_ids_counting)
        self.agent_ids_counting += 1
        a.repeating = True
        self.schedule_once(a, the_time, the_ordering)
        return True

    def distributed_schedule_repeating(self, agent, the_time, the_ordering):
        """:305-313"""
        ok = self.schedule_repeating(agent, the_time, the_ordering)
        return self.agent_ids_counting - 1, ok

    def get_all_events(self):
        """:316-322"""
        return [e[3].agent for e in sorted(self.events, key=lambda e: e[2])]

    def dequeue(self, agent, my_id):
        """:329-341"""
        for i, e in enumerate(self.events):
            if e[3].id == my_id:
                self.events.pop(i)
                heapq.heapify(self.events)
                return True
        return False

    def step_once(self, state):
        """Schedule::step  schedule.rs:347-413"""
        if self.step == 0:
            state.update(self.step)
        state.before_step(self)
        if not self.events:
            print("No agent in the queue to schedule. Terminating.")
            state.after_step(self)
            self.step += 1
            state.update(self.step)
            return
        self.time = self.events[0][0]
        cevents = []
        while self.events and not self.events[0][0] > self.time:
            cevents.append(heapq.heappop(self.events))
        for time, ordering, _, item in cevents:
            item.agent.before_step(state)
            item.agent.step(state)
            item.agent.after_step(state)
            if item.repeating and not item.agent.is_stopped(state):
                self.schedule_once(item, time + 1.0, ordering)
        state.after_step(self)
        self.step += 1
        state.update(self.step)
