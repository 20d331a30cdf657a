"""AECEnv stand-in, written as a MODEL of PettingZoo 1.22.3's agent-environment cycle, not as
a class hierarchy of behaviours: the bookkeeping of one env is five per-agent columns plus a cursor, and every
AEC operation is one row of the transition table below applied to that record.  TEST INFRASTRUCTURE ONLY --
deliberately a different construction from the product's gobblet_rl_b200/_aec.py so that the two can check
each other (tests/test_boundary_and_host_logic.py::test_aec_semantics_table).
"""

PER_AGENT = ("rewards", "_cumulative_rewards", "terminations", "truncations", "infos")


def cycle_is_over(agents):
    return len(agents) == 0


def finished(env, agent):
    return bool(env.terminations[agent] or env.truncations[agent])


def first_finished(env):
    """PettingZoo serves dead agents before live ones, in `agents` order."""
    for a in env.agents:
        if finished(env, a):
            return a
    return None


# ---- the transition table: operation -> function(env, *args) mutating the bookkeeping record -----------------
def op_accumulate(env):                       # AECEnv._accumulate_rewards
    for a in env.rewards:
        env._cumulative_rewards[a] = env._cumulative_rewards[a] + env.rewards[a]


def op_clear(env):                            # AECEnv._clear_rewards
    env.rewards = dict.fromkeys(env.rewards, 0)


def op_serve_dead_first(env):                 # AECEnv._deads_step_first
    d = first_finished(env)
    if d is not None:
        env._skip_agent_selection = env.agent_selection
        env.agent_selection = d
    return env.agent_selection


def op_bury(env, action):                     # AECEnv._was_dead_step
    if action is not None:
        raise ValueError("when an agent is dead, the only valid action is None")
    gone = env.agent_selection
    if not finished(env, gone):
        raise AssertionError("an agent that is not dead cannot be removed")
    for column in PER_AGENT:
        getattr(env, column).pop(gone)
    env.agents.remove(gone)                   # IN PLACE: wrappers mirror `agents` by reference, so the raw env sees the
                                              # removal too -- observable: raw_env.observe uses self.agents.index(agent)
                                              # (gobblet.py:182, :199), so after player_1 is buried player_2's last view
                                              # has index 0 = no sign flip, plane 12 zero (recorded in env_wrapped.npz)
    nxt = first_finished(env)
    parked = getattr(env, "_skip_agent_selection", None)
    if nxt is not None:                       # more dead agents queue up; remember who was really next
        env._skip_agent_selection = parked if parked is not None else env.agent_selection
        env.agent_selection = nxt
    else:                                     # queue empty: hand the turn back to the parked live agent, if any
        if parked is not None:
            env.agent_selection = parked
        env._skip_agent_selection = None
    op_clear(env)


TRANSITIONS = {"accumulate": op_accumulate, "clear": op_clear, "deads_first": op_serve_dead_first, "dead_step": op_bury}


class AECEnv:
    """The methods PettingZoo's AECEnv offers to an environment and its callers, each a one-line lookup into
    TRANSITIONS (reset / step / observe / render are the environment's own)."""

    def __init__(self):
        pass

    # -- what gobblet.py calls on itself (gobblet.py:236, :269) --------------------------------------------------
    def _accumulate_rewards(self):
        TRANSITIONS["accumulate"](self)

    def _clear_rewards(self):
        TRANSITIONS["clear"](self)

    def _deads_step_first(self):
        return TRANSITIONS["deads_first"](self)

    def _was_dead_step(self, action):
        TRANSITIONS["dead_step"](self, action)

    # -- what callers use (example_basic.py:50-67) ------------------------------------------------------------------
    def last(self, observe=True):
        who = self.agent_selection
        assert who
        view = self.observe(who) if observe else None
        return view, self._cumulative_rewards[who], self.terminations[who], self.truncations[who], self.infos[who]

    def agent_iter(self, max_iter=2**63):
        def gen(budget=max_iter):
            while budget > 0 and not cycle_is_over(self.agents):
                budget -= 1
                yield self.agent_selection
        return gen()

    def observation_space(self, agent):
        return self.observation_spaces[agent]

    def action_space(self, agent):
        return self.action_spaces[agent]

    num_agents = property(lambda self: len(self.agents))
    max_num_agents = property(lambda self: len(self.possible_agents))
    unwrapped = property(lambda self: self)

    def step(self, action):
        raise NotImplementedError

    def reset(self, seed=None, return_info=False, options=None):
        raise NotImplementedError

    def observe(self, agent):
        raise NotImplementedError

    def render(self):
        raise NotImplementedError

    def close(self):
        pass
