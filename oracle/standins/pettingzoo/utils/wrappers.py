"""The wrapper stack of gobblet.py:110-117 as ONE generic forwarding layer configured by a rule table.

PettingZoo 1.22.3 ships a class per wrapper; this stand-in (TEST INFRASTRUCTURE ONLY) models
them as data: a layer forwards every call to the env below it and mirrors the AEC bookkeeping back, and each
named wrapper is a row of RULES saying which guards run before `step` / `observe` / ... and which attributes are
hidden before `reset`.  Deliberately unlike gobblet_rl_b200/_aec.py in construction, so the two are independent
statements of the same published behaviour.
"""
from .env import AECEnv, TRANSITIONS, finished

MIRRORED = ("agent_selection", "rewards", "terminations", "truncations", "infos", "agents", "_cumulative_rewards")
STATIC = ("possible_agents", "metadata", "observation_spaces", "action_spaces")


# ---- guards: (layer, action) -> True when the call has been fully handled and must NOT be forwarded ----------
def guard_reset_first(layer, what):
    if not layer._state.get("has_reset"):
        raise AssertionError(f"reset() needs to be called before {what}")
    return False


def guard_cycle_over(layer, action):
    return not layer.agents                       # OrderEnforcingWrapper.step on an exhausted env is a no-op


def guard_bounds(layer, action):
    who = layer.agent_selection
    ok = (action is None and finished(layer, who)) or layer.action_space(who).contains(action)
    assert ok, "action is not in action space"
    return False


def guard_illegal(layer, action):
    """TerminateIllegalWrapper.step: the mask seen by the last observe(agent_selection) decides."""
    st, who = layer._state, layer.agent_selection
    if st.get("seen") is None:
        layer.observe(who)
    mask, st["seen"] = st["seen"]["action_mask"], None
    if st.get("ended"):
        TRANSITIONS["dead_step"](layer, action)
        return True
    if finished(layer, who) or mask[action]:
        return False                              # legal (or a regular dead step): forward
    layer._cumulative_rewards[who] = 0
    layer.terminations = dict.fromkeys(layer.agents, True)
    layer.truncations = dict.fromkeys(layer.agents, True)
    layer.rewards = dict.fromkeys(layer.truncations, 0)
    layer.rewards[who] = float(st["illegal_reward"])
    TRANSITIONS["accumulate"](layer)
    TRANSITIONS["deads_first"](layer)
    st["ended"] = True
    return True


RULES = {
    "BaseWrapper": {},
    "CaptureStdoutWrapper": {},
    "TerminateIllegalWrapper": {"step": [guard_illegal], "remember_observation": True,
                                "on_reset": {"ended": False, "seen": None}},
    "AssertOutOfBoundsWrapper": {"step": [guard_bounds]},
    "OrderEnforcingWrapper": {"step": [lambda l, a: guard_reset_first(l, "step"), guard_cycle_over],
                              "observe": [lambda l, a: guard_reset_first(l, "observe")],
                              "render": [lambda l, a: guard_reset_first(l, "render")],
                              "agent_iter": [lambda l, a: guard_reset_first(l, "agent_iter")],
                              "hidden_before_reset": ("rewards", "terminations", "truncations", "infos", "agent_selection",
                                                      "num_agents", "agents"),
                              "on_reset": {"has_reset": True}},
}


class _Layer(AECEnv):
    RULE = "BaseWrapper"

    def __init__(self, env, illegal_reward=None):
        self.__dict__["_state"] = {"illegal_reward": illegal_reward}
        self.__dict__["env"] = env
        for name in STATIC:
            if hasattr(env, name):
                setattr(self, name, getattr(env, name))

    def _rule(self, key, default=()):
        return RULES[self.RULE].get(key, default)

    def _guards(self, op, arg=None):
        return any(g(self, arg) for g in self._rule(op))

    def _mirror(self):
        for name in MIRRORED:
            setattr(self, name, getattr(self.env, name))

    def __getattr__(self, name):
        if name in self._rule("hidden_before_reset") and not self._state.get("has_reset"):
            raise AttributeError(f"{name} cannot be accessed before reset")
        if name.startswith("_"):
            raise AttributeError(f"accessing private attribute '{name}' is prohibited")
        return getattr(self.env, name)

    unwrapped = property(lambda self: self.env.unwrapped)

    def reset(self, seed=None, return_info=False, options=None):
        self._state.update(self._rule("on_reset", {}))
        self.env.reset(seed=seed, options=options)
        self._mirror()

    def step(self, action):
        if self._guards("step", action):
            return
        self.env.step(action)
        self._mirror()

    def observe(self, agent):
        self._guards("observe")
        obs = self.env.observe(agent)
        if self._rule("remember_observation", False) and agent == self.agent_selection:
            self._state["seen"] = obs
        return obs

    def render(self):
        self._guards("render")
        return self.env.render()

    def agent_iter(self, max_iter=2**63):
        self._guards("agent_iter")
        return super().agent_iter(max_iter)

    def close(self):
        self.env.close()

    def observation_space(self, agent):
        return self.env.observation_space(agent)

    def action_space(self, agent):
        return self.env.action_space(agent)


def _named(rule):
    return type(rule, (_Layer,), {"RULE": rule, "__doc__": f"pettingzoo.utils.wrappers.{rule} (stand-in, rule-table driven)"})


BaseWrapper = _named("BaseWrapper")
CaptureStdoutWrapper = _named("CaptureStdoutWrapper")
TerminateIllegalWrapper = _named("TerminateIllegalWrapper")
AssertOutOfBoundsWrapper = _named("AssertOutOfBoundsWrapper")
OrderEnforcingWrapper = _named("OrderEnforcingWrapper")
