"""BaseWrapper + the three wrappers of gobblet.py:110-117 (+ CaptureStdoutWrapper name)."""
from .env import AECEnv

_MIRRORED = ("agent_selection", "rewards", "terminations", "truncations", "infos", "agents",
             "_cumulative_rewards")


class BaseWrapper(AECEnv):
    def __init__(self, env):
        super().__init__()
        self.env = env
        for name in ("possible_agents", "metadata", "observation_spaces", "action_spaces"):
            if hasattr(env, name):
                setattr(self, name, getattr(env, name))

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(f"accessing private attribute '{name}' is prohibited")
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def _mirror(self):
        for name in _MIRRORED:
            setattr(self, name, getattr(self.env, name))

    def close(self):
        self.env.close()

    def render(self):
        return self.env.render()

    def reset(self, seed=None, return_info=False, options=None):
        self.env.reset(seed=seed, options=options)
        self._mirror()

    def observe(self, agent):
        return self.env.observe(agent)

    def step(self, action):
        self.env.step(action)
        self._mirror()

    def observation_space(self, agent):
        return self.env.observation_space(agent)

    def action_space(self, agent):
        return self.env.action_space(agent)


class CaptureStdoutWrapper(BaseWrapper):
    pass


class TerminateIllegalWrapper(BaseWrapper):
    def __init__(self, env, illegal_reward):
        super().__init__(env)
        self._illegal_value = illegal_reward
        self._prev_obs = None

    def reset(self, seed=None, return_info=False, options=None):
        self._terminated = False
        self._prev_obs = None
        super().reset(seed=seed, options=options)

    def observe(self, agent):
        obs = super().observe(agent)
        if agent == self.agent_selection:
            self._prev_obs = obs
        return obs

    def step(self, action):
        current = self.agent_selection
        if self._prev_obs is None:
            self.observe(current)
        mask = self._prev_obs["action_mask"]
        self._prev_obs = None
        if self._terminated:
            self._was_dead_step(action)
        elif (not self.terminations[current] and not self.truncations[current] and not mask[action]):
            self._cumulative_rewards[current] = 0
            self.terminations = {a: True for a in self.agents}
            self.truncations = {a: True for a in self.agents}
            self.rewards = {a: 0 for a in self.truncations}
            self.rewards[current] = float(self._illegal_value)
            self._accumulate_rewards()
            self._deads_step_first()
            self._terminated = True
        else:
            super().step(action)


class AssertOutOfBoundsWrapper(BaseWrapper):
    def step(self, action):
        sel = self.agent_selection
        dead = self.terminations[sel] or self.truncations[sel]
        assert (action is None and dead) or self.action_space(sel).contains(action), \
            "action is not in action space"
        super().step(action)


class OrderEnforcingWrapper(BaseWrapper):
    def __init__(self, env):
        self._has_reset = False
        super().__init__(env)

    def __getattr__(self, name):
        if name in ("rewards", "terminations", "truncations", "infos", "agent_selection",
                    "num_agents", "agents"):
            raise AttributeError(f"{name} cannot be accessed before reset")
        return super().__getattr__(name)

    def _need_reset(self, what):
        if not self._has_reset:
            raise AssertionError(f"reset() needs to be called before {what}")

    def render(self):
        self._need_reset("render")
        return super().render()

    def step(self, action):
        self._need_reset("step")
        if not self.agents:
            return None
        super().step(action)

    def observe(self, agent):
        self._need_reset("observe")
        return super().observe(agent)

    def agent_iter(self, max_iter=2**63):
        self._need_reset("agent_iter")
        return super().agent_iter(max_iter)

    def reset(self, seed=None, return_info=False, options=None):
        self._has_reset = True
        super().reset(seed=seed, options=options)
