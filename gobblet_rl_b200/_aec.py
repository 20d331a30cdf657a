"""Agent-Environment-Cycle plumbing for the `gobblet_v1.env()` facade.

When PettingZoo is installed its own AECEnv / agent_selector / wrappers are used, exactly as the
reference does (gobblet.py:102-104, :110-117).  The build image has no PettingZoo, so this module
carries the small subset of its 1.22.3 behaviour the facade relies on: `last()`,
reward accumulation, dead steps, and the illegal-move / bounds / ordering wrappers.
"""

try:  # pragma: no cover - pettingzoo is not in the build image
    from pettingzoo import AECEnv  # noqa: F401
    from pettingzoo.utils import agent_selector  # noqa: F401
    from pettingzoo.utils.wrappers import (AssertOutOfBoundsWrapper, BaseWrapper, OrderEnforcingWrapper,  # noqa: F401
                                           TerminateIllegalWrapper)

    HAVE_PETTINGZOO = True
except ImportError:
    HAVE_PETTINGZOO = False

    class agent_selector:
        """Round-robin over a fixed agent order; reset() returns the first agent."""

        def __init__(self, agent_order):
            self.reinit(agent_order)

        def reinit(self, agent_order):
            self.agent_order = agent_order
            self._idx = 0
            self.selected_agent = 0

        def reset(self):
            self.reinit(self.agent_order)
            return self.next()

        def next(self):
            self.selected_agent = self.agent_order[self._idx]
            self._idx = (self._idx + 1) % len(self.agent_order)
            return self.selected_agent

        def is_first(self):
            return self.selected_agent == self.agent_order[0]

        def is_last(self):
            return self.selected_agent == self.agent_order[-1]

    class _AgentIter:
        def __init__(self, env, max_iter):
            self._env, self._budget = env, max_iter

        def __iter__(self):
            return self

        def __next__(self):
            if self._budget <= 0 or not self._env.agents:
                raise StopIteration
            self._budget -= 1
            return self._env.agent_selection

    class AECEnv:
        def __init__(self):
            pass

        # -- interface a concrete env provides ----------------------------------------------------
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

        def observation_space(self, agent):
            return self.observation_spaces[agent]

        def action_space(self, agent):
            return self.action_spaces[agent]

        # -- derived behaviour ---------------------------------------------------------------------
        @property
        def num_agents(self):
            return len(self.agents)

        @property
        def max_num_agents(self):
            return len(self.possible_agents)

        @property
        def unwrapped(self):
            return self

        def agent_iter(self, max_iter=2**63):
            return _AgentIter(self, max_iter)

        def last(self, observe=True):
            a = self.agent_selection
            assert a
            return (self.observe(a) if observe else None, self._cumulative_rewards[a],
                    self.terminations[a], self.truncations[a], self.infos[a])

        def _clear_rewards(self):
            for a in self.rewards:
                self.rewards[a] = 0

        def _accumulate_rewards(self):
            for a, r in self.rewards.items():
                self._cumulative_rewards[a] += r

        def _first_dead(self):
            for a in self.agents:
                if self.terminations[a] or self.truncations[a]:
                    return a
            return None

        def _deads_step_first(self):
            dead = self._first_dead()
            if dead is not None:
                self._skip_agent_selection = self.agent_selection
                self.agent_selection = dead
            return self.agent_selection

        def _was_dead_step(self, action):
            if action is not None:
                raise ValueError("when an agent is dead, the only valid action is None")
            agent = self.agent_selection
            assert self.terminations[agent] or self.truncations[agent], \
                "an agent that was not dead was attempted to be removed"
            for table in (self.terminations, self.truncations, self.rewards, self._cumulative_rewards, self.infos):
                del table[agent]
            self.agents.remove(agent)
            dead = self._first_dead()
            if dead is not None:
                if getattr(self, "_skip_agent_selection", None) is None:
                    self._skip_agent_selection = self.agent_selection
                self.agent_selection = dead
            else:
                if getattr(self, "_skip_agent_selection", None) is not None:
                    self.agent_selection = self._skip_agent_selection
                self._skip_agent_selection = None
            self._clear_rewards()

    class BaseWrapper(AECEnv):
        _SHARED = ("agent_selection", "rewards", "terminations", "truncations", "infos", "agents",
                   "_cumulative_rewards")

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

        def _sync(self):
            for name in self._SHARED:
                setattr(self, name, getattr(self.env, name))

        def reset(self, seed=None, return_info=False, options=None):
            self.env.reset(seed=seed, options=options)
            self._sync()

        def step(self, action):
            self.env.step(action)
            self._sync()

        def observe(self, agent):
            return self.env.observe(agent)

        def render(self):
            return self.env.render()

        def close(self):
            self.env.close()

        def observation_space(self, agent):
            return self.env.observation_space(agent)

        def action_space(self, agent):
            return self.env.action_space(agent)

    class TerminateIllegalWrapper(BaseWrapper):
        """An action outside the last observed mask ends the game: mover gets `illegal_reward`,
        everyone is terminated and truncated, the board is untouched (gobblet.py:50-51, :114)."""

        def __init__(self, env, illegal_reward):
            super().__init__(env)
            self._illegal_value = illegal_reward
            self._prev_obs = None
            self._terminated = False

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
            mover = self.agent_selection
            if self._prev_obs is None:
                self.observe(mover)
            mask = self._prev_obs["action_mask"]
            self._prev_obs = None
            if self._terminated:
                self._was_dead_step(action)
            elif not self.terminations[mover] and not self.truncations[mover] and not mask[action]:
                self._cumulative_rewards[mover] = 0
                self.terminations = {a: True for a in self.agents}
                self.truncations = {a: True for a in self.agents}
                self.rewards = {a: 0 for a in self.agents}
                self.rewards[mover] = float(self._illegal_value)
                self._accumulate_rewards()
                self._deads_step_first()
                self._terminated = True
            else:
                super().step(action)

    class AssertOutOfBoundsWrapper(BaseWrapper):
        def step(self, action):
            a = self.agent_selection
            dead = self.terminations[a] or self.truncations[a]
            assert (action is None and dead) or self.action_space(a).contains(action), \
                "action is not in action space"
            super().step(action)

    class OrderEnforcingWrapper(BaseWrapper):
        _NEEDS_RESET = ("rewards", "terminations", "truncations", "infos", "agent_selection", "num_agents",
                        "agents")

        def __init__(self, env):
            self._has_reset = False
            super().__init__(env)

        def __getattr__(self, name):
            if name in self._NEEDS_RESET:
                raise AttributeError(f"{name} cannot be accessed before reset")
            return super().__getattr__(name)

        def _require_reset(self, what):
            if not self._has_reset:
                raise AssertionError(f"reset() needs to be called before {what}")

        def reset(self, seed=None, return_info=False, options=None):
            self._has_reset = True
            super().reset(seed=seed, options=options)

        def step(self, action):
            self._require_reset("step")
            if not self.agents:
                return None
            super().step(action)

        def observe(self, agent):
            self._require_reset("observe")
            return super().observe(agent)

        def render(self):
            self._require_reset("render")
            return super().render()

        def agent_iter(self, max_iter=2**63):
            self._require_reset("agent_iter")
            return super().agent_iter(max_iter)
