class _AECIterator:
    def __init__(self, env, max_iter):
        self.env, self.left = env, max_iter

    def __iter__(self):
        return self

    def __next__(self):
        if not self.env.agents or self.left <= 0:
            raise StopIteration
        self.left -= 1
        return self.env.agent_selection


class AECEnv:
    """Agent-environment-cycle base: last(), dead steps, reward accumulation."""

    def __init__(self):
        pass

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

    @property
    def num_agents(self):
        return len(self.agents)

    @property
    def max_num_agents(self):
        return len(self.possible_agents)

    @property
    def unwrapped(self):
        return self

    def _dead_agents(self):
        return [a for a in self.agents if self.terminations[a] or self.truncations[a]]

    def _deads_step_first(self):
        dead = self._dead_agents()
        if dead:
            self._skip_agent_selection = self.agent_selection
            self.agent_selection = dead[0]
        return self.agent_selection

    def _clear_rewards(self):
        for a in self.rewards:
            self.rewards[a] = 0

    def _accumulate_rewards(self):
        for a, r in self.rewards.items():
            self._cumulative_rewards[a] += r

    def agent_iter(self, max_iter=2**63):
        return _AECIterator(self, max_iter)

    def last(self, observe=True):
        agent = self.agent_selection
        assert agent
        obs = self.observe(agent) if observe else None
        return (obs, self._cumulative_rewards[agent], self.terminations[agent],
                self.truncations[agent], self.infos[agent])

    def _was_dead_step(self, action):
        if action is not None:
            raise ValueError("when an agent is dead, the only valid action is None")
        agent = self.agent_selection
        assert self.terminations[agent] or self.truncations[agent]
        for d in (self.terminations, self.truncations, self.rewards, self._cumulative_rewards, self.infos):
            del d[agent]
        self.agents.remove(agent)
        dead = self._dead_agents()
        if dead:
            if getattr(self, "_skip_agent_selection", None) is None:
                self._skip_agent_selection = self.agent_selection
            self.agent_selection = dead[0]
        else:
            if getattr(self, "_skip_agent_selection", None) is not None:
                self.agent_selection = self._skip_agent_selection
            self._skip_agent_selection = None
        self._clear_rewards()
