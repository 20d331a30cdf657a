from . import wrappers  # noqa: F401
from .agent_selector import agent_selector  # noqa: F401
from .env import AECEnv  # noqa: F401
