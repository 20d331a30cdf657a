"""pettingzoo stand-in restating 1.22.3 AEC semantics. TEST INFRASTRUCTURE."""
from .utils.env import AECEnv  # noqa: F401

__version__ = "1.22.3-standin"
