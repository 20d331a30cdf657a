"""gymnasium stand-in (spaces + logger only). TEST INFRASTRUCTURE."""
from . import logger, spaces  # noqa: F401
