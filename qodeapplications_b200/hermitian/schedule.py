"""The dependency analysis / stream placement of recorded launch sequences lives one level up (it serves both paths)."""
from ..schedule import *            # noqa: F401,F403
from ..schedule import dependencies, assign_streams, footprint_of     # noqa: F401
