"""Helper tools (type families, operator checks)."""
from .types import *      # noqa: F401,F403
from .utils import *      # noqa: F401,F403
