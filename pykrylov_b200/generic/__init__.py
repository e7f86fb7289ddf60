"""Generic Krylov method template."""
from .generic import KrylovMethod, null_log     # noqa: F401
