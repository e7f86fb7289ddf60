"""Problem gallery (matrix-free definitions + device-resident CSR builders)."""
from .gallery import *      # noqa: F401,F403
