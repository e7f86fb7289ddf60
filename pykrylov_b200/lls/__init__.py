"""Iterative Methods for Linear Least-Squares Problems"""
from .lsqr import LSQRFramework, LSQR      # noqa: F401
from .lsmr import LSMRFramework            # noqa: F401
from .craig import CRAIGFramework          # noqa: F401
from .craigmr import CRAIGMRFramework      # noqa: F401
