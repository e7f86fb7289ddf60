"""Iterative Methods for Linear Least-Squares Problems"""
from .lsqr import LSQRFramework, LSQR      # noqa: F401
