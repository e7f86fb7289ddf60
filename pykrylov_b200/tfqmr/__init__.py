"""Transpose-Free Quasi-Minimum Residual Algorithm"""
from .tfqmr import TFQMR     # noqa: F401
