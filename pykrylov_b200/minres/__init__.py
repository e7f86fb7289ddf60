"""Minimum Residual Algorithm"""
from .minres import Minres     # noqa: F401
