"""Conjugate-Gradient Algorithm"""
from .cg import CG     # noqa: F401
