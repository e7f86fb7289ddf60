"""Conjugate Gradient Squared Algorithm"""
from .cgs import CGS     # noqa: F401
