"""Bi-Conjugate Gradient Stabilized Algorithm"""
from .bicgstab import BiCGSTAB     # noqa: F401
