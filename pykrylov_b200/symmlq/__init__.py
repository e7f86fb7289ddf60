"""Symmetric Indefinite Lanczos with Orthogonal Factorization"""
from .symmlq import Symmlq      # noqa: F401
