"""NumPy scalar-type families accepted as operator dtypes.

Mirrors the role of the reference's pykrylov/tools/types.py:3-16 (the removed
NumPy aliases np.int / np.float / np.complex are spelled with their modern names).
"""
import numpy as np

unsigned_integer_types = [np.uint, np.uint8, np.uint16, np.uint32, np.uint64]
signed_integer_types = [int, np.int_, np.intc, np.intp, np.int8, np.int16, np.int32, np.int64]
integer_types = unsigned_integer_types + signed_integer_types

real_types = [float, np.float16, np.float32, np.float64]
if hasattr(np, "float128"):
    real_types.append(np.float128)

complex_types = [complex, np.complex64, np.complex128]
if hasattr(np, "complex256"):
    complex_types.append(np.complex256)

numeric_types = integer_types + real_types + complex_types
allowed_types = numeric_types


def is_complex(dtype):
    return np.issubdtype(np.dtype(dtype), np.complexfloating)
