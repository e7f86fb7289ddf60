"""pykrylov_b200 -- B200-native Krylov iteration engine behind the pykrylov API.

The per-iteration hot path (CSR SpMV, inner products, AXPY updates; fp64) runs as
hand-written sm_100a CUDA kernels in libkrylov_b200.so, reached through a ctypes
C-ABI shim (include/krylov_b200.h).  Host code is pure Python + NumPy.
"""
__version__ = "0.1.0"
