#!/usr/bin/env python3
"""Build tests/emu/_build/libkrylov_emu.so: the library's own csrc/solvers.cu and csrc/ops.cu
(with the headers they include) compiled for the HOST by g++ -DKRY_EMULATE, next to a host
stand-in for context.cu / comm.cu and the CUDA runtime calls (emu_context.cpp).

TEST INFRASTRUCTURE ONLY: the product never loads this library; tests/test_emulated_device_logic.py
swaps it in behind the ctypes layer to run the GPU parity tests' logic on a machine without a GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pykrylov_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libkrylov_emu.so")
SOURCES = [os.path.join(CSRC, "solvers.cu"), os.path.join(CSRC, "ops.cu"), os.path.join(CSRC, "lls.cu"),
           os.path.join(CSRC, "comm.cu"),
           os.path.join(HERE, "emu_context.cpp")]
DEPS = SOURCES + [os.path.join(CSRC, h) for h in ("common.cuh", "spmv.cuh", "launch.cuh", "solver.cuh")] + \
    [os.path.join(HERE, "emu_device.h"), os.path.join(ROOT, "include", "krylov_b200.h")]


def cuda_include():
    for d in (os.environ.get("CUDA_HOME", ""), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include", "cuda_runtime.h")):
            return os.path.join(d, "include")
    return None


def build(force=False):
    """Returns the path of the emulation library, or None when it cannot be built here
    (no g++ / no CUDA headers)."""
    inc = cuda_include()
    if inc is None:
        return None
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-Wl,-Bsymbolic",
           # un-fused, individually rounded multiplies and adds: the contract of __dmul_rn / __dadd_rn
           "-ffp-contract=off", "-fno-fast-math",
           "-DKRY_EMULATE", "-Wno-unknown-pragmas", "-include", os.path.join(HERE, "emu_device.h"),
           "-I" + inc, "-I" + CSRC, "-x", "c++"] + SOURCES + ["-o", OUT, "-lpthread"]
    try:
        subprocess.check_call(cmd)
    except (OSError, subprocess.CalledProcessError) as exc:
        print("emulation build failed: %s" % exc, file=sys.stderr)
        return None
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
