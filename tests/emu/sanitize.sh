#!/bin/bash
# Run the emulated device-logic tests (fast mode) under AddressSanitizer + UndefinedBehaviorSanitizer:
# the library's own solvers.cu / ops.cu / lls.cu / comm.cu, built for the host with -fsanitize, behind the
# unchanged Python layer.  Fibers (SIMT mode) and ASan do not mix, so the SIMT modules are left out.
#   bash tests/emu/sanitize.sh            (from the repo root; ~4 min)
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
cd "$ROOT"
python tests/emu/build_emu.py > /dev/null
OUT=tests/emu/_build/libkrylov_emu.so
cp "$OUT" "$OUT.plain"
trap 'mv "$OUT.plain" "$OUT"; touch "$OUT"' EXIT
g++ -std=c++17 -O1 -g -fPIC -shared -Wl,-Bsymbolic -fsanitize=address,undefined -fno-omit-frame-pointer \
    -ffp-contract=off -fno-fast-math -DKRY_EMULATE -Wno-unknown-pragmas -include tests/emu/emu_device.h \
    -I"${CUDA_HOME:-/usr/local/cuda}/include" -Ipykrylov_b200/csrc -x c++ \
    pykrylov_b200/csrc/solvers.cu pykrylov_b200/csrc/ops.cu pykrylov_b200/csrc/lls.cu pykrylov_b200/csrc/comm.cu tests/emu/emu_context.cpp \
    -o "$OUT" -lpthread
touch "$OUT"
LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" \
ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1 \
python -m pytest tests/test_emulated_device_logic.py tests/test_emulated_solvers_fuzz.py \
    tests/test_emulated_cg_plans_fuzz.py tests/test_emulated_multi_rank.py tests/test_emulated_scalar_planes.py \
    tests/test_emulated_minres_plans_fuzz.py -q -p no:cacheprovider -k "not (2-1 or 3-1)"
