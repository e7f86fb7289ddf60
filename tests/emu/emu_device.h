// emu_device.h -- host stand-ins for the CUDA device environment (TEST INFRASTRUCTURE ONLY).
//
// tests/emu/build_emu.py compiles the library's own sources -- csrc/solvers.cu, csrc/ops.cu and
// the headers they include, unmodified -- for the host with g++ -DKRY_EMULATE, force-including
// this file.  What runs is the device *logic* of the product (kernel loops of spmv_row_kernel /
// vec_pass_kernel / vec_map_kernel, every solver functor, the launch sequences, the settle
// logic, the C ABI), executed by one host thread that plays every CUDA thread in turn.  What is
// NOT exercised: the memory system, warp shuffles, atomics, PTX, CUDA graphs, NCCL.  It lets the
// CPU test-suite check the solver state machines against the oracle when no GPU is at hand; it is
// never loaded by the product (pykrylov_b200/_lib.py knows nothing about it).
#pragma once
#include <cuda_runtime.h>

#include <stdint.h>

#include <cmath>
#include <cstring>

#define __launch_bounds__(...)

// ---- thread coordinates of the CUDA thread being played
struct EmuDim {
    unsigned x, y, z;
};
extern thread_local EmuDim threadIdx, blockIdx, blockDim, gridDim;

// ---- arithmetic intrinsics: the file is compiled with -ffp-contract=off, so a + b and a * b are
// the individually rounded operations __dadd_rn / __dmul_rn stand for
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
template <class T>
static inline T __ldg(const T *p) { return *p; }

// ---- L2 hint helpers of common.cuh: plain accesses
static inline uint64_t l2_policy_evict_first() { return 0; }
static inline uint64_t l2_policy_evict_last() { return 0; }
static inline double ldnc_hint(const double *a, uint64_t) { return *a; }
static inline int ldnc_hint(const int *a, uint64_t) { return *a; }
static inline double ld_hint(const double *a, uint64_t) { return *a; }
static inline double2 ld2_hint(const double *a, int i2, uint64_t) { return reinterpret_cast<const double2 *>(a)[i2]; }
static inline void st2_hint(double *a, int i2, double2 v, uint64_t) { reinterpret_cast<double2 *>(a)[i2] = v; }
static inline void st_hint(double *a, double v, uint64_t) { *a = v; }

// ---- reduction: every played thread adds its accumulators to the launch totals, in thread
// order; the launcher hands the totals to the finalize functor after the last thread
constexpr int EMU_MAX_DOTS = 4;
extern thread_local double emu_tot[EMU_MAX_DOTS];
extern thread_local int    emu_reduced;

struct ReduceWs;
template <int ND, class Fin>
static inline void block_reduce_finalize(double (&acc)[ND], const ReduceWs &, Fin &)
{
    for (int d = 0; d < ND; ++d) emu_tot[d] = emu_tot[d] + acc[d];
    emu_reduced = 1;
}

// ---- launcher: play grid x block threads one after the other
template <int ND, class Ws, class Fin, class Kernel>
static inline void emu_launch(int grid, int block, const Ws &ws, Fin fin, Kernel kernel)
{
    for (int d = 0; d < EMU_MAX_DOTS; ++d) emu_tot[d] = 0.0;
    emu_reduced = 0;
    gridDim = EmuDim{(unsigned)grid, 1, 1};
    blockDim = EmuDim{(unsigned)block, 1, 1};
    for (int b = 0; b < grid; ++b) {
        blockIdx = EmuDim{(unsigned)b, 0, 0};
        for (int t = 0; t < block; ++t) {
            threadIdx = EmuDim{(unsigned)t, 0, 0};
            kernel();
        }
    }
    if constexpr (ND > 0) {
        if (emu_reduced) {
            for (int d = 0; d < ND; ++d) ws.sums[d] = emu_tot[d];
            if (!ws.defer) fin(emu_tot);
        }
    }
}
