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
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
template <class T>
static inline T __ldcg(const T *p) { return *reinterpret_cast<const volatile T *>(p); }

// ---- L2 hint helpers of common.cuh: plain accesses
static inline uint64_t l2_policy_evict_first() { return 0; }
static inline uint64_t l2_policy_evict_last() { return 0; }
static inline double ldnc_hint(const double *a, uint64_t) { return *a; }
static inline int ldnc_hint(const int *a, uint64_t) { return *a; }
static inline double ld_hint(const double *a, uint64_t) { return *a; }
static inline double2 ld2_hint(const double *a, int i2, uint64_t) { return reinterpret_cast<const double2 *>(a)[i2]; }
static inline void st2_hint(double *a, int i2, double2 v, uint64_t) { reinterpret_cast<double2 *>(a)[i2] = v; }
static inline void st_hint(double *a, double v, uint64_t) { *a = v; }

// ---- block-level CUDA, played by fibers (emu_context.cpp): when the SIMT mode is on
// (kry_emu_set_fibers), every thread of a block is a fiber, __syncthreads() and the warp shuffles
// are real barriers between them, __shared__ variables are shared by the block, and the
// library's genuine block_reduce_finalize (warp butterflies, per-warp partials, last-CTA ticket,
// ordered final sum) runs as written.  The default mode plays the threads one after the other
// -- only valid for kernels without block-level synchronisation -- and replaces the reduction
// by a sequential stand-in; it is ~100x faster and is what most emulated tests use.
#undef __shared__
#define __shared__ static thread_local      // per host thread: the ranks of a multi-GPU run are threads

extern int emu_fibers_on;
void emu_yield();
void emu_block_barrier();
void emu_warp_barrier(int warp);
extern thread_local unsigned char emu_shfl_slots[32][32][8];

static inline void __syncthreads() { emu_block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_barrier((int)(threadIdx.x >> 5)); }
static inline void __threadfence() {}
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
void emu_spin_wait();
static inline void __nanosleep(unsigned) { emu_spin_wait(); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline int atomicMax(int *p, int v) { const int o = *p; if (v > o) *p = v; return o; }
static inline int atomicMin(int *p, int v) { const int o = *p; if (v < o) *p = v; return o; }

template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask)
{
    static_assert(sizeof(T) <= 8, "shuffle of up to 8 bytes");
    const int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5);
    memcpy(emu_shfl_slots[warp][lane], &v, sizeof(T));
    emu_warp_barrier(warp);
    T r;
    memcpy(&r, emu_shfl_slots[warp][lane ^ lane_mask], sizeof(T));
    emu_warp_barrier(warp);
    return r;
}

// ---- reduction: the genuine code (common.cuh, compiled as block_reduce_finalize_real) in SIMT
// mode; otherwise every played thread adds its accumulators to the launch totals, in thread
// order, and the launcher hands the totals to the finalize functor after the last thread
constexpr int EMU_MAX_DOTS = 4;
extern thread_local double emu_tot[EMU_MAX_DOTS];
extern thread_local int    emu_reduced;

struct ReduceWs;
template <int ND, class Fin>
void block_reduce_finalize_real(double (&acc)[ND], const ReduceWs &ws, Fin &fin);

template <int ND, class Fin>
static inline void block_reduce_finalize(double (&acc)[ND], const ReduceWs &ws, Fin &fin)
{
    if (emu_fibers_on) {
        block_reduce_finalize_real<ND>(acc, ws, fin);
        return;
    }
    for (int d = 0; d < ND; ++d) emu_tot[d] = emu_tot[d] + acc[d];
    emu_reduced = 1;
}

// ---- dynamic shared memory of a launch (extern __shared__ in CUDA)
extern unsigned char *emu_dynamic_smem;
void emu_set_dynamic_smem(size_t bytes);

// ---- launchers
void emu_p2p_allreduce(const ReduceWs &ws, double *tot, int nd);
void emu_launch_fibers(int grid, int block, const void *kernel_closure, void (*invoke)(const void *));
void emu_launch_fibers_mode(int grid, int block, const void *kernel_closure, void (*invoke)(const void *), int cooperative);

template <int ND, class Ws, class Fin, class Kernel>
static inline void emu_launch(int grid, int block, const Ws &ws, Fin fin, Kernel kernel)
{
    if (emu_fibers_on && ND > 0) {
        // SIMT mode: blocks one after the other, the threads of a block as fibers; the kernel's own
        // last-CTA code finalises
        emu_launch_fibers(grid, block, &kernel, [](const void *k) { (*static_cast<const Kernel *>(k))(); });
        return;
    }
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
            if (ws.p2p) emu_p2p_allreduce(ws, emu_tot, ND);      // sharded runs: all-reduce over the inboxes
            for (int d = 0; d < ND; ++d) ws.sums[d] = emu_tot[d];
            if (!ws.defer) fin(emu_tot);
        }
    }
}
