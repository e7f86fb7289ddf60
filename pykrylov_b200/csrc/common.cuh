// common.cuh -- shared host/device plumbing of libkrylov_b200 (sm_100a only).
//
// * error reporting for the C ABI (include/krylov_b200.h)
// * the context / vector / CSR handle structs
// * the deterministic two-stage reduction with fused "last block finalises"
//   scalar recurrence, used by every kernel on the hot path:
//     stage 1: warp-shuffle butterfly -> per-warp smem -> one partial per CTA
//     stage 2: the CTA that retires last (threadfence + atomic ticket) sums the
//              partials in a fixed order and runs the solver's scalar update
//              (alpha, beta, rho, omega, Givens ...) on device, so no host
//              round-trip and no extra launch is needed per inner product.
//   The summation order depends only on (grid, block) -- never on which CTA
//   happens to be last -- so results are reproducible run to run.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <type_traits>

#include "../../include/krylov_b200.h"

// ---------------------------------------------------------------- errors
void kry_set_error(const char *fmt, ...);

#define KRY_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e_ = (call);                                                    \
        if (e_ != cudaSuccess) {                                                    \
            kry_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,             \
                          cudaGetErrorString(e_));                                  \
            return (e_ == cudaErrorMemoryAllocation) ? KRY_ERR_NOMEM : KRY_ERR_CUDA;\
        }                                                                           \
    } while (0)

#define KRY_REQUIRE(cond, code, ...)                                                \
    do {                                                                            \
        if (!(cond)) {                                                              \
            kry_set_error(__VA_ARGS__);                                             \
            return (code);                                                          \
        }                                                                           \
    } while (0)

#define KRY_TRY(expr)                                                               \
    do {                                                                            \
        int rc_ = (expr);                                                           \
        if (rc_ != KRY_OK) return rc_;                                              \
    } while (0)

// ---------------------------------------------------------------- handles
constexpr int KRY_MAX_DOTS = 4;
constexpr int KRY_MAX_RANKS = 16;   // one NVSwitch domain
constexpr int KRY_HALO_FLAG_OFFSET = 1024;   // doubles into the inbox allocation: flag of source rank q at +8q

struct ReduceWs {
    double   *partials;   // [KRY_MAX_DOTS][stride] one partial per CTA
    double   *sums;       // [KRY_MAX_DOTS] totals (input of the NCCL all-reduce when sharded)
    unsigned *counter;    // retirement ticket, self-resetting
    int       stride;
    int       defer;      // 1: only write sums[]; the finalize functor runs after the all-reduce
    // fused all-reduce over NVLink peer memory (sharded runs, see block_reduce_finalize)
    int                 p2p;      // 1: exchange the totals through the peers' inboxes in-kernel
    int                 nranks, rank;
    double             *inbox;    // local  [2][nranks][8 packets of {32 data bits | 32-bit seq}]
    double *const      *peers;    // device array: peers[q] = rank q's inbox (IPC-mapped)
    unsigned long long *seq;      // reduction sequence number (identical on every rank)
};

struct kry_ctx {
    int          device;
    cudaStream_t stream;
    cudaStream_t copy_stream;  // small D2H reads that must not queue behind the compute stream (lazy)
    cudaEvent_t  ev0, ev1;
    int          sm_count;
    int64_t      l2_bytes;
    int64_t      smem_optin;
    double      *scalars;      // KRY_NUM_SLOTS user-visible scalar slots
    double      *partials;
    double      *sums;
    unsigned    *counter;
    int          partial_stride;
    int          partials_gen; // bumped whenever `partials` is reallocated (kry_ctx_ensure_partials)
    int         *never_done;   // device int == 0: "done" flag of the stand-alone ops
    int         *gate;         // optional: `done` flag of a device-resident scalar plane (lls.cu) gating the stand-alone ops
    void        *flush_buf;
    size_t       flush_bytes;
    int64_t      launches;
    // NVLink peer-memory all-reduce (comm.cu): inboxes of all ranks mapped through CUDA IPC
    int                 p2p_on;
    double             *p2p_inbox;
    double            **p2p_peers_dev;
    unsigned long long *p2p_seq;
    void               *p2p_peer_ptr[16];
    unsigned long long *halo_trace; // device [16], KRY_HALO_TRACE only
    unsigned long long  halo_seq;  // tag of the last fused halo exchange (host counter, same on every rank)
    int          l2_hints;     // bit 0: CG vector kernels use L2 eviction-priority hints (default 1)
    int          use_graphs;   // 1: solver loops replay CUDA graphs of 12 iterations (default)
    int          cg_fuse;      // KRY_OPT_CG_FUSE: CG launch plan (0: 3 launches, 1/2: fused 2-launch forms)
    int          cg_fuse_shards;   // KRY_OPT_CG_FUSE_SHARDS: the plan also applies to row shards
    int          halo_p2p;         // KRY_OPT_HALO_P2P: halo entries travel through peer memory from inside the SpMV launch
    int          cg_one_cta;       // KRY_OPT_CG_ONE_CTA: small problems iterate inside one CTA
    int          minres_fuse;      // KRY_OPT_MINRES_FUSE: 2-launch MINRES plan
    int          minres_persistent;    // KRY_OPT_MINRES_PERSISTENT: one cooperative kernel per iterate call
    // lifetime: vectors / operators / solvers hold a reference; kry_ctx_destroy releases the
    // device resources at once but the struct itself lives until the last child is destroyed,
    // so handles may be destroyed in any order (interpreter shutdown does exactly that)
    int          refs;
    int          closed;
    // optional per-launch timing of the dominant kernel (kry_prof_*)
    cudaEvent_t *prof_ev;      // 2 * prof_cap events
    int          prof_cap, prof_n;
    // multi-GPU
    void        *nccl;         // ncclComm_t
    int          nranks, rank;
};

struct CsrDev {
    int     *rowptr = nullptr;
    int     *col    = nullptr;
    double  *val    = nullptr;
    int     *rowblk = nullptr;   // first row of each nnz tile, [nblocks+1]
    int      nblocks = 0;
    int      tile_nnz = 0;       // tile the rowblk partition was built for
    int      max_row = 0;        // longest row
    int64_t  nrows = 0, ncols = 0, nnz = 0;
};

struct HaloPlan {                // 1-D row sharding
    bool     active = false;
    int64_t  n_global = 0, row_begin = 0;
    int      n_send = 0;         // boundary entries this rank publishes
    int      max_send = 0;       // padded per-rank slot in the all-gather
    int     *send_idx = nullptr; // local indices to pack
    double  *send_buf = nullptr; // [max_send]
    // peer-memory halo (KRY_OPT_HALO_P2P): who reads my boundary entries, whose entries I read,
    // and the row order that puts the rows touching halo columns last
    int      n_to = 0, n_from = 0;
    int      to_rank[KRY_MAX_RANKS], from_rank[KRY_MAX_RANKS];
    int      rot = 0;            // virtual row v maps to row (v + rot) mod nrows ...
    int      v_wait = 0;         // ... and only rows with v >= v_wait may touch halo columns
};

// Per gathered vector of a solver (device memory): where this rank's boundary entries land in
// each reader's copy of that vector, and the arrival flags (one 64-bit tag per directed pair of
// ranks, living in the IPC-mapped inbox allocation of the reader).
struct HaloTable {
    int                       n_to, n_from;
    double                   *to_tail[KRY_MAX_RANKS];     // reader q's tail slot for my entries
    unsigned long long       *to_flag[KRY_MAX_RANKS];     // reader q's flag word for me
    const unsigned long long *from_flag[KRY_MAX_RANKS];   // my flag words for the ranks I read from
};

struct HaloArgs {                // by value into the sharded SpMV launch
    const HaloTable   *tbl;
    const int         *send_idx;
    int                n_send, push_ctas;   // the LAST push_ctas CTAs of the grid only publish, they own no rows
    int                row_ctas;            // CTAs that walk rows: gridDim.x - push_ctas
    unsigned          *ticket;   // self-resetting retirement ticket of the push CTAs
    unsigned long long tag;      // monotonic per context, identical on every rank
    int                rot, v_wait;
    int                skip_push;   // bit 0: no push (the entries travelled by ncclAllGather, or tests/emu played
                                    // the push in the launcher); bit 1: no wait
    int                emu_phase;   // tests/emu SIMT mode only: 1 = publishing CTAs act, 2 = row CTAs act (0: all)
    unsigned long long *trace;      // optional (KRY_HALO_TRACE): in-kernel %globaltimer statistics, see kry_halo_trace_read
};

struct kry_csr {
    kry_ctx *ctx;
    CsrDev   A, T;
    bool     has_T = false;
    uint32_t flags = 0;
    int      kind = KRY_SPMV_AUTO;
    int      tile_nnz = 0, threads = 0;
    HaloPlan halo;
};

struct kry_vec {
    kry_ctx *ctx;
    double  *d;
    int64_t  n;        // logical length
    int64_t  cap;      // allocated length (>= n; sharded x vectors carry the halo tail)
    bool     owned;
};

// CSR plumbing shared by context.cu and assemble.cu
void csr_dev_free(CsrDev &m);
int  csr_dev_alloc(CsrDev &m, int64_t nrows, int64_t ncols, int64_t nnz);
int  csr_validate(kry_ctx *c, const CsrDev &m);
int  csr_finish(kry_ctx *c, CsrDev &m);
int  check_sizes(int64_t nrows, int64_t ncols, int64_t nnz);
int  kry_ctx_ensure_partials(kry_ctx *ctx, int nblocks);
void kry_ctx_retain(kry_ctx *ctx);
void kry_ctx_release(kry_ctx *ctx);     // child destroyed: frees a closed context with the last one
// every entry point that touches the device through a child handle starts with this
#define KRY_CTX_LIVE(ctx, who)                                                      \
    KRY_REQUIRE(!(ctx)->closed, KRY_ERR_STATE, "%s: the context of this handle was destroyed", who)
int  kry_alloc(void **p, size_t bytes);
static inline const int *kry_gate(const kry_ctx *c) { return c->gate ? c->gate : c->never_done; }
ReduceWs kry_ws(kry_ctx *ctx);

// ---------------------------------------------------------------- device
// KRY_EMULATE: tests/emu compiles the device *logic* of this library (kernel loops, solver
// functors, launch sequences) for the host with g++ to check it against the oracle without a
// GPU.  It is test infrastructure: nothing in the product defines the macro, and the blocks it
// switches off below (PTX) have their stand-ins in tests/emu/emu_device.h.
#if defined(__CUDACC__) || defined(KRY_EMULATE)

#ifdef KRY_EMULATE      // the genuine two-stage reduction below is compiled as ..._real: the emulation runs
#define block_reduce_finalize block_reduce_finalize_real   // it on SIMT fibers, or a sequential stand-in
#endif                  // in its place for speed (tests/emu/emu_device.h picks)
__device__ __forceinline__ double warp_sum(double v)
{
    // xor butterfly: every lane ends with the same, order-fixed sum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of ND accumulators; result valid in thread 0.
template <int ND>
__device__ __forceinline__ void block_sum(double (&acc)[ND], double (*s_warp)[32])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        acc[d] = warp_sum(acc[d]);
        if (lane == 0) s_warp[d][warp] = acc[d];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            double v = (lane < nwarps) ? s_warp[d][lane] : 0.0;
            acc[d] = warp_sum(v);
        }
    }
}

// Stage 1 + stage 2 + fused scalar recurrence. Must be reached by all threads of
// all CTAs of the launch. `fin(const double *totals)` runs in exactly one thread.
template <int ND, class Fin>
__device__ __forceinline__ void block_reduce_finalize(double (&acc)[ND], const ReduceWs &ws, Fin &fin)
{
    __shared__ double s_warp[ND][32];
    __shared__ int    s_last;
    block_sum<ND>(acc, s_warp);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) ws.partials[(size_t)d * ws.stride + blockIdx.x] = acc[d];
        __threadfence();
        const unsigned ticket = atomicAdd(ws.counter, 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double tot[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        double a = 0.0;
        const volatile double *p = ws.partials + (size_t)d * ws.stride;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) a = __dadd_rn(a, p[i]);
        tot[d] = a;
    }
    __syncthreads();            // s_warp reuse
    block_sum<ND>(tot, s_warp);
    if (ws.p2p) {
        // One-shot all-reduce fused into this kernel, over NVLink peer memory: thread q of the
        // last CTA stores this rank's totals into rank q's inbox and spins until rank q's
        // contribution with the same sequence number has landed in the local inbox.  Every rank
        // then sums the nranks contributions in rank order, so all ranks hold bit-identical totals
        // and run the same scalar recurrence -- no NCCL call, no extra launch.
        // Wire format: each double travels as two self-validating 8-byte packets
        // {32 data bits | 32-bit sequence number} (aligned 8-byte stores are single-copy atomic,
        // also over NVLink -- the scheme of NCCL's LL protocol), so neither side needs a
        // system-scope fence (measured at 3.5 us each on B200, profiles/r2d_halo_trace_n8.txt).
        // Slots alternate with the sequence parity: a rank can be at most one reduction ahead of
        // its slowest peer.
        __shared__ double             s_tot[ND];
        __shared__ double             s_in[KRY_MAX_RANKS][ND];
        __shared__ unsigned long long s_seq;
        if (threadIdx.x == 0) {
#pragma unroll
            for (int d = 0; d < ND; ++d) s_tot[d] = tot[d];
            s_seq = *ws.seq + 1ull;
            *ws.seq = s_seq;
        }
        __syncthreads();
        const unsigned long long seq = s_seq;
        const unsigned long long tag = (seq & 0xffffffffull) << 32;
        const size_t slot = (size_t)(seq & 1ull) * ws.nranks;
        if ((int)threadIdx.x < ws.nranks) {
            volatile unsigned long long *dst =
                reinterpret_cast<volatile unsigned long long *>(ws.peers[threadIdx.x] + (slot + ws.rank) * 8);
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const unsigned long long bits = (unsigned long long)__double_as_longlong(s_tot[d]);
                dst[2 * d] = tag | (bits & 0xffffffffull);
                dst[2 * d + 1] = tag | (bits >> 32);
            }
            const volatile unsigned long long *src =
                reinterpret_cast<const volatile unsigned long long *>(ws.inbox + (slot + threadIdx.x) * 8);
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                unsigned long long lo, hi;
                while (((lo = src[2 * d]) & 0xffffffff00000000ull) != tag) __nanosleep(20);
                while (((hi = src[2 * d + 1]) & 0xffffffff00000000ull) != tag) __nanosleep(20);
                s_in[threadIdx.x][d] = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                double a = 0.0;
                for (int q = 0; q < ws.nranks; ++q) a = __dadd_rn(a, s_in[q][d]);
                tot[d] = a;
                ws.sums[d] = a;
            }
            *ws.counter = 0u;
            fin(tot);
        }
        return;
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) ws.sums[d] = tot[d];
        *ws.counter = 0u;
        if (!ws.defer) fin(tot);
    }
}
#ifdef KRY_EMULATE
#undef block_reduce_finalize
#endif

// Bodies that also provide pair(i2[, acc]) -- elements 2*i2 and 2*i2+1 through one
// 16-byte access per vector -- are run in that form by the vector kernels.
template <class B, class = void>
struct has_pair : std::false_type {};
template <class B>
struct has_pair<B, std::void_t<decltype(B::kPair)>> : std::true_type {};

__device__ __forceinline__ double2 ld2(const double *p, int i2)
{
    return reinterpret_cast<const double2 *>(p)[i2];
}
__device__ __forceinline__ void st2(double *p, int i2, double2 v)
{
    reinterpret_cast<double2 *>(p)[i2] = v;
}

#ifndef KRY_EMULATE
// L2 eviction-priority hints (createpolicy + ld/st .L2::cache_hint).  The three CG
// launches hand 80 MB vectors to each other (Ap: K1->K2, r: K2->K3, p: K3->K1); marking
// the producer's stores evict_last and the pure streams (CSR arrays, x) evict_first lets
// part of that traffic be served from the 126 MB L2 instead of HBM.
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double ldnc_hint(const double *a, uint64_t pol)
{
    double d;
    asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(d) : "l"(a), "l"(pol));
    return d;
}
__device__ __forceinline__ int ldnc_hint(const int *a, uint64_t pol)
{
    int d;
    asm("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(d) : "l"(a), "l"(pol));
    return d;
}
__device__ __forceinline__ double ld_hint(const double *a, uint64_t pol)   // coherent path (data written in-kernel)
{
    double d;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(d) : "l"(a), "l"(pol));
    return d;
}
__device__ __forceinline__ double2 ld2_hint(const double *a, int i2, uint64_t pol)
{
    double2 d;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(d.x), "=d"(d.y)
                 : "l"(reinterpret_cast<const double2 *>(a) + i2), "l"(pol));
    return d;
}
__device__ __forceinline__ void st2_hint(double *a, int i2, double2 v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(
                     reinterpret_cast<double2 *>(a) + i2),
                 "d"(v.x), "d"(v.y), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void st_hint(double *a, double v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(a), "d"(v), "l"(pol) : "memory");
}
#endif  // !KRY_EMULATE

struct NoFin {
    __device__ void operator()(const double *) const {}
};

// Finalize that publishes the totals into the context's scalar slots.
struct SlotFin {
    double *slots;
    int     n;
    __device__ void operator()(const double *t) const
    {
        for (int d = 0; d < n; ++d) slots[d] = t[d];
    }
};

// Deferred scalar update (after an ncclAllReduce of ws.sums on sharded runs).
template <class Fin>
__global__ void finalize_kernel(Fin fin, const double *sums, const int *done)
{
    if (*done) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) fin(sums);
}

// Resident CTAs per SM a body is compiled for (register cap = 65536 / (256 * n)):
// 6 (<= 40 registers) unless the body asks for fewer with `static constexpr int kMinBlocks`.
template <class B, class = void>
struct body_min_blocks : std::integral_constant<int, 6> {};
template <class B>
struct body_min_blocks<B, std::void_t<decltype(B::kMinBlocks)>> : std::integral_constant<int, B::kMinBlocks> {};

// Generic fused vector pass: body(i, acc) over all elements + reduction + finalize.
template <int ND, class Body, class Fin>
__global__ void __launch_bounds__(256, body_min_blocks<Body>::value)   // grid = one resident wave
vec_pass_kernel(int64_t n, Body body, ReduceWs ws, Fin fin, const int *done)
{
    static_assert(ND >= 1 && ND <= KRY_MAX_DOTS, "1..KRY_MAX_DOTS fused inner products");
    if (*done) return;
    body.init();
    double acc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) acc[d] = 0.0;
    // n < 2^31 - 2^20 (check_sizes) and stride < 2^20: 32-bit indices cannot overflow
    const int stride = (int)(gridDim.x * blockDim.x), nn = (int)n;
    const int t0 = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if constexpr (has_pair<Body>::value) {
        // 16-byte (double2) accesses: two consecutive elements per thread per trip
        const int half = nn >> 1;
        for (int i = t0; i < half; i += stride) body.pair(i, acc);
        if ((nn & 1) && t0 == 0) body(nn - 1, acc);
    } else {
        for (int i = t0; i < nn; i += stride) body(i, acc);
    }
    block_reduce_finalize<ND>(acc, ws, fin);
}

// Same without inner products (pure element-wise update).
template <class Body>
__global__ void __launch_bounds__(256, body_min_blocks<Body>::value)
vec_map_kernel(int64_t n, Body body, const int *done)
{
    if (*done) return;
    body.init();
    const int stride = (int)(gridDim.x * blockDim.x), nn = (int)n;
    const int t0 = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if constexpr (has_pair<Body>::value) {
        const int half = nn >> 1;
        for (int i = t0; i < half; i += stride) body.pair(i);
        if ((nn & 1) && t0 == 0) body(nn - 1);
    } else {
        for (int i = t0; i < nn; i += stride) body(i);
    }
}

#endif  // __CUDACC__ || KRY_EMULATE
