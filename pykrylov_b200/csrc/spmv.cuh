// spmv.cuh -- CSR SpMV kernels with fused row-local epilogue + inner products.
//
// Replaces the operator closure behind `op * x` (reference
// pykrylov/linop/linop.py:362-369 -> :356-360 -> :271-298 -> :697-706) together
// with the np.dot that follows it in every solver loop (cg/cg.py:115-117,
// bicgstab/bicgstab.py:101-103,125-127, minres/minres.py:239-245).
//
// Parity contract: each row is summed sequentially in
// storage order with an un-fused multiply and add,
//     sum = 0;  sum = sum + (val[k] * x[col[k]])   for k in row,
// which is bit-for-bit scipy's csr_matvec.  All kernels below keep that order;
// they differ only in how the nnz stream reaches the SM.
//
//   spmv_row_kernel    one thread per row, direct loads (fallback for very long rows)
//   spmv_stream_kernel one CTA per nnz tile: val/col are read fully coalesced,
//                      products are staged in shared memory, then one thread per
//                      row sums its (short) row from smem.  HBM sees only
//                      contiguous 128-byte streams; the x gather hits L1/L2.
//   spmv_tma_kernel    persistent CTAs; val/col tiles are fetched by the TMA unit
//                      (cp.async.bulk + mbarrier, multi-stage ring) while the
//                      previous tile is being reduced.
//
// Gather functor  : double operator()(int col)   -- x[col], possibly transformed
// Epilogue functor: void operator()(int row, double Ax, double *acc) -- writes y
//                   (and anything row-local) and accumulates fused dot terms.
#pragma once

#include "common.cuh"

struct CsrView {
    const int    *rowptr;
    const int    *col;
    const double *val;
    const int    *rowblk;
    int           nrows;
    int           nblocks;
    int           hints;      // L2 evict_first on the CSR streams (row kernel)
};

static inline CsrView csr_view(const CsrDev &m)
{
    CsrView v;
    v.rowptr = m.rowptr; v.col = m.col; v.val = m.val; v.rowblk = m.rowblk;
    v.nrows = (int)m.nrows; v.nblocks = m.nblocks; v.hints = 0;
    return v;
}

#if defined(__CUDACC__) || defined(KRY_EMULATE)

// Besides operator()(col) every gather functor states, for the sharded launch with the halo
// exchange fused in (spmv_row_shard_kernel):
//   boundary(j)  what this rank publishes for its local entry j -- the raw vector entry, such
//                that the reader's gather of the column it lands in gives operator()'s value;
//   coherent(c)  operator() with the gathered vector read by ordinary (weak) loads instead of the
//                read-only path of __ldg.  Halo entries are written by peer GPUs while the kernel
//                runs; a thread reads them only after it has seen the peers' flags and executed a
//                system-scope fence (halo_wait), which is the acquire pattern of the PTX memory
//                model for weak loads -- ld.global.nc gives no such guarantee.  One unconditional
//                load for local and halo columns alike: a predicated pair of loads (read-only path
//                for local, ld.cg for halo columns) measured 2.6 % slower on the fused CG launch
//                (profiles/r2m_*).
struct GatherPlain {
    const double *x;
    __device__ void   init() {}
    __device__ double operator()(int c) const { return __ldg(x + c); }
    __device__ double boundary(int j) const { return x[j]; }
    __device__ double coherent(int c) const { return x[c]; }
};

// x[c] * s with s read from device memory (MINRES: v = (1/beta) * y on the fly)
struct GatherScaled {
    const double *x;
    const double *s_ptr;
    double        s;
    __device__ void   init() { s = *s_ptr; }
    __device__ double operator()(int c) const { return __dmul_rn(s, __ldg(x + c)); }
    __device__ double boundary(int j) const { return x[j]; }
    __device__ double coherent(int c) const { return __dmul_rn(s, x[c]); }
};

// Halo-aware view of a gather for row shards (columns >= n_local address the halo tail).
template <class G>
struct HaloGather {
    G   g;
    int n_local;
    __device__ void   init() { g.init(); }
    __device__ double operator()(int c) const { return g.coherent(c); }
    __device__ double boundary(int j) const { return g.boundary(j); }
};

// ------------------------------------------------------------ thread per row
// Resident CTAs per SM an epilogue asks the row kernel to be compiled for (register cap
// 65536 / (256 * n)); 0 = no constraint beyond the 256-thread block (same as omitting it).
template <class E, class = void>
struct epi_min_blocks : std::integral_constant<int, 0> {};
template <class E>
struct epi_min_blocks<E, std::void_t<decltype(E::kMinBlocks)>> : std::integral_constant<int, E::kMinBlocks> {};

template <int ND, class Gather, class Epi, class Fin>
__global__ void __launch_bounds__(256, epi_min_blocks<Epi>::value)
spmv_row_kernel(CsrView A, Gather g, Epi epi, ReduceWs ws, Fin fin, const int *done)
{
    if (*done) return;
    g.init();
    epi.init();
    double acc[ND > 0 ? ND : 1];
#pragma unroll
    for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
    const int stride = gridDim.x * blockDim.x;
    if (A.hints) {
        // CSR arrays are pure streams: evict_first in L2 (still L1-cached for the row walk)
        const uint64_t ef = l2_policy_evict_first();
        for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < A.nrows; row += stride) {
            const int s = ldnc_hint(A.rowptr + row, ef), e = ldnc_hint(A.rowptr + row + 1, ef);
            double sum = 0.0;
            for (int k = s; k < e; ++k)
                sum = __dadd_rn(sum, __dmul_rn(ldnc_hint(A.val + k, ef), g(ldnc_hint(A.col + k, ef))));
            epi(row, sum, acc);
        }
    } else {
        for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < A.nrows; row += stride) {
            const int s = __ldg(A.rowptr + row), e = __ldg(A.rowptr + row + 1);
            double sum = 0.0;
            for (int k = s; k < e; ++k)
                sum = __dadd_rn(sum, __dmul_rn(__ldg(A.val + k), g(__ldg(A.col + k))));
            epi(row, sum, acc);
        }
    }
    if constexpr (ND > 0) block_reduce_finalize<ND>(acc, ws, fin);
}

// ------------------------------------- thread per row, software-pipelined
// Same mapping and summation order as spmv_row_kernel.  A warp-trip of that kernel is a chain
// of three dependent memory latencies (rowptr -> col/val -> gathers); here the row pointers
// are loaded two trips ahead (registers) and, one trip ahead, every lane asks L2 for the line
// that holds the start of its next row (prefetch.global.L2: no register, no dependency), so
// the next trip finds rowptr in a register and col/val in L2.
__device__ __forceinline__ void prefetch_l2(const void *p)
{
#ifndef KRY_EMULATE
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

template <int ND, int DEPTH, class Gather, class Epi, class Fin>
__global__ void __launch_bounds__(256, epi_min_blocks<Epi>::value)
spmv_rowpf_kernel(CsrView A, Gather g, Epi epi, ReduceWs ws, Fin fin, const int *done)
{
    static_assert(DEPTH == 1 || DEPTH == 2, "row pointers one or two trips ahead");
    if (*done) return;
    g.init();
    epi.init();
    double acc[ND > 0 ? ND : 1];
#pragma unroll
    for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
    const int stride = gridDim.x * blockDim.x;      // < 2^20, rows < 2^31 - 2^20: no overflow below
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    int s = 0, e = 0, s1 = 0, e1 = 0;
    if (row < A.nrows) {
        s = __ldg(A.rowptr + row);
        e = __ldg(A.rowptr + row + 1);
    }
    if constexpr (DEPTH == 2) {
        if (row + stride < A.nrows) {
            s1 = __ldg(A.rowptr + row + stride);
            e1 = __ldg(A.rowptr + row + stride + 1);
        }
    }
    while (row < A.nrows) {
        int s2 = 0, e2 = 0;
        if constexpr (DEPTH == 2) {
            // trip t+1: its CSR window towards L2 now; trip t+2: its row pointers into registers
            if (row + stride < A.nrows) {
                prefetch_l2(A.val + s1);
                prefetch_l2(A.col + s1);
                if (row + 2 * stride < A.nrows) {
                    s2 = __ldg(A.rowptr + row + 2 * stride);
                    e2 = __ldg(A.rowptr + row + 2 * stride + 1);
                }
            }
        } else {
            // trip t+1: its row pointers into registers while this trip's loads are in flight
            if (row + stride < A.nrows) {
                s1 = __ldg(A.rowptr + row + stride);
                e1 = __ldg(A.rowptr + row + stride + 1);
            }
        }
        double sum = 0.0;
        for (int k = s; k < e; ++k)
            sum = __dadd_rn(sum, __dmul_rn(__ldg(A.val + k), g(__ldg(A.col + k))));
        epi(row, sum, acc);
        row += stride;
        s = s1;
        e = e1;
        if constexpr (DEPTH == 2) {
            s1 = s2;
            e1 = e2;
        }
    }
    if constexpr (ND > 0) block_reduce_finalize<ND>(acc, ws, fin);
}

// ------------------------------------- thread per row, halo exchange fused in
// Row shards with KRY_OPT_HALO_P2P: ONE launch does the exchange of the boundary entries and
// the SpMV (+ epilogue + fused inner products + in-kernel all-reduce).
//   1. A few dedicated CTAs at the end of the grid (they own no rows: a CTA that pushed first and
//      walked rows afterwards started ~10 us late and ended the launch ~10 us late, measured)
//      store this rank's boundary entries g.boundary(send_idx[i]) straight into the halo tails of
//      the ranks that read them, over NVLink peer memory (CUDA IPC mapping of the readers' solver
//      slabs); the one that finishes last publishes the launch's tag in every reader's flag word
//      (release fence in between).  They take part in the reduction with zero partial sums.
//   2. All CTAs walk the rows in the rotated order v -> (v + rot) mod nrows, which puts the rows
//      that may touch halo columns last (HaloPlan::rot / v_wait, found at kry_csr_shard_finalize;
//      for a banded operator they are two thin slabs at the ends of the shard): the interior is
//      multiplied while the peers' entries are in flight, and every CTA stays in step with its
//      neighbours (measured: a CTA that stalls 25 us at the start of the launch falls four trips
//      behind the wave, out of the L2 window the others share, and never catches up -- the launch
//      then takes 12 % longer; profiles/r2e_*).
//   3. A thread that reaches v >= v_wait first waits until every rank it reads from has
//      published this tag and fences; the gather reads with weak loads (Gather::coherent).
// Every sharded row launch uses this kernel and this row order -- also when the entries travelled
// by pack kernel + ncclAllGather (skip bits set) -- so the fused inner products are the same
// bits whichever way the halo travelled.
// Ordering argument (why a tail is never overwritten while it is still being read, and why no
// wait can be circular): DESIGN.md section 6.  No pack launch, no ncclAllGather.
#ifndef KRY_EMULATE
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#else
static inline unsigned long long global_ns() { return 0; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; if (v > o) *p = v; return o; }
#endif

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long *p)
{
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

template <class Gather>
__device__ __forceinline__ void halo_push_entries(const HaloArgs &h, const Gather &g)
{
    const HaloTable *T = h.tbl;
    const int n_to = T->n_to;
    const int first = (int)gridDim.x - h.push_ctas;       // the last push_ctas CTAs of the grid publish
    for (int i = ((int)blockIdx.x - first) * blockDim.x + threadIdx.x; i < h.n_send; i += h.push_ctas * blockDim.x) {
        const double v = g.boundary(__ldg(h.send_idx + i));
        for (int q = 0; q < n_to; ++q) T->to_tail[q][i] = v;
    }
}

// the load that observes a peer's flag is an acquire at system scope: the weak loads of the halo
// entries that follow are ordered behind it (measured: a separate fence.sys costs 3.4 us per warp)
__device__ __forceinline__ unsigned long long ld_flag_acquire(const unsigned long long *p)
{
#ifdef KRY_EMULATE
    const unsigned long long v = *reinterpret_cast<const volatile unsigned long long *>(p);
    __threadfence_system();
    return v;
#else
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#endif
}

// release / acquire fences at system scope (fence.acq_rel.sys): all the message passing below needs;
// __threadfence_system() is the sequentially consistent fence.sc.sys
__device__ __forceinline__ void fence_acq_rel_sys()
{
#ifdef KRY_EMULATE
    __threadfence_system();
#else
    asm volatile("fence.acq_rel.sys;" ::: "memory");
#endif
}

__device__ __forceinline__ void halo_publish(const HaloArgs &h)
{
    const HaloTable *T = h.tbl;
    fence_acq_rel_sys();
    for (int q = 0; q < T->n_to; ++q)
        *reinterpret_cast<volatile unsigned long long *>(T->to_flag[q]) = h.tag;
}

template <class Gather>
__device__ __forceinline__ void halo_push(const HaloArgs &h, const Gather &g, const unsigned long long &t_entry)
{
    halo_push_entries(h, g);
    __syncthreads();                  // the CTA's stores happen-before thread 0's fence (cumulative)
    if (threadIdx.x == 0) {
        fence_acq_rel_sys();
        const unsigned ticket = atomicAdd(h.ticket, 1u);
        if (ticket == (unsigned)h.push_ctas - 1u) {
            *h.ticket = 0u;
            halo_publish(h);
            if (h.trace) {                       // [0] sum of (publish - this CTA's entry), [1] launches
                atomicAdd(h.trace + 0, global_ns() - t_entry);
                atomicAdd(h.trace + 1, 1ull);
            }
        }
    }
}

#ifdef KRY_EMULATE
// tests/emu, fast mode (threads played one after the other): the launcher plays the whole push
// -- every entry, then the flags -- before the row loops
template <class Gather>
static inline void emu_halo_push_all(const HaloArgs &h, const Gather &g)
{
    for (int b = 0; b < h.push_ctas; ++b)
        for (int t = 0; t < 256; ++t) {
            blockIdx = EmuDim{(unsigned)((int)gridDim.x - h.push_ctas + b), 0, 0};
            threadIdx = EmuDim{(unsigned)t, 0, 0};
            halo_push_entries(h, g);
        }
    if (h.push_ctas > 0) halo_publish(h);
}
#endif

__device__ __forceinline__ void halo_wait(const HaloArgs &h, const unsigned long long &t_entry)
{
    const HaloTable *T = h.tbl;
    unsigned long long t0 = 0;
    if (h.trace) t0 = global_ns();
    for (int q = 0; q < T->n_from; ++q)
        while (ld_flag_acquire(T->from_flag[q]) < h.tag) __nanosleep(20);     // acquire load: no separate fence
    unsigned long long t1 = 0;
    if (h.trace) t1 = global_ns();
    if (h.trace && (threadIdx.x & 31) == 0) {    // per waiting warp: [2] sum spin, [3] sum fence, [4] warps,
        const unsigned long long t2 = global_ns();   // [5] max (end of wait - entry), [6] sum (start of wait - entry)
        atomicAdd(h.trace + 2, t1 - t0);
        atomicAdd(h.trace + 3, t2 - t1);
        atomicAdd(h.trace + 4, 1ull);
        atomicMax(h.trace + 5, t2 - t_entry);
        atomicAdd(h.trace + 6, t0 - t_entry);
    }
}

template <int ND, class Gather, class Epi, class Fin>
__global__ void __launch_bounds__(256, epi_min_blocks<Epi>::value)
spmv_row_shard_kernel(CsrView A, Gather g, Epi epi, ReduceWs ws, Fin fin, const int *done, HaloArgs h)
{
    if (*done) return;
    g.init();
    epi.init();
    __shared__ unsigned long long t_entry;      // KRY_HALO_TRACE only (shared: no register across the row loops)
    if (h.trace) {
        if (threadIdx.x == 0) t_entry = global_ns();
        __syncthreads();
    }
    double acc[ND > 0 ? ND : 1];
#pragma unroll
    for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
#ifdef KRY_EMULATE
    // tests/emu, SIMT mode: __shared__ is one copy per host thread there, so a publishing block must not
    // run its reduction while warps of a row block sit in theirs; the launcher plays the two kinds of
    // blocks one after the other (launch.cuh)
    if (h.emu_phase == 1 && (int)blockIdx.x < h.row_ctas) return;
    if (h.emu_phase == 2 && (int)blockIdx.x >= h.row_ctas) return;
#endif
    if ((int)blockIdx.x >= (int)gridDim.x - h.push_ctas) {
        if (!(h.skip_push & 1)) halo_push(h, g, t_entry);
        if ((int)blockIdx.x >= h.row_ctas) {              // dedicated publishing CTA: no rows
            if constexpr (ND > 0) block_reduce_finalize<ND>(acc, ws, fin);
            return;
        }
    }
    const int stride = h.row_ctas * blockDim.x;
    auto one_row = [&](int v) {
        int row = v + h.rot;
        if (row >= A.nrows) row -= A.nrows;
        const int s = __ldg(A.rowptr + row), e = __ldg(A.rowptr + row + 1);
        double sum = 0.0;
        for (int k = s; k < e; ++k)
            sum = __dadd_rn(sum, __dmul_rn(__ldg(A.val + k), g(__ldg(A.col + k))));
        epi(row, sum, acc);
    };
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int v_int = h.v_wait < A.nrows ? h.v_wait : A.nrows;
    for (; v < v_int; v += stride) one_row(v);            // interior: no halo column in these rows
    if (v < A.nrows) {
        if (!(h.skip_push & 2)) halo_wait(h, t_entry);    // the peers' boundary entries have landed
        for (; v < A.nrows; v += stride) one_row(v);
    }
    if constexpr (ND > 0) block_reduce_finalize<ND>(acc, ws, fin);
}

// -------------------------------------------- thread per row, batched loads
// Same mapping as spmv_row_kernel, but the loads of a row are issued in three
// waves of independent requests -- all (col,val) pairs of a CHUNK, then all x
// gathers, then the (sequential, order-preserving) sum -- instead of one
// dependent col -> x chain per entry.  More bytes in flight per warp for a
// latency-bound kernel, at the price of registers.
template <int ND, int CHUNK, class Gather, class Epi, class Fin>
__global__ void __launch_bounds__(256)
spmv_rowb_kernel(CsrView A, Gather g, Epi epi, ReduceWs ws, Fin fin, const int *done)
{
    if (*done) return;
    g.init();
    epi.init();
    double acc[ND > 0 ? ND : 1];
#pragma unroll
    for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
    const int stride = gridDim.x * blockDim.x;
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < A.nrows; row += stride) {
        const int s = __ldg(A.rowptr + row), e = __ldg(A.rowptr + row + 1);
        double sum = 0.0;
        for (int k = s; k < e; k += CHUNK) {
            int    c[CHUNK];
            double v[CHUNK], xv[CHUNK];
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) {
                const bool on = k + j < e;
                c[j] = on ? __ldg(A.col + k + j) : 0;
                v[j] = on ? __ldg(A.val + k + j) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) xv[j] = (k + j < e) ? g(c[j]) : 0.0;
#pragma unroll
            for (int j = 0; j < CHUNK; ++j)
                if (k + j < e) sum = __dadd_rn(sum, __dmul_rn(v[j], xv[j]));
        }
        epi(row, sum, acc);
    }
    if constexpr (ND > 0) block_reduce_finalize<ND>(acc, ws, fin);
}

#ifndef KRY_EMULATE      // shared-memory / TMA variants: CUDA only
// ------------------------------------------------------- coalesced nnz stream
// Dynamic smem: (tile_nnz + max_row) doubles of staged products.
template <int ND, class Gather, class Epi, class Fin>
__global__ void
spmv_stream_kernel(CsrView A, Gather g, Epi epi, ReduceWs ws, Fin fin, const int *done)
{
    if (*done) return;
    extern __shared__ double s_prod[];
    g.init();
    epi.init();
    const int tid = threadIdx.x, nt = blockDim.x;
    const int row_lo = __ldg(A.rowblk + blockIdx.x);
    const int row_hi = __ldg(A.rowblk + blockIdx.x + 1);
    const int k_lo = __ldg(A.rowptr + row_lo);
    const int k_hi = __ldg(A.rowptr + row_hi);

    // phase 1: stream val/col (coalesced), gather x, stage products.
    // 4 independent element loads in flight per thread per trip.
    int k = k_lo + tid;
    for (; k + 3 * nt < k_hi; k += 4 * nt) {
        const int    c0 = __ldg(A.col + k), c1 = __ldg(A.col + k + nt);
        const int    c2 = __ldg(A.col + k + 2 * nt), c3 = __ldg(A.col + k + 3 * nt);
        const double v0 = __ldg(A.val + k), v1 = __ldg(A.val + k + nt);
        const double v2 = __ldg(A.val + k + 2 * nt), v3 = __ldg(A.val + k + 3 * nt);
        const double x0 = g(c0), x1 = g(c1), x2 = g(c2), x3 = g(c3);
        s_prod[k - k_lo]          = __dmul_rn(v0, x0);
        s_prod[k - k_lo + nt]     = __dmul_rn(v1, x1);
        s_prod[k - k_lo + 2 * nt] = __dmul_rn(v2, x2);
        s_prod[k - k_lo + 3 * nt] = __dmul_rn(v3, x3);
    }
    for (; k < k_hi; k += nt) s_prod[k - k_lo] = __dmul_rn(__ldg(A.val + k), g(__ldg(A.col + k)));
    __syncthreads();

    // phase 2: one thread per row, sequential sum in storage order.
    double acc[ND > 0 ? ND : 1];
#pragma unroll
    for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
    for (int row = row_lo + tid; row < row_hi; row += nt) {
        const int s = __ldg(A.rowptr + row) - k_lo, e = __ldg(A.rowptr + row + 1) - k_lo;
        double sum = 0.0;
        for (int j = s; j < e; ++j) sum = __dadd_rn(sum, s_prod[j]);
        epi(row, sum, acc);
    }
    if constexpr (ND > 0) block_reduce_finalize<ND>(acc, ws, fin);
}

// --------------------------------------------------------- TMA bulk pipeline
// Persistent CTAs walk the nnz tiles with stride gridDim.x.  One elected thread
// drives the TMA unit: for every tile it issues two 1-D bulk copies
// (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes; SASS
// UBLKCP) for the 16-byte aligned val and col windows of the tile into a
// kStages-deep smem ring; all threads wait on the stage's mbarrier, turn the
// staged (val, col) pairs into products in place, sum rows, run the epilogue,
// and hand the stage back.  HBM latency of the streaming part is hidden behind
// the reduction of the previous tile instead of behind occupancy.
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned phase)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned phase)
{
    while (!mbar_try_wait(bar, phase)) {
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes,
                                         uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// non-blocking probe (mbarrier.try_wait may suspend the thread up to a system time limit)
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, unsigned phase)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bulk prefetch of a global window into L2 by the TMA unit (no destination, no register cost)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

}  // namespace tma

// smem per stage: cap doubles (val, overwritten by products) + cap ints (col),
// cap = tile_nnz + max_row + 4 rounded up to a multiple of 4.
template <int ND, int kStages, class Gather, class Epi, class Fin>
__global__ void
spmv_tma_kernel(CsrView A, int cap, Gather g, Epi epi, ReduceWs ws, Fin fin, const int *done)
{
    if (*done) return;
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) uint64_t s_full[kStages];
    double *s_val = reinterpret_cast<double *>(s_raw);                         // [kStages][cap]
    int    *s_col = reinterpret_cast<int *>(s_raw + (size_t)kStages * cap * 8);  // [kStages][cap]

    g.init();
    epi.init();
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) tma::mbar_init(&s_full[s], 1);
        tma::fence_barrier_init();
    }
    __syncthreads();

    // issue(tile t into stage s): thread 0 only
    auto issue = [&](int tile, int s) {
        const int row_lo = __ldg(A.rowblk + tile), row_hi = __ldg(A.rowblk + tile + 1);
        const int k_lo = __ldg(A.rowptr + row_lo), k_hi = __ldg(A.rowptr + row_hi);
        const int a_lo = k_lo & ~3;                 // 16-byte aligned for both arrays
        const int a_hi = (k_hi + 3) & ~3;           // arrays are padded by >= 4 entries
        const unsigned n = (unsigned)(a_hi - a_lo);
        if (n == 0) {                               // empty tile: complete the phase by hand
            tma::mbar_expect_tx(&s_full[s], 0);
            return;
        }
        tma::mbar_expect_tx(&s_full[s], n * 12u);
        tma::bulk_g2s(s_val + (size_t)s * cap, A.val + a_lo, n * 8u, &s_full[s]);
        tma::bulk_g2s(s_col + (size_t)s * cap, A.col + a_lo, n * 4u, &s_full[s]);
    };

    const int first = blockIdx.x, step = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            const int t = first + s * step;
            if (t < A.nblocks) issue(t, s);
        }
    }

    double acc[ND > 0 ? ND : 1];
#pragma unroll
    for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;

    int it = 0;
    for (int tile = first; tile < A.nblocks; tile += step, ++it) {
        const int      s     = it % kStages;
        const unsigned phase = (unsigned)(it / kStages) & 1u;
        const int row_lo = __ldg(A.rowblk + tile), row_hi = __ldg(A.rowblk + tile + 1);
        const int k_lo = __ldg(A.rowptr + row_lo), k_hi = __ldg(A.rowptr + row_hi);
        const int a_lo = k_lo & ~3;
        double *pv = s_val + (size_t)s * cap;
        int    *pc = s_col + (size_t)s * cap;

        tma::mbar_wait(&s_full[s], phase);

        // products in place (each thread touches only its own slots)
        for (int j = (k_lo - a_lo) + tid; j < k_hi - a_lo; j += nt)
            pv[j] = __dmul_rn(pv[j], g(pc[j]));
        __syncthreads();

        for (int row = row_lo + tid; row < row_hi; row += nt) {
            const int rs = __ldg(A.rowptr + row) - a_lo, re = __ldg(A.rowptr + row + 1) - a_lo;
            double sum = 0.0;
            for (int j = rs; j < re; ++j) sum = __dadd_rn(sum, pv[j]);
            epi(row, sum, acc);
        }
        tma::fence_proxy_async();        // my generic-proxy smem writes before the async-proxy refill
        __syncthreads();                 // stage s fully consumed by every thread

        if (tid == 0) {
            const int nxt = tile + kStages * step;
            if (nxt < A.nblocks) issue(nxt, s);
        }
    }
    if constexpr (ND > 0) block_reduce_finalize<ND>(acc, ws, fin);
}
// ------------------------------------------- thread per row, TMA-staged row pointers
// The row kernel's per-row chain is rowptr -> (col, val) -> gathers, three dependent memory
// latencies (ncu: 0.26 eligible warps per scheduler, long-scoreboard stalls on exactly these
// instructions).  Here the TMA unit runs ahead of the threads, at no register cost:
//   * the row pointers of a CTA's next tiles (256 rows + 1) arrive in shared memory through
//     cp.async.bulk + mbarrier, two trips ahead (ring of kStages tiles): a thread reads its
//     [s, e) from shared memory and its col/val loads start at once;
//   * as soon as the row pointers of the next tile have landed, one thread asks the TMA unit to
//     pull that tile's col / val windows (contiguous: the rows of a tile are consecutive) into
//     L2 (cp.async.bulk.prefetch.L2), so the second link of the chain is an L2 hit.
// Row -> thread mapping and order inside a row are those of spmv_row_kernel: same bits.
template <int ND, class Gather, class Epi, class Fin>
__global__ void __launch_bounds__(256, epi_min_blocks<Epi>::value)
spmv_rowtma_kernel(CsrView A, Gather g, Epi epi, ReduceWs ws, Fin fin, const int *done)
{
    constexpr int kStages = 4, kLag = 2, kTile = 256, kCap = 264;
    if (*done) return;
    __shared__ __align__(16) int s_rp[kStages][kCap];
    __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];
    g.init();
    epi.init();
    const int tid = threadIdx.x;
    const int ntiles = (A.nrows + kTile - 1) / kTile;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            tma::mbar_init(&s_full[s], 1);
            tma::mbar_init(&s_empty[s], 8);              // one arrival per warp
        }
        tma::fence_barrier_init();
    }
    __syncthreads();
    auto rows_of = [&](int tile) {
        const int left = A.nrows - tile * kTile;
        return left < kTile ? left : kTile;
    };
    auto issue = [&](int tile, int s) {                  // thread 0 only
        const unsigned bytes = (unsigned)(((rows_of(tile) + 1 + 3) & ~3) * 4);    // rowptr is padded by 8 entries
        tma::mbar_expect_tx(&s_full[s], bytes);
        tma::bulk_g2s(s_rp[s], A.rowptr + (size_t)tile * kTile, bytes, &s_full[s]);
    };
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            const int t = blockIdx.x + s * gridDim.x;
            if (t < ntiles) issue(t, s);
        }
    }
    double acc[ND > 0 ? ND : 1];
#pragma unroll
    for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % kStages;
        tma::mbar_wait(&s_full[s], (unsigned)(it / kStages) & 1u);
        const int row = tile * kTile + tid;
        int rs = 0, re = 0;
        if (row < A.nrows) {
            rs = s_rp[s][tid];
            re = s_rp[s][tid + 1];
        }
        __syncwarp();
        if ((tid & 31) == 0) tma::mbar_arrive(&s_empty[s]);          // this warp has read its row pointers
        if (tid == 0 && it >= kLag) {
            // refill the stage every warp left kLag trips ago with the tile kStages - kLag trips ahead
            const int ip = it - kLag, nxt = tile + (kStages - kLag) * gridDim.x;
            if (nxt < ntiles) {
                tma::mbar_wait(&s_empty[ip % kStages], (unsigned)(ip / kStages) & 1u);
                issue(nxt, ip % kStages);
            }
        }
        if (tid == 32) {
            // next tile: once its row pointers are in shared memory, pull its col / val windows into L2
            const int nt = tile + gridDim.x;
            if (nt < ntiles && tma::mbar_test_wait(&s_full[(it + 1) % kStages], (unsigned)((it + 1) / kStages) & 1u)) {
                const int *rp = s_rp[(it + 1) % kStages];
                const int k0 = rp[0] & ~3, k1 = (rp[rows_of(nt)] + 3) & ~3;          // 16-byte aligned for both arrays
                if (k1 > k0) {
                    tma::bulk_prefetch_l2(A.val + k0, (unsigned)(k1 - k0) * 8u);
                    tma::bulk_prefetch_l2(A.col + k0, (unsigned)(k1 - k0) * 4u);
                }
            }
        }
        if (row < A.nrows) {
            double sum = 0.0;
            for (int k = rs; k < re; ++k)
                sum = __dadd_rn(sum, __dmul_rn(__ldg(A.val + k), g(__ldg(A.col + k))));
            epi(row, sum, acc);
        }
    }
    if constexpr (ND > 0) block_reduce_finalize<ND>(acc, ws, fin);
}
#endif  // !KRY_EMULATE

#endif  // __CUDACC__ || KRY_EMULATE
