// emu_context.cpp -- host stand-in for csrc/context.cu and csrc/comm.cu (TEST INFRASTRUCTURE ONLY,
// see emu_device.h).  Contexts, vectors and CSR operators live in host memory; the hot-path entry
// points (kry_spmv*, kry_multi_axpy_dot, kry_solver_*) come from the product's own ops.cu and
// solvers.cu compiled next to this file.  Also defines the handful of CUDA runtime calls those
// two files make, as plain host operations.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <new>
#include <utility>
#include <vector>

#include "common.cuh"

thread_local EmuDim threadIdx, blockIdx, blockDim, gridDim;
thread_local double emu_tot[EMU_MAX_DOTS];
thread_local int    emu_reduced;

// ------------------------------------------------------------------ SIMT mode: one fiber per CUDA thread
// Blocks run one after the other; inside a block the threads are ucontext fibers resumed
// round-robin.  __syncthreads() and the warp-level barrier behind the shuffles count arrivals
// (threads that have returned no longer count, as on the hardware) and spin by yielding.
#include <sched.h>
#include <ucontext.h>

int emu_fibers_on = 0;
thread_local unsigned char emu_shfl_slots[32][32][8];

namespace {
constexpr size_t kFiberStack = 128 * 1024;
struct Fiber {
    ucontext_t uc;
    int        tid, blk;
    bool       done;
};
struct Block {
    int      live, arrived;
    unsigned gen;
    int      wlive[32], warrived[32];
    unsigned wgen[32];
};
// one set per host thread: the ranks of an emulated multi-GPU run are threads of this process
thread_local std::vector<Block> g_blks;
thread_local std::vector<Fiber> g_fibers;
thread_local std::vector<char>  g_stacks;
thread_local ucontext_t         g_sched;
thread_local Fiber             *g_cur = nullptr;
thread_local const void        *g_closure = nullptr;
thread_local void (*g_invoke)(const void *) = nullptr;
thread_local bool               g_grid_wait = false;

Block &cur_block() { return g_blks[(size_t)g_cur->blk]; }

void release_if_complete(Block &b)
{
    if (b.live > 0 && b.arrived == b.live) {
        b.arrived = 0;
        b.gen++;
    }
}
void warp_release_if_complete(Block &b, int w)
{
    if (b.wlive[w] > 0 && b.warrived[w] == b.wlive[w]) {
        b.warrived[w] = 0;
        b.wgen[w]++;
    }
}
void fiber_main()
{
    g_invoke(g_closure);
    Fiber *me = g_cur;
    Block &b = g_blks[(size_t)me->blk];
    me->done = true;
    b.live--;
    release_if_complete(b);
    const int w = me->tid >> 5;
    b.wlive[w]--;
    warp_release_if_complete(b, w);
    // returning resumes uc_link = the scheduler
}
}  // namespace

void emu_yield()
{
    if (!g_cur) return;                         // sequential mode: nothing to switch to
    swapcontext(&g_cur->uc, &g_sched);
}

// a thread that spins on memory another block (or another rank) will write: tell the scheduler of
// a cooperative launch that this block cannot advance, so that it runs the other blocks
void emu_spin_wait()
{
    g_grid_wait = true;
    emu_yield();
}

void emu_block_barrier()
{
    if (!g_cur) {
        fprintf(stderr, "emulation: __syncthreads() reached outside the SIMT mode (kry_emu_set_fibers)\n");
        abort();
    }
    Block &b = cur_block();
    const unsigned gen = b.gen;
    b.arrived++;
    release_if_complete(b);
    while (b.gen == gen) emu_yield();
}

void emu_warp_barrier(int w)
{
    if (!g_cur) {
        fprintf(stderr, "emulation: warp-level synchronisation reached outside the SIMT mode\n");
        abort();
    }
    Block &b = cur_block();
    const unsigned gen = b.wgen[w];
    b.warrived[w]++;
    warp_release_if_complete(b, w);
    while (b.wgen[w] == gen) emu_yield();
}

// cooperative = 0: the blocks run one after the other.  cooperative = 1 (cudaLaunchCooperativeKernel):
// all blocks are alive at once; a block runs until all its threads have returned or it waits for
// the other blocks (emu_spin_wait), then the next block runs.  __shared__ variables are static,
// i.e. one copy per host thread: that is safe as long as no __shared__ value is live across a
// grid-wide wait, which holds for the kernels of this library.
void emu_launch_fibers_mode(int grid, int block, const void *closure, void (*invoke)(const void *), int cooperative)
{
    const int wave = cooperative ? grid : 1;
    const size_t nf = (size_t)wave * block;
    if (g_fibers.size() < nf) g_fibers.resize(nf);
    if (g_stacks.size() < nf * kFiberStack) g_stacks.resize(nf * kFiberStack);
    if ((int)g_blks.size() < grid) g_blks.resize((size_t)grid);
    g_closure = closure;
    g_invoke = invoke;
    gridDim = EmuDim{(unsigned)grid, 1, 1};
    blockDim = EmuDim{(unsigned)block, 1, 1};
    for (int b0 = 0; b0 < grid; b0 += wave) {
        for (int bb = 0; bb < wave; ++bb) {
            Block &B = g_blks[(size_t)(b0 + bb)];
            memset(&B, 0, sizeof(B));
            B.live = block;
            for (int t = 0; t < block; ++t) B.wlive[t >> 5]++;
            for (int t = 0; t < block; ++t) {
                Fiber &f = g_fibers[(size_t)bb * block + t];
                f.tid = t;
                f.blk = b0 + bb;
                f.done = false;
                getcontext(&f.uc);
                f.uc.uc_stack.ss_sp = g_stacks.data() + ((size_t)bb * block + t) * kFiberStack;
                f.uc.uc_stack.ss_size = kFiberStack;
                f.uc.uc_link = &g_sched;
                makecontext(&f.uc, fiber_main, 0);
            }
        }
        int blocks_left = wave;
        std::vector<int> left((size_t)wave, block);
        while (blocks_left > 0) {
            for (int bb = 0; bb < wave; ++bb) {
                if (left[(size_t)bb] == 0) continue;
                // run this block until it is finished or waits for the others
                for (;;) {
                    g_grid_wait = false;
                    for (int t = 0; t < block; ++t) {
                        Fiber &f = g_fibers[(size_t)bb * block + t];
                        if (f.done) continue;
                        g_cur = &f;
                        threadIdx = EmuDim{(unsigned)t, 0, 0};
                        blockIdx = EmuDim{(unsigned)(b0 + bb), 0, 0};
                        swapcontext(&g_sched, &f.uc);
                        if (f.done) left[(size_t)bb]--;
                    }
                    if (left[(size_t)bb] == 0) {
                        blocks_left--;
                        break;
                    }
                    if (g_grid_wait && wave > 1) break;          // let the other blocks run
                }
            }
        }
        g_cur = nullptr;
    }
}

void emu_launch_fibers(int grid, int block, const void *closure, void (*invoke)(const void *))
{
    emu_launch_fibers_mode(grid, block, closure, invoke, 0);
}

unsigned char *emu_dynamic_smem = nullptr;
void emu_set_dynamic_smem(size_t bytes)
{
    static std::vector<double> buf;                   // 8-byte aligned
    if (buf.size() * 8 < bytes + 16) buf.resize(bytes / 8 + 4);
    emu_dynamic_smem = reinterpret_cast<unsigned char *>(buf.data());
}

// emulation-only switch (not part of the product ABI): 1 = SIMT fibers + the genuine reduction
extern "C" int kry_emu_set_fibers(int on)
{
    emu_fibers_on = on ? 1 : 0;
    return KRY_OK;
}

// ------------------------------------------------------------------ CUDA runtime, host edition
extern "C" {
cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *b, const void *, int, size_t) { *b = 2; return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int *b, const void *, int, size_t, unsigned) { *b = 2; return cudaSuccess; }
cudaError_t cudaFuncSetAttribute(const void *, cudaFuncAttribute, int) { return cudaSuccess; }
// graphs are never used under emulation (use_graphs = 0); the symbols only have to link
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *g) { *g = nullptr; return cudaErrorNotSupported; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *, cudaGraph_t, unsigned long long) { return cudaErrorNotSupported; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
}

// ------------------------------------------------------------------ errors / misc
static thread_local char g_err[512] = "";

void kry_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *kry_last_error(void) { return g_err; }
extern "C" int kry_abi_version(void) { return KRY_ABI_VERSION; }
extern "C" int kry_device_count(int *count) { *count = 1; return KRY_OK; }

int kry_alloc(void **p, size_t bytes)
{
    KRY_CUDA(cudaMalloc(p, bytes ? bytes : 256));
    return KRY_OK;
}

int kry_ctx_ensure_partials(kry_ctx *, int) { return KRY_OK; }

ReduceWs kry_ws(kry_ctx *c)
{
    ReduceWs ws;
    memset(&ws, 0, sizeof(ws));
    ws.partials = c->partials;
    ws.sums = c->sums;
    ws.counter = c->counter;
    ws.stride = c->partial_stride;
    ws.nranks = c->nranks;
    ws.rank = c->rank;
    ws.inbox = c->p2p_inbox;
    ws.peers = c->p2p_peers_dev;
    ws.seq = c->p2p_seq;
    return ws;
}

void kry_ctx_retain(kry_ctx *c) { c->refs++; }
void kry_ctx_release(kry_ctx *c)
{
    if (--c->refs <= 0 && c->closed) delete c;
}

extern "C" int kry_comm_destroy(kry_ctx *ctx);

// ------------------------------------------------------------------ context
extern "C" int kry_ctx_create(int device, kry_ctx **out)
{
    KRY_REQUIRE(out && device == 0, KRY_ERR_INVALID, "kry_ctx_create (emulation): device 0 only");
    kry_ctx *c = new kry_ctx();
    memset(c, 0, sizeof(*c));
    c->sm_count = 2;                       // small grids: the host plays every thread
    c->l2_bytes = 126 << 20;
    c->smem_optin = 227 * 1024;
    c->nranks = 1;
    c->l2_hints = 1;
    c->use_graphs = 0;
    c->cg_fuse = 2;
    c->cg_fuse_shards = 1;
    c->halo_p2p = 1;
#ifdef KRY_OPT_MINRES_FUSE
    c->minres_fuse = 1;
#endif
#ifdef KRY_OPT_MINRES_PERSISTENT
    c->minres_persistent = 0;
#endif
    KRY_TRY(kry_alloc((void **)&c->scalars, KRY_NUM_SLOTS * sizeof(double)));
    KRY_TRY(kry_alloc((void **)&c->sums, 2 * KRY_MAX_DOTS * sizeof(double)));
    KRY_TRY(kry_alloc((void **)&c->counter, 256));
    KRY_TRY(kry_alloc((void **)&c->never_done, 256));
    c->partial_stride = 1024;
    KRY_TRY(kry_alloc((void **)&c->partials, (size_t)KRY_MAX_DOTS * c->partial_stride * sizeof(double)));
    *out = c;
    return KRY_OK;
}

extern "C" int kry_ctx_destroy(kry_ctx *c)
{
    if (!c || c->closed) return KRY_OK;
    if (c->nccl) kry_comm_destroy(c);
    free(c->scalars);
    free(c->sums);
    free(c->counter);
    free(c->never_done);
    free(c->partials);
    c->partials = nullptr;
    c->scalars = c->sums = nullptr;
    c->counter = nullptr;
    c->never_done = nullptr;
    c->closed = 1;
    if (c->refs <= 0) delete c;
    return KRY_OK;
}

extern "C" int kry_ctx_sync(kry_ctx *) { return KRY_OK; }
extern "C" int kry_ctx_props(kry_ctx *c, int64_t p[6])
{
    p[0] = c->sm_count; p[1] = p[2] = (int64_t)1 << 34; p[3] = 100; p[4] = c->l2_bytes; p[5] = c->smem_optin;
    return KRY_OK;
}
static thread_local std::chrono::steady_clock::time_point g_t0;
extern "C" int kry_timer_start(kry_ctx *) { g_t0 = std::chrono::steady_clock::now(); return KRY_OK; }
extern "C" int kry_timer_stop(kry_ctx *, double *ms)
{
    *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_t0).count();
    return KRY_OK;
}
extern "C" int kry_flush_l2(kry_ctx *) { return KRY_OK; }
extern "C" int kry_launch_count(kry_ctx *c, int64_t *n) { *n = c->launches; return KRY_OK; }
extern "C" int kry_halo_trace_read(kry_ctx *, uint64_t *out16) { memset(out16, 0, 16 * sizeof(uint64_t)); return KRY_OK; }
extern "C" int kry_prof_enable(kry_ctx *, int) { return KRY_OK; }
extern "C" int kry_prof_read(kry_ctx *, int64_t *n, double *ms) { *n = 0; *ms = 0.0; return KRY_OK; }
extern "C" int kry_host_alloc(int64_t bytes, void **out) { *out = malloc(bytes > 0 ? bytes : 1); return *out ? KRY_OK : KRY_ERR_NOMEM; }
extern "C" int kry_host_free(void *p) { free(p); return KRY_OK; }

extern "C" int kry_ctx_set_option(kry_ctx *c, int option, int value)
{
    switch (option) {
        case KRY_OPT_L2_HINTS: c->l2_hints = value; return KRY_OK;
        case KRY_OPT_GRAPHS: return KRY_OK;                   // no graphs on the host
        case KRY_OPT_CG_FUSE:
            KRY_REQUIRE(value >= 0 && value <= 2, KRY_ERR_INVALID, "CG_FUSE=%d not in 0..2", value);
            c->cg_fuse = value;
            return KRY_OK;
        case KRY_OPT_CG_FUSE_SHARDS: c->cg_fuse_shards = value ? 1 : 0; return KRY_OK;
        case KRY_OPT_HALO_P2P: c->halo_p2p = value ? 1 : 0; return KRY_OK;
        case KRY_OPT_P2P:
            KRY_REQUIRE(!value || c->p2p_inbox, KRY_ERR_STATE, "peer-memory all-reduce was not set up");
            c->p2p_on = value ? 1 : 0;
            return KRY_OK;
#ifdef KRY_OPT_MINRES_FUSE
        case KRY_OPT_MINRES_FUSE: c->minres_fuse = value ? 1 : 0; return KRY_OK;
#endif
#ifdef KRY_OPT_MINRES_PERSISTENT
        case KRY_OPT_MINRES_PERSISTENT: c->minres_persistent = (value && emu_fibers_on) ? 1 : 0; return KRY_OK;   // SIMT mode only
#endif
#ifdef KRY_OPT_CG_ONE_CTA
        case KRY_OPT_CG_ONE_CTA: c->cg_one_cta = (value && emu_fibers_on) ? 1 : 0; return KRY_OK;   // SIMT mode only
#endif
        default: kry_set_error("kry_ctx_set_option (emulation): option %d", option); return KRY_ERR_INVALID;
    }
}

extern "C" int kry_ctx_get_option(kry_ctx *c, int option, int *value)
{
    switch (option) {
        case KRY_OPT_L2_HINTS: *value = c->l2_hints; return KRY_OK;
        case KRY_OPT_GRAPHS: *value = 0; return KRY_OK;
        case KRY_OPT_P2P: *value = c->p2p_on; return KRY_OK;
        case KRY_OPT_CG_FUSE: *value = c->cg_fuse; return KRY_OK;
        case KRY_OPT_CG_FUSE_SHARDS: *value = c->cg_fuse_shards; return KRY_OK;
        case KRY_OPT_HALO_P2P: *value = c->halo_p2p; return KRY_OK;
#ifdef KRY_OPT_MINRES_FUSE
        case KRY_OPT_MINRES_FUSE: *value = c->minres_fuse; return KRY_OK;
#endif
#ifdef KRY_OPT_MINRES_PERSISTENT
        case KRY_OPT_MINRES_PERSISTENT: *value = c->minres_persistent; return KRY_OK;
#endif
#ifdef KRY_OPT_CG_ONE_CTA
        case KRY_OPT_CG_ONE_CTA: *value = c->cg_one_cta; return KRY_OK;
#endif
        default: kry_set_error("kry_ctx_get_option (emulation): option %d", option); return KRY_ERR_INVALID;
    }
}

// ------------------------------------------------------------------ vectors
extern "C" int kry_vec_create_cap(kry_ctx *c, int64_t n, int64_t cap, kry_vec **out)
{
    KRY_REQUIRE(c && out && n >= 0 && cap >= n, KRY_ERR_INVALID, "kry_vec_create: bad argument");
    kry_vec *v = new kry_vec();
    v->ctx = c; v->n = n; v->cap = cap; v->owned = true;
    KRY_TRY(kry_alloc((void **)&v->d, (size_t)(cap + 4) * sizeof(double)));
    kry_ctx_retain(c);
    *out = v;
    return KRY_OK;
}
extern "C" int kry_vec_create(kry_ctx *c, int64_t n, kry_vec **out) { return kry_vec_create_cap(c, n, n, out); }
extern "C" int kry_vec_destroy(kry_vec *v)
{
    if (!v) return KRY_OK;
    free(v->d);
    kry_ctx_release(v->ctx);
    delete v;
    return KRY_OK;
}
extern "C" int kry_vec_size(const kry_vec *v, int64_t *n) { *n = v->n; return KRY_OK; }
extern "C" int kry_vec_upload(kry_vec *v, const double *h, int64_t n)
{
    KRY_REQUIRE(v && h, KRY_ERR_INVALID, "kry_vec_upload: NULL argument");
    KRY_CTX_LIVE(v->ctx, "kry_vec_upload");
    KRY_REQUIRE(n == v->n, KRY_ERR_SHAPE, "kry_vec_upload: host has %lld entries, vector %lld", (long long)n, (long long)v->n);
    memcpy(v->d, h, (size_t)n * sizeof(double));
    return KRY_OK;
}
extern "C" int kry_vec_download(const kry_vec *v, double *h, int64_t n)
{
    KRY_REQUIRE(v && h, KRY_ERR_INVALID, "kry_vec_download: NULL argument");
    KRY_CTX_LIVE(v->ctx, "kry_vec_download");
    KRY_REQUIRE(n == v->n, KRY_ERR_SHAPE, "kry_vec_download: host has %lld entries, vector %lld", (long long)n, (long long)v->n);
    memcpy(h, v->d, (size_t)n * sizeof(double));
    return KRY_OK;
}
extern "C" int kry_vec_read(const kry_vec *v, int64_t off, int64_t cnt, double *h)
{
    KRY_REQUIRE(v && h && off >= 0 && cnt >= 0 && off + cnt <= v->n, KRY_ERR_SHAPE, "kry_vec_read: bad range");
    memcpy(h, v->d + off, (size_t)cnt * sizeof(double));
    return KRY_OK;
}
extern "C" int kry_vec_fill(kry_vec *v, double value)
{
    KRY_REQUIRE(v, KRY_ERR_INVALID, "kry_vec_fill: NULL vector");
    for (int64_t i = 0; i < v->n; ++i) v->d[i] = value;
    return KRY_OK;
}
extern "C" int kry_vec_copy(kry_vec *dst, const kry_vec *src)
{
    KRY_REQUIRE(dst && src, KRY_ERR_INVALID, "kry_vec_copy: NULL vector");
    KRY_REQUIRE(dst->n == src->n, KRY_ERR_SHAPE, "kry_vec_copy: sizes %lld != %lld", (long long)dst->n, (long long)src->n);
    memcpy(dst->d, src->d, (size_t)src->n * sizeof(double));
    return KRY_OK;
}
extern "C" int kry_scalars_read(kry_ctx *c, int first, int count, double *h)
{
    KRY_REQUIRE(c && h && first >= 0 && count >= 0 && first + count <= KRY_NUM_SLOTS, KRY_ERR_INVALID, "kry_scalars_read: range");
    memcpy(h, c->scalars + first, (size_t)count * sizeof(double));
    return KRY_OK;
}
extern "C" int kry_scalars_write(kry_ctx *c, int first, int count, const double *h)
{
    KRY_REQUIRE(c && h && first >= 0 && count >= 0 && first + count <= KRY_NUM_SLOTS, KRY_ERR_INVALID, "kry_scalars_write: range");
    memcpy(c->scalars + first, h, (size_t)count * sizeof(double));
    return KRY_OK;
}

// ------------------------------------------------------------------ operators
static void csr_free(CsrDev &m)
{
    free(m.rowptr); free(m.col); free(m.val); free(m.rowblk);
    m = CsrDev();
}

static int csr_from_host(CsrDev &m, int64_t nrows, int64_t ncols, const std::vector<int> &rp,
                         const std::vector<int> &col, const std::vector<double> &val)
{
    const int64_t nnz = (int64_t)col.size();
    m.nrows = nrows; m.ncols = ncols; m.nnz = nnz;
    KRY_TRY(kry_alloc((void **)&m.rowptr, (size_t)(nrows + 1 + 8) * sizeof(int)));
    KRY_TRY(kry_alloc((void **)&m.col, (size_t)(nnz + 8) * sizeof(int)));
    KRY_TRY(kry_alloc((void **)&m.val, (size_t)(nnz + 8) * sizeof(double)));
    memcpy(m.rowptr, rp.data(), (size_t)(nrows + 1) * sizeof(int));
    if (nnz) {
        memcpy(m.col, col.data(), (size_t)nnz * sizeof(int));
        memcpy(m.val, val.data(), (size_t)nnz * sizeof(double));
    }
    int mx = 0;
    for (int64_t i = 0; i < nrows; ++i) mx = std::max(mx, rp[i + 1] - rp[i]);
    m.max_row = mx;
    return KRY_OK;
}

// stable counting sort by column: inside a row of A^T the entries keep ascending original-row
// order, which is what the device build (stable radix sort of (col, k)) produces
static int csr_transpose(const CsrDev &A, CsrDev &T)
{
    std::vector<int> rp((size_t)A.ncols + 1, 0), col((size_t)A.nnz);
    std::vector<double> val((size_t)A.nnz);
    for (int64_t k = 0; k < A.nnz; ++k) rp[(size_t)A.col[k] + 1]++;
    for (int64_t c = 0; c < A.ncols; ++c) rp[(size_t)c + 1] += rp[(size_t)c];
    std::vector<int> next(rp.begin(), rp.end() - 1);
    for (int64_t r = 0; r < A.nrows; ++r)
        for (int k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k) {
            const int dst = next[(size_t)A.col[k]]++;
            col[(size_t)dst] = (int)r;
            val[(size_t)dst] = A.val[k];
        }
    return csr_from_host(T, A.ncols, A.nrows, rp, col, val);
}

int csr_build_partition(kry_ctx *, CsrDev &, int) { return KRY_OK; }    // nnz-tiled kernels: CUDA only

static int new_csr(kry_ctx *c, uint32_t flags, int64_t nrows, int64_t ncols, const std::vector<int> &rp,
                   const std::vector<int> &col, const std::vector<double> &val, kry_csr **out)
{
    kry_csr *M = new kry_csr();
    M->ctx = c;
    M->flags = flags;
    KRY_TRY(csr_from_host(M->A, nrows, ncols, rp, col, val));
    if ((flags & KRY_CSR_BUILD_TRANSPOSE) && !(flags & KRY_CSR_SYMMETRIC)) {
        KRY_TRY(csr_transpose(M->A, M->T));
        M->has_T = true;
    }
    kry_ctx_retain(c);
    *out = M;
    return KRY_OK;
}

extern "C" int kry_csr_create(kry_ctx *c, int64_t nrows, int64_t ncols, int64_t nnz, const int32_t *rowptr,
                              const int32_t *col, const double *val, uint32_t flags, kry_csr **out)
{
    KRY_REQUIRE(c && out && rowptr && (nnz == 0 || (col && val)), KRY_ERR_INVALID, "kry_csr_create: NULL argument");
    KRY_REQUIRE(nrows >= 0 && ncols >= 0 && nnz >= 0, KRY_ERR_INVALID, "csr: negative size");
    KRY_REQUIRE(rowptr[0] == 0 && rowptr[nrows] == nnz, KRY_ERR_INVALID, "kry_csr_create: rowptr[0]=%d, rowptr[nrows]=%d but nnz=%lld",
                rowptr[0], rowptr[nrows], (long long)nnz);
    std::vector<int> rp(rowptr, rowptr + nrows + 1), cc(col, col + nnz);
    std::vector<double> vv(val, val + nnz);
    return new_csr(c, flags, nrows, ncols, rp, cc, vv, out);
}

// host stand-ins of csrc/assemble.cu (same orders: arrival order inside a row; A, then B, then diagonal)
static void coo_rows(int64_t nrows, int64_t nnz, const int32_t *rows, const int32_t *cols, const double *vals, int sym,
                     std::vector<int> &rp, std::vector<int> &cc, std::vector<double> &vv)
{
    std::vector<std::vector<std::pair<int, double>>> R((size_t)nrows);
    for (int64_t k = 0; k < nnz; ++k) {
        R[(size_t)rows[k]].push_back({cols[k], vals[k]});
        if (sym && rows[k] != cols[k]) R[(size_t)cols[k]].push_back({rows[k], vals[k]});
    }
    rp.assign((size_t)nrows + 1, 0);
    cc.clear();
    vv.clear();
    for (int64_t i = 0; i < nrows; ++i) {
        for (auto &e : R[(size_t)i]) { cc.push_back(e.first); vv.push_back(e.second); }
        rp[(size_t)i + 1] = (int)cc.size();
    }
}

extern "C" int kry_csr_create_coo(kry_ctx *c, int64_t nrows, int64_t ncols, int64_t nnz, const int32_t *rows,
                                  const int32_t *cols, const double *vals, uint32_t flags, kry_csr **out)
{
    KRY_REQUIRE(c && out && (nnz == 0 || (rows && cols && vals)), KRY_ERR_INVALID, "kry_csr_create_coo: NULL argument");
    const int sym = (flags & KRY_CSR_SYMMETRIC) ? 1 : 0;
    int bad = 0;
    for (int64_t k = 0; k < nnz; ++k)
        bad += !(rows[k] >= 0 && rows[k] < nrows && cols[k] >= 0 && cols[k] < ncols &&
                 (!sym || (rows[k] < ncols && cols[k] < nrows)));
    KRY_REQUIRE(bad == 0, KRY_ERR_INVALID, "kry_csr_create_coo: %d coordinate(s) outside a (%lld, %lld) operator", bad,
                (long long)nrows, (long long)ncols);
    std::vector<int> rp, cc;
    std::vector<double> vv;
    coo_rows(nrows, nnz, rows, cols, vals, sym, rp, cc, vv);
    KRY_TRY(new_csr(c, flags & KRY_CSR_SYMMETRIC, nrows, ncols, rp, cc, vv, out));
    if ((flags & KRY_CSR_BUILD_TRANSPOSE) && !sym) {
        coo_rows(ncols, nnz, cols, rows, vals, 0, rp, cc, vv);
        kry_csr *T = nullptr;
        KRY_TRY(new_csr(c, 0, ncols, nrows, rp, cc, vv, &T));
        (*out)->T = T->A;
        (*out)->has_T = true;
        T->A = CsrDev();
        kry_csr_destroy(T);
    }
    return KRY_OK;
}

extern "C" int kry_csr_combine(kry_ctx *c, const kry_csr *A, double alpha, const kry_csr *B, double beta,
                               const double *diag, double gamma, uint32_t flags, kry_csr **out)
{
    KRY_REQUIRE(c && A && out, KRY_ERR_INVALID, "kry_csr_combine: NULL argument");
    KRY_REQUIRE(!B || (B->A.nrows == A->A.nrows && B->A.ncols == A->A.ncols), KRY_ERR_SHAPE, "kry_csr_combine: shapes differ");
    KRY_REQUIRE(!diag || A->A.nrows == A->A.ncols, KRY_ERR_SHAPE, "kry_csr_combine: diagonal term on a rectangular operator");
    std::vector<int> rp((size_t)A->A.nrows + 1, 0), cc;
    std::vector<double> vv;
    for (int64_t i = 0; i < A->A.nrows; ++i) {
        for (int k = A->A.rowptr[i]; k < A->A.rowptr[i + 1]; ++k) { cc.push_back(A->A.col[k]); vv.push_back(alpha * A->A.val[k]); }
        if (B)
            for (int k = B->A.rowptr[i]; k < B->A.rowptr[i + 1]; ++k) { cc.push_back(B->A.col[k]); vv.push_back(beta * B->A.val[k]); }
        if (diag) { cc.push_back((int)i); vv.push_back(gamma * diag[i]); }
        rp[(size_t)i + 1] = (int)cc.size();
    }
    return new_csr(c, flags & KRY_CSR_SYMMETRIC, A->A.nrows, A->A.ncols, rp, cc, vv, out);
}

extern "C" int kry_csr_to_dense(const kry_csr *A, double *dense)
{
    KRY_REQUIRE(A && dense, KRY_ERR_INVALID, "kry_csr_to_dense: NULL argument");
    memset(dense, 0, (size_t)(A->A.nrows * A->A.ncols) * sizeof(double));
    for (int64_t i = 0; i < A->A.nrows; ++i)
        for (int k = A->A.rowptr[i]; k < A->A.rowptr[i + 1]; ++k) dense[i * A->A.ncols + A->A.col[k]] += A->A.val[k];
    return KRY_OK;
}

extern "C" int kry_csr_destroy(kry_csr *M)
{
    if (!M) return KRY_OK;
    csr_free(M->A);
    csr_free(M->T);
    kry_ctx_release(M->ctx);
    delete M;
    return KRY_OK;
}

extern "C" int kry_csr_shape(const kry_csr *M, int64_t *nrows, int64_t *ncols, int64_t *nnz)
{
    KRY_REQUIRE(M, KRY_ERR_INVALID, "kry_csr_shape: NULL operator");
    if (nrows) *nrows = M->A.nrows;
    if (ncols) *ncols = M->A.ncols;
    if (nnz) *nnz = M->A.nnz;
    return KRY_OK;
}

extern "C" int kry_csr_build_transpose(kry_csr *M)
{
    KRY_REQUIRE(M, KRY_ERR_INVALID, "kry_csr_build_transpose: NULL operator");
    if (M->has_T || (M->flags & KRY_CSR_SYMMETRIC)) return KRY_OK;
    KRY_TRY(csr_transpose(M->A, M->T));
    M->has_T = true;
    return KRY_OK;
}

extern "C" int kry_csr_download(const kry_csr *M, int transposed, int32_t *rowptr, int32_t *col, double *val)
{
    KRY_REQUIRE(M, KRY_ERR_INVALID, "kry_csr_download: NULL operator");
    const CsrDev *m = &M->A;
    if (transposed && !(M->flags & KRY_CSR_SYMMETRIC)) {
        KRY_REQUIRE(M->has_T, KRY_ERR_STATE, "kry_csr_download: transpose not built");
        m = &M->T;
    }
    memcpy(rowptr, m->rowptr, (size_t)(m->nrows + 1) * sizeof(int));
    memcpy(col, m->col, (size_t)m->nnz * sizeof(int));
    memcpy(val, m->val, (size_t)m->nnz * sizeof(double));
    return KRY_OK;
}

extern "C" int kry_csr_diagonal(const kry_csr *M, double *diag)
{
    KRY_REQUIRE(M && diag, KRY_ERR_INVALID, "kry_csr_diagonal: NULL argument");
    for (int64_t r = 0; r < M->A.nrows; ++r) {
        double d = 0.0;
        for (int k = M->A.rowptr[r]; k < M->A.rowptr[r + 1]; ++k)
            if (M->A.col[k] == r) d += M->A.val[k];
        diag[r] = d;
    }
    return KRY_OK;
}

extern "C" int kry_csr_set_kernel(kry_csr *M, int kind, int tile_nnz, int threads)
{
    KRY_REQUIRE(M, KRY_ERR_INVALID, "kry_csr_set_kernel: NULL operator");
    M->kind = kind; M->tile_nnz = tile_nnz; M->threads = threads;
    return KRY_OK;
}

// gallery: the stencils of csrc/context.cu restated as host loops (sorted columns)
template <class Emit>
static int stencil_rows(kry_ctx *c, int64_t n, int64_t row_begin, int64_t row_end, uint32_t flags, Emit emit,
                        kry_csr **out)
{
    KRY_REQUIRE(c && out, KRY_ERR_INVALID, "stencil: NULL argument");
    if (row_end < 0) row_end = n;
    KRY_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= n, KRY_ERR_INVALID, "stencil: bad row range");
    std::vector<int> rp(1, 0), col;
    std::vector<double> val;
    for (int64_t r = row_begin; r < row_end; ++r) {
        emit(r, col, val);
        rp.push_back((int)col.size());
    }
    KRY_TRY(new_csr(c, flags, row_end - row_begin, n, rp, col, val, out));
    (*out)->halo.n_global = n;
    (*out)->halo.row_begin = row_begin;
    return KRY_OK;
}

extern "C" int kry_csr_create_poisson1d(kry_ctx *c, int64_t n, int64_t rb, int64_t re, uint32_t flags, kry_csr **out)
{
    return stencil_rows(c, n, rb, re, flags, [n](int64_t r, std::vector<int> &col, std::vector<double> &val) {
        if (r > 0) { col.push_back((int)(r - 1)); val.push_back(-1.0); }
        col.push_back((int)r); val.push_back(2.0);
        if (r < n - 1) { col.push_back((int)(r + 1)); val.push_back(-1.0); }
    }, out);
}

extern "C" int kry_csr_create_poisson2d(kry_ctx *c, int64_t g, int64_t rb, int64_t re, uint32_t flags, kry_csr **out)
{
    return stencil_rows(c, g * g, rb, re, flags | KRY_CSR_SYMMETRIC, [g](int64_t r, std::vector<int> &col, std::vector<double> &val) {
        const int64_t i = r / g, j = r % g;
        if (i > 0) { col.push_back((int)(r - g)); val.push_back(-1.0); }
        if (j > 0) { col.push_back((int)(r - 1)); val.push_back(-1.0); }
        col.push_back((int)r); val.push_back(4.0);
        if (j < g - 1) { col.push_back((int)(r + 1)); val.push_back(-1.0); }
        if (i < g - 1) { col.push_back((int)(r + g)); val.push_back(-1.0); }
    }, out);
}

extern "C" int kry_csr_create_convdiff3d(kry_ctx *c, int64_t m, double gamma, int64_t rb, int64_t re, uint32_t flags,
                                         kry_csr **out)
{
    const int64_t m2 = m * m;
    return stencil_rows(c, m * m2, rb, re, flags, [=](int64_t r, std::vector<int> &col, std::vector<double> &val) {
        const int64_t i = r / m2, j = (r / m) % m, k = r % m;
        if (i > 0) { col.push_back((int)(r - m2)); val.push_back(-1.0 - gamma); }
        if (j > 0) { col.push_back((int)(r - m)); val.push_back(-1.0 - gamma); }
        if (k > 0) { col.push_back((int)(r - 1)); val.push_back(-1.0 - gamma); }
        col.push_back((int)r); val.push_back(6.0 + 3.0 * gamma);
        if (k < m - 1) { col.push_back((int)(r + 1)); val.push_back(-1.0); }
        if (j < m - 1) { col.push_back((int)(r + m)); val.push_back(-1.0); }
        if (i < m - 1) { col.push_back((int)(r + m2)); val.push_back(-1.0); }
    }, out);
}

// ------------------------------------------------------------------ multi-GPU: ranks are threads
// csrc/comm.cu is compiled as it is (shard finalisation, halo exchange, peer-memory inboxes); what
// it gets from NCCL and CUDA IPC comes from here.  Every rank of an emulated run is a host thread
// with its own context; the collectives are barriers between those threads, "peer memory" is
// simply the other thread's allocation.
#include <nccl.h>

#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>

namespace {
struct EmuComm {
    int                       nranks = 0, joined = 0, arrived = 0;
    unsigned                  gen = 0;
    std::mutex                m;
    std::condition_variable   cv;
    std::vector<const void *> send;
    void barrier()
    {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = gen;
        if (++arrived == nranks) {
            arrived = 0;
            gen++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g; });
        }
    }
};
struct EmuRank {
    std::shared_ptr<EmuComm> comm;
    int                      rank;
};
std::mutex                                      g_comms_m;
std::map<std::string, std::shared_ptr<EmuComm>> g_comms;
unsigned long long                              g_next_id = 1;

size_t dtype_size(ncclDataType_t t) { return t == ncclDouble ? 8 : 1; }

ncclResult_t emu_get_unique_id(ncclUniqueId *id)
{
    std::lock_guard<std::mutex> lk(g_comms_m);
    memset(id, 0, sizeof(*id));
    snprintf(id->internal, sizeof(id->internal), "emu-comm-%llu", g_next_id++);
    return ncclSuccess;
}

ncclResult_t emu_comm_init_rank(ncclComm_t *out, int nranks, ncclUniqueId id, int rank)
{
    std::shared_ptr<EmuComm> c;
    {
        std::lock_guard<std::mutex> lk(g_comms_m);
        auto &slot = g_comms[std::string(id.internal, sizeof(id.internal))];
        if (!slot) {
            slot = std::make_shared<EmuComm>();
            slot->nranks = nranks;
            slot->send.assign((size_t)nranks, nullptr);
        }
        c = slot;
    }
    if (c->nranks != nranks) return ncclInvalidArgument;
    *out = reinterpret_cast<ncclComm_t>(new EmuRank{c, rank});
    c->barrier();                                   // every rank has joined
    return ncclSuccess;
}

ncclResult_t emu_comm_destroy(ncclComm_t comm)
{
    delete reinterpret_cast<EmuRank *>(comm);
    return ncclSuccess;
}

ncclResult_t emu_all_gather(const void *send, void *recv, size_t count, ncclDataType_t t, ncclComm_t comm, cudaStream_t)
{
    EmuRank *r = reinterpret_cast<EmuRank *>(comm);
    EmuComm &c = *r->comm;
    const size_t bytes = count * dtype_size(t);
    c.send[(size_t)r->rank] = send;
    c.barrier();
    for (int q = 0; q < c.nranks; ++q) memcpy((char *)recv + (size_t)q * bytes, c.send[(size_t)q], bytes);
    c.barrier();
    return ncclSuccess;
}

ncclResult_t emu_all_reduce(const void *send, void *recv, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t comm,
                            cudaStream_t)
{
    if (t != ncclDouble) return ncclInvalidArgument;
    EmuRank *r = reinterpret_cast<EmuRank *>(comm);
    EmuComm &c = *r->comm;
    c.send[(size_t)r->rank] = send;
    c.barrier();
    std::vector<double> acc(count);
    for (size_t i = 0; i < count; ++i) {            // rank order on every rank: identical bits everywhere
        double a = ((const double *)c.send[0])[i];
        for (int q = 1; q < c.nranks; ++q) {
            const double v = ((const double *)c.send[(size_t)q])[i];
            a = (op == ncclMax) ? (v > a ? v : a) : a + v;
        }
        acc[i] = a;
    }
    c.barrier();                                    // everyone has read the inputs (recv may alias send)
    memcpy(recv, acc.data(), count * sizeof(double));
    return ncclSuccess;
}

const char *emu_get_error_string(ncclResult_t) { return "emulated NCCL error"; }
}  // namespace

extern "C" void emu_nccl_table(void **get_unique_id, void **comm_init_rank, void **comm_destroy, void **all_reduce,
                               void **all_gather, void **get_error_string)
{
    *get_unique_id = (void *)emu_get_unique_id;
    *comm_init_rank = (void *)emu_comm_init_rank;
    *comm_destroy = (void *)emu_comm_destroy;
    *all_reduce = (void *)emu_all_reduce;
    *all_gather = (void *)emu_all_gather;
    *get_error_string = (void *)emu_get_error_string;
}

// CUDA IPC inside one process: the handle carries the pointer
extern "C" {
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p)
{
    memset(h, 0, sizeof(*h));
    memcpy(h, &p, sizeof(p));
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, &h, sizeof(*p)); return cudaSuccess; }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
}

// Fast mode stand-in for the in-kernel all-reduce over the peers' inboxes (common.cuh,
// block_reduce_finalize, ws.p2p): the same protocol -- totals then sequence number into every
// peer's inbox, wait for every peer's, sum in rank order -- spoken by the launching host thread.
void emu_p2p_allreduce(const ReduceWs &ws, double *tot, int nd)
{
    const unsigned long long seq = *ws.seq + 1ull;
    *ws.seq = seq;
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    const size_t slot = (size_t)(seq & 1ull) * ws.nranks;
    for (int q = 0; q < ws.nranks; ++q) {
        volatile unsigned long long *dst = reinterpret_cast<volatile unsigned long long *>(ws.peers[q] + (slot + ws.rank) * 8);
        for (int d = 0; d < nd; ++d) {
            unsigned long long bits;
            memcpy(&bits, &tot[d], 8);
            __atomic_store_n(const_cast<unsigned long long *>(dst + 2 * d), tag | (bits & 0xffffffffull), __ATOMIC_RELEASE);
            __atomic_store_n(const_cast<unsigned long long *>(dst + 2 * d + 1), tag | (bits >> 32), __ATOMIC_RELEASE);
        }
    }
    double in[KRY_MAX_RANKS][EMU_MAX_DOTS];
    for (int q = 0; q < ws.nranks; ++q) {
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(ws.inbox + (slot + q) * 8);
        for (int d = 0; d < nd; ++d) {
            unsigned long long lo, hi;
            while (((lo = __atomic_load_n(src + 2 * d, __ATOMIC_ACQUIRE)) & 0xffffffff00000000ull) != tag) sched_yield();
            while (((hi = __atomic_load_n(src + 2 * d + 1, __ATOMIC_ACQUIRE)) & 0xffffffff00000000ull) != tag) sched_yield();
            const unsigned long long bits = (hi << 32) | (lo & 0xffffffffull);
            memcpy(&in[q][d], &bits, 8);
        }
    }
    for (int d = 0; d < nd; ++d) {
        double a = 0.0;
        for (int q = 0; q < ws.nranks; ++q) a = a + in[q][d];
        tot[d] = a;
    }
}
