// context.cu -- contexts, device vectors, CSR operators (upload, device-side
// gallery generation, tiling partition, transpose) for libkrylov_b200.
//
// Nothing here is on the per-iteration path; it is the data layout that the
// hot kernels (spmv.cuh, solvers_*.cu) stream from HBM:
//   * val fp64[nnz+8], col int32[nnz+8], rowptr int32[nrows+1]  (256-B aligned,
//     padded so that 16-byte aligned TMA windows never leave the allocation)
//   * rowblk int32[nblocks+1]: rows are grouped into tiles of ~tile_nnz stored
//     entries; tile b owns rows [rowblk[b], rowblk[b+1]) -- whole rows only, so
//     each row is summed by one thread in storage order (bit-exact contract).
#include <cub/cub.cuh>
#include <stdarg.h>
#include <string.h>

#include <new>

#include "common.cuh"

// ------------------------------------------------------------------ errors
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

extern "C" int kry_device_count(int *count)
{
    KRY_REQUIRE(count, KRY_ERR_INVALID, "kry_device_count: NULL output");
    *count = 0;
    KRY_CUDA(cudaGetDeviceCount(count));
    return KRY_OK;
}

int kry_alloc(void **p, size_t bytes)
{
    *p = nullptr;
    if (bytes == 0) bytes = 256;
    KRY_CUDA(cudaMalloc(p, bytes));
    return KRY_OK;
}

// ----------------------------------------------------------------- context
extern "C" int kry_ctx_create(int device, kry_ctx **out)
{
    KRY_REQUIRE(out, KRY_ERR_INVALID, "kry_ctx_create: NULL output");
    *out = nullptr;
    int ndev = 0;
    KRY_CUDA(cudaGetDeviceCount(&ndev));
    KRY_REQUIRE(device >= 0 && device < ndev, KRY_ERR_INVALID,
                "kry_ctx_create: device %d out of range (%d visible)", device, ndev);
    KRY_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    KRY_CUDA(cudaGetDeviceProperties(&prop, device));
    kry_ctx *c = new (std::nothrow) kry_ctx();
    KRY_REQUIRE(c, KRY_ERR_NOMEM, "kry_ctx_create: host allocation failed");
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->l2_bytes = prop.l2CacheSize;
    c->smem_optin = (int64_t)prop.sharedMemPerBlockOptin;
    c->nranks = 1;
    c->l2_hints = 1;
    c->use_graphs = 1;
    c->cg_one_cta = 1;
    c->minres_fuse = 1;
    c->minres_persistent = 0;  // measured on B200, config 3: 51.5 us/iteration against 50.9 for the 2-launch plan and 54.4
                               // for 3 launches (profiles/r2f_minres_ab.json): kept as an option, not the default
    c->cg_fuse = 2;        // measured on B200, 10^7-row 5-pt Laplacian: 0.222 ms/iteration against 0.233
                           // (form 1) and 0.249 (form 0) -- profiles/r1b_ab_cgfuse*.json, r1_final_bench_n1.json
    c->cg_fuse_shards = 1; // row shards use the same plan: 2 x B200, 10^8 rows: 825.7 vs 803.8 it/s with the same
                           // residual bits (profiles/r1e_*); 2-8 emulated ranks: tests/test_emulated_multi_rank.py
    c->halo_p2p = 1;       // row shards: boundary entries travel through peer memory from inside the SpMV launch
                           // (takes effect where the CUDA IPC mappings of kry_comm_init / kry_halo_link succeeded)
    KRY_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    KRY_CUDA(cudaEventCreate(&c->ev0));
    KRY_CUDA(cudaEventCreate(&c->ev1));
    KRY_TRY(kry_alloc((void **)&c->scalars, KRY_NUM_SLOTS * sizeof(double)));
    KRY_CUDA(cudaMemsetAsync(c->scalars, 0, KRY_NUM_SLOTS * sizeof(double), c->stream));
    KRY_TRY(kry_alloc((void **)&c->sums, 2 * KRY_MAX_DOTS * sizeof(double)));
    KRY_TRY(kry_alloc((void **)&c->counter, 256));
    KRY_CUDA(cudaMemsetAsync(c->counter, 0, 256, c->stream));
    KRY_TRY(kry_alloc((void **)&c->never_done, 256));
    KRY_CUDA(cudaMemsetAsync(c->never_done, 0, 256, c->stream));
    KRY_TRY(kry_ctx_ensure_partials(c, 16384));
    if (getenv("KRY_HALO_TRACE")) {
        KRY_TRY(kry_alloc((void **)&c->halo_trace, 16 * sizeof(uint64_t)));
        KRY_CUDA(cudaMemset(c->halo_trace, 0, 16 * sizeof(uint64_t)));
    }
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    *out = c;
    return KRY_OK;
}

int kry_ctx_ensure_partials(kry_ctx *c, int nblocks)
{
    if (nblocks <= c->partial_stride) return KRY_OK;
    int stride = 16384;
    while (stride < nblocks) stride *= 2;
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    if (c->partials) cudaFree(c->partials);
    c->partials = nullptr;
    KRY_TRY(kry_alloc((void **)&c->partials, (size_t)KRY_MAX_DOTS * stride * sizeof(double)));
    c->partial_stride = stride;
    c->partials_gen++;      // captured graphs have the old pointer / stride baked in: solvers re-capture
    return KRY_OK;
}

ReduceWs kry_ws(kry_ctx *c)
{
    ReduceWs ws;
    ws.partials = c->partials;
    ws.sums = c->sums;
    ws.counter = c->counter;
    ws.stride = c->partial_stride;
    ws.defer = 0;
    ws.p2p = 0;
    ws.nranks = c->nranks;
    ws.rank = c->rank;
    ws.inbox = c->p2p_inbox;
    ws.peers = c->p2p_peers_dev;
    ws.seq = c->p2p_seq;
    return ws;
}

extern "C" int kry_comm_destroy(kry_ctx *ctx);

void kry_ctx_retain(kry_ctx *c) { c->refs++; }

void kry_ctx_release(kry_ctx *c)
{
    if (--c->refs <= 0 && c->closed) delete c;
}

extern "C" int kry_ctx_destroy(kry_ctx *c)
{
    if (!c || c->closed) return KRY_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->nccl) kry_comm_destroy(c);
    cudaFree(c->scalars);
    cudaFree(c->partials);
    cudaFree(c->sums);
    cudaFree(c->counter);
    cudaFree(c->never_done);
    if (c->flush_buf) cudaFree(c->flush_buf);
    for (int i = 0; i < 2 * c->prof_cap; ++i) cudaEventDestroy(c->prof_ev[i]);
    delete[] c->prof_ev;
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    c->copy_stream = nullptr;
    c->stream = nullptr;
    c->scalars = c->partials = c->sums = nullptr;
    c->counter = nullptr;
    c->never_done = nullptr;
    c->flush_buf = nullptr;
    c->prof_ev = nullptr;
    c->prof_cap = c->prof_n = 0;
    c->closed = 1;
    if (c->refs <= 0) delete c;
    return KRY_OK;
}

extern "C" int kry_ctx_sync(kry_ctx *c)
{
    KRY_REQUIRE(c, KRY_ERR_INVALID, "kry_ctx_sync: NULL context");
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    KRY_CUDA(cudaGetLastError());
    return KRY_OK;
}

extern "C" int kry_ctx_props(kry_ctx *c, int64_t props[6])
{
    KRY_REQUIRE(c && props, KRY_ERR_INVALID, "kry_ctx_props: NULL argument");
    size_t fr = 0, tot = 0;
    KRY_CUDA(cudaSetDevice(c->device));
    KRY_CUDA(cudaMemGetInfo(&fr, &tot));
    cudaDeviceProp prop;
    KRY_CUDA(cudaGetDeviceProperties(&prop, c->device));
    props[0] = c->sm_count;
    props[1] = (int64_t)tot;
    props[2] = (int64_t)fr;
    props[3] = prop.major * 10 + prop.minor;
    props[4] = c->l2_bytes;
    props[5] = c->smem_optin;
    return KRY_OK;
}

extern "C" int kry_timer_start(kry_ctx *c)
{
    KRY_REQUIRE(c, KRY_ERR_INVALID, "kry_timer_start: NULL context");
    KRY_CUDA(cudaEventRecord(c->ev0, c->stream));
    return KRY_OK;
}

extern "C" int kry_timer_stop(kry_ctx *c, double *ms)
{
    KRY_REQUIRE(c && ms, KRY_ERR_INVALID, "kry_timer_stop: NULL argument");
    KRY_CUDA(cudaEventRecord(c->ev1, c->stream));
    KRY_CUDA(cudaEventSynchronize(c->ev1));
    float f = 0.f;
    KRY_CUDA(cudaEventElapsedTime(&f, c->ev0, c->ev1));
    *ms = (double)f;
    return KRY_OK;
}

__global__ void flush_kernel(double *buf, size_t n, double v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = v;
}

extern "C" int kry_flush_l2(kry_ctx *c)
{
    KRY_REQUIRE(c, KRY_ERR_INVALID, "kry_flush_l2: NULL context");
    if (!c->flush_buf) {
        c->flush_bytes = (size_t)(c->l2_bytes > 0 ? c->l2_bytes : (128 << 20)) * 2;
        KRY_TRY(kry_alloc(&c->flush_buf, c->flush_bytes));
    }
    flush_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>((double *)c->flush_buf,
                                                         c->flush_bytes / 8, 1.0);
    KRY_CUDA(cudaGetLastError());
    return KRY_OK;
}

// Diagnostics of the fused halo exchange (env KRY_HALO_TRACE=1 at context creation): sums of
// in-kernel %globaltimer intervals since the last read -- see spmv.cuh halo_push / halo_wait.
extern "C" int kry_halo_trace_read(kry_ctx *c, uint64_t *out16)
{
    KRY_REQUIRE(c && out16, KRY_ERR_INVALID, "kry_halo_trace_read: NULL argument");
    memset(out16, 0, 16 * sizeof(uint64_t));
    if (!c->halo_trace) return KRY_OK;
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    KRY_CUDA(cudaMemcpy(out16, c->halo_trace, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    KRY_CUDA(cudaMemset(c->halo_trace, 0, 16 * sizeof(uint64_t)));
    return KRY_OK;
}

extern "C" int kry_launch_count(kry_ctx *c, int64_t *count)
{
    KRY_REQUIRE(c && count, KRY_ERR_INVALID, "kry_launch_count: NULL argument");
    *count = c->launches;
    return KRY_OK;
}

extern "C" int kry_ctx_set_option(kry_ctx *c, int option, int value)
{
    KRY_REQUIRE(c, KRY_ERR_INVALID, "kry_ctx_set_option: NULL context");
    KRY_REQUIRE(option == KRY_OPT_L2_HINTS || option == KRY_OPT_GRAPHS || option == KRY_OPT_P2P ||
                    option == KRY_OPT_CG_FUSE || option == KRY_OPT_CG_FUSE_SHARDS ||
                    option == KRY_OPT_CG_ONE_CTA || option == KRY_OPT_MINRES_FUSE ||
                    option == KRY_OPT_MINRES_PERSISTENT || option == KRY_OPT_HALO_P2P,
                KRY_ERR_INVALID, "kry_ctx_set_option: unknown option %d", option);
    if (option == KRY_OPT_L2_HINTS) c->l2_hints = value;
    else if (option == KRY_OPT_CG_FUSE_SHARDS) c->cg_fuse_shards = value ? 1 : 0;
    else if (option == KRY_OPT_CG_ONE_CTA) c->cg_one_cta = value ? 1 : 0;
    else if (option == KRY_OPT_MINRES_FUSE) c->minres_fuse = value ? 1 : 0;
    else if (option == KRY_OPT_MINRES_PERSISTENT) c->minres_persistent = value ? 1 : 0;
    else if (option == KRY_OPT_HALO_P2P) c->halo_p2p = value ? 1 : 0;
    else if (option == KRY_OPT_CG_FUSE) {
        KRY_REQUIRE(value >= 0 && value <= 2, KRY_ERR_INVALID, "kry_ctx_set_option: CG_FUSE=%d not in 0..2", value);
        c->cg_fuse = value;
    }
    else if (option == KRY_OPT_GRAPHS) c->use_graphs = value ? 1 : 0;
    else {
        KRY_REQUIRE(!value || c->p2p_inbox, KRY_ERR_STATE,
                    "kry_ctx_set_option: peer-memory all-reduce was not set up (kry_comm_init)");
        c->p2p_on = value ? 1 : 0;
    }
    return KRY_OK;
}

extern "C" int kry_ctx_get_option(kry_ctx *c, int option, int *value)
{
    KRY_REQUIRE(c && value, KRY_ERR_INVALID, "kry_ctx_get_option: NULL argument");
    switch (option) {
        case KRY_OPT_L2_HINTS: *value = c->l2_hints; break;
        case KRY_OPT_GRAPHS: *value = c->use_graphs; break;
        case KRY_OPT_P2P: *value = c->p2p_on; break;
        case KRY_OPT_CG_FUSE: *value = c->cg_fuse; break;
        case KRY_OPT_CG_FUSE_SHARDS: *value = c->cg_fuse_shards; break;
        case KRY_OPT_CG_ONE_CTA: *value = c->cg_one_cta; break;
        case KRY_OPT_MINRES_FUSE: *value = c->minres_fuse; break;
        case KRY_OPT_MINRES_PERSISTENT: *value = c->minres_persistent; break;
        case KRY_OPT_HALO_P2P: *value = c->halo_p2p; break;
        default: kry_set_error("kry_ctx_get_option: unknown option %d", option); return KRY_ERR_INVALID;
    }
    return KRY_OK;
}

extern "C" int kry_prof_enable(kry_ctx *c, int max_samples)
{
    KRY_REQUIRE(c && max_samples >= 0 && max_samples <= (1 << 20), KRY_ERR_INVALID,
                "kry_prof_enable: bad argument");
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 2 * c->prof_cap; ++i) cudaEventDestroy(c->prof_ev[i]);
    delete[] c->prof_ev;
    c->prof_ev = nullptr;
    c->prof_cap = c->prof_n = 0;
    if (max_samples == 0) return KRY_OK;
    c->prof_ev = new (std::nothrow) cudaEvent_t[2 * (size_t)max_samples];
    KRY_REQUIRE(c->prof_ev, KRY_ERR_NOMEM, "kry_prof_enable: host allocation failed");
    for (int i = 0; i < 2 * max_samples; ++i) KRY_CUDA(cudaEventCreate(&c->prof_ev[i]));
    c->prof_cap = max_samples;
    return KRY_OK;
}

extern "C" int kry_prof_read(kry_ctx *c, int64_t *samples, double *total_ms)
{
    KRY_REQUIRE(c && samples && total_ms, KRY_ERR_INVALID, "kry_prof_read: NULL argument");
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    double tot = 0.0;
    for (int i = 0; i < c->prof_n; ++i) {
        float ms = 0.f;
        KRY_CUDA(cudaEventElapsedTime(&ms, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]));
        tot += ms;
    }
    *samples = c->prof_n;
    *total_ms = tot;
    c->prof_n = 0;
    return KRY_OK;
}

extern "C" int kry_host_alloc(int64_t bytes, void **out)
{
    KRY_REQUIRE(out && bytes >= 0, KRY_ERR_INVALID, "kry_host_alloc: bad argument");
    KRY_CUDA(cudaMallocHost(out, (size_t)(bytes > 0 ? bytes : 1)));
    return KRY_OK;
}

extern "C" int kry_host_free(void *p)
{
    if (p) KRY_CUDA(cudaFreeHost(p));
    return KRY_OK;
}

// ----------------------------------------------------------------- vectors
extern "C" int kry_vec_create(kry_ctx *c, int64_t n, kry_vec **out)
{
    return kry_vec_create_cap(c, n, n, out);
}

extern "C" int kry_vec_create_cap(kry_ctx *c, int64_t n, int64_t cap, kry_vec **out)
{
    KRY_REQUIRE(c && out && n >= 0 && cap >= n, KRY_ERR_INVALID, "kry_vec_create: bad argument");
    *out = nullptr;
    KRY_CUDA(cudaSetDevice(c->device));
    kry_vec *v = new (std::nothrow) kry_vec();
    KRY_REQUIRE(v, KRY_ERR_NOMEM, "kry_vec_create: host allocation failed");
    v->ctx = c;
    v->n = n;
    v->cap = cap;
    v->owned = true;
    int rc = kry_alloc((void **)&v->d, (size_t)(cap + 4) * sizeof(double));
    if (rc != KRY_OK) {
        delete v;
        return rc;
    }
    kry_ctx_retain(c);
    *out = v;
    return KRY_OK;
}

extern "C" int kry_vec_destroy(kry_vec *v)
{
    if (!v) return KRY_OK;
    if (!v->ctx->closed) cudaStreamSynchronize(v->ctx->stream);
    if (v->owned) cudaFree(v->d);
    kry_ctx_release(v->ctx);
    delete v;
    return KRY_OK;
}

extern "C" int kry_vec_size(const kry_vec *v, int64_t *n)
{
    KRY_REQUIRE(v && n, KRY_ERR_INVALID, "kry_vec_size: NULL argument");
    *n = v->n;
    return KRY_OK;
}

extern "C" int kry_vec_upload(kry_vec *v, const double *host, int64_t n)
{
    KRY_REQUIRE(v && host, KRY_ERR_INVALID, "kry_vec_upload: NULL argument");
    KRY_CTX_LIVE(v->ctx, "kry_vec_upload");
    KRY_REQUIRE(n == v->n, KRY_ERR_SHAPE, "kry_vec_upload: host has %lld entries, vector %lld",
                (long long)n, (long long)v->n);
    KRY_CUDA(cudaMemcpyAsync(v->d, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice,
                             v->ctx->stream));
    KRY_CUDA(cudaStreamSynchronize(v->ctx->stream));   // host buffer is never retained
    return KRY_OK;
}

extern "C" int kry_vec_download(const kry_vec *v, double *host, int64_t n)
{
    KRY_REQUIRE(v && host, KRY_ERR_INVALID, "kry_vec_download: NULL argument");
    KRY_CTX_LIVE(v->ctx, "kry_vec_download");
    KRY_REQUIRE(n == v->n, KRY_ERR_SHAPE, "kry_vec_download: host has %lld entries, vector %lld",
                (long long)n, (long long)v->n);
    KRY_CUDA(cudaMemcpyAsync(host, v->d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost,
                             v->ctx->stream));
    KRY_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return KRY_OK;
}

extern "C" int kry_vec_read(const kry_vec *v, int64_t offset, int64_t count, double *host)
{
    KRY_REQUIRE(v && host, KRY_ERR_INVALID, "kry_vec_read: NULL argument");
    KRY_CTX_LIVE(v->ctx, "kry_vec_read");
    KRY_REQUIRE(offset >= 0 && count >= 0 && offset + count <= v->n, KRY_ERR_SHAPE,
                "kry_vec_read: [%lld,%lld) outside a vector of %lld entries", (long long)offset,
                (long long)(offset + count), (long long)v->n);
    KRY_CUDA(cudaMemcpyAsync(host, v->d + offset, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost,
                             v->ctx->stream));
    KRY_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return KRY_OK;
}

__global__ void fill_kernel(double *d, int64_t n, double v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) d[i] = v;
}

extern "C" int kry_vec_fill(kry_vec *v, double value)
{
    KRY_REQUIRE(v, KRY_ERR_INVALID, "kry_vec_fill: NULL vector");
    KRY_CTX_LIVE(v->ctx, "kry_vec_fill");
    if (v->n == 0) return KRY_OK;
    fill_kernel<<<v->ctx->sm_count * 4, 256, 0, v->ctx->stream>>>(v->d, v->n, value);
    v->ctx->launches++;
    KRY_CUDA(cudaGetLastError());
    return KRY_OK;
}

extern "C" int kry_vec_copy(kry_vec *dst, const kry_vec *src)
{
    KRY_REQUIRE(dst && src, KRY_ERR_INVALID, "kry_vec_copy: NULL vector");
    KRY_CTX_LIVE(dst->ctx, "kry_vec_copy");
    KRY_REQUIRE(dst->n == src->n, KRY_ERR_SHAPE, "kry_vec_copy: sizes %lld != %lld",
                (long long)dst->n, (long long)src->n);
    KRY_CUDA(cudaMemcpyAsync(dst->d, src->d, (size_t)src->n * sizeof(double),
                             cudaMemcpyDeviceToDevice, dst->ctx->stream));
    return KRY_OK;
}

extern "C" int kry_scalars_read(kry_ctx *c, int first, int count, double *host)
{
    KRY_REQUIRE(c && host, KRY_ERR_INVALID, "kry_scalars_read: NULL argument");
    KRY_REQUIRE(first >= 0 && count >= 0 && first + count <= KRY_NUM_SLOTS, KRY_ERR_INVALID,
                "kry_scalars_read: slots [%d,%d) out of range", first, first + count);
    KRY_CUDA(cudaMemcpyAsync(host, c->scalars + first, count * sizeof(double),
                             cudaMemcpyDeviceToHost, c->stream));
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    return KRY_OK;
}

extern "C" int kry_scalars_write(kry_ctx *c, int first, int count, const double *host)
{
    KRY_REQUIRE(c && host, KRY_ERR_INVALID, "kry_scalars_write: NULL argument");
    KRY_REQUIRE(first >= 0 && count >= 0 && first + count <= KRY_NUM_SLOTS, KRY_ERR_INVALID,
                "kry_scalars_write: slots [%d,%d) out of range", first, first + count);
    KRY_CUDA(cudaMemcpyAsync(c->scalars + first, host, count * sizeof(double),
                             cudaMemcpyHostToDevice, c->stream));
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    return KRY_OK;
}

// --------------------------------------------------------------------- CSR
void csr_dev_free(CsrDev &m)
{
    cudaFree(m.rowptr);
    cudaFree(m.col);
    cudaFree(m.val);
    cudaFree(m.rowblk);
    m = CsrDev();
}

int csr_dev_alloc(CsrDev &m, int64_t nrows, int64_t ncols, int64_t nnz)
{
    m.nrows = nrows;
    m.ncols = ncols;
    m.nnz = nnz;
    KRY_TRY(kry_alloc((void **)&m.rowptr, (size_t)(nrows + 1 + 8) * sizeof(int)));
    KRY_TRY(kry_alloc((void **)&m.col, (size_t)(nnz + 8) * sizeof(int)));
    KRY_TRY(kry_alloc((void **)&m.val, (size_t)(nnz + 8) * sizeof(double)));
    return KRY_OK;
}

__global__ void max_row_kernel(const int *rowptr, int nrows, int *out)
{
    int m = 0;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += stride)
        m = max(m, rowptr[i + 1] - rowptr[i]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// rowblk[b] = first row r with rowptr[r] >= b*tile  (b < nblocks);  rowblk[nblocks] = nrows
__global__ void rowblk_kernel(const int *rowptr, int nrows, int tile, int nblocks, int *rowblk)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nblocks) return;
    if (b == nblocks) {
        rowblk[b] = nrows;
        return;
    }
    const long long target = (long long)b * tile;
    int lo = 0, hi = nrows;               // search in rowptr[0..nrows]
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if ((long long)rowptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    rowblk[b] = lo;
}

// (Re)build the tile partition of one CSR for tile_nnz entries per tile.
int csr_build_partition(kry_ctx *c, CsrDev &m, int tile_nnz)
{
    if (m.rowblk && m.tile_nnz == tile_nnz) return KRY_OK;
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    if (m.rowblk) cudaFree(m.rowblk);
    m.rowblk = nullptr;
    int nblocks = (int)((m.nnz + tile_nnz - 1) / tile_nnz);
    if (nblocks < 1) nblocks = 1;
    KRY_TRY(kry_alloc((void **)&m.rowblk, (size_t)(nblocks + 2) * sizeof(int)));
    rowblk_kernel<<<(nblocks + 1 + 255) / 256, 256, 0, c->stream>>>(m.rowptr, (int)m.nrows,
                                                                    tile_nnz, nblocks, m.rowblk);
    c->launches++;
    KRY_CUDA(cudaGetLastError());
    m.nblocks = nblocks;
    m.tile_nnz = tile_nnz;
    return KRY_OK;
}

// One pass over an uploaded CSR: row pointers must not decrease and every column index
// must address the input vector -- a malformed matrix would otherwise read out of bounds in
// every SpMV.  bad[0] counts offending rows, bad[1] offending entries.
__global__ void csr_validate_kernel(const int *rowptr, const int *col, int nrows, int ncols, int nnz, int *bad)
{
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    int br = 0, bc = 0;
    for (int i = t0; i < nrows; i += stride) br += (rowptr[i] > rowptr[i + 1]) || rowptr[i] < 0;
    for (int k = t0; k < nnz; k += stride) bc += (col[k] < 0 || col[k] >= ncols);
    if (br) atomicAdd(bad, br);
    if (bc) atomicAdd(bad + 1, bc);
}

int csr_validate(kry_ctx *c, const CsrDev &m)
{
    int *d_bad = (int *)c->counter + 12;    // scratch inside the 256-byte counter block
    KRY_CUDA(cudaMemsetAsync(d_bad, 0, 2 * sizeof(int), c->stream));
    csr_validate_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(m.rowptr, m.col, (int)m.nrows, (int)m.ncols,
                                                              (int)m.nnz, d_bad);
    c->launches++;
    KRY_CUDA(cudaGetLastError());
    int h[2] = {0, 0};
    KRY_CUDA(cudaMemcpyAsync(h, d_bad, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    KRY_REQUIRE(h[0] == 0, KRY_ERR_INVALID, "kry_csr_create: rowptr decreases at %d row(s)", h[0]);
    KRY_REQUIRE(h[1] == 0, KRY_ERR_INVALID, "kry_csr_create: %d column index(es) outside [0, %lld)", h[1],
                (long long)m.ncols);
    return KRY_OK;
}

int csr_finish(kry_ctx *c, CsrDev &m)
{
    int *d_max = (int *)c->counter + 8;     // scratch inside the 256-byte counter block
    KRY_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), c->stream));
    if (m.nrows > 0) {
        max_row_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(m.rowptr, (int)m.nrows, d_max);
        c->launches++;
        KRY_CUDA(cudaGetLastError());
    }
    int h = 0;
    KRY_CUDA(cudaMemcpyAsync(&h, d_max, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    m.max_row = h;
    return KRY_OK;
}

int csr_build_transpose_dev(kry_ctx *c, const CsrDev &A, CsrDev &T);

int check_sizes(int64_t nrows, int64_t ncols, int64_t nnz)
{
    KRY_REQUIRE(nrows >= 0 && ncols >= 0 && nnz >= 0, KRY_ERR_INVALID,
                "csr: negative size (%lld x %lld, nnz %lld)", (long long)nrows,
                (long long)ncols, (long long)nnz);
    const int64_t lim = (int64_t)INT32_MAX - (1 << 20);
    KRY_REQUIRE(nrows < lim && ncols < lim && nnz < lim, KRY_ERR_UNSUPPORTED,
                "csr: int32 index space exceeded (%lld x %lld, nnz %lld)", (long long)nrows,
                (long long)ncols, (long long)nnz);
    return KRY_OK;
}

extern "C" int kry_csr_create(kry_ctx *c, int64_t nrows, int64_t ncols, int64_t nnz,
                              const int32_t *rowptr, const int32_t *col, const double *val,
                              uint32_t flags, kry_csr **out)
{
    KRY_REQUIRE(c && out && rowptr && (nnz == 0 || (col && val)), KRY_ERR_INVALID,
                "kry_csr_create: NULL argument");
    *out = nullptr;
    KRY_TRY(check_sizes(nrows, ncols, nnz));
    KRY_REQUIRE(rowptr[0] == 0 && rowptr[nrows] == nnz, KRY_ERR_INVALID,
                "kry_csr_create: rowptr[0]=%d, rowptr[nrows]=%d but nnz=%lld", rowptr[0],
                rowptr[nrows], (long long)nnz);
    KRY_CUDA(cudaSetDevice(c->device));
    kry_csr *M = new (std::nothrow) kry_csr();
    KRY_REQUIRE(M, KRY_ERR_NOMEM, "kry_csr_create: host allocation failed");
    M->ctx = c;
    M->flags = flags;
    int rc = csr_dev_alloc(M->A, nrows, ncols, nnz);
    if (rc == KRY_OK) {
        cudaError_t e = cudaMemcpyAsync(M->A.rowptr, rowptr, (size_t)(nrows + 1) * sizeof(int),
                                        cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess && nnz > 0)
            e = cudaMemcpyAsync(M->A.col, col, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice,
                                c->stream);
        if (e == cudaSuccess && nnz > 0)
            e = cudaMemcpyAsync(M->A.val, val, (size_t)nnz * sizeof(double),
                                cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(M->A.col + nnz, 0, 8 * sizeof(int), c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(M->A.val + nnz, 0, 8 * sizeof(double), c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            kry_set_error("kry_csr_create: upload failed: %s", cudaGetErrorString(e));
            rc = KRY_ERR_CUDA;
        }
    }
    if (rc == KRY_OK) rc = csr_validate(c, M->A);
    if (rc == KRY_OK) rc = csr_finish(c, M->A);
    if (rc == KRY_OK && (flags & KRY_CSR_BUILD_TRANSPOSE) && !(flags & KRY_CSR_SYMMETRIC)) {
        rc = csr_build_transpose_dev(c, M->A, M->T);
        M->has_T = (rc == KRY_OK);
    }
    if (rc != KRY_OK) {
        csr_dev_free(M->A);
        csr_dev_free(M->T);
        delete M;
        return rc;
    }
    kry_ctx_retain(c);
    *out = M;
    return KRY_OK;
}

extern "C" int kry_csr_destroy(kry_csr *M)
{
    if (!M) return KRY_OK;
    if (!M->ctx->closed) cudaStreamSynchronize(M->ctx->stream);
    csr_dev_free(M->A);
    csr_dev_free(M->T);
    cudaFree(M->halo.send_idx);
    cudaFree(M->halo.send_buf);
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
    KRY_CTX_LIVE(M->ctx, "kry_csr_build_transpose");
    if (M->has_T || (M->flags & KRY_CSR_SYMMETRIC)) return KRY_OK;
    KRY_CUDA(cudaSetDevice(M->ctx->device));
    KRY_TRY(csr_build_transpose_dev(M->ctx, M->A, M->T));
    M->has_T = true;
    return KRY_OK;
}

extern "C" int kry_csr_download(const kry_csr *M, int transposed, int32_t *rowptr, int32_t *col,
                                double *val)
{
    KRY_REQUIRE(M, KRY_ERR_INVALID, "kry_csr_download: NULL operator");
    KRY_CTX_LIVE(M->ctx, "kry_csr_download");
    const CsrDev *m = &M->A;
    if (transposed && !(M->flags & KRY_CSR_SYMMETRIC)) {
        KRY_REQUIRE(M->has_T, KRY_ERR_STATE, "kry_csr_download: transpose was not built");
        m = &M->T;
    }
    cudaStream_t st = M->ctx->stream;
    if (rowptr)
        KRY_CUDA(cudaMemcpyAsync(rowptr, m->rowptr, (size_t)(m->nrows + 1) * sizeof(int),
                                 cudaMemcpyDeviceToHost, st));
    if (col && m->nnz)
        KRY_CUDA(cudaMemcpyAsync(col, m->col, (size_t)m->nnz * sizeof(int),
                                 cudaMemcpyDeviceToHost, st));
    if (val && m->nnz)
        KRY_CUDA(cudaMemcpyAsync(val, m->val, (size_t)m->nnz * sizeof(double),
                                 cudaMemcpyDeviceToHost, st));
    KRY_CUDA(cudaStreamSynchronize(st));
    return KRY_OK;
}

__global__ void diag_kernel(const int *rowptr, const int *col, const double *val, int nrows,
                            int col_offset, double *diag)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += stride) {
        double d = 0.0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
            if (col[k] == i + col_offset) d += val[k];
        diag[i] = d;
    }
}

extern "C" int kry_csr_diagonal(const kry_csr *M, double *diag_host)
{
    KRY_REQUIRE(M && diag_host, KRY_ERR_INVALID, "kry_csr_diagonal: NULL argument");
    KRY_CTX_LIVE(M->ctx, "kry_csr_diagonal");
    kry_ctx *c = M->ctx;
    double *d = nullptr;
    KRY_TRY(kry_alloc((void **)&d, (size_t)(M->A.nrows + 1) * sizeof(double)));
    // a finalized shard addresses its own rows as local columns [0, nrows)
    const int off = M->halo.active ? 0 : (int)M->halo.row_begin;
    diag_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(M->A.rowptr, M->A.col, M->A.val,
                                                        (int)M->A.nrows, off, d);
    c->launches++;
    cudaError_t e = cudaMemcpyAsync(diag_host, d, (size_t)M->A.nrows * sizeof(double),
                                    cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    KRY_CUDA(e);
    return KRY_OK;
}

extern "C" int kry_csr_set_kernel(kry_csr *M, int kind, int tile_nnz, int threads)
{
    KRY_REQUIRE(M, KRY_ERR_INVALID, "kry_csr_set_kernel: NULL operator");
    KRY_REQUIRE(kind >= KRY_SPMV_AUTO && kind <= KRY_SPMV_ROWPF2, KRY_ERR_INVALID,
                "kry_csr_set_kernel: unknown kernel %d", kind);
    KRY_REQUIRE(tile_nnz == 0 || (tile_nnz >= 256 && tile_nnz <= 16384 && tile_nnz % 4 == 0),
                KRY_ERR_INVALID, "kry_csr_set_kernel: tile_nnz %d not in [256,16384] / 4", tile_nnz);
    KRY_REQUIRE(threads == 0 || (threads >= 64 && threads <= 1024 && threads % 32 == 0),
                KRY_ERR_INVALID, "kry_csr_set_kernel: threads %d", threads);
    M->kind = kind;
    M->tile_nnz = tile_nnz;
    M->threads = threads;
    return KRY_OK;
}

// ---------------------------------------------------------- transpose (device)
// Stable counting sort by column: radix-sort (col, k) pairs -- stable, so inside
// a row of A^T the entries keep ascending original-row order, which is the
// accumulation order of scipy's csc_matvec for (A.T) @ x  (oracle/csr_ref.c).
__global__ void expand_rows_kernel(const int *rowptr, int nrows, int *rowidx)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += stride)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) rowidx[k] = i;
}

__global__ void iota_kernel(int *a, int n)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = i;
}

__global__ void count_cols_kernel(const int *col, int nnz, int *counts)
{
    const int stride = gridDim.x * blockDim.x;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride)
        atomicAdd(counts + col[k], 1);
}

__global__ void permute_kernel(const int *perm, const int *rowidx, const double *val, int nnz,
                               int *tcol, double *tval)
{
    const int stride = gridDim.x * blockDim.x;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += stride) {
        const int k = perm[j];
        tcol[j] = rowidx[k];
        tval[j] = val[k];
    }
}

int csr_build_transpose_dev(kry_ctx *c, const CsrDev &A, CsrDev &T)
{
    cudaStream_t st = c->stream;
    const int nnz = (int)A.nnz, nrows = (int)A.nrows, ncols = (int)A.ncols;
    KRY_TRY(csr_dev_alloc(T, A.ncols, A.nrows, A.nnz));
    int *rowidx = nullptr, *keys_out = nullptr, *perm_in = nullptr, *perm_out = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0, scan_bytes = 0;
    int rc = KRY_OK;
    const int g = c->sm_count * 8;
    do {
        if ((rc = kry_alloc((void **)&rowidx, (size_t)(nnz + 1) * sizeof(int)))) break;
        if ((rc = kry_alloc((void **)&keys_out, (size_t)(nnz + 1) * sizeof(int)))) break;
        if ((rc = kry_alloc((void **)&perm_in, (size_t)(nnz + 1) * sizeof(int)))) break;
        if ((rc = kry_alloc((void **)&perm_out, (size_t)(nnz + 1) * sizeof(int)))) break;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, A.col, keys_out, perm_in, perm_out,
                                        nnz, 0, 32, st);
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, T.rowptr, T.rowptr, ncols + 1, st);
        if (scan_bytes > tmp_bytes) tmp_bytes = scan_bytes;
        if ((rc = kry_alloc(&tmp, tmp_bytes + 256))) break;
        cudaMemsetAsync(T.rowptr, 0, (size_t)(ncols + 1) * sizeof(int), st);
        if (nnz > 0) {
            expand_rows_kernel<<<g, 256, 0, st>>>(A.rowptr, nrows, rowidx);
            iota_kernel<<<g, 256, 0, st>>>(perm_in, nnz);
            count_cols_kernel<<<g, 256, 0, st>>>(A.col, nnz, T.rowptr);
            c->launches += 3;
            cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, A.col, keys_out, perm_in, perm_out,
                                            nnz, 0, 32, st);
            permute_kernel<<<g, 256, 0, st>>>(perm_out, rowidx, A.val, nnz, T.col, T.val);
            c->launches++;
        }
        cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, T.rowptr, T.rowptr, ncols + 1, st);
        cudaMemsetAsync(T.col + nnz, 0, 8 * sizeof(int), st);
        cudaMemsetAsync(T.val + nnz, 0, 8 * sizeof(double), st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            kry_set_error("transpose build failed: %s", cudaGetErrorString(e));
            rc = KRY_ERR_CUDA;
        }
    } while (0);
    cudaFree(rowidx);
    cudaFree(keys_out);
    cudaFree(perm_in);
    cudaFree(perm_out);
    cudaFree(tmp);
    if (rc == KRY_OK) rc = csr_finish(c, T);
    if (rc != KRY_OK) csr_dev_free(T);
    return rc;
}

// -------------------------------------------------------- device-side gallery
// Stencil functors: count(i) = stored entries of global row i; fill(i, c, v)
// writes them in ascending column order.
struct Poisson1dStencil {
    long long n;
    __device__ int count(long long i) const { return 1 + (i > 0) + (i < n - 1); }
    __device__ void fill(long long i, int *c, double *v) const
    {
        int k = 0;
        if (i > 0) { c[k] = (int)(i - 1); v[k++] = -1.0; }
        c[k] = (int)i; v[k++] = 2.0;
        if (i < n - 1) { c[k] = (int)(i + 1); v[k++] = -1.0; }
    }
};

struct Poisson2dStencil {       // reference gallery/gallery.py:10-29
    long long g, n;
    __device__ int count(long long i) const
    {
        const long long cx = i % g;
        return 1 + (i >= g) + (cx > 0) + (cx < g - 1) + (i < n - g);
    }
    __device__ void fill(long long i, int *c, double *v) const
    {
        const long long cx = i % g;
        int k = 0;
        if (i >= g) { c[k] = (int)(i - g); v[k++] = -1.0; }
        if (cx > 0) { c[k] = (int)(i - 1); v[k++] = -1.0; }
        c[k] = (int)i; v[k++] = 4.0;
        if (cx < g - 1) { c[k] = (int)(i + 1); v[k++] = -1.0; }
        if (i < n - g) { c[k] = (int)(i + g); v[k++] = -1.0; }
    }
};

struct ConvDiff3dStencil {      // BASELINE.json config 4
    long long m, n;
    double gamma;
    __device__ int count(long long i) const
    {
        const long long ix = i % m, iy = (i / m) % m, iz = i / (m * m);
        return 1 + (iz > 0) + (iy > 0) + (ix > 0) + (ix < m - 1) + (iy < m - 1) + (iz < m - 1);
    }
    __device__ void fill(long long i, int *c, double *v) const
    {
        const long long ix = i % m, iy = (i / m) % m, iz = i / (m * m);
        const double up = -1.0 - gamma;
        int k = 0;
        if (iz > 0) { c[k] = (int)(i - m * m); v[k++] = up; }
        if (iy > 0) { c[k] = (int)(i - m); v[k++] = up; }
        if (ix > 0) { c[k] = (int)(i - 1); v[k++] = up; }
        c[k] = (int)i; v[k++] = 6.0 + 3.0 * gamma;
        if (ix < m - 1) { c[k] = (int)(i + 1); v[k++] = -1.0; }
        if (iy < m - 1) { c[k] = (int)(i + m); v[k++] = -1.0; }
        if (iz < m - 1) { c[k] = (int)(i + m * m); v[k++] = -1.0; }
    }
};

template <class S>
__global__ void stencil_count_kernel(S s, long long row_begin, int nrows, int *counts)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= nrows; i += stride)
        counts[i] = (i < nrows) ? s.count(row_begin + i) : 0;
}

template <class S>
__global__ void stencil_fill_kernel(S s, long long row_begin, int nrows, const int *rowptr,
                                    int *col, double *val)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += stride) {
        const int k = rowptr[i];
        s.fill(row_begin + i, col + k, val + k);
    }
}

template <class S>
static int stencil_create(kry_ctx *c, S s, int64_t n, int64_t row_begin, int64_t row_end,
                          int max_per_row, uint32_t flags, kry_csr **out)
{
    KRY_REQUIRE(c && out, KRY_ERR_INVALID, "stencil: NULL argument");
    *out = nullptr;
    if (row_end < 0) row_end = n;
    KRY_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= n, KRY_ERR_INVALID,
                "stencil: rows [%lld,%lld) not inside [0,%lld)", (long long)row_begin,
                (long long)row_end, (long long)n);
    const int64_t nrows = row_end - row_begin;
    KRY_TRY(check_sizes(nrows, n, nrows * max_per_row));
    KRY_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    kry_csr *M = new (std::nothrow) kry_csr();
    KRY_REQUIRE(M, KRY_ERR_NOMEM, "stencil: host allocation failed");
    M->ctx = c;
    M->flags = flags;
    M->halo.n_global = n;
    M->halo.row_begin = row_begin;
    int rc = KRY_OK;
    void *tmp = nullptr;
    do {
        M->A.nrows = nrows;
        M->A.ncols = n;
        if ((rc = kry_alloc((void **)&M->A.rowptr, (size_t)(nrows + 1 + 8) * sizeof(int)))) break;
        const int g = c->sm_count * 8;
        stencil_count_kernel<<<g, 256, 0, st>>>(s, (long long)row_begin, (int)nrows, M->A.rowptr);
        c->launches++;
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, M->A.rowptr, M->A.rowptr,
                                      (int)nrows + 1, st);
        if ((rc = kry_alloc(&tmp, tmp_bytes + 256))) break;
        cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, M->A.rowptr, M->A.rowptr, (int)nrows + 1, st);
        int nnz = 0;
        cudaError_t e = cudaMemcpyAsync(&nnz, M->A.rowptr + nrows, sizeof(int),
                                        cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            kry_set_error("stencil: scan failed: %s", cudaGetErrorString(e));
            rc = KRY_ERR_CUDA;
            break;
        }
        M->A.nnz = nnz;
        if ((rc = kry_alloc((void **)&M->A.col, (size_t)(nnz + 8) * sizeof(int)))) break;
        if ((rc = kry_alloc((void **)&M->A.val, (size_t)(nnz + 8) * sizeof(double)))) break;
        cudaMemsetAsync(M->A.col + nnz, 0, 8 * sizeof(int), st);
        cudaMemsetAsync(M->A.val + nnz, 0, 8 * sizeof(double), st);
        if (nrows > 0) {
            stencil_fill_kernel<<<g, 256, 0, st>>>(s, (long long)row_begin, (int)nrows,
                                                   M->A.rowptr, M->A.col, M->A.val);
            c->launches++;
        }
        e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            kry_set_error("stencil: fill failed: %s", cudaGetErrorString(e));
            rc = KRY_ERR_CUDA;
            break;
        }
        rc = csr_finish(c, M->A);
        if (rc == KRY_OK && (flags & KRY_CSR_BUILD_TRANSPOSE) && !(flags & KRY_CSR_SYMMETRIC) &&
            nrows == n) {
            rc = csr_build_transpose_dev(c, M->A, M->T);
            M->has_T = (rc == KRY_OK);
        }
    } while (0);
    cudaFree(tmp);
    if (rc != KRY_OK) {
        csr_dev_free(M->A);
        csr_dev_free(M->T);
        delete M;
        return rc;
    }
    kry_ctx_retain(c);
    *out = M;
    return KRY_OK;
}

extern "C" int kry_csr_create_poisson1d(kry_ctx *c, int64_t n, int64_t row_begin, int64_t row_end,
                                        uint32_t flags, kry_csr **out)
{
    KRY_REQUIRE(n >= 1, KRY_ERR_INVALID, "poisson1d: n=%lld", (long long)n);
    Poisson1dStencil s{(long long)n};
    return stencil_create(c, s, n, row_begin, row_end, 3, flags | KRY_CSR_SYMMETRIC, out);
}

extern "C" int kry_csr_create_poisson2d(kry_ctx *c, int64_t g, int64_t row_begin, int64_t row_end,
                                        uint32_t flags, kry_csr **out)
{
    KRY_REQUIRE(g >= 1 && g < 46340, KRY_ERR_INVALID, "poisson2d: grid %lld", (long long)g);
    Poisson2dStencil s{(long long)g, (long long)g * g};
    return stencil_create(c, s, g * g, row_begin, row_end, 5, flags | KRY_CSR_SYMMETRIC, out);
}

extern "C" int kry_csr_create_convdiff3d(kry_ctx *c, int64_t m, double gamma, int64_t row_begin,
                                         int64_t row_end, uint32_t flags, kry_csr **out)
{
    KRY_REQUIRE(m >= 1 && m < 1290, KRY_ERR_INVALID, "convdiff3d: grid %lld", (long long)m);
    ConvDiff3dStencil s{(long long)m, (long long)m * m * m, gamma};
    return stencil_create(c, s, m * m * m, row_begin, row_end, 7, flags, out);
}
