// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL
// over NVLink/NVSwitch.  NCCL is dlopen()ed so that the single-GPU library has
// no link-time dependency on it.
//
// Data path per sharded SpMV:  pack boundary entries of x (one small kernel) ->
// ONE ncclAllGather of the packed boundary sets into the halo tail of x ->
// local SpMV over columns remapped to [local | halo].  Per fused inner product:
// ONE ncclAllReduce of <= 3 doubles, then a one-thread kernel runs the scalar
// recurrence.  Nothing else crosses NVLink.
#ifndef KRY_EMULATE
#include <cub/cub.cuh>
#endif
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "solver.cuh"

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

#ifdef KRY_EMULATE
extern "C" void emu_nccl_table(void **get_unique_id, void **comm_init_rank, void **comm_destroy, void **all_reduce,
                               void **all_gather, void **get_error_string);
#endif

int nccl_load()
{
    if (g_nccl.lib) return KRY_OK;
#ifdef KRY_EMULATE      // tests/emu: ranks are threads of one process, the collectives are host barriers
    emu_nccl_table((void **)&g_nccl.GetUniqueId, (void **)&g_nccl.CommInitRank, (void **)&g_nccl.CommDestroy,
                   (void **)&g_nccl.AllReduce, (void **)&g_nccl.AllGather, (void **)&g_nccl.GetErrorString);
    g_nccl.lib = (void *)&g_nccl;
    return KRY_OK;
#endif
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    KRY_REQUIRE(h, KRY_ERR_COMM, "NCCL not found: %s", dlerror());
#define LOAD(field, sym)                                                     \
    *(void **)(&g_nccl.field) = dlsym(h, sym);                               \
    KRY_REQUIRE(g_nccl.field, KRY_ERR_COMM, "NCCL symbol %s missing", sym)
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(AllReduce, "ncclAllReduce");
    LOAD(AllGather, "ncclAllGather");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    g_nccl.lib = h;
    return KRY_OK;
}

#define KRY_NCCL(call)                                                                \
    do {                                                                              \
        ncclResult_t r_ = (call);                                                     \
        if (r_ != ncclSuccess) {                                                      \
            kry_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,               \
                          g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?");   \
            return KRY_ERR_COMM;                                                      \
        }                                                                             \
    } while (0)

}  // namespace

static_assert(sizeof(ncclUniqueId) == KRY_COMM_ID_BYTES, "ncclUniqueId size");

// ------------------------------------------- NVLink peer-memory inboxes (CUDA IPC)
// Every rank owns a small inbox [2 slots][nranks][8 doubles]; all inboxes are mapped into
// every process, so the last CTA of a fused reduction can deposit its totals directly in
// its peers' memory (common.cuh, block_reduce_finalize).
static void p2p_teardown(kry_ctx *c)
{
    for (int q = 0; q < 16; ++q) {
        if (c->p2p_peer_ptr[q] && q != c->rank) cudaIpcCloseMemHandle(c->p2p_peer_ptr[q]);
        c->p2p_peer_ptr[q] = nullptr;
    }
    if (c->p2p_inbox) cudaFree(c->p2p_inbox);
    if (c->p2p_peers_dev) cudaFree(c->p2p_peers_dev);
    if (c->p2p_seq) cudaFree(c->p2p_seq);
    c->p2p_inbox = nullptr;
    c->p2p_peers_dev = nullptr;
    c->p2p_seq = nullptr;
    c->p2p_on = 0;
    cudaGetLastError();
}

extern "C" int kry_comm_allgather_host(kry_ctx *c, const void *send, void *recv, int64_t bytes);

static int p2p_setup(kry_ctx *c)
{
    const int P = c->nranks;
    // cudaIpcGetMemHandle describes the *base* of the driver allocation a pointer lives in and
    // cudaIpcOpenMemHandle returns that base: small cudaMalloc blocks are sub-allocated, so the
    // inbox gets a dedicated 2 MiB allocation (its own base) and the mapping is verified below
    // with a per-rank magic word before it is trusted.
    const size_t inbox_bytes = (size_t)2 << 20;
    const size_t magic_at = 2 * 16 * 8;                       // first double behind the slots
    KRY_TRY(kry_alloc((void **)&c->p2p_inbox, inbox_bytes));
    KRY_CUDA(cudaMemset(c->p2p_inbox, 0, inbox_bytes));
    const double my_magic = 7777.0 + c->rank;
    KRY_CUDA(cudaMemcpy(c->p2p_inbox + magic_at, &my_magic, sizeof(double), cudaMemcpyHostToDevice));
    KRY_TRY(kry_alloc((void **)&c->p2p_peers_dev, 16 * sizeof(double *)));
    KRY_TRY(kry_alloc((void **)&c->p2p_seq, 256));
    KRY_CUDA(cudaMemset(c->p2p_seq, 0, 256));
    cudaIpcMemHandle_t mine;
    KRY_CUDA(cudaIpcGetMemHandle(&mine, c->p2p_inbox));
    std::vector<cudaIpcMemHandle_t> all((size_t)P);
    KRY_TRY(kry_comm_allgather_host(c, &mine, all.data(), sizeof(mine)));
    int ok = 1;
    for (int q = 0; q < P && ok; ++q) {
        if (q == c->rank) {
            c->p2p_peer_ptr[q] = c->p2p_inbox;
            continue;
        }
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
            break;
        }
        c->p2p_peer_ptr[q] = ptr;
        double seen = 0.0;                                    // does the mapping address rank q's inbox?
        if (cudaMemcpy(&seen, (double *)ptr + magic_at, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
            seen != 7777.0 + q) {
            cudaGetLastError();
            ok = 0;
        }
    }
    // all ranks must agree, otherwise some would wait for peers that use NCCL
    double flag = ok ? 1.0 : 0.0, neg = -flag;
    double mm[2] = {flag, neg};
    KRY_TRY(kry_comm_allreduce_host(c, mm, 2, 1));           // max(flag), max(-flag) = -min(flag)
    KRY_REQUIRE(-mm[1] >= 1.0, KRY_ERR_COMM, "peer mapping failed on some rank");
    KRY_CUDA(cudaMemcpy(c->p2p_peers_dev, c->p2p_peer_ptr, (size_t)P * sizeof(void *), cudaMemcpyHostToDevice));
    KRY_CUDA(cudaDeviceSynchronize());
    c->p2p_on = 1;
    return KRY_OK;
}

extern "C" int kry_comm_unique_id(void *id128)
{
    KRY_REQUIRE(id128, KRY_ERR_INVALID, "kry_comm_unique_id: NULL output");
    KRY_TRY(nccl_load());
    ncclUniqueId id;
    KRY_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return KRY_OK;
}

extern "C" int kry_comm_init(kry_ctx *c, int nranks, int rank, const void *id128)
{
    KRY_REQUIRE(c && id128, KRY_ERR_INVALID, "kry_comm_init: NULL argument");
    KRY_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, KRY_ERR_INVALID,
                "kry_comm_init: rank %d of %d", rank, nranks);
    KRY_REQUIRE(!c->nccl, KRY_ERR_STATE, "kry_comm_init: communicator already initialised");
    KRY_TRY(nccl_load());
    KRY_CUDA(cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    KRY_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    c->nccl = (void *)comm;
    c->nranks = nranks;
    c->rank = rank;
    if (nranks > 1 && nranks <= 16) {
        // best effort: without peer mapping the NCCL all-reduce path stays in use
        if (p2p_setup(c) != KRY_OK) p2p_teardown(c);
    }
    return KRY_OK;
}

extern "C" int kry_comm_destroy(kry_ctx *c)
{
    if (!c || !c->nccl) return KRY_OK;
    cudaStreamSynchronize(c->stream);
    p2p_teardown(c);
    g_nccl.CommDestroy((ncclComm_t)c->nccl);
    c->nccl = nullptr;
    c->nranks = 1;
    c->rank = 0;
    return KRY_OK;
}

extern "C" int kry_comm_size(kry_ctx *c, int *nranks, int *rank)
{
    KRY_REQUIRE(c, KRY_ERR_INVALID, "kry_comm_size: NULL context");
    if (nranks) *nranks = c->nranks;
    if (rank) *rank = c->rank;
    return KRY_OK;
}

// all-reduce of the fused-reduction totals (c->sums[0..n)) in place, on the stream
int kry_allreduce_sums(kry_ctx *c, int n)
{
    if (c->nranks <= 1) return KRY_OK;
    KRY_REQUIRE(c->nccl, KRY_ERR_COMM, "all-reduce: communicator not initialised");
    KRY_NCCL(g_nccl.AllReduce(c->sums, c->sums, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)c->nccl,
                              c->stream));
    return KRY_OK;
}

extern "C" int kry_comm_allreduce_host(kry_ctx *c, double *inout, int count, int op)
{
    KRY_REQUIRE(c && inout && count >= 0 && count <= 64, KRY_ERR_INVALID,
                "kry_comm_allreduce_host: bad argument");
    if (c->nranks <= 1 || count == 0) return KRY_OK;
    KRY_REQUIRE(c->nccl, KRY_ERR_COMM, "kry_comm_allreduce_host: communicator not initialised");
    double *d = nullptr;
    KRY_TRY(kry_alloc((void **)&d, 64 * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d, inout, count * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    ncclResult_t r = ncclSuccess;
    if (e == cudaSuccess)
        r = g_nccl.AllReduce(d, d, (size_t)count, ncclDouble, op == 1 ? ncclMax : ncclSum,
                             (ncclComm_t)c->nccl, c->stream);
    if (e == cudaSuccess && r == ncclSuccess)
        e = cudaMemcpyAsync(inout, d, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    KRY_REQUIRE(r == ncclSuccess, KRY_ERR_COMM, "kry_comm_allreduce_host: %s", g_nccl.GetErrorString(r));
    KRY_CUDA(e);
    return KRY_OK;
}

extern "C" int kry_comm_barrier(kry_ctx *c)
{
    KRY_REQUIRE(c, KRY_ERR_INVALID, "kry_comm_barrier: NULL context");
    double z = 0.0;
    KRY_TRY(kry_comm_allreduce_host(c, &z, 1, 0));
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    return KRY_OK;
}

extern "C" int kry_comm_allgather_host(kry_ctx *c, const void *send, void *recv, int64_t bytes)
{
    KRY_REQUIRE(c && send && recv && bytes >= 0, KRY_ERR_INVALID, "kry_comm_allgather_host: bad argument");
    if (c->nranks <= 1) {
        memcpy(recv, send, (size_t)bytes);
        return KRY_OK;
    }
    KRY_REQUIRE(c->nccl, KRY_ERR_COMM, "kry_comm_allgather_host: communicator not initialised");
    if (bytes == 0) return KRY_OK;
    char *d_s = nullptr, *d_r = nullptr;
    KRY_TRY(kry_alloc((void **)&d_s, (size_t)bytes));
    int rc = kry_alloc((void **)&d_r, (size_t)bytes * c->nranks);
    if (rc != KRY_OK) {
        cudaFree(d_s);
        return rc;
    }
    cudaError_t e = cudaMemcpyAsync(d_s, send, (size_t)bytes, cudaMemcpyHostToDevice, c->stream);
    ncclResult_t r = ncclSuccess;
    if (e == cudaSuccess)
        r = g_nccl.AllGather(d_s, d_r, (size_t)bytes, ncclChar, (ncclComm_t)c->nccl, c->stream);
    if (e == cudaSuccess && r == ncclSuccess)
        e = cudaMemcpyAsync(recv, d_r, (size_t)bytes * c->nranks, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_s);
    cudaFree(d_r);
    KRY_REQUIRE(r == ncclSuccess, KRY_ERR_COMM, "kry_comm_allgather_host: %s", g_nccl.GetErrorString(r));
    KRY_CUDA(e);
    return KRY_OK;
}

// ------------------------------------------------------------ halo exchange
__global__ void halo_pack_kernel(const double *x, const int *idx, int n_send, int max_send, double *buf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < max_send) buf[i] = (i < n_send) ? x[idx[i]] : 0.0;
}

// Fused CG forms on shards: the boundary entries travel already updated,
// buf = beta * p[idx] - r[idx]  (cg.py:150-151; same two rounded operations as everywhere else).
__global__ void halo_pack_dir_kernel(const double *p, const double *r, const double *beta_ptr,
                                     const int *idx, int n_send, int max_send, double *buf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= max_send) return;
    double v = 0.0;
    if (i < n_send) {
        const int j = idx[i];
        v = __dsub_rn(__dmul_rn(*beta_ptr, p[j]), r[j]);
    }
    buf[i] = v;
}

// x_dev: [n_local | nranks*max_send] -- fills the tail.  No-op for unsharded operators.
// With r_dev/beta_dev the packed entries are beta*x - r instead of x (see above).
int kry_halo_exchange_dir(kry_csr *M, double *x_dev, const double *r_dev, const double *beta_dev)
{
    HaloPlan &h = M->halo;
    if (!h.active) return KRY_OK;
    kry_ctx *c = M->ctx;
    if (h.max_send == 0) return KRY_OK;
#ifdef KRY_EMULATE
    if (r_dev)
        emu_launch<0>((h.max_send + 255) / 256, 256, ReduceWs(), NoFin(), [&] {
            halo_pack_dir_kernel(x_dev, r_dev, beta_dev, h.send_idx, h.n_send, h.max_send, h.send_buf); });
    else
        emu_launch<0>((h.max_send + 255) / 256, 256, ReduceWs(), NoFin(), [&] {
            halo_pack_kernel(x_dev, h.send_idx, h.n_send, h.max_send, h.send_buf); });
#else
    if (r_dev)
        halo_pack_dir_kernel<<<(h.max_send + 255) / 256, 256, 0, c->stream>>>(
            x_dev, r_dev, beta_dev, h.send_idx, h.n_send, h.max_send, h.send_buf);
    else
        halo_pack_kernel<<<(h.max_send + 255) / 256, 256, 0, c->stream>>>(x_dev, h.send_idx, h.n_send,
                                                                          h.max_send, h.send_buf);
#endif
    c->launches++;
    KRY_CUDA(cudaGetLastError());
    double *tail = x_dev + M->A.nrows;
    if (c->nranks <= 1) {
        KRY_CUDA(cudaMemcpyAsync(tail, h.send_buf, (size_t)h.max_send * sizeof(double),
                                 cudaMemcpyDeviceToDevice, c->stream));
        return KRY_OK;
    }
    KRY_REQUIRE(c->nccl, KRY_ERR_COMM, "halo exchange: communicator not initialised");
    KRY_NCCL(g_nccl.AllGather(h.send_buf, tail, (size_t)h.max_send, ncclDouble, (ncclComm_t)c->nccl,
                              c->stream));
    return KRY_OK;
}

int kry_halo_exchange(kry_csr *M, double *x_dev)
{
    return kry_halo_exchange_dir(M, x_dev, nullptr, nullptr);
}

// ------------------------------------------------------- shard finalisation
struct OffShard {
    int lo, hi;
    __device__ bool operator()(int c) const { return c < lo || c >= hi; }
};

__global__ void remap_cols_kernel(int *col, int nnz, int lo, int hi, const int *halo_cols,
                                  const int *halo_map, int n_halo)
{
    const int stride = gridDim.x * blockDim.x;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
        const int c = col[k];
        if (c >= lo && c < hi) {
            col[k] = c - lo;
        } else {
            int a = 0, b = n_halo;          // lower_bound in the sorted unique halo columns
            while (a < b) {
                const int m = (a + b) >> 1;
                if (halo_cols[m] < c) a = m + 1; else b = m;
            }
            col[k] = halo_map[a];
        }
    }
}

// edge[0] = 1 + last row of the first half that touches a halo column (c >= n_local),
// edge[1] = first such row of the second half
__global__ void halo_rows_kernel(const int *rowptr, const int *col, int nrows, int *edge)
{
    const int stride = gridDim.x * blockDim.x, half = nrows / 2;
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += stride) {
        bool touches = false;
        for (int k = rowptr[row]; k < rowptr[row + 1] && !touches; ++k) touches = col[k] >= nrows;
        if (touches) {
            if (row < half) atomicMax(edge, row + 1);
            else atomicMin(edge + 1, row);
        }
    }
}

extern "C" int kry_csr_shard_finalize(kry_csr *M, int64_t n_global, int64_t row_begin)
{
    KRY_REQUIRE(M, KRY_ERR_INVALID, "kry_csr_shard_finalize: NULL operator");
    KRY_CTX_LIVE(M->ctx, "kry_csr_shard_finalize");
    KRY_REQUIRE(!M->halo.active, KRY_ERR_STATE, "kry_csr_shard_finalize: already finalised");
    kry_ctx *c = M->ctx;
    const int P = c->nranks, me = c->rank;
    const int64_t n_local = M->A.nrows;
    KRY_REQUIRE(M->A.ncols == n_global && row_begin >= 0 && row_begin + n_local <= n_global,
                KRY_ERR_SHAPE, "kry_csr_shard_finalize: shard rows [%lld,%lld) of %lld, operator has %lld columns",
                (long long)row_begin, (long long)(row_begin + n_local), (long long)n_global,
                (long long)M->A.ncols);
    KRY_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int nnz = (int)M->A.nnz;
    const int lo = (int)row_begin, hi = (int)(row_begin + n_local);

    // 1. unique off-shard columns (device: select -> sort -> unique)
    int *d_sel = nullptr, *d_sorted = nullptr, *d_count = nullptr;
    void *tmp = nullptr;
    std::vector<int> need;
#ifdef KRY_EMULATE      // tests/emu: the same set with the host's sort / unique
    (void)d_sel; (void)d_sorted; (void)d_count; (void)tmp;
    for (int k = 0; k < nnz; ++k)
        if (M->A.col[k] < lo || M->A.col[k] >= hi) need.push_back(M->A.col[k]);
    std::sort(need.begin(), need.end());
    need.erase(std::unique(need.begin(), need.end()), need.end());
#else
    {
        KRY_TRY(kry_alloc((void **)&d_sel, (size_t)(nnz + 1) * sizeof(int)));
        KRY_TRY(kry_alloc((void **)&d_sorted, (size_t)(nnz + 1) * sizeof(int)));
        KRY_TRY(kry_alloc((void **)&d_count, 256));
        size_t b1 = 0, b2 = 0, b3 = 0;
        OffShard pred{lo, hi};
        cub::DeviceSelect::If(nullptr, b1, M->A.col, d_sel, d_count, nnz, pred, st);
        cub::DeviceRadixSort::SortKeys(nullptr, b2, d_sel, d_sorted, nnz, 0, 32, st);
        cub::DeviceSelect::Unique(nullptr, b3, d_sorted, d_sel, d_count, nnz, st);
        size_t bytes = std::max(b1, std::max(b2, b3)) + 256;
        KRY_TRY(kry_alloc(&tmp, bytes));
        int n_off = 0, n_uni = 0;
        cub::DeviceSelect::If(tmp, bytes, M->A.col, d_sel, d_count, nnz, pred, st);
        KRY_CUDA(cudaMemcpyAsync(&n_off, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        KRY_CUDA(cudaStreamSynchronize(st));
        if (n_off > 0) {
            cub::DeviceRadixSort::SortKeys(tmp, bytes, d_sel, d_sorted, n_off, 0, 32, st);
            cub::DeviceSelect::Unique(tmp, bytes, d_sorted, d_sel, d_count, n_off, st);
            KRY_CUDA(cudaMemcpyAsync(&n_uni, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
            KRY_CUDA(cudaStreamSynchronize(st));
            need.resize(n_uni);
            KRY_CUDA(cudaMemcpyAsync(need.data(), d_sel, (size_t)n_uni * sizeof(int),
                                     cudaMemcpyDeviceToHost, st));
            KRY_CUDA(cudaStreamSynchronize(st));
        }
        cudaFree(tmp);
        cudaFree(d_sorted);
        cudaFree(d_count);
        cudaFree(d_sel);
    }
#endif

    // 2. everyone learns every shard's row range and need list
    std::vector<int64_t> ranges((size_t)2 * P);
    int64_t mine[2] = {row_begin, row_begin + n_local};
    KRY_TRY(kry_comm_allgather_host(c, mine, ranges.data(), sizeof(mine)));
    std::vector<int64_t> counts(P);
    int64_t my_count = (int64_t)need.size();
    KRY_TRY(kry_comm_allgather_host(c, &my_count, counts.data(), sizeof(int64_t)));
    int64_t max_need = 0;
    for (int q = 0; q < P; ++q) max_need = std::max(max_need, counts[q]);
    std::vector<int> need_pad((size_t)std::max<int64_t>(max_need, 1), -1), all_need;
    std::copy(need.begin(), need.end(), need_pad.begin());
    all_need.resize(need_pad.size() * P);
    KRY_TRY(kry_comm_allgather_host(c, need_pad.data(), all_need.data(),
                                    (int64_t)need_pad.size() * sizeof(int)));

    // 3. my boundary set = union of what the others need from my rows (sorted, unique)
    std::vector<int> send;
    for (int q = 0; q < P; ++q) {
        if (q == me) continue;
        const int *lst = all_need.data() + (size_t)q * need_pad.size();
        for (int64_t i = 0; i < counts[q]; ++i)
            if (lst[i] >= lo && lst[i] < hi) send.push_back(lst[i]);
    }
    std::sort(send.begin(), send.end());
    send.erase(std::unique(send.begin(), send.end()), send.end());

    // 4. everyone learns every boundary set
    std::vector<int64_t> scounts(P);
    int64_t my_send = (int64_t)send.size();
    KRY_TRY(kry_comm_allgather_host(c, &my_send, scounts.data(), sizeof(int64_t)));
    int64_t max_send = 0;
    for (int q = 0; q < P; ++q) max_send = std::max(max_send, scounts[q]);
    std::vector<int> send_pad((size_t)std::max<int64_t>(max_send, 1), -1), all_send;
    std::copy(send.begin(), send.end(), send_pad.begin());
    all_send.resize(send_pad.size() * P);
    KRY_TRY(kry_comm_allgather_host(c, send_pad.data(), all_send.data(),
                                    (int64_t)send_pad.size() * sizeof(int)));
    KRY_REQUIRE(n_local + (int64_t)P * max_send < (int64_t)INT32_MAX - (1 << 20), KRY_ERR_UNSUPPORTED,
                "kry_csr_shard_finalize: [local | halo] index space exceeds int32");

    // 5. halo column -> index in [n_local + owner*max_send + position in owner's set]
    std::vector<int> halo_map(need.size());
    for (size_t i = 0; i < need.size(); ++i) {
        const int col = need[i];
        int owner = -1;
        for (int q = 0; q < P; ++q)
            if (col >= ranges[2 * q] && col < ranges[2 * q + 1]) { owner = q; break; }
        KRY_REQUIRE(owner >= 0 && owner != me, KRY_ERR_INVALID,
                    "kry_csr_shard_finalize: column %d is owned by no other shard", col);
        const int *lst = all_send.data() + (size_t)owner * send_pad.size();
        const int *pos = std::lower_bound(lst, lst + scounts[owner], col);
        KRY_REQUIRE(pos != lst + scounts[owner] && *pos == col, KRY_ERR_COMM,
                    "kry_csr_shard_finalize: column %d missing from shard %d's boundary set", col, owner);
        halo_map[i] = (int)(n_local + (int64_t)owner * max_send + (pos - lst));
    }

    // 6. remap on device, install the plan
    int *d_need = nullptr, *d_map = nullptr;
    KRY_TRY(kry_alloc((void **)&d_need, (need.size() + 1) * sizeof(int)));
    KRY_TRY(kry_alloc((void **)&d_map, (need.size() + 1) * sizeof(int)));
    if (!need.empty()) {
        KRY_CUDA(cudaMemcpyAsync(d_need, need.data(), need.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        KRY_CUDA(cudaMemcpyAsync(d_map, halo_map.data(), need.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    if (nnz > 0) {
#ifdef KRY_EMULATE
        emu_launch<0>(c->sm_count * 8, 256, ReduceWs(), NoFin(), [&] {
            remap_cols_kernel(M->A.col, nnz, lo, hi, d_need, d_map, (int)need.size()); });
#else
        remap_cols_kernel<<<c->sm_count * 8, 256, 0, st>>>(M->A.col, nnz, lo, hi, d_need, d_map,
                                                           (int)need.size());
#endif
        c->launches++;
    }
    KRY_CUDA(cudaStreamSynchronize(st));
    KRY_CUDA(cudaGetLastError());
    cudaFree(d_need);
    cudaFree(d_map);

    HaloPlan &h = M->halo;
    h.n_global = n_global;
    h.row_begin = row_begin;
    h.n_send = (int)send.size();
    h.max_send = (int)max_send;
    KRY_TRY(kry_alloc((void **)&h.send_idx, (send.size() + 1) * sizeof(int)));
    KRY_TRY(kry_alloc((void **)&h.send_buf, (size_t)(max_send + 1) * sizeof(double)));
    if (!send.empty()) {
        std::vector<int> local(send.size());
        for (size_t i = 0; i < send.size(); ++i) local[i] = send[i] - lo;
        KRY_CUDA(cudaMemcpyAsync(h.send_idx, local.data(), local.size() * sizeof(int),
                                 cudaMemcpyHostToDevice, st));
        KRY_CUDA(cudaStreamSynchronize(st));
    }
    M->A.ncols = n_local + (int64_t)P * max_send;

    // peer-memory halo (KRY_OPT_HALO_P2P): readers of my boundary entries, owners of my halo columns
    h.n_to = h.n_from = 0;
    for (int q = 0; q < P && P <= KRY_MAX_RANKS; ++q) {
        if (q == me) continue;
        const int *lst = all_need.data() + (size_t)q * need_pad.size();
        bool reads_me = false;
        for (int64_t i = 0; i < counts[q] && !reads_me; ++i) reads_me = lst[i] >= lo && lst[i] < hi;
        if (reads_me) h.to_rank[h.n_to++] = q;
        bool i_read = false;
        for (size_t i = 0; i < need.size() && !i_read; ++i)
            i_read = need[i] >= ranges[2 * q] && need[i] < ranges[2 * q + 1];
        if (i_read) h.from_rank[h.n_from++] = q;
    }
    // rows that touch halo columns: for a banded operator two thin slabs at the ends of the shard; the
    // sharded launch walks the rows in between first (rot / v_wait) and waits for the peers only then
    h.rot = 0;
    h.v_wait = (int)n_local;
    if (n_local > 0 && nnz > 0) {
        int *d_edge = nullptr;
        KRY_TRY(kry_alloc((void **)&d_edge, 256));
        const int init[2] = {0, (int)n_local};
        KRY_CUDA(cudaMemcpyAsync(d_edge, init, sizeof(init), cudaMemcpyHostToDevice, st));
#ifdef KRY_EMULATE
        emu_launch<0>(c->sm_count * 8, 256, ReduceWs(), NoFin(), [&] {
            halo_rows_kernel(M->A.rowptr, M->A.col, (int)n_local, d_edge); });
#else
        halo_rows_kernel<<<c->sm_count * 8, 256, 0, st>>>(M->A.rowptr, M->A.col, (int)n_local, d_edge);
#endif
        c->launches++;
        int edge[2] = {0, 0};
        KRY_CUDA(cudaMemcpyAsync(edge, d_edge, sizeof(edge), cudaMemcpyDeviceToHost, st));
        KRY_CUDA(cudaStreamSynchronize(st));
        cudaFree(d_edge);
        // lo = edge[0] leading and n - edge[1] trailing rows may touch halo columns: the launch walks
        // rows [lo, edge[1]) first, then the trailing and (wrapped around) the leading slab
        h.rot = edge[0];
        h.v_wait = edge[1] - edge[0];
        if (h.v_wait < 0) { h.rot = 0; h.v_wait = 0; }
    }
    h.active = true;
    return KRY_OK;
}

// ------------------------------------------------- peer-memory halo: solver linkage
// Collective over the ranks (every rank creates its solvers in the same order).  Each rank exports
// its solver slab through CUDA IPC together with the layout of its vectors; every rank then
// builds, per gathered vector, the table of remote tail slots its boundary entries go to and of
// the flag words it signals / waits on.  Any failure on any rank leaves every rank on the
// pack + ncclAllGather exchange.
struct SlabInfo {
    cudaIpcMemHandle_t handle;
    int64_t            n_local, max_send, total;
    int64_t            voff[16];      // offset of vector i in the slab, in doubles
};

int kry_halo_link(kry_solver *S)
{
    kry_ctx *c = S->ctx;
    const HaloPlan &hp = S->A->halo;
    const int P = c->nranks, me = c->rank;
    S->halo_linked = false;
    if (!S->sharded || !c->p2p_inbox || P > KRY_MAX_RANKS) return KRY_OK;
    KRY_CUDA(cudaSetDevice(c->device));
    const double magic = 424242.0 + me;
    KRY_CUDA(cudaMemcpyAsync(S->slab + S->slab_doubles, &magic, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    SlabInfo mine;
    memset(&mine, 0, sizeof(mine));
    int ok = cudaIpcGetMemHandle(&mine.handle, S->slab) == cudaSuccess;
    cudaGetLastError();
    mine.n_local = S->n;
    mine.max_send = hp.max_send;
    mine.total = S->slab_doubles;
    for (int i = 0; i < S->nvecs; ++i) mine.voff[i] = S->vecs[i].d - S->slab;
    std::vector<SlabInfo> all((size_t)P);
    KRY_TRY(kry_comm_allgather_host(c, &mine, all.data(), sizeof(mine)));
    for (int k = 0; k < hp.n_to && ok; ++k) {
        const int q = hp.to_rank[k];
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
            break;
        }
        S->peer_slab[q] = ptr;
        double seen = 0.0;                      // does the mapping address rank q's slab?
        if (cudaMemcpy(&seen, (double *)ptr + all[q].total, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
            seen != 424242.0 + q || all[q].max_send != hp.max_send) {
            cudaGetLastError();
            ok = 0;
        }
    }
    double mm[2] = {ok ? 1.0 : 0.0, ok ? -1.0 : 0.0};
    KRY_TRY(kry_comm_allreduce_host(c, mm, 2, 1));            // max(ok), max(-ok) = -min(ok)
    if (-mm[1] < 1.0) {
        kry_halo_unlink(S);
        return KRY_OK;                                        // every rank stays on the NCCL exchange
    }
    std::vector<HaloTable> tbl((size_t)S->nvecs);
    memset(tbl.data(), 0, tbl.size() * sizeof(HaloTable));
    unsigned long long *my_flags = reinterpret_cast<unsigned long long *>(c->p2p_inbox + KRY_HALO_FLAG_OFFSET);
    for (int i = 0; i < S->nvecs; ++i) {
        HaloTable &T = tbl[(size_t)i];
        T.n_to = hp.n_to;
        T.n_from = hp.n_from;
        for (int k = 0; k < hp.n_to; ++k) {
            const int q = hp.to_rank[k];
            T.to_tail[k] = (double *)S->peer_slab[q] + all[q].voff[i] + all[q].n_local + (int64_t)me * hp.max_send;
            T.to_flag[k] = reinterpret_cast<unsigned long long *>((double *)c->p2p_peer_ptr[q] + KRY_HALO_FLAG_OFFSET) + 8 * me;
        }
        for (int k = 0; k < hp.n_from; ++k) T.from_flag[k] = my_flags + 8 * hp.from_rank[k];
    }
    KRY_TRY(kry_alloc((void **)&S->halo_tbl, tbl.size() * sizeof(HaloTable)));
    KRY_CUDA(cudaMemcpy(S->halo_tbl, tbl.data(), tbl.size() * sizeof(HaloTable), cudaMemcpyHostToDevice));
    S->halo_linked = true;
    return KRY_OK;
}

void kry_halo_unlink(kry_solver *S)
{
    for (int q = 0; q < KRY_MAX_RANKS; ++q) {
        if (S->peer_slab[q]) cudaIpcCloseMemHandle(S->peer_slab[q]);
        S->peer_slab[q] = nullptr;
    }
    if (S->halo_tbl) cudaFree(S->halo_tbl);
    S->halo_tbl = nullptr;
    S->halo_linked = false;
    cudaGetLastError();
}
