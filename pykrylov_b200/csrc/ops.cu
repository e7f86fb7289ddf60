// ops.cu -- stand-alone hot-path entry points of the C ABI:
//   kry_spmv, kry_spmv_dot, kry_multi_axpy_dot
// (the device-resident solver loops in solvers.cu use the same kernels with
//  solver-specific epilogues; these generic forms back LinearOperator.__mul__
//  and user-composed iterations).
#include <new>

#include "vecops.cuh"

// y = A x, plus up to 3 fused  w_k . y  (w_k == nullptr means y . y)
template <int ND>
struct EpiStoreDots {
    double       *y;
    const double *w[ND > 0 ? ND : 1];
    __device__ void init() {}
    __device__ void operator()(int row, double ax, double *acc) const
    {
        y[row] = ax;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const double wv = w[d] ? w[d][row] : ax;
            acc[d] = __dadd_rn(acc[d], __dmul_rn(wv, ax));
        }
    }
};

static int check_spmv_args(const char *who, kry_csr *A, int trans, const kry_vec *x, const kry_vec *y)
{
    KRY_REQUIRE(A && x && y, KRY_ERR_INVALID, "%s: NULL argument", who);
    KRY_CTX_LIVE(A->ctx, who);
    KRY_REQUIRE(x->ctx == A->ctx && y->ctx == A->ctx, KRY_ERR_INVALID,
                "%s: operands belong to different contexts", who);
    KRY_REQUIRE(x->d != y->d, KRY_ERR_INVALID, "%s: x and y must not alias", who);
    const int64_t nin = trans ? A->A.nrows : (A->halo.active ? A->A.nrows : A->A.ncols);
    const int64_t nout = trans ? A->A.ncols : A->A.nrows;
    // the reference raises ValueError on a size mismatch (linop/linop.py:283-296)
    KRY_REQUIRE(x->n == nin, KRY_ERR_SHAPE, "%s: input has %lld entries, operator expects %lld",
                who, (long long)x->n, (long long)nin);
    KRY_REQUIRE(y->n == nout, KRY_ERR_SHAPE, "%s: output has %lld entries, operator produces %lld",
                who, (long long)y->n, (long long)nout);
    KRY_REQUIRE(!(trans && A->halo.active), KRY_ERR_UNSUPPORTED,
                "%s: transpose product of a row shard is not supported", who);
    if (A->halo.active)
        KRY_REQUIRE(x->cap >= A->A.ncols, KRY_ERR_SHAPE,
                    "%s: sharded operator needs an x vector with halo capacity %lld (has %lld)",
                    who, (long long)A->A.ncols, (long long)x->cap);
    return KRY_OK;
}

extern "C" int kry_spmv(kry_csr *A, int trans, const kry_vec *x, kry_vec *y)
{
    KRY_TRY(check_spmv_args("kry_spmv", A, trans, x, y));
    KRY_CUDA(cudaSetDevice(A->ctx->device));
    KRY_TRY(kry_halo_exchange(A, x->d));
    GatherPlain g{x->d};
    EpiStoreDots<0> e;
    e.y = y->d;
    e.w[0] = nullptr;
    return spmv_launch<0>(A, trans != 0, g, e, NoFin(), kry_gate(A->ctx), 0);
}

template <int ND>
static int spmv_dot_nd(kry_csr *A, int trans, const kry_vec *x, kry_vec *y,
                       const kry_vec *const *dot_with, int slot0)
{
    kry_ctx *c = A->ctx;
    GatherPlain g{x->d};
    EpiStoreDots<ND> e;
    e.y = y->d;
    for (int d = 0; d < ND; ++d) e.w[d] = (dot_with && dot_with[d]) ? dot_with[d]->d : nullptr;
    SlotFin fin{c->scalars + slot0, ND};
    if (c->nranks > 1 && A->halo.active) {
        // sharded: local sums -> all-reduce -> publish
        extern int kry_allreduce_sums(kry_ctx * c, int n);
        KRY_TRY(spmv_launch<ND>(A, trans != 0, g, e, fin, kry_gate(c), 1));
        KRY_TRY(kry_allreduce_sums(c, ND));
        return finalize_launch(c, fin, kry_gate(c));
    }
    return spmv_launch<ND>(A, trans != 0, g, e, fin, kry_gate(c), 0);
}

extern "C" int kry_spmv_dot(kry_csr *A, int trans, const kry_vec *x, kry_vec *y, int n_dots,
                            const kry_vec *const *dot_with, int slot0)
{
    KRY_TRY(check_spmv_args("kry_spmv_dot", A, trans, x, y));
    KRY_REQUIRE(n_dots >= 0 && n_dots <= 3, KRY_ERR_INVALID, "kry_spmv_dot: n_dots=%d", n_dots);
    KRY_REQUIRE(slot0 >= 0 && slot0 + n_dots <= KRY_NUM_SLOTS, KRY_ERR_INVALID,
                "kry_spmv_dot: scalar slots [%d,%d) out of range", slot0, slot0 + n_dots);
    for (int d = 0; d < n_dots; ++d)
        if (dot_with && dot_with[d])
            KRY_REQUIRE(dot_with[d]->n == y->n, KRY_ERR_SHAPE,
                        "kry_spmv_dot: dot operand %d has %lld entries, y has %lld", d,
                        (long long)dot_with[d]->n, (long long)y->n);
    if (n_dots == 0) return kry_spmv(A, trans, x, y);
    KRY_CUDA(cudaSetDevice(A->ctx->device));
    KRY_TRY(kry_halo_exchange(A, x->d));
    switch (n_dots) {
        case 1: return spmv_dot_nd<1>(A, trans, x, y, dot_with, slot0);
        case 2: return spmv_dot_nd<2>(A, trans, x, y, dot_with, slot0);
        default: return spmv_dot_nd<3>(A, trans, x, y, dot_with, slot0);
    }
}

int multi_axpy_check(const char *who, kry_ctx *c, int n_ops, const kry_axpby *ops, int n_dots,
                     const kry_dotspec *dots, int slot0, int64_t *n_out)
{
    KRY_REQUIRE(c, KRY_ERR_INVALID, "%s: NULL context", who);
    KRY_REQUIRE(n_ops >= 0 && n_ops <= 4 && n_dots >= 0 && n_dots <= 3 && n_ops + n_dots > 0,
                KRY_ERR_INVALID, "%s: n_ops=%d n_dots=%d", who, n_ops, n_dots);
    KRY_REQUIRE((n_ops == 0 || ops) && (n_dots == 0 || dots), KRY_ERR_INVALID, "%s: NULL op/dot array", who);
    KRY_REQUIRE(slot0 >= 0 && slot0 + n_dots <= KRY_NUM_SLOTS, KRY_ERR_INVALID,
                "%s: scalar slots [%d,%d) out of range", who, slot0, slot0 + n_dots);
    int64_t n = -1;
    auto chk = [&](const kry_vec *v) -> bool {
        if (!v) return true;
        if (v->ctx != c) return false;
        if (n < 0) n = v->n;
        return v->n == n;
    };
    for (int k = 0; k < n_ops; ++k) {
        KRY_REQUIRE(ops[k].z, KRY_ERR_INVALID, "%s: op %d has no output", who, k);
        KRY_REQUIRE(chk(ops[k].z) && chk(ops[k].u) && chk(ops[k].w), KRY_ERR_SHAPE,
                    "%s: op %d operand size/context mismatch", who, k);
        KRY_REQUIRE(ops[k].a_slot < KRY_NUM_SLOTS && ops[k].b_slot < KRY_NUM_SLOTS, KRY_ERR_INVALID,
                    "%s: op %d scalar slot out of range", who, k);
    }
    for (int d = 0; d < n_dots; ++d) {
        KRY_REQUIRE(dots[d].u && dots[d].w, KRY_ERR_INVALID, "%s: dot %d NULL", who, d);
        KRY_REQUIRE(chk(dots[d].u) && chk(dots[d].w), KRY_ERR_SHAPE, "%s: dot %d operand size/context mismatch", who, d);
    }
    *n_out = n;
    return KRY_OK;
}

extern "C" int kry_multi_axpy_dot(kry_ctx *c, int n_ops, const kry_axpby *ops, int n_dots,
                                  const kry_dotspec *dots, int slot0)
{
    int64_t n = -1;
    KRY_TRY(multi_axpy_check("kry_multi_axpy_dot", c, n_ops, ops, n_dots, dots, slot0, &n));
    KRY_CUDA(cudaSetDevice(c->device));
    if (n <= 0) {                       // empty vectors: inner products are exactly 0
        if (n_dots > 0) KRY_CUDA(cudaMemsetAsync(c->scalars + slot0, 0, n_dots * sizeof(double), c->stream));
        return KRY_OK;
    }
    SlotFin fin{c->scalars + slot0, n_dots};
    switch (n_dots) {
        case 0: return multi_axpy_run<0>(c, n, n_ops, ops, dots, fin);
        case 1: return multi_axpy_run<1>(c, n, n_ops, ops, dots, fin);
        case 2: return multi_axpy_run<2>(c, n, n_ops, ops, dots, fin);
        default: return multi_axpy_run<3>(c, n, n_ops, ops, dots, fin);
    }
}

int spmv_axpby_check(const char *who, kry_csr *A, int trans, const kry_vec *x, const kry_axpby *op,
                     int n_dots, const kry_vec *dot_with, int slot0)
{
    KRY_REQUIRE(op && op->z, KRY_ERR_INVALID, "%s: NULL op / output", who);
    KRY_TRY(check_spmv_args(who, A, trans, x, op->z));
    KRY_REQUIRE(!op->u, KRY_ERR_INVALID, "%s: op->u must be NULL (the product A x takes its place)", who);
    KRY_REQUIRE(n_dots == 0 || n_dots == 1, KRY_ERR_INVALID, "%s: n_dots=%d", who, n_dots);
    KRY_REQUIRE(slot0 >= 0 && slot0 + n_dots <= KRY_NUM_SLOTS, KRY_ERR_INVALID,
                "%s: scalar slot %d out of range", who, slot0);
    KRY_REQUIRE(op->a_slot < KRY_NUM_SLOTS && op->b_slot < KRY_NUM_SLOTS, KRY_ERR_INVALID,
                "%s: coefficient slot out of range", who);
    if (op->w)
        KRY_REQUIRE(op->w->ctx == A->ctx && op->w->n == op->z->n && op->w->d != x->d, KRY_ERR_SHAPE,
                    "%s: w must live on the operator's context, have the length of z and not alias x", who);
    if (dot_with)
        KRY_REQUIRE(n_dots == 1 && dot_with->ctx == A->ctx && dot_with->n == op->z->n, KRY_ERR_SHAPE,
                    "%s: dot operand size/context mismatch", who);
    return KRY_OK;
}

extern "C" int kry_spmv_axpby_dot(kry_csr *A, int trans, const kry_vec *x, const kry_axpby *op, int n_dots,
                                  const kry_vec *dot_with, int slot0)
{
    KRY_TRY(spmv_axpby_check("kry_spmv_axpby_dot", A, trans, x, op, n_dots, dot_with, slot0));
    KRY_CUDA(cudaSetDevice(A->ctx->device));
    KRY_TRY(kry_halo_exchange(A, x->d));
    SlotFin fin{A->ctx->scalars + slot0, n_dots};
    if (n_dots == 0) return spmv_axpby_run<0>(A, trans, x, op, nullptr, NoFin());
    return spmv_axpby_run<1>(A, trans, x, op, dot_with, fin);
}

// ------------------------------------------------------------ launch-sequence graphs
// A static sequence of stand-alone launches on a context's stream (kry_spmv, kry_spmv_dot,
// kry_multi_axpy_dot, kry_lls_step: one trip of an lls / SYMMLQ loop) captured once and replayed
// with one call per trip.  The sequence must have run un-captured before (first-use allocations
// and attribute calls do not belong into a capture); coefficients and stopping flags live in
// device memory, so a replay does exactly what the enqueued sequence would do.
struct kry_graph {
    kry_ctx        *ctx;
    cudaGraphExec_t exec;
    int64_t         launches;     // kernel launches inside one replay
    int64_t         l0;           // launch counter when the capture began
};

extern "C" int kry_graph_begin(kry_ctx *c, kry_graph **out)
{
    KRY_REQUIRE(c && out, KRY_ERR_INVALID, "kry_graph_begin: NULL argument");
    *out = nullptr;
    KRY_REQUIRE(!c->closed, KRY_ERR_STATE, "kry_graph_begin: the context was destroyed");
#ifdef KRY_EMULATE
    kry_set_error("kry_graph_begin: no CUDA graphs on the host emulation");
    return KRY_ERR_UNSUPPORTED;
#else
    KRY_REQUIRE(c->prof_cap == 0, KRY_ERR_STATE, "kry_graph_begin: per-launch profiling is on");
    KRY_CUDA(cudaSetDevice(c->device));
    kry_graph *g = new (std::nothrow) kry_graph();
    KRY_REQUIRE(g, KRY_ERR_NOMEM, "kry_graph_begin: host allocation failed");
    g->ctx = c;
    g->exec = nullptr;
    g->launches = 0;
    g->l0 = c->launches;
    cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) {
        delete g;
        KRY_CUDA(e);
    }
    kry_ctx_retain(c);
    *out = g;
    return KRY_OK;
#endif
}

extern "C" int kry_graph_end(kry_graph *g)
{
    KRY_REQUIRE(g && !g->exec, KRY_ERR_INVALID, "kry_graph_end: not a graph under capture");
#ifdef KRY_EMULATE
    return KRY_ERR_UNSUPPORTED;
#else
    kry_ctx *c = g->ctx;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    g->launches = c->launches - g->l0;
    c->launches = g->l0;                     // nothing has executed
    if (e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        kry_set_error("kry_graph_end: capture failed: %s", cudaGetErrorString(e));
        return KRY_ERR_CUDA;
    }
    e = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    KRY_CUDA(e);
    return KRY_OK;
#endif
}

extern "C" int kry_graph_launch(kry_graph *g, int times)
{
    KRY_REQUIRE(g && g->exec && times >= 0, KRY_ERR_INVALID, "kry_graph_launch: bad argument");
    KRY_CTX_LIVE(g->ctx, "kry_graph_launch");
#ifndef KRY_EMULATE
    for (int k = 0; k < times; ++k) {
        KRY_CUDA(cudaGraphLaunch(g->exec, g->ctx->stream));
        g->ctx->launches += g->launches;
    }
#endif
    return KRY_OK;
}

extern "C" int kry_graph_destroy(kry_graph *g)
{
    if (!g) return KRY_OK;
#ifndef KRY_EMULATE
    if (!g->ctx->closed) cudaStreamSynchronize(g->ctx->stream);
    if (g->exec) cudaGraphExecDestroy(g->exec);
#endif
    kry_ctx_release(g->ctx);
    delete g;
    return KRY_OK;
}
