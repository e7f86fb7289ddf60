// assemble.cu -- device-side operator construction.
//
//   kry_csr_create_coo   coordinate triplets -> CSR in HBM, optionally expanding one stored
//                        triangle of a symmetric operator.  Inside every row the entries keep
//                        the order in which the reference's CoordLinearOperator loop
//                        (linop/linop.py:657-664) accumulates them -- entry k adds
//                        vals[k]*x[cols[k]] to y[rows[k]] and, when symmetric and off the
//                        diagonal, vals[k]*x[rows[k]] to y[cols[k]] -- so the CSR row sums of
//                        the SpMV kernels are that loop's sums bit for bit.
//   kry_csr_combine      alpha*A [+ beta*B] [+ gamma*diag(d)] as a new CSR in HBM: per row the
//                        (scaled) entries of A, then those of B, then the diagonal entry.
//                        Operator algebra on device operators (A + sigma*I, -A, 2*A, A + B)
//                        therefore stays a device operator and the solver loops stay on the GPU.
//                        A + sigma*I and A +- D reproduce the reference's closure
//                        `A(v) + (sigma*I)(v)` (linop.py:378-396) bit for bit: the appended entry
//                        is added last, exactly like the second summand.
//   kry_csr_to_dense     dense row-major copy (LinearOperator.to_array, linop.py:256-269): each
//                        row scatters its entries in order, which equals the reference's n
//                        products with unit vectors (sums of exact terms v*1 and zeros).
#include <cub/cub.cuh>
#include <string.h>

#include <new>

#include "common.cuh"

namespace {

// slot 2k: entry k itself; slot 2k+1: its mirror image (symmetric storage, off-diagonal only)
__global__ void coo_expand_kernel(const int *rows, const int *cols, const double *vals, int nnz, int sym,
                                  int nrows, int ncols, int *erow, int *ecol, double *eval, int *keep, int *bad)
{
    const int stride = gridDim.x * blockDim.x;
    int nbad = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
        const int r = rows[k], c = cols[k];
        const bool ok = r >= 0 && r < nrows && c >= 0 && c < ncols && (!sym || (r < ncols && c < nrows));
        nbad += !ok;
        erow[2 * k] = r;
        ecol[2 * k] = c;
        eval[2 * k] = vals[k];
        keep[2 * k] = ok;
        const bool mirror = ok && sym && r != c;
        erow[2 * k + 1] = c;
        ecol[2 * k + 1] = r;
        eval[2 * k + 1] = vals[k];
        keep[2 * k + 1] = mirror;
    }
    if (nbad) atomicAdd(bad, nbad);
}

__global__ void compact_kernel(const int *erow, const int *ecol, const double *eval, const int *keep,
                               const int *pos, int n, int *row, int *col, double *val, int *counts)
{
    const int stride = gridDim.x * blockDim.x;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        if (!keep[j]) continue;
        const int p = pos[j];
        row[p] = erow[j];
        col[p] = ecol[j];
        val[p] = eval[j];
        atomicAdd(counts + erow[j], 1);
    }
}

__global__ void iota2_kernel(int *a, int n)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = i;
}

__global__ void gather_kernel(const int *perm, const int *col, const double *val, int n, int *ocol, double *oval)
{
    const int stride = gridDim.x * blockDim.x;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int k = perm[j];
        ocol[j] = col[k];
        oval[j] = val[k];
    }
}

__global__ void combine_count_kernel(const int *rpa, const int *rpb, int has_diag, int nrows, int *counts)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += stride)
        counts[i] = (rpa[i + 1] - rpa[i]) + (rpb ? rpb[i + 1] - rpb[i] : 0) + (has_diag ? 1 : 0);
}

__global__ void combine_fill_kernel(const int *rpa, const int *ca, const double *va, double alpha,
                                    const int *rpb, const int *cb, const double *vb, double beta,
                                    const double *diag, double gamma, int nrows,
                                    const int *rpc, int *cc, double *vc)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += stride) {
        int o = rpc[i];
        for (int k = rpa[i]; k < rpa[i + 1]; ++k, ++o) {
            cc[o] = ca[k];
            vc[o] = __dmul_rn(alpha, va[k]);
        }
        if (rpb)
            for (int k = rpb[i]; k < rpb[i + 1]; ++k, ++o) {
                cc[o] = cb[k];
                vc[o] = __dmul_rn(beta, vb[k]);
            }
        if (diag) {
            cc[o] = i;
            vc[o] = __dmul_rn(gamma, diag[i]);
        }
    }
}

__global__ void to_dense_kernel(const int *rp, const int *col, const double *val, int nrows, int64_t ncols, double *out)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += stride)
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            double *o = out + (int64_t)i * ncols + col[k];
            *o = __dadd_rn(*o, val[k]);       // duplicates add up in row order, like the products A*e_j
        }
}

struct Scratch {                 // frees whatever was allocated, on every exit path
    static constexpr int kMax = 32;
    void *p[kMax];
    int   n = 0;
    template <class T>
    int get(T **out, size_t bytes)
    {
        *out = nullptr;
        KRY_REQUIRE(n < kMax, KRY_ERR_STATE, "assemble: scratch table full");
        void *q = nullptr;
        int rc = kry_alloc(&q, bytes ? bytes : 8);
        if (rc == KRY_OK) p[n++] = q;
        *out = (T *)q;
        return rc;
    }
    ~Scratch()
    {
        for (int i = 0; i < n; ++i) cudaFree(p[i]);
    }
};

int new_operator(kry_ctx *c, uint32_t flags, kry_csr **out)
{
    kry_csr *M = new (std::nothrow) kry_csr();
    KRY_REQUIRE(M, KRY_ERR_NOMEM, "operator: host allocation failed");
    M->ctx = c;
    M->flags = flags;
    *out = M;
    return KRY_OK;
}

int finish_operator(kry_ctx *c, kry_csr *M, int rc, kry_csr **out)
{
    if (rc == KRY_OK) {
        cudaMemsetAsync(M->A.col + M->A.nnz, 0, 8 * sizeof(int), c->stream);
        cudaMemsetAsync(M->A.val + M->A.nnz, 0, 8 * sizeof(double), c->stream);
        rc = csr_finish(c, M->A);
    }
    if (rc != KRY_OK) {
        csr_dev_free(M->A);
        delete M;
        *out = nullptr;
        return rc;
    }
    kry_ctx_retain(c);
    *out = M;
    return KRY_OK;
}

}  // namespace

// Triplets already in HBM -> CSR `m` (allocated here).  `bad` receives the number of
// out-of-range coordinates; nothing is built when it is non-zero.
static int coo_to_csr_dev(kry_ctx *c, int64_t nrows, int64_t ncols, int64_t nnz, const int *d_rows, const int *d_cols,
                          const double *d_vals, int sym, CsrDev &m, int *bad)
{
    cudaStream_t st = c->stream;
    const int n2 = (int)(2 * nnz), g = c->sm_count * 8;
    Scratch S;
    int *erow, *ecol, *keep, *pos, *crow, *ccol, *perm_in, *perm_out, *keys_out, *d_meta;
    double *eval, *cval;
    *bad = 0;
    KRY_TRY(S.get(&d_meta, 256));
    KRY_CUDA(cudaMemsetAsync(d_meta, 0, 256, st));
    KRY_TRY(S.get(&erow, (size_t)n2 * 4));
    KRY_TRY(S.get(&ecol, (size_t)n2 * 4));
    KRY_TRY(S.get(&eval, (size_t)n2 * 8));
    KRY_TRY(S.get(&keep, (size_t)n2 * 4));
    KRY_TRY(S.get(&pos, (size_t)(n2 + 1) * 4));
    KRY_TRY(kry_alloc((void **)&m.rowptr, (size_t)(nrows + 1 + 8) * sizeof(int)));
    int *counts = m.rowptr;
    KRY_CUDA(cudaMemsetAsync(counts, 0, (size_t)(nrows + 1) * sizeof(int), st));
    int kept = 0;
    if (nnz > 0) {
        coo_expand_kernel<<<g, 256, 0, st>>>(d_rows, d_cols, d_vals, (int)nnz, sym, (int)nrows, (int)ncols,
                                             erow, ecol, eval, keep, d_meta);
        c->launches++;
        size_t b1 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, b1, keep, pos, n2, st);
        void *tmp = nullptr;
        KRY_TRY(S.get(&tmp, b1 + 256));
        cub::DeviceScan::ExclusiveSum(tmp, b1, keep, pos, n2, st);
        int last_pos = 0, last_keep = 0;
        KRY_CUDA(cudaMemcpyAsync(bad, d_meta, sizeof(int), cudaMemcpyDeviceToHost, st));
        KRY_CUDA(cudaMemcpyAsync(&last_pos, pos + n2 - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        KRY_CUDA(cudaMemcpyAsync(&last_keep, keep + n2 - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        KRY_CUDA(cudaStreamSynchronize(st));
        if (*bad) return KRY_OK;
        kept = last_pos + last_keep;
    }
    m.nrows = nrows;
    m.ncols = ncols;
    m.nnz = kept;
    KRY_TRY(kry_alloc((void **)&m.col, (size_t)(kept + 8) * sizeof(int)));
    KRY_TRY(kry_alloc((void **)&m.val, (size_t)(kept + 8) * sizeof(double)));
    if (kept > 0) {
        KRY_TRY(S.get(&crow, (size_t)kept * 4));
        KRY_TRY(S.get(&ccol, (size_t)kept * 4));
        KRY_TRY(S.get(&cval, (size_t)kept * 8));
        KRY_TRY(S.get(&perm_in, (size_t)kept * 4));
        KRY_TRY(S.get(&perm_out, (size_t)kept * 4));
        KRY_TRY(S.get(&keys_out, (size_t)kept * 4));
        compact_kernel<<<g, 256, 0, st>>>(erow, ecol, eval, keep, pos, n2, crow, ccol, cval, counts);
        iota2_kernel<<<g, 256, 0, st>>>(perm_in, kept);
        c->launches += 2;
        // stable sort by row: inside a row the arrival order (slot index) survives
        size_t b2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, b2, crow, keys_out, perm_in, perm_out, kept, 0, 32, st);
        void *tmp2 = nullptr;
        KRY_TRY(S.get(&tmp2, b2 + 256));
        cub::DeviceRadixSort::SortPairs(tmp2, b2, crow, keys_out, perm_in, perm_out, kept, 0, 32, st);
        gather_kernel<<<g, 256, 0, st>>>(perm_out, ccol, cval, kept, m.col, m.val);
        c->launches++;
    }
    size_t b3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b3, counts, counts, (int)nrows + 1, st);
    void *tmp3 = nullptr;
    KRY_TRY(S.get(&tmp3, b3 + 256));
    cub::DeviceScan::ExclusiveSum(tmp3, b3, counts, counts, (int)nrows + 1, st);
    KRY_CUDA(cudaMemsetAsync(m.col + kept, 0, 8 * sizeof(int), st));
    KRY_CUDA(cudaMemsetAsync(m.val + kept, 0, 8 * sizeof(double), st));
    KRY_CUDA(cudaStreamSynchronize(st));
    KRY_CUDA(cudaGetLastError());
    return csr_finish(c, m);
}

// flags: KRY_CSR_SYMMETRIC -- the triplets are one triangle of a symmetric operator (expanded here);
// KRY_CSR_BUILD_TRANSPOSE -- also assemble A^T from the same triplets with the roles of rows and
// columns swapped, i.e. in the accumulation order of the reference's matvec_transp (linop.py:666-681).
extern "C" int kry_csr_create_coo(kry_ctx *c, int64_t nrows, int64_t ncols, int64_t nnz, const int32_t *rows,
                                  const int32_t *cols, const double *vals, uint32_t flags, kry_csr **out)
{
    KRY_REQUIRE(c && out && (nnz == 0 || (rows && cols && vals)), KRY_ERR_INVALID, "kry_csr_create_coo: NULL argument");
    *out = nullptr;
    KRY_REQUIRE(!c->closed, KRY_ERR_STATE, "kry_csr_create_coo: the context was destroyed");
    const int sym = (flags & KRY_CSR_SYMMETRIC) ? 1 : 0;
    KRY_TRY(check_sizes(nrows, ncols, 2 * nnz));
    KRY_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    Scratch S;
    int *d_rows, *d_cols;
    double *d_vals;
    KRY_TRY(S.get(&d_rows, (size_t)nnz * 4));
    KRY_TRY(S.get(&d_cols, (size_t)nnz * 4));
    KRY_TRY(S.get(&d_vals, (size_t)nnz * 8));
    if (nnz > 0) {
        KRY_CUDA(cudaMemcpyAsync(d_rows, rows, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
        KRY_CUDA(cudaMemcpyAsync(d_cols, cols, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
        KRY_CUDA(cudaMemcpyAsync(d_vals, vals, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    }
    kry_csr *M = nullptr;
    KRY_TRY(new_operator(c, flags & KRY_CSR_SYMMETRIC, &M));
    int bad = 0;
    int rc = coo_to_csr_dev(c, nrows, ncols, nnz, d_rows, d_cols, d_vals, sym, M->A, &bad);
    if (rc == KRY_OK && bad) {
        kry_set_error("kry_csr_create_coo: %d coordinate(s) outside a (%lld, %lld) operator", bad, (long long)nrows,
                      (long long)ncols);
        rc = KRY_ERR_INVALID;
    }
    if (rc == KRY_OK && (flags & KRY_CSR_BUILD_TRANSPOSE) && !sym) {
        rc = coo_to_csr_dev(c, ncols, nrows, nnz, d_cols, d_rows, d_vals, 0, M->T, &bad);
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

extern "C" int kry_csr_combine(kry_ctx *c, const kry_csr *A, double alpha, const kry_csr *B, double beta,
                               const double *diag_host, double gamma, uint32_t flags, kry_csr **out)
{
    KRY_REQUIRE(c && A && out, KRY_ERR_INVALID, "kry_csr_combine: NULL argument");
    *out = nullptr;
    KRY_CTX_LIVE(c, "kry_csr_combine");
    KRY_REQUIRE(A->ctx == c && (!B || B->ctx == c), KRY_ERR_INVALID, "kry_csr_combine: operators of another context");
    KRY_REQUIRE(!A->halo.active && (!B || !B->halo.active), KRY_ERR_UNSUPPORTED,
                "kry_csr_combine: row shards are combined before kry_csr_shard_finalize");
    const int64_t nrows = A->A.nrows, ncols = A->A.ncols;
    KRY_REQUIRE(!B || (B->A.nrows == nrows && B->A.ncols == ncols), KRY_ERR_SHAPE,
                "kry_csr_combine: %lld x %lld and %lld x %lld", (long long)nrows, (long long)ncols,
                B ? (long long)B->A.nrows : 0LL, B ? (long long)B->A.ncols : 0LL);
    KRY_REQUIRE(!diag_host || nrows == ncols, KRY_ERR_SHAPE, "kry_csr_combine: diagonal term on a rectangular operator");
    const int64_t nnz = A->A.nnz + (B ? B->A.nnz : 0) + (diag_host ? nrows : 0);
    KRY_TRY(check_sizes(nrows, ncols, nnz));
    KRY_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int g = c->sm_count * 8;
    Scratch S;
    double *d_diag = nullptr;
    if (diag_host) {
        KRY_TRY(S.get(&d_diag, (size_t)nrows * 8));
        KRY_CUDA(cudaMemcpyAsync(d_diag, diag_host, (size_t)nrows * 8, cudaMemcpyHostToDevice, st));
    }
    kry_csr *M = nullptr;
    KRY_TRY(new_operator(c, flags & KRY_CSR_SYMMETRIC, &M));
    int rc = csr_dev_alloc(M->A, nrows, ncols, nnz);
    if (rc == KRY_OK) {
        cudaMemsetAsync(M->A.rowptr, 0, (size_t)(nrows + 1) * sizeof(int), st);
        combine_count_kernel<<<g, 256, 0, st>>>(A->A.rowptr, B ? B->A.rowptr : nullptr, diag_host != nullptr, (int)nrows,
                                                M->A.rowptr);
        size_t b = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, b, M->A.rowptr, M->A.rowptr, (int)nrows + 1, st);
        void *tmp = nullptr;
        rc = S.get(&tmp, b + 256);
        if (rc == KRY_OK) {
            cub::DeviceScan::ExclusiveSum(tmp, b, M->A.rowptr, M->A.rowptr, (int)nrows + 1, st);
            combine_fill_kernel<<<g, 256, 0, st>>>(A->A.rowptr, A->A.col, A->A.val, alpha,
                                                   B ? B->A.rowptr : nullptr, B ? B->A.col : nullptr, B ? B->A.val : nullptr,
                                                   beta, d_diag, gamma, (int)nrows, M->A.rowptr, M->A.col, M->A.val);
            c->launches += 2;
            cudaError_t e = cudaStreamSynchronize(st);
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e != cudaSuccess) { kry_set_error("kry_csr_combine: %s", cudaGetErrorString(e)); rc = KRY_ERR_CUDA; }
        }
    }
    return finish_operator(c, M, rc, out);
}

extern "C" int kry_csr_to_dense(const kry_csr *A, double *dense_host)
{
    KRY_REQUIRE(A && dense_host, KRY_ERR_INVALID, "kry_csr_to_dense: NULL argument");
    KRY_CTX_LIVE(A->ctx, "kry_csr_to_dense");
    KRY_REQUIRE(!A->halo.active, KRY_ERR_UNSUPPORTED, "kry_csr_to_dense: not for row shards");
    kry_ctx *c = A->ctx;
    const int64_t nrows = A->A.nrows, ncols = A->A.ncols;
    KRY_REQUIRE(nrows * ncols <= ((int64_t)1 << 31), KRY_ERR_UNSUPPORTED, "kry_csr_to_dense: %lld x %lld is too large",
                (long long)nrows, (long long)ncols);
    KRY_CUDA(cudaSetDevice(c->device));
    Scratch S;
    double *d = nullptr;
    const size_t bytes = (size_t)(nrows * ncols) * sizeof(double);
    KRY_TRY(S.get(&d, bytes));
    KRY_CUDA(cudaMemsetAsync(d, 0, bytes ? bytes : 8, c->stream));
    if (nrows > 0 && A->A.nnz > 0) {
        to_dense_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(A->A.rowptr, A->A.col, A->A.val, (int)nrows, ncols, d);
        c->launches++;
    }
    if (bytes) KRY_CUDA(cudaMemcpyAsync(dense_host, d, bytes, cudaMemcpyDeviceToHost, c->stream));
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    KRY_CUDA(cudaGetLastError());
    return KRY_OK;
}
