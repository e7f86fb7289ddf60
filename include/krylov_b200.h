/* krylov_b200.h -- C ABI of the B200-native Krylov iteration engine.
 *
 * This is the drop-in boundary for the per-iteration hot path of
 * PythonOptimizers/pykrylov.  The reference has no FFI of
 * its own (it is pure Python + NumPy); each entry point below therefore cites
 * the *reference code whose work it absorbs* (paths relative to
 * /root/reference/pykrylov/).  The Python host side (pykrylov_b200/) binds
 * these symbols with ctypes and nothing else; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain C: opaque handles, pointers and sizes; no C++/torch types.
 *   - every function returns 0 on success and a negative kry_status on error;
 *     kry_last_error() returns the message of the last failure on this thread.
 *   - no exceptions and no callbacks cross the ABI.
 *   - the caller owns every host buffer; the library copies on upload and never
 *     retains a host pointer after the call returns.
 *   - one CUDA stream per context; a context is not thread-safe.
 *   - all floating point data is IEEE fp64; all CSR indices are int32
 *     (what scipy.sparse produces for the configs of BASELINE.json).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with KRY_ERR_CUDA.
 */
#ifndef KRYLOV_B200_H
#define KRYLOV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KRY_ABI_VERSION 1

typedef enum kry_status {
    KRY_OK              =  0,
    KRY_ERR_INVALID     = -1,   /* bad argument (NULL handle, negative size, ...)     */
    KRY_ERR_SHAPE       = -2,   /* operand sizes do not match (linop.py:283-296)       */
    KRY_ERR_CUDA        = -3,   /* CUDA runtime error / no device                      */
    KRY_ERR_NOMEM       = -4,   /* device allocation failed                            */
    KRY_ERR_UNSUPPORTED = -5,   /* e.g. nnz >= 2^31                                    */
    KRY_ERR_COMM        = -6,   /* NCCL error / communicator not initialised           */
    KRY_ERR_STATE       = -7    /* call out of order (iterate before setup, ...)       */
} kry_status;

typedef struct kry_ctx    kry_ctx;     /* device + stream + reduction workspace      */
typedef struct kry_csr    kry_csr;     /* device-resident CSR operator               */
typedef struct kry_vec    kry_vec;     /* device-resident fp64 vector                */
typedef struct kry_solver kry_solver;  /* device-resident Krylov iteration state     */
typedef struct kry_lls kry_lls;        /* device-resident scalar plane of the lls solvers / SYMMLQ */
typedef struct kry_graph kry_graph;    /* a captured sequence of stand-alone launches             */

/* ------------------------------------------------------------------ misc */
int         kry_abi_version(void);
const char *kry_last_error(void);
int         kry_device_count(int *count);

/* ---------------------------------------------------------------- context */
int kry_ctx_create(int device, kry_ctx **out);
int kry_ctx_destroy(kry_ctx *ctx);
int kry_ctx_sync(kry_ctx *ctx);                       /* cudaStreamSynchronize        */
/* props[0]=SM count, [1]=total bytes, [2]=free bytes, [3]=cc major*10+minor,
 * [4]=L2 bytes, [5]=max smem/block optin                                            */
int kry_ctx_props(kry_ctx *ctx, int64_t props[6]);
/* CUDA-event timing on the context's stream (bench.py measures with these).         */
int kry_timer_start(kry_ctx *ctx);
int kry_timer_stop(kry_ctx *ctx, double *elapsed_ms); /* synchronises                 */
/* Overwrite a scratch buffer larger than L2 (timing hygiene between samples).       */
int kry_flush_l2(kry_ctx *ctx);
/* Number of kernels this library has launched on this context so far.               */
int kry_launch_count(kry_ctx *ctx, int64_t *count);
/* Diagnostics of the halo exchange fused into the sharded SpMV launch: with KRY_HALO_TRACE set in
 * the environment when the context is created, the kernels accumulate %globaltimer intervals (ns):
 * out[0] sum over launches of (flags published - kernel entry), out[1] launches, out[2] sum of the
 * spin time and out[3] of the fence time over waiting warps, out[4] waiting warps, out[5] max
 * (end of wait - kernel entry), out[6] sum (start of wait - kernel entry).  Reading resets them. */
int kry_halo_trace_read(kry_ctx *ctx, uint64_t *out16);
/* Per-kernel device timing of the dominant kernel (the fused SpMV+dot of the solver
 * loops): when enabled, every such launch is bracketed by a CUDA event pair on the
 * context's stream (up to max_samples launches; 0 disables).  kry_prof_read
 * synchronises and returns the number of completed samples and their summed
 * duration -- this is what bench.py's roofline.achieved is computed from.          */
int kry_prof_enable(kry_ctx *ctx, int max_samples);
int kry_prof_read(kry_ctx *ctx, int64_t *samples, double *total_ms);

/* Engine options (A/B switches for measurement). */
#define KRY_OPT_L2_HINTS 1   /* L2 eviction-priority hints in the CG kernels (default 1) */
#define KRY_OPT_P2P      3   /* sharded runs: all-reduce the fused inner products through NVLink
                                peer memory inside the kernel instead of ncclAllReduce + a
                                finalize launch (default 1 when CUDA IPC mapping succeeded;
                                must be set identically on every rank)                       */
#define KRY_OPT_GRAPHS   2   /* replay the solver loops as CUDA graphs of 12 iterations (default 1) */
#define KRY_OPT_CG_FUSE  4   /* CG launch plan on unsharded operators (latched at kry_solver_setup):
                                0: 3 launches/iteration  SpMV+dot | x,r update+dot | p update
                                1: 2 launches  [p = beta p - r]+SpMV+dot | x,r update+dot
                                2: 2 launches  [x += alpha p ; p = beta p - r]+SpMV+dot | r update+dot
                                The fused forms carry the p (and x) update of a trip into the SpMV of
                                the next one; results are bit-identical to form 0.                  */
#define KRY_OPT_CG_ONE_CTA 6  /* 1 (default): CG on an unsharded operator whose CSR and vectors fit the shared
                                memory of one SM runs as ONE persistent CTA -- the whole loop inside a single
                                launch per kry_solver_iterate call (candidate; latched at kry_solver_setup) */
#define KRY_OPT_MINRES_FUSE 7 /* 1 (default): MINRES runs 2 launches per iteration -- the w / x update of a trip
                                (minres.py:294-297, no reduction in it) rides in the second launch of the next
                                trip; 0: 3 launches.  Candidate; latched at kry_solver_setup.               */
#define KRY_OPT_MINRES_PERSISTENT 8 /* 1 (default 0: measured no faster than the 2-launch plan): unsharded, unpreconditioned MINRES runs as ONE cooperative
                                persistent kernel per kry_solver_iterate call: one CTA wave, the three phases of
                                a trip separated by two grid-wide barriers that carry the reductions -- no launch
                                between trips.  Candidate; latched at kry_solver_setup.                        */
#define KRY_OPT_CG_FUSE_SHARDS 5 /* 1 (default): row shards use the CG_FUSE plan too (the packed halo then
                                carries beta p - r of the boundary entries); 0: shards keep plan 0 */
#define KRY_OPT_HALO_P2P 9     /* sharded runs with KRY_OPT_P2P on: the SpMV launch itself writes this rank's
                                boundary entries into the peers' halo tails over NVLink peer memory (CUDA IPC)
                                and waits for the peers' flags only when it reaches its boundary rows, which it
                                processes last -- no pack launch, no ncclAllGather.  0: pack kernel + one
                                ncclAllGather ahead of the SpMV.  Must be set identically on every rank.       */
int kry_ctx_set_option(kry_ctx *ctx, int option, int value);
int kry_ctx_get_option(kry_ctx *ctx, int option, int *value);

/* Pinned host staging memory (for the end-to-end H2D/D2H legs). */
int kry_host_alloc(int64_t bytes, void **out);
int kry_host_free(void *p);

/* ---------------------------------------------------------------- vectors */
int kry_vec_create(kry_ctx *ctx, int64_t n, kry_vec **out);
/* Same with spare capacity behind the n logical entries: the input vector of a
 * row-sharded operator carries the gathered halo there (kry_csr_shape's ncols). */
int kry_vec_create_cap(kry_ctx *ctx, int64_t n, int64_t capacity, kry_vec **out);
int kry_vec_destroy(kry_vec *v);
int kry_vec_size(const kry_vec *v, int64_t *n);
int kry_vec_upload(kry_vec *v, const double *host, int64_t n);
int kry_vec_download(const kry_vec *v, double *host, int64_t n);
/* Read `count` entries starting at `offset` (e.g. x[0] for a log line). */
int kry_vec_read(const kry_vec *v, int64_t offset, int64_t count, double *host);
int kry_vec_fill(kry_vec *v, double value);
int kry_vec_copy(kry_vec *dst, const kry_vec *src);

/* -------------------------------------------------------------- operators */
#define KRY_CSR_SYMMETRIC        1u   /* A^T == A: op.T is op (linop.py:148-152)      */
#define KRY_CSR_BUILD_TRANSPOSE  2u   /* build the CSR of A^T on device at creation   */

/* Device CSR operator: replaces the user matvec closure behind
 * LinearOperator.__mul__ (linop/linop.py:362-369 -> :356-360 -> :271-298), i.e.
 * PysparseLinearOperator's `A*x` / `y*A` (linop.py:697-717).  Column indices
 * must be in [0, ncols); rows need not be sorted (the summation order of a row
 * is its storage order -- sorted columns reproduce scipy's csr_matvec
 * bit-for-bit).  ncols_ext >= ncols may be passed through kry_csr_create_ext for
 * a row shard whose columns address [local | halo] (section 8e).                */
int kry_csr_create(kry_ctx *ctx, int64_t nrows, int64_t ncols, int64_t nnz,
                   const int32_t *rowptr, const int32_t *col, const double *val,
                   uint32_t flags, kry_csr **out);
/* Operator from coordinate triplets, assembled on the device (replaces the Python loop of the
 * reference's CoordLinearOperator, linop/linop.py:638-685).  Inside every CSR row the entries keep
 * the order in which that loop accumulates them, so SpMV row sums are its sums bit for bit.
 * flags: KRY_CSR_SYMMETRIC = the triplets are one triangle (the mirror images are generated);
 * KRY_CSR_BUILD_TRANSPOSE = also assemble A^T from the triplets in the accumulation order of the
 * reference's matvec_transp (linop.py:666-681).  Out-of-range coordinates -> KRY_ERR_INVALID.   */
int kry_csr_create_coo(kry_ctx *ctx, int64_t nrows, int64_t ncols, int64_t nnz,
                       const int32_t *rows, const int32_t *cols, const double *vals,
                       uint32_t flags, kry_csr **out);
/* Operator algebra that stays in HBM (reference linop.py:307-345, 378-410 builds host closures):
 * C = alpha*A [+ beta*B] [+ gamma*diag(d)] as a new CSR; per row the scaled entries of A, then of B,
 * then the diagonal entry.  B and diag_host may be NULL.  A + sigma*I and A +- D equal the
 * reference's two-product expression bit for bit; alpha*A and A + B to rounding (one fused row sum
 * instead of two products).  flags: KRY_CSR_SYMMETRIC if the result is symmetric.                */
int kry_csr_combine(kry_ctx *ctx, const kry_csr *A, double alpha, const kry_csr *B, double beta,
                    const double *diag_host, double gamma, uint32_t flags, kry_csr **out);
/* Dense row-major copy (LinearOperator.to_array, linop.py:256-269), scattered on the device. */
int kry_csr_to_dense(const kry_csr *A, double *dense_host);
int kry_csr_destroy(kry_csr *A);
int kry_csr_shape(const kry_csr *A, int64_t *nrows, int64_t *ncols, int64_t *nnz);
/* Copy the device CSR back (integer-parity tests: device-built == scipy).           */
int kry_csr_download(const kry_csr *A, int transposed,
                     int32_t *rowptr, int32_t *col, double *val);
int kry_csr_build_transpose(kry_csr *A);
/* Main diagonal (PysparseMatrix.takeDiagonal, examples/bmark.py:19).                */
int kry_csr_diagonal(const kry_csr *A, double *diag_host);

/* Device-side gallery: CSR of the reference's matrix-free stencils, rows
 * [row_begin,row_end) with *global* column ids, generated directly in HBM.
 *   poisson2d : gallery/gallery.py:10-29 on a g x g grid (diag 4, off -1).
 *   poisson1d : gallery/gallery.py:3-8 (diag 2, off -1).
 *   convdiff3d: 7-pt convection-diffusion of BASELINE.json config 4 (not in the
 *               reference): diag 6+3*gamma, upstream -1-gamma, downstream -1.    */
int kry_csr_create_poisson1d(kry_ctx *ctx, int64_t n, int64_t row_begin, int64_t row_end,
                             uint32_t flags, kry_csr **out);
int kry_csr_create_poisson2d(kry_ctx *ctx, int64_t g, int64_t row_begin, int64_t row_end,
                             uint32_t flags, kry_csr **out);
int kry_csr_create_convdiff3d(kry_ctx *ctx, int64_t m, double gamma,
                              int64_t row_begin, int64_t row_end,
                              uint32_t flags, kry_csr **out);

/* SpMV kernel selection (tuning / A-B measurement; default KRY_SPMV_AUTO). */
#define KRY_SPMV_AUTO    0
#define KRY_SPMV_ROW     1   /* one thread per row, direct global loads              */
#define KRY_SPMV_STREAM  2   /* coalesced nnz stream staged in smem, row-sum pass    */
#define KRY_SPMV_TMA     3   /* persistent CTAs, cp.async.bulk (TMA) multi-stage     */
#define KRY_SPMV_ROWB8   4   /* one thread per row, loads batched 8 entries at a time */
#define KRY_SPMV_ROWB4   5   /* one thread per row, loads batched 4 entries at a time */
#define KRY_SPMV_ROWPF   6   /* one thread per row, row pointers loaded one trip ahead (candidate)  */
#define KRY_SPMV_ROWPF2  7   /* one thread per row; the TMA unit stages the row pointers of the next tiles in
                                shared memory (cp.async.bulk + mbarrier, two trips ahead) and bulk-prefetches
                                the next tile's col/val windows into L2 -- no register cost (spmv_rowtma_kernel) */
#define KRY_SPMV_ROWTMA  KRY_SPMV_ROWPF2
int kry_csr_set_kernel(kry_csr *A, int kind, int tile_nnz, int threads);

/* ------------------------------------------------------- hot-path kernels */
/* y = A x (trans=0) or y = A^T x (trans=1).  Replaces `self.op * p`
 * (cg/cg.py:115, bicgstab/bicgstab.py:101,125, minres/minres.py:239,
 *  cgs/cgs.py:84,96, tfqmr/tfqmr.py:84,114,146) and `A.T * u`
 * (lls/lsqr.py:200,264, lls/lsmr.py:224,322, lls/craig.py:224,321).              */
int kry_spmv(kry_csr *A, int trans, const kry_vec *x, kry_vec *y);

/* y = A x fused with up to 3 row-local inner products in the same launch:
 * dot k = sum_i w_k[i] * y[i], with w_k = dot_with[k], or y itself when
 * dot_with[k] == NULL.  Results land in the context's scalar slots
 * [slot0, slot0+n_dots).  Replaces `Ap = op*p; pAp = dot(p,Ap)` (cg.py:115-117),
 * `v = A q; dot(r0,v)` (bicgstab.py:101-103), `t = A z; dot(t,s), dot(t,t),
 * dot(r0,t)` (bicgstab.py:125-127), `y = A v; alfa = dot(v,y)` (minres.py:239-245). */
int kry_spmv_dot(kry_csr *A, int trans, const kry_vec *x, kry_vec *y,
                 int n_dots, const kry_vec *const *dot_with, int slot0);

/* One pass over up to 4 fused updates  z_k <- a_k*u_k + b_k*w_k  (executed in
 * order per element, un-fused multiply then add exactly like the NumPy
 * expressions they replace) followed by up to 3 inner products u.w in the same
 * launch.  Coefficients are immediate doubles, or read on device from scalar
 * slot `a_slot`/`b_slot` when that slot index is >= 0 (optionally negated).
 * Replaces `x += alpha*p; r += alpha*Ap; dot(r,r)` (cg.py:130-146),
 * `p *= beta; p -= r` (cg.py:150-151) and the AXPY groups of bicgstab.py:91-93,
 * 104-107, 130-139, cgs.py:86-114, tfqmr.py:92-99,133-150, minres.py:246-251.   */
typedef struct kry_axpby {
    kry_vec       *z;        /* output (may alias u or w)                         */
    const kry_vec *u;        /* may be NULL: term a*u omitted                     */
    const kry_vec *w;        /* may be NULL: term b*w omitted                     */
    double a, b;             /* immediates, used when the slot is < 0             */
    int    a_slot, b_slot;   /* scalar-slot index or -1                           */
    int    a_neg,  b_neg;    /* bit 0: negate the coefficient; bit 1: the coefficient
                                divides its vector (`u /= beta`) instead of scaling it */
} kry_axpby;
typedef struct kry_dotspec { const kry_vec *u, *w; } kry_dotspec;
int kry_multi_axpy_dot(kry_ctx *ctx, int n_ops, const kry_axpby *ops,
                       int n_dots, const kry_dotspec *dots, int slot0);

/* SpMV with its y-side update fused in:  z = a*(A x) + b*w  (op->u must be NULL: the product takes
 * its place; coefficient rules of kry_axpby), and for n_dots == 1 the inner product
 * dot_with . z (dot_with == NULL: z . z) into scalar slot slot0.  z may alias w, not x.  Replaces
 * kry_spmv into a temporary followed by kry_multi_axpy_dot (lls/lsqr.py:279-281,297-299: u = A v - alpha u,
 * v = A'u - beta v): no bit of z changes, 16 bytes per row less traffic, one launch less.          */
int kry_spmv_axpby_dot(kry_csr *A, int trans, const kry_vec *x, const kry_axpby *op, int n_dots,
                       const kry_vec *dot_with, int slot0);

/* The only per-check-interval device->host traffic: read scalar slots. */
#define KRY_NUM_SLOTS 64
int kry_scalars_read(kry_ctx *ctx, int first, int count, double *host);
int kry_scalars_write(kry_ctx *ctx, int first, int count, const double *host);

/* -------------------------------------------- device-resident iterations */
typedef enum kry_method {
    KRY_CG       = 1,   /* cg/cg.py:113-158                                         */
    KRY_BICGSTAB = 2,   /* bicgstab/bicgstab.py:85-145                              */
    KRY_CGS      = 3,   /* cgs/cgs.py:76-117                                        */
    KRY_TFQMR    = 4,   /* tfqmr/tfqmr.py:85-153                                    */
    KRY_MINRES   = 5    /* minres/minres.py:218-383                                 */
} kry_method;

typedef struct kry_solver_params {
    double  abstol;          /* generic/generic.py:74  (default 1e-8)               */
    double  reltol;          /* generic/generic.py:75  (default 1e-6)               */
    int64_t matvec_max;      /* solve kwarg, default 2n (cg.py:82); MINRES: itnlim  */
    int32_t check_curvature; /* CG only (cg.py:119-124)                             */
    int32_t guess_supplied;  /* whether `guess` was passed (matvec counting quirk)  */
    double  shift;           /* MINRES (minres.py:122)                              */
    double  rtol;            /* MINRES (minres.py:126)                              */
    double  etol;            /* MINRES (minres.py:127)                              */
    int32_t window;          /* MINRES (minres.py:130), <= 16                       */
    int32_t reserved;
} kry_solver_params;

/* Snapshot of the device scalar block (one small D2H per check interval).
 * Field use per method is documented in DESIGN.md.                                  */
typedef struct kry_solver_status {
    int32_t done;            /* the loop condition of the reference turned false    */
    int32_t converged;
    int32_t definite;        /* CG: 0 after non-positive curvature (cg.py:119-124)  */
    int32_t istop;           /* MINRES (minres.py:349-361)                          */
    int64_t n_matvec;        /* reference nMatvec counter                           */
    int64_t n_iter;          /* iterations executed on device                       */
    int64_t hist_count;      /* residual-history entries produced so far            */
    double  resid_norm0;
    double  resid_norm;
    double  threshold;
    double  aux[16];         /* method specific (CG: ry,pAp,alpha,beta; MINRES:
                                Anorm,Acond,ynorm,Arnorm,beta1,...)                 */
} kry_solver_status;

int kry_solver_create(kry_ctx *ctx, kry_method method, kry_csr *A, kry_solver **out);
int kry_solver_destroy(kry_solver *S);
/* Optional diagonal preconditioner applied on device.
 *   mode 1: y = d .* r   (precon = DiagonalOperator(d), linop.py:473-503)
 *   mode 2: y = r ./ d   (examples/bmark.py:14-22 DiagonalPrec, doc/source/bmark.rst:91-93)
 * NULL or mode 0 clears it.                                                         */
int kry_solver_set_precon_diag(kry_solver *S, const double *diag_host, int mode);
/* Pre-loop part of solve(): initial residual, residNorm0, threshold, p/r0/...
 * (cg.py:61-111, bicgstab.py:52-83, cgs.py:49-74, tfqmr.py:48-85,
 * minres.py:132-209).  `guess` may be NULL.                                        */
int kry_solver_setup(kry_solver *S, const double *rhs_host, const double *guess_host,
                     const kry_solver_params *params);
/* Same, with rhs/guess already resident in HBM (bench `value` leg, shards).        */
int kry_solver_setup_dev(kry_solver *S, const kry_vec *rhs, const kry_vec *guess,
                         const kry_solver_params *params);
/* Enqueue up to n_iters iterations on the context's stream.  No host
 * synchronisation: the stopping tests run on device with the reference's
 * formulas and latch `done`; launches after that are no-ops.                       */
int kry_solver_iterate(kry_solver *S, int64_t n_iters);
int kry_solver_status_read(kry_solver *S, kry_solver_status *out);   /* synchronises */
/* Pipelined form of the same read (candidate): `enqueue` queues a copy of the device status
 * block into pinned host memory (slot 0 or 1) behind everything enqueued so far and returns at
 * once; `wait` blocks until that copy has landed -- not until later work has finished -- so the
 * host can keep one chunk of iterations in flight while it replays the history of the previous
 * one.  kry_solver_history_nowait reads history entries on a separate copy stream for the same
 * reason (only entries counted by a status that has been waited for are final).             */
int kry_solver_status_enqueue(kry_solver *S, int slot);
int kry_solver_status_wait(kry_solver *S, int slot, kry_solver_status *out);
int kry_solver_history_nowait(kry_solver *S, int64_t first, int64_t count, double *host,
                              int32_t *width);
/* Per-iteration scalars recorded on device (residHistory replay, log lines):
 * `width` doubles per entry (CG: residNorm,pAp; others: residNorm).                */
int kry_solver_history(kry_solver *S, int64_t first, int64_t count, double *host,
                       int32_t *width);
int kry_solver_solution(kry_solver *S, double *x_host);
/* Named state vectors for single-step parity tests.      */
int kry_solver_get_vector(kry_solver *S, const char *name, double *host);
int kry_solver_set_vector(kry_solver *S, const char *name, const double *host);
int kry_solver_set_scalar(kry_solver *S, const char *name, double value);
int kry_solver_get_scalar(kry_solver *S, const char *name, double *value);

/* ------------------------------------------------------------- multi-GPU */
/* One process per GPU; rank 0 creates the id, the host side broadcasts the
 * 128 bytes out of band (pykrylov_b200/comm.py), every rank then joins.            */
#define KRY_COMM_ID_BYTES 128
int kry_comm_unique_id(void *id128);
int kry_comm_init(kry_ctx *ctx, int nranks, int rank, const void *id128);
int kry_comm_destroy(kry_ctx *ctx);
int kry_comm_size(kry_ctx *ctx, int *nranks, int *rank);
int kry_comm_barrier(kry_ctx *ctx);
/* Host-buffer collectives for setup-time plumbing (halo analysis, timing max).     */
int kry_comm_allgather_host(kry_ctx *ctx, const void *send, void *recv, int64_t bytes_per_rank);
int kry_comm_allreduce_host(kry_ctx *ctx, double *inout, int count, int op /*0 sum,1 max*/);

/* Row shard of a square operator (section 8e): local rows [row_begin,row_end)
 * with global column ids.  The library analyses the off-shard columns,
 * exchanges the boundary sets once, remaps columns to [local | halo] and, per
 * SpMV, packs + ncclAllGathers only the boundary entries of x.                     */
int kry_csr_shard_finalize(kry_csr *A_local, int64_t n_global, int64_t row_begin);

/* ------------------------------------------------- LSQR / LSMR / CRAIG / CRAIG-MR / SYMMLQ
 * The vector work of these solvers is SpMV with A and A^T + fused multi-AXPY passes; their scalar
 * plane -- plane rotations, norm estimates, stopping tests (reference lls/lsqr.py:277-390,
 * lls/lsmr.py:337-475, lls/craig.py:314-455, lls/craigmr.py:159-215, symmlq/symmlq.py:235-355) --
 * runs on the device, one phase at a time: in the finalize of the launch that produced the inner
 * product it consumes (kry_lls_spmv_axpby_dot, kry_lls_multi_axpy_dot) or as a one-thread launch
 * of its own (kry_lls_step).  A phase reads the inner products from the context's scalar slots
 * 0..2 and writes the coefficients of the next vector launches into slots 8.. (kry_axpby::a_slot /
 * b_slot).  The host enqueues whole iterations (or replays them: kry_graph_*) and reads one status
 * block per check interval; after the reference's stopping test has fired every later launch,
 * vector kernels included, is a no-op (kry_lls_setup installs the context's gate).
 * Coefficient slots: 8 alpha, 9 u-divisor, 10/11 (A^T u, Nv) coefficients, 12 v-divisor,
 * 13..20 method specific C0..C7 (see csrc/lls.cu).                                            */
#define KRY_LLS_LSQR    0
#define KRY_LLS_LSMR    1
#define KRY_LLS_CRAIG   2
#define KRY_LLS_CRAIGMR 3
#define KRY_LLS_SYMMLQ  4
#define KRY_LLS_HIST_WIDTH 4
typedef struct kry_lls_params {
    int32_t window, istop;
    int64_t itn, nmatvec, itnlim;     /* SYMMLQ: itnlim carries matvec_max */
    double  damp, atol, btol, ctol, etol, rtol, shift, eps;
} kry_lls_params;
typedef struct kry_lls_status_t {
    int32_t done, istop;
    int64_t itn, nmatvec, hist_count;
} kry_lls_status_t;
int kry_lls_create(kry_ctx *ctx, int method, kry_lls **out);
int kry_lls_destroy(kry_lls *L);
const char *kry_lls_scalar_name(int method, int index);   /* NULL past the end */
int kry_lls_setup(kry_lls *L, const kry_lls_params *params, const double *scalars, int n_scalars);
int kry_lls_step(kry_lls *L, int phase);                   /* enqueue one scalar step (no sync) */
int kry_lls_status(kry_lls *L, kry_lls_status_t *out, double *scalars, int n_scalars);
int kry_lls_history(kry_lls *L, int64_t first, int64_t count, double *host);   /* 4 doubles per entry */
/* Fused forms: the launch's inner products go to slots 0.. and phase `phase` of the recurrence runs in
 * the launch's finalize (what kry_multi_axpy_dot / kry_spmv_axpby_dot followed by kry_lls_step do).   */
int kry_lls_multi_axpy_dot(kry_lls *L, int phase, int n_ops, const kry_axpby *ops, int n_dots,
                           const kry_dotspec *dots);
int kry_lls_spmv_axpby_dot(kry_lls *L, int phase, kry_csr *A, int trans, const kry_vec *x,
                           const kry_axpby *op, const kry_vec *dot_with);
int kry_lls_release_gate(kry_lls *L);                      /* stand-alone launches run unconditionally again */

/* A static sequence of stand-alone launches on the context's stream (one trip of an lls / SYMMLQ loop:
 * kry_spmv, kry_multi_axpy_dot, kry_lls_step ...) captured as a CUDA graph and replayed with one call
 * per trip.  Everything issued between kry_graph_begin and kry_graph_end is recorded, not executed;
 * the sequence must have run un-captured once before.                                              */
int kry_graph_begin(kry_ctx *ctx, kry_graph **out);
int kry_graph_end(kry_graph *g);
int kry_graph_launch(kry_graph *g, int times);
int kry_graph_destroy(kry_graph *g);

#ifdef __cplusplus
}
#endif
#endif /* KRYLOV_B200_H */
