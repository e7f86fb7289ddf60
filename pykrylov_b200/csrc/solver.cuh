// solver.cuh -- shared state of the device-resident Krylov iterations.
//
// Every solver keeps (a) its vectors in one HBM slab, (b) one DevScalars block
// holding the scalar recurrence, the reference's counters and the latched
// `done` flag, (c) a ring buffer of per-iteration scalars that the host drains
// at its check interval to replay residHistory / log lines.  The host never
// synchronises inside kry_solver_iterate().
#pragma once

#include "launch.cuh"

constexpr int KRY_HIST_CAP = 1 << 16;   // ring entries; the host drains more often than this

struct DevScalars {
    int       done;          // reference loop condition turned false (latched)
    int       converged;
    int       definite;      // CG curvature flag (cg/cg.py:119-124)
    int       istop;         // MINRES
    int       skip_half;     // Bi-CGSTAB early exit pending (bicgstab.py:110-113)
    int       pad0;
    long long n_matvec, n_iter, hist_count, matvec_max;
    double    resid0, resid, threshold, abstol, reltol;
    int       check_curvature, window;
    double    shift, rtol, etol;
    double    s[64];         // method specific scalars, indices below
    double    derr[16];      // MINRES truncated direct-error window
};

// method specific slots in DevScalars::s
enum {
    // CG (cg/cg.py); ALPHA/BETA are shared by all methods
    S_RY = 0, S_PAP, S_ALPHA, S_BETA,
    // Bi-CGSTAB / CGS / TFQMR
    S_RHO, S_RHO_NEXT, S_OMEGA, S_R0V, S_TS, S_TT, S_R0T, S_SIGMA,
    S_THETA, S_ETA, S_TAU, S_K, S_M, S_DCOEF,
    // MINRES (minres/minres.py:201-205 and the loop scalars)
    M_BETA1, M_BETA, M_OLDB, M_ALFA, M_DBAR, M_EPSLN, M_OLDEPS, M_DELTA, M_GBAR, M_PHIBAR,
    M_RHS1, M_RHS2, M_TNORM2, M_YNORM2, M_CS, M_SN, M_GMAX, M_GMIN, M_PHI, M_DENOM,
    M_ANORM, M_ACOND, M_YNORM, M_ARNORM, M_XNRG2, M_INVBETA, M_TEST1, M_TEST2,
    M_C_R1, M_C_R2, M_TRNC,
    // fused CG forms (KRY_OPT_CG_FUSE): what is still owed to p and x, see cg_settle()
    //   S_PSTATE 0: p is materialised in P[(n_iter+1)&1]
    //            1: p = beta * P[(n_iter-1)&1] - r is pending (normal state between trips)
    //            2: p was materialised into P[n_iter&1] by a trip that then stopped (curvature exit)
    //   S_XPEND  1: x += alpha * P[(n_iter-1)&1] is pending (form 2)
    S_PSTATE, S_XPEND,
    // fused MINRES plan (KRY_OPT_MINRES_FUSE): 1 = the w / x update of trip n_iter-1 is still owed
    M_WPEND,
    // persistent MINRES kernel: 1 = the reference left the trip before the w / x update (beta < 0)
    M_STOPNOW,
    S_COUNT
};
constexpr int KRY_NSCAL = 64;
static_assert(S_COUNT <= KRY_NSCAL, "DevScalars::s too small");

struct NamedVec {
    const char *name;
    double     *d;
    int64_t     n;
};

struct kry_solver {
    kry_ctx          *ctx;
    kry_csr          *A;
    kry_method        method;
    int64_t           n;        // local rows
    int64_t           ncap;     // length of vectors that are SpMV inputs (n + halo)
    double           *slab;
    NamedVec          vecs[16];
    int               nvecs;
    DevScalars       *ds;
    double           *hist;
    int               hist_width;
    double           *dinv;     // optional diagonal preconditioner
    int               precon_mode;  // 0 none, 1 y = d .* r, 2 y = r ./ d
    long long         rot;      // MINRES: trips enqueued (buffer rotation schedule)
    kry_solver_params params;
    bool              ready;
    bool              sharded;
    // peer-memory halo (KRY_OPT_HALO_P2P, comm.cu kry_halo_link): the peers' solver slabs are mapped
    // through CUDA IPC; halo_tbl[i] says where vector i's boundary entries land in each reader
    bool              halo_linked;
    HaloTable        *halo_tbl;                 // device, one per vector of the solver
    void             *peer_slab[KRY_MAX_RANKS];
    int64_t           slab_doubles;             // vectors + preconditioner diagonal (a magic word sits behind)
    // CUDA-graph replay of KRY_GRAPH_ITERS iterations (launch-latency bound problems)
    cudaGraphExec_t   graph_exec;
    int64_t           graph_launches;   // kernel launches inside one replay
    uint64_t          graph_key;        // solver_graph_key() the graph was captured under
    bool              warm;             // at least one iteration ran un-captured
    int               cg_fuse;          // CG launch plan latched at setup (KRY_OPT_CG_FUSE)
    bool              fresh;            // fused CG: nothing pending, p sits in the next trip's source buffer
    bool              one_cta;          // CG: the whole loop runs inside one CTA (KRY_OPT_CG_ONE_CTA)
    int               minres_fuse;      // MINRES launch plan latched at setup (KRY_OPT_MINRES_FUSE)
    bool              minres_persistent;    // ... or the cooperative persistent kernel (KRY_OPT_MINRES_PERSISTENT)
    DevScalars       *snap_host[2];     // pinned status snapshots (kry_solver_status_enqueue / _wait)
    cudaEvent_t       snap_ev[2];
    bool              snap_pending[2];
    size_t            one_cta_smem;     // its dynamic shared memory
};

constexpr int KRY_GRAPH_ITERS = 12;     // multiple of 6 = lcm of the MINRES buffer rotations

static inline double *solver_vec(kry_solver *S, const char *name)
{
    for (int i = 0; i < S->nvecs; ++i)
        if (strcmp(S->vecs[i].name, name) == 0) return S->vecs[i].d;
    return nullptr;
}

int kry_allreduce_sums(kry_ctx *c, int n);
int kry_halo_link(kry_solver *S);       // comm.cu: collective over the ranks, at kry_solver_create
void kry_halo_unlink(kry_solver *S);

// table of the gathered vector `x_dev`, or nullptr when the fused exchange does not apply
static inline const HaloTable *solver_halo_table(kry_solver *S, const double *x_dev)
{
    if (!S->halo_linked || !spmv_shard_fusable(S->A)) return nullptr;
    for (int i = 0; i < S->nvecs; ++i)
        if (S->vecs[i].d == x_dev) return S->halo_tbl + i;
    return nullptr;
}

#if defined(__CUDACC__) || defined(KRY_EMULATE)
// append one history entry (width doubles)
__device__ __forceinline__ void hist_push(DevScalars *s, double *hist, int width, double a, double b)
{
    const long long slot = s->hist_count % KRY_HIST_CAP;
    hist[slot * width] = a;
    if (width > 1) hist[slot * width + 1] = b;
    s->hist_count++;
}
#endif

// Launch wrappers that route the fused reduction through NCCL on sharded runs.
// Row shards (thread-per-row kernel): one kernel and one row order for every way the halo and
// the scalars can travel; `tbl` non-null = halo exchange fused into the launch.
template <int ND, class Gather, class Epi, class Fin>
int solver_spmv_shard(kry_solver *S, Gather g, Epi e, Fin f, const int *done, const HaloTable *tbl)
{
    if (S->ctx->p2p_on) return spmv_shard_launch<ND>(S->A, g, e, f, done, tbl, 2);     // in-kernel all-reduce
    KRY_TRY((spmv_shard_launch<ND>(S->A, g, e, f, done, nullptr, 1)));
    KRY_TRY(kry_allreduce_sums(S->ctx, ND));
    return finalize_launch(S->ctx, f, done);
}

template <int ND, class Gather, class Epi, class Fin>
int solver_spmv(kry_solver *S, Gather g, Epi e, Fin f, const int *done, double *x_dev)
{
    if (S->sharded) {
        if (spmv_shard_row_kind(S->A)) {
            // halo exchange fused into the launch (peer-memory stores + flags), or pack + ncclAllGather
            const HaloTable *tbl = x_dev ? solver_halo_table(S, x_dev) : nullptr;
            if (!tbl && x_dev) KRY_TRY(kry_halo_exchange(S->A, x_dev));     // x_dev == nullptr: the caller did it
            return solver_spmv_shard<ND>(S, HaloGather<Gather>{g, (int)S->n}, e, f, done, tbl);
        }
        if (x_dev) KRY_TRY(kry_halo_exchange(S->A, x_dev));
        if (S->ctx->p2p_on) return spmv_launch<ND>(S->A, false, g, e, f, done, 2);   // in-kernel all-reduce
        KRY_TRY((spmv_launch<ND>(S->A, false, g, e, f, done, 1)));
        KRY_TRY(kry_allreduce_sums(S->ctx, ND));
        return finalize_launch(S->ctx, f, done);
    }
    return spmv_launch<ND>(S->A, false, g, e, f, done, 0);
}

template <int ND, class Body, class Fin>
int solver_pass(kry_solver *S, Body b, Fin f, const int *done)
{
    if (S->sharded) {
        if (S->ctx->p2p_on) return vec_pass_launch<ND>(S->ctx, S->n, b, f, done, 2);
        KRY_TRY((vec_pass_launch<ND>(S->ctx, S->n, b, f, done, 1)));
        KRY_TRY(kry_allreduce_sums(S->ctx, ND));
        return finalize_launch(S->ctx, f, done);
    }
    return vec_pass_launch<ND>(S->ctx, S->n, b, f, done, 0);
}
