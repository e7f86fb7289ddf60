// solvers.cu -- device-resident Krylov iterations (C ABI: kry_solver_*).
//
// Each method restates the *recurrence* of the reference loop (same operation
// order per element, un-fused multiply/add, same scalar formulas, same
// stopping tests) but is laid out for the GPU: every pass over HBM is one
// fused kernel (SpMV + row-local epilogue + inner products, or several AXPYs +
// inner products), and the scalar step that follows an inner product runs in
// the last CTA of the same launch.  Reference line numbers are relative to
// /root/reference/pykrylov/.
#include <stdlib.h>
#include <string.h>

#include <new>

#include "solver.cuh"

static void solver_drop_graph(kry_solver *S);
static size_t cg_one_cta_bytes(int64_t n, int64_t nnz);

// ======================================================================= CG
// cg/cg.py:113-158.   3 launches per iteration:
//   K1  Ap = A p ; pAp = p.Ap ; alpha = ry/pAp (+ curvature test)       [spmv]
//   K2  x += alpha p ; r += alpha Ap ; y = M r ; ry' = r.y ; beta, residNorm, stop test
//   K3  p = beta p - r
struct CgEpiAp {
    double       *Ap;
    const double *p;
    int           hints;
    uint64_t      el;
    __device__ void init() { el = l2_policy_evict_last(); }
    __device__ void operator()(int row, double ax, double *acc) const
    {
        if (hints) st_hint(Ap + row, ax, el);                        // K2 reads it next: keep in L2
        else Ap[row] = ax;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(p[row], ax));          // cg.py:117
    }
};

struct CgFinAp {
    DevScalars *s;
    __device__ void operator()(const double *t) const
    {
        const double pAp = t[0];
        s->s[S_PAP] = pAp;
        s->n_matvec++;                                               // cg.py:116
        if (s->check_curvature && pAp <= 0.0) {                      // cg.py:119-124
            s->definite = 0;
            s->done = 1;
            return;
        }
        s->s[S_ALPHA] = s->s[S_RY] / pAp;                            // cg.py:127
    }
};

// precon_mode: 0 none, 1 y = d .* r (DiagonalOperator), 2 y = r ./ d (bmark.py DiagonalPrec)
__device__ __forceinline__ double apply_diag(const double *d, int mode, int i, double r)
{
    if (mode == 1) return __dmul_rn(d[i], r);
    if (mode == 2) return __ddiv_rn(r, d[i]);
    return r;
}

struct CgUpdateBody {
    double       *x, *r;
    const double *p, *Ap, *pd;
    int           pmode;
    DevScalars   *s;
    double        alpha;
    int           hints;
    uint64_t      ef, el;
    __device__ void init()
    {
        alpha = s->s[S_ALPHA];
        ef = l2_policy_evict_first();
        el = l2_policy_evict_last();
    }
    __device__ void operator()(int i, double *acc) const
    {
        x[i] = __dadd_rn(x[i], __dmul_rn(alpha, p[i]));              // cg.py:130
        const double rn = __dadd_rn(r[i], __dmul_rn(alpha, Ap[i]));  // cg.py:131
        r[i] = rn;
        const double y = apply_diag(pd, pmode, i, rn);               // cg.py:137-140
        acc[0] = __dadd_rn(acc[0], __dmul_rn(rn, y));                // cg.py:146
    }
    static constexpr bool kPair = true;
    __device__ void pair(int i2, double *acc) const
    {
        double2 xv, rv, pv, av;
        if (hints) {
            // x, r(old), Ap are dead after this pass: evict_first; p is re-read by K3
            xv = ld2_hint(x, i2, ef); rv = ld2_hint(r, i2, ef);
            pv = ld2_hint(p, i2, el); av = ld2_hint(Ap, i2, ef);
        } else {
            xv = ld2(x, i2); rv = ld2(r, i2);
            pv = ld2(p, i2); av = ld2(Ap, i2);
        }
        xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, pv.x));
        xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, pv.y));
        rv.x = __dadd_rn(rv.x, __dmul_rn(alpha, av.x));
        rv.y = __dadd_rn(rv.y, __dmul_rn(alpha, av.y));
        if (hints) {
            st2_hint(x, i2, xv, ef);
            st2_hint(r, i2, rv, el);                                 // K3 reads r next
        } else {
            st2(x, i2, xv);
            st2(r, i2, rv);
        }
        const double y0 = apply_diag(pd, pmode, 2 * i2, rv.x), y1 = apply_diag(pd, pmode, 2 * i2 + 1, rv.y);
        acc[0] = __dadd_rn(acc[0], __dmul_rn(rv.x, y0));
        acc[0] = __dadd_rn(acc[0], __dmul_rn(rv.y, y1));
    }
};

struct CgFinRy {
    DevScalars *s;
    double     *hist;
    int         fuse;        // fused forms: the p (and x) update of this trip is now owed
    int         late;        // 3-launch plan: K3 (p = beta p - r) still runs in the trip that meets the
                             // stopping test and latches `done` itself (CgFinDir), as cg.py:149-151 does
    __device__ void operator()(const double *t) const
    {
        const double ry_next = t[0];
        if (fuse) s->s[S_PSTATE] = 1.0;
        if (fuse == 2) s->s[S_XPEND] = 1.0;
        s->s[S_BETA] = ry_next / s->s[S_RY];                         // cg.py:149
        s->s[S_RY] = ry_next;                                        // cg.py:153
        const double resid = fabs(sqrt(ry_next));                    // cg.py:154
        s->resid = resid;
        s->n_iter++;
        hist_push(s, hist, 2, resid, s->s[S_PAP]);                   // cg.py:155-158
        if (!(resid > s->threshold && s->n_matvec < s->matvec_max)) {              // cg.py:113
            if (late) s->skip_half = 1;
            else s->done = 1;
        }
    }
};

struct CgFinDir {                 // 3-launch plan, after K3: the loop condition of cg.py:113 is re-tested here
    DevScalars *s;
    __device__ void operator()(const double *) const
    {
        if (s->skip_half) {
            s->skip_half = 0;
            s->done = 1;
        }
    }
};

struct CgDirBody {
    double       *p;
    const double *r;
    DevScalars   *s;
    double        beta;
    int           hints;
    uint64_t      ef, el;
    __device__ void init()
    {
        beta = s->s[S_BETA];
        ef = l2_policy_evict_first();
        el = l2_policy_evict_last();
    }
    __device__ void operator()(int i, double *) const
    {
        p[i] = __dsub_rn(__dmul_rn(beta, p[i]), r[i]);               // cg.py:150-151
    }
    static constexpr bool kPair = true;
    __device__ void pair(int i2, double *) const
    {
        double2 pv, rv;
        if (hints) {
            pv = ld2_hint(p, i2, ef);
            rv = ld2_hint(r, i2, ef);
        } else {
            pv = ld2(p, i2);
            rv = ld2(r, i2);
        }
        pv.x = __dsub_rn(__dmul_rn(beta, pv.x), rv.x);
        pv.y = __dsub_rn(__dmul_rn(beta, pv.y), rv.y);
        if (hints) st2_hint(p, i2, pv, el);                          // gathered by the next K1
        else st2(p, i2, pv);
    }
};

// ---- fused forms (KRY_OPT_CG_FUSE = 1, 2): 2 launches per iteration.
// The direction update of trip i-1 (cg.py:150-151) -- and in form 2 its x update
// (cg.py:130) -- ride in the SpMV launch of trip i: the gather forms
// p_i[c] = beta * p_{i-1}[c] - r_i[c] on the fly for every column it touches (same
// expression, so every copy of p_i[c] is the same bits as the stand-alone update), the
// row's own entry is written to the other p buffer (ping-pong: neighbours still gather
// the old one).  Per row this moves 96 B (form 1) / 112 B (form 2) in the SpMV launch
// and 48 / 24 B in the second one, against 80 + 48 + 24 B for the 3-launch form.
//   trip i reads P[(i+1)&1] and writes P[i&1];  PEND = false on the first trip after a
//   setup / settle (p is already materialised there: plain copy).
template <bool PEND, bool SHARD = false>
struct CgGatherDir {
    const double *p_old, *r;
    DevScalars   *s;
    double        beta;
    int           n_local;     // SHARD: columns >= n_local are halo entries, which arrive
                               // already updated (halo_pack_dir_kernel)
    __device__ void   init() { beta = s->s[S_BETA]; }
    __device__ double operator()(int c) const
    {
        if constexpr (SHARD) {
            // one weak load for local and halo columns alike: halo entries are written by peer GPUs
            // during the launch and read after the flag wait + fence (spmv.cuh, Gather::coherent)
            const double pq = p_old[c];
            if constexpr (PEND) {
                if (c >= n_local) return pq;            // arrives already updated
                return __dsub_rn(__dmul_rn(beta, pq), __ldg(r + c));               // cg.py:150-151
            } else {
                return pq;
            }
        } else {
            const double po = __ldg(p_old + c);
            if constexpr (PEND) {
                return __dsub_rn(__dmul_rn(beta, po), __ldg(r + c));               // cg.py:150-151
            } else {
                return po;
            }
        }
    }
    // what this rank publishes for its boundary entry j (spmv_row_shard_kernel): the updated
    // direction, same two rounded operations as everywhere else
    __device__ double boundary(int j) const
    {
        if constexpr (PEND) return __dsub_rn(__dmul_rn(beta, p_old[j]), r[j]);
        else return p_old[j];
    }
    __device__ double coherent(int c) const { return (*this)(c); }
};

template <bool PEND, bool XLAG>
struct CgEpiFused {
    static constexpr int kMinBlocks = 8;     // keep the row kernel at 32 registers (full occupancy)
    double       *Ap, *p_new, *x;
    const double *p_old, *r;
    DevScalars   *s;
    int           hints;      // bit 0: x is a pure stream (evict_first); bit 2: keep Ap in L2
    double        beta, alpha;
    uint64_t      ef, el;
    __device__ void init()
    {
        beta = s->s[S_BETA];
        alpha = s->s[S_ALPHA];
        ef = l2_policy_evict_first();
        el = l2_policy_evict_last();
    }
    __device__ void operator()(int row, double ax, double *acc) const
    {
        const double po = __ldg(p_old + row);
        double pn = po;
        if constexpr (PEND) {
            if constexpr (XLAG) {                                    // cg.py:130 of the previous trip
                if (hints & 1) st_hint(x + row, __dadd_rn(ld_hint(x + row, ef), __dmul_rn(alpha, po)), ef);
                else x[row] = __dadd_rn(x[row], __dmul_rn(alpha, po));
            }
            pn = __dsub_rn(__dmul_rn(beta, po), __ldg(r + row));     // cg.py:150-151 of the previous trip
        }
        p_new[row] = pn;
        if (hints & 4) st_hint(Ap + row, ax, el);                    // the second launch reads it next
        else Ap[row] = ax;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(pn, ax));               // cg.py:117
    }
};

struct CgFinApFused {
    DevScalars *s;
    __device__ void operator()(const double *t) const
    {
        s->s[S_PSTATE] = 2.0;        // p of this trip now sits in P[n_iter & 1] ...
        s->s[S_XPEND] = 0.0;         // ... and x is up to date (until the second launch runs)
        CgFinAp{s}(t);
    }
};

struct CgUpdateRBody {            // form 2, second launch: r += alpha Ap ; y = M r ; ry' = r.y
    double       *r;
    const double *Ap, *pd;
    int           pmode;
    DevScalars   *s;
    double        alpha;
    int           hints;
    uint64_t      ef, el;
    __device__ void init()
    {
        alpha = s->s[S_ALPHA];
        ef = l2_policy_evict_first();
        el = l2_policy_evict_last();
    }
    __device__ void operator()(int i, double *acc) const
    {
        const double rn = __dadd_rn(r[i], __dmul_rn(alpha, Ap[i]));  // cg.py:131
        r[i] = rn;
        const double y = apply_diag(pd, pmode, i, rn);               // cg.py:137-140
        acc[0] = __dadd_rn(acc[0], __dmul_rn(rn, y));                // cg.py:146
    }
    static constexpr bool kPair = true;
    __device__ void pair(int i2, double *acc) const
    {
        double2 rv, av;
        if (hints) {
            rv = ld2_hint(r, i2, ef);
            av = ld2_hint(Ap, i2, ef);
        } else {
            rv = ld2(r, i2);
            av = ld2(Ap, i2);
        }
        rv.x = __dadd_rn(rv.x, __dmul_rn(alpha, av.x));
        rv.y = __dadd_rn(rv.y, __dmul_rn(alpha, av.y));
        if (hints) st2_hint(r, i2, rv, el);                          // gathered by the next SpMV launch
        else st2(r, i2, rv);
        const double y0 = apply_diag(pd, pmode, 2 * i2, rv.x), y1 = apply_diag(pd, pmode, 2 * i2 + 1, rv.y);
        acc[0] = __dadd_rn(acc[0], __dmul_rn(rv.x, y0));
        acc[0] = __dadd_rn(acc[0], __dmul_rn(rv.y, y1));
    }
};

// Pay what the fused forms still owe (driven by the device flags, so it is right
// whichever launch latched `done`): afterwards x is current and p is materialised
// in P[(n_iter+1)&1], i.e. exactly the state of the 3-launch form.
struct CgSettleBody {
    double       *x, *P0, *P1;
    const double *r;
    DevScalars   *s;
    double       *base, *other;
    double        alpha, beta;
    int           pstate, xpend;
    __device__ void init()
    {
        const long long n = s->n_iter;
        pstate = (int)s->s[S_PSTATE];
        xpend = (int)s->s[S_XPEND];
        alpha = s->s[S_ALPHA];
        beta = s->s[S_BETA];
        // pstate 1: base = P[(n-1)&1] == P[(n+1)&1], updated in place
        // pstate 2: base = P[n&1] is copied to other = P[(n+1)&1]
        base = (pstate == 2) ? ((n & 1) ? P1 : P0) : (((n + 1) & 1) ? P1 : P0);
        other = (base == P0) ? P1 : P0;
    }
    __device__ void operator()(int i) const
    {
        if (pstate == 1) {
            const double po = base[i];
            if (xpend) x[i] = __dadd_rn(x[i], __dmul_rn(alpha, po));           // cg.py:130
            base[i] = __dsub_rn(__dmul_rn(beta, po), r[i]);                    // cg.py:150-151
        } else if (pstate == 2) {
            other[i] = base[i];
        }
    }
};

// setup, cg.py:85-104:  r = -rhs (+ A x);  y = M r;  ry = r.y;  p = -r
struct CgSetupEpi {               // with an initial guess: runs as the SpMV epilogue of A x
    double       *r, *p;
    const double *rhs, *pd;
    int           pmode;
    __device__ void init() {}
    __device__ void operator()(int row, double ax, double *acc) const
    {
        const double rn = __dadd_rn(-rhs[row], ax);                  // cg.py:85,87
        r[row] = rn;
        p[row] = -rn;                                                // cg.py:104
        const double y = apply_diag(pd, pmode, row, rn);
        acc[0] = __dadd_rn(acc[0], __dmul_rn(rn, y));                // cg.py:99
    }
};

struct CgSetupBody {              // zero initial guess
    double       *r, *p;
    const double *rhs, *pd;
    int           pmode;
    __device__ void init() {}
    __device__ void operator()(int i, double *acc) const
    {
        const double rn = -rhs[i];
        r[i] = rn;
        p[i] = -rn;
        const double y = apply_diag(pd, pmode, i, rn);
        acc[0] = __dadd_rn(acc[0], __dmul_rn(rn, y));
    }
};

struct CgSetupFin {
    DevScalars *s;
    double     *hist;
    int         guess;
    __device__ void operator()(const double *t) const
    {
        s->s[S_RY] = t[0];
        const double resid = fabs(sqrt(t[0]));                       // cg.py:100
        s->resid0 = s->resid = resid;
        s->threshold = fmax(s->abstol, s->reltol * resid);           // cg.py:102
        s->n_matvec = guess ? 1 : 0;                                 // cg.py:88
        hist_push(s, hist, 2, resid, nan(""));
        if (!(resid > s->threshold && s->n_matvec < s->matvec_max)) s->done = 1;
    }
};

static int cg_setup(kry_solver *S, int guess)
{
    double *x = solver_vec(S, "x"), *r = solver_vec(S, "r"), *p = solver_vec(S, "p");
    double *rhs = solver_vec(S, "rhs");
    S->cg_fuse = (S->sharded && !S->ctx->cg_fuse_shards) ? 0 : S->ctx->cg_fuse;
    S->one_cta_smem = cg_one_cta_bytes(S->n, S->A->A.nnz);
    S->one_cta = S->ctx->cg_one_cta && !S->sharded && !S->A->halo.active &&
                 S->one_cta_smem + 2048 <= (size_t)S->ctx->smem_optin;
    if (S->one_cta) S->cg_fuse = 0;     // its HBM state is that of the 3-launch plan
    S->fresh = true;
    S->rot = 0;
    CgSetupFin fin{S->ds, S->hist, guess};
    if (guess) {
        CgSetupEpi e{r, p, rhs, S->dinv, S->precon_mode};
        return solver_spmv<1>(S, GatherPlain{x}, e, fin, &S->ds->done, x);
    }
    CgSetupBody b{r, p, rhs, S->dinv, S->precon_mode};
    return solver_pass<1>(S, b, fin, &S->ds->done);
}

template <bool PEND, bool XLAG>
static int cg_fused_spmv(kry_solver *S, double *p_old, double *p_new, int opt)
{
    double *x = solver_vec(S, "x"), *r = solver_vec(S, "r"), *Ap = solver_vec(S, "Ap");
    CgEpiFused<PEND, XLAG> e{Ap, p_new, x, p_old, r, S->ds, opt & 5, 0.0, 0.0, 0, 0};
    if (S->sharded) {
        // boundary entries of p travel already updated; the local ones are updated in the gather
        CgGatherDir<PEND, true> g{p_old, r, S->ds, 0.0, (int)S->n};
        const HaloTable *tbl = solver_halo_table(S, p_old);          // non-null: exchange fused into the launch
        if (!tbl) KRY_TRY(kry_halo_exchange_dir(S->A, p_old, PEND ? r : nullptr, PEND ? &S->ds->s[S_BETA] : nullptr));
        if (spmv_shard_row_kind(S->A)) return solver_spmv_shard<1>(S, g, e, CgFinApFused{S->ds}, &S->ds->done, tbl);
        return solver_spmv<1>(S, g, e, CgFinApFused{S->ds}, &S->ds->done, nullptr);
    }
    CgGatherDir<PEND> g{p_old, r, S->ds, 0.0, 0};
    return solver_spmv<1>(S, g, e, CgFinApFused{S->ds}, &S->ds->done, p_old);
}

static int cg_iterate(kry_solver *S)
{
    double *x = solver_vec(S, "x"), *r = solver_vec(S, "r"), *p = solver_vec(S, "p");
    double *Ap = solver_vec(S, "Ap");
    const int *done = &S->ds->done;
    // option bits: 1 = vector kernels, 4 = Ap store of the SpMV epilogue (2 = CSR streams, launch.cuh)
    // measured: the vector hints pay only while one vector fits the L2 (config 2: +2 %;
    // 5e7 / 1e8 rows: -1.5 / -2.5 %), so they are switched off for larger vectors
    const int opt = ((int64_t)S->n * 8 <= S->ctx->l2_bytes) ? S->ctx->l2_hints : (S->ctx->l2_hints & ~1);
    if (S->cg_fuse) {
        // trip i = S->rot reads P[(i+1)&1] and writes P[i&1];  P[1] = "p", P[0] = "p2"
        double *P[2] = {solver_vec(S, "p2"), p};
        double *p_old = P[(S->rot + 1) & 1], *p_new = P[S->rot & 1];
        const bool pend = !S->fresh, xlag = (S->cg_fuse == 2);
        int rc;
        if (!pend) rc = cg_fused_spmv<false, false>(S, p_old, p_new, opt);
        else if (xlag) rc = cg_fused_spmv<true, true>(S, p_old, p_new, opt);
        else rc = cg_fused_spmv<true, false>(S, p_old, p_new, opt);
        KRY_TRY(rc);
        S->fresh = false;
        S->rot++;
        CgFinRy fin{S->ds, S->hist, S->cg_fuse, 0};
        if (xlag) {
            CgUpdateRBody ub{r, Ap, S->dinv, S->precon_mode, S->ds, 0.0, opt & 1, 0, 0};
            return solver_pass<1>(S, ub, fin, done);
        }
        CgUpdateBody ub{x, r, p_new, Ap, S->dinv, S->precon_mode, S->ds, 0.0, opt & 1, 0, 0};
        return solver_pass<1>(S, ub, fin, done);
    }
    KRY_TRY((solver_spmv<1>(S, GatherPlain{p}, CgEpiAp{Ap, p, (opt & 4) ? 1 : 0, 0}, CgFinAp{S->ds}, done, p)));
    CgUpdateBody ub{x, r, p, Ap, S->dinv, S->precon_mode, S->ds, 0.0, opt & 1, 0, 0};
    KRY_TRY((solver_pass<1>(S, ub, CgFinRy{S->ds, S->hist, 0, 1}, done)));
    // K3 carries no inner product; its finalize only latches `done` (every rank holds the same
    // flag, so there is nothing to all-reduce on shards)
    CgDirBody db{p, r, S->ds, 0.0, opt & 1, 0, 0};
    return vec_pass_launch<1>(S->ctx, S->n, db, CgFinDir{S->ds}, done, 0);
}

// ---- CG inside one CTA (KRY_OPT_CG_ONE_CTA): problems whose CSR and four vectors fit the
// shared memory of one SM (BASELINE config 0: 1138bus = 90 kB).  The 3-launch plan spends
// ~10 us per iteration there on launch and grid-reduction latency; here the whole loop of
// cg.py:113-158 runs in one launch out of shared memory, the phases separated by
// __syncthreads() only.  Per element and per row the arithmetic is the reference's (same
// expressions as the multi-CTA kernels), inner products are summed by a fixed block tree, the
// scalar recurrence and stopping tests are the same functors (CgFinAp, CgFinRy) run by thread 0
// on the same device scalar block, so status / history / done behave exactly as before.  State
// in HBM is that of the 3-launch plan (x current, p materialised in "p").
constexpr int KRY_ONE_CTA_THREADS = 1024;

__global__ void __launch_bounds__(KRY_ONE_CTA_THREADS, 1)
cg_one_cta_kernel(CsrView A, double *gx, double *gr, double *gp, double *gAp, const double *pd, int pmode,
                  DevScalars *s, double *hist, long long n_iters)
{
#ifdef KRY_EMULATE
    unsigned char *smem_raw = emu_dynamic_smem;          // tests/emu: the launcher sized it
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    __shared__ double s_warp[1][32];
    __shared__ double sh_scalar;
    __shared__ int    sh_done;
    const int n = A.nrows, tid = threadIdx.x, nt = blockDim.x;
    if (s->done) return;
    const int nnz = __ldg(A.rowptr + n);
    // layout: val[nnz] | x[n] | r[n] | p[n] | Ap[n] | rowptr[n+1] | col[nnz]
    double *val = reinterpret_cast<double *>(smem_raw);
    double *x = val + nnz, *r = x + n, *p = r + n, *Ap = p + n;
    int *rowptr = reinterpret_cast<int *>(Ap + n);
    int *col = rowptr + n + 1;
    for (int k = tid; k < nnz; k += nt) {
        val[k] = __ldg(A.val + k);
        col[k] = __ldg(A.col + k);
    }
    for (int i = tid; i <= n; i += nt) rowptr[i] = __ldg(A.rowptr + i);
    for (int i = tid; i < n; i += nt) {
        x[i] = gx[i];
        r[i] = gr[i];
        p[i] = gp[i];
        Ap[i] = gAp[i];
    }
    __syncthreads();

    for (long long it = 0; it < n_iters; ++it) {
        // Ap = A p ; pAp = p.Ap ; alpha, curvature test                          cg.py:115-127
        double acc[1] = {0.0};
        for (int row = tid; row < n; row += nt) {
            double sum = 0.0;
            for (int k = rowptr[row]; k < rowptr[row + 1]; ++k)
                sum = __dadd_rn(sum, __dmul_rn(val[k], p[col[k]]));
            Ap[row] = sum;
            acc[0] = __dadd_rn(acc[0], __dmul_rn(p[row], sum));
        }
        block_sum<1>(acc, s_warp);
        if (tid == 0) {
            CgFinAp{s}(acc);
            sh_scalar = s->s[S_ALPHA];
            sh_done = s->done;
        }
        __syncthreads();
        if (sh_done) break;                     // x is left un-updated (cg.py:119-124)
        const double alpha = sh_scalar;
        // x += alpha p ; r += alpha Ap ; y = M r ; ry' = r.y ; beta, residNorm, loop test   cg.py:130-158
        acc[0] = 0.0;
        for (int i = tid; i < n; i += nt) {
            x[i] = __dadd_rn(x[i], __dmul_rn(alpha, p[i]));
            const double rn = __dadd_rn(r[i], __dmul_rn(alpha, Ap[i]));
            r[i] = rn;
            acc[0] = __dadd_rn(acc[0], __dmul_rn(rn, apply_diag(pd, pmode, i, rn)));
        }
        __syncthreads();                        // s_warp reuse
        block_sum<1>(acc, s_warp);
        if (tid == 0) {
            CgFinRy{s, hist, 0, 0}(acc);
            sh_scalar = s->s[S_BETA];
            sh_done = s->done;
        }
        __syncthreads();
        const double beta = sh_scalar;
        // p = beta p - r   (the reference updates p before it re-tests the loop condition)   cg.py:150-151
        for (int i = tid; i < n; i += nt) p[i] = __dsub_rn(__dmul_rn(beta, p[i]), r[i]);
        __syncthreads();
        if (sh_done) break;
    }
    for (int i = tid; i < n; i += nt) {
        gx[i] = x[i];
        gr[i] = r[i];
        gp[i] = p[i];
        gAp[i] = Ap[i];
    }
}

static size_t cg_one_cta_bytes(int64_t n, int64_t nnz)
{
    return (size_t)nnz * 12 + (size_t)n * 32 + (size_t)(n + 1) * 4 + 16;
}

static int cg_one_cta_iterate(kry_solver *S, int64_t n_iters)
{
    kry_ctx *c = S->ctx;
#ifdef KRY_EMULATE
    // tests/emu, SIMT mode: one block of fibers plays the CTA
    KRY_REQUIRE(emu_fibers_on, KRY_ERR_UNSUPPORTED, "cg_one_cta needs the SIMT mode of the emulation");
    CsrView Ae = csr_view(S->A->A);
    double *ex = solver_vec(S, "x"), *er = solver_vec(S, "r"), *ep = solver_vec(S, "p"), *eAp = solver_vec(S, "Ap");
    emu_set_dynamic_smem(S->one_cta_smem);
    auto body = [&] { cg_one_cta_kernel(Ae, ex, er, ep, eAp, S->dinv, S->precon_mode, S->ds, S->hist, (long long)n_iters); };
    emu_launch_fibers(1, KRY_ONE_CTA_THREADS, &body, [](const void *k) { (*static_cast<const decltype(body) *>(k))(); });
    c->launches++;
    return KRY_OK;
#else
    static bool attr_set = false;
    if (!attr_set) {
        KRY_CUDA(cudaFuncSetAttribute(cg_one_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(c->smem_optin - 1024)));
        attr_set = true;
    }
    CsrView A = csr_view(S->A->A);
    cg_one_cta_kernel<<<1, KRY_ONE_CTA_THREADS, S->one_cta_smem, c->stream>>>(
        A, solver_vec(S, "x"), solver_vec(S, "r"), solver_vec(S, "p"), solver_vec(S, "Ap"), S->dinv,
        S->precon_mode, S->ds, S->hist, (long long)n_iters);
    c->launches++;
    KRY_CUDA(cudaGetLastError());
    return KRY_OK;
#endif
}

// Fused forms only: bring x and p to the state the 3-launch form would hold (see
// CgSettleBody).  Called before any vector of the solver is read or written from
// outside; the next trip then starts from a materialised p again.
static int cg_settle(kry_solver *S)
{
    if (S->method != KRY_CG || !S->cg_fuse || !S->ready) return KRY_OK;
    CgSettleBody b{solver_vec(S, "x"), solver_vec(S, "p2"), solver_vec(S, "p"), solver_vec(S, "r"), S->ds,
                   nullptr, nullptr, 0.0, 0.0, 0, 0};
    KRY_TRY(vec_map_launch(S->ctx, S->n, b, &S->ctx->never_done[0]));
    KRY_CUDA(cudaMemsetAsync(&S->ds->s[S_PSTATE], 0, 2 * sizeof(double), S->ctx->stream));
    S->fresh = true;
    S->warm = false;      // the next trip must run un-captured (it is the PEND = false variant)
    return KRY_OK;
}

// ================================================================ Bi-CGSTAB
// bicgstab/bicgstab.py:85-145.  4 launches per iteration (the p update of the
// next iteration rides in the last one):
//   KB  v = A q ; r0.v ; alpha = rho/(r0.v)                               [spmv]
//   KC  s = r - alpha v ; |s| ; early-exit test
//   KD  t = A z ; t.s, t.t, r0.t ; omega, rho', beta                      [spmv]
//   KE  r = s - omega t ; z *= omega ; x += z ; x += alpha q ; |r| ; stop test ;
//       p = beta p - beta omega v + r ; q = M p
struct BcgEpiV {
    double       *v;
    const double *r0;
    __device__ void init() {}
    __device__ void operator()(int row, double ax, double *acc) const
    {
        v[row] = ax;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(r0[row], ax));          // bicgstab.py:103
    }
};

struct BcgFinV {
    DevScalars *s;
    __device__ void operator()(const double *t) const
    {
        s->n_matvec++;                                               // :101
        s->s[S_R0V] = t[0];
        s->s[S_ALPHA] = s->s[S_RHO] / t[0];                          // :103
    }
};

struct BcgBodyS {
    double       *sv, *z;
    const double *r, *v, *pd;
    int           pmode;
    DevScalars   *s;
    double        alpha;
    __device__ void init() { alpha = s->s[S_ALPHA]; }
    __device__ void operator()(int i, double *acc) const
    {
        const double si = __dsub_rn(r[i], __dmul_rn(alpha, v[i]));   // :104
        sv[i] = si;
        if (pmode) z[i] = apply_diag(pd, pmode, i, si);              // :120-123
        acc[0] = __dadd_rn(acc[0], __dmul_rn(si, si));               // :107
    }
    static constexpr bool kPair = true;
    __device__ void pair(int i2, double *acc) const
    {
        const double2 rv = ld2(r, i2), vv = ld2(v, i2);
        double2 sn;
        sn.x = __dsub_rn(rv.x, __dmul_rn(alpha, vv.x));
        sn.y = __dsub_rn(rv.y, __dmul_rn(alpha, vv.y));
        st2(sv, i2, sn);
        if (pmode) {
            double2 zn;
            zn.x = apply_diag(pd, pmode, 2 * i2, sn.x);
            zn.y = apply_diag(pd, pmode, 2 * i2 + 1, sn.y);
            st2(z, i2, zn);
        }
        acc[0] = __dadd_rn(acc[0], __dmul_rn(sn.x, sn.x));
        acc[0] = __dadd_rn(acc[0], __dmul_rn(sn.y, sn.y));
    }
};

struct BcgFinS {
    DevScalars *s;
    double     *hist;
    __device__ void operator()(const double *t) const
    {
        const double resid = sqrt(t[0]);
        s->resid = resid;
        hist_push(s, hist, 1, resid, 0.0);                           // :109
        if (resid <= s->threshold) {                                 // :110-113
            s->skip_half = 1;      // KE applies x += alpha q and latches done
        } else if (s->n_matvec >= s->matvec_max) {                   // :115-117
            s->skip_half = 2;      // nothing more to do; KE latches done
        }
    }
};

struct BcgEpiT {
    static constexpr int kMinBlocks = 8;     // 32 registers: the row kernel needs full occupancy
    double       *t;
    const double *sv, *r0;
    __device__ void init() {}
    __device__ void operator()(int row, double ax, double *acc) const
    {
        t[row] = ax;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(ax, sv[row]));          // t.s  :126
        acc[1] = __dadd_rn(acc[1], __dmul_rn(ax, ax));               // t.t  :126
        acc[2] = __dadd_rn(acc[2], __dmul_rn(r0[row], ax));          // r0.t :127
    }
};

struct BcgFinT {
    DevScalars *s;
    __device__ void operator()(const double *t) const
    {
        s->n_matvec++;                                               // :125
        s->s[S_TS] = t[0];
        s->s[S_TT] = t[1];
        s->s[S_R0T] = t[2];
        const double omega = t[0] / t[1];                            // :126
        s->s[S_OMEGA] = omega;
        const double rho_next = -omega * t[2];                       // :127
        s->s[S_RHO_NEXT] = rho_next;
        // scalars of the next trip, bicgstab.py:87-88
        s->s[S_BETA] = rho_next / s->s[S_RHO] * s->s[S_ALPHA] / omega;
        s->s[S_RHO] = rho_next;
    }
};

struct BcgBodyX {
    static constexpr int kMinBlocks = 4;       // 10 vectors in flight: allow 64 registers
    double       *x, *r, *p, *q, *sv, *z;
    const double *t, *v, *pd;
    int           pmode;
    DevScalars   *s;
    double        alpha, omega, beta, bo;
    int           half;
    __device__ void init()
    {
        alpha = s->s[S_ALPHA];
        omega = s->s[S_OMEGA];
        beta = s->s[S_BETA];
        bo = __dmul_rn(beta, omega);
        half = s->skip_half;
    }
    __device__ void operator()(int i, double *acc) const
    {
        const double qi = pmode ? q[i] : p[i];
        if (half) {
            if (half == 1) x[i] = __dadd_rn(x[i], __dmul_rn(alpha, qi));     // :111
            return;
        }
        const double ri = __dsub_rn(sv[i], __dmul_rn(omega, t[i]));          // :130
        r[i] = ri;
        const double zi = __dmul_rn(pmode ? z[i] : sv[i], omega);            // :135
        double xi = __dadd_rn(x[i], zi);                                     // :136
        xi = __dadd_rn(xi, __dmul_rn(alpha, qi));                            // :137
        x[i] = xi;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(ri, ri));                       // :139
        // next trip's direction, bicgstab.py:91-93
        double pi = __dmul_rn(p[i], beta);
        pi = __dsub_rn(pi, __dmul_rn(bo, v[i]));
        pi = __dadd_rn(pi, ri);
        p[i] = pi;
        if (pmode) q[i] = apply_diag(pd, pmode, i, pi);                      // :96-99
    }
    static constexpr bool kPair = true;
    __device__ void pair(int i2, double *acc) const
    {
        double2 pv = ld2(p, i2), xv = ld2(x, i2);
        const double2 qv = pmode ? ld2(q, i2) : pv;
        if (half) {
            if (half == 1) {
                xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, qv.x));
                xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, qv.y));
                st2(x, i2, xv);
            }
            return;
        }
        const double2 sn = ld2(sv, i2), tv = ld2(t, i2), vv = ld2(v, i2);
        const double2 zs = pmode ? ld2(z, i2) : sn;
        double2 rn;
        rn.x = __dsub_rn(sn.x, __dmul_rn(omega, tv.x));
        rn.y = __dsub_rn(sn.y, __dmul_rn(omega, tv.y));
        st2(r, i2, rn);
        xv.x = __dadd_rn(__dadd_rn(xv.x, __dmul_rn(zs.x, omega)), __dmul_rn(alpha, qv.x));
        xv.y = __dadd_rn(__dadd_rn(xv.y, __dmul_rn(zs.y, omega)), __dmul_rn(alpha, qv.y));
        st2(x, i2, xv);
        acc[0] = __dadd_rn(acc[0], __dmul_rn(rn.x, rn.x));
        acc[0] = __dadd_rn(acc[0], __dmul_rn(rn.y, rn.y));
        pv.x = __dadd_rn(__dsub_rn(__dmul_rn(pv.x, beta), __dmul_rn(bo, vv.x)), rn.x);
        pv.y = __dadd_rn(__dsub_rn(__dmul_rn(pv.y, beta), __dmul_rn(bo, vv.y)), rn.y);
        st2(p, i2, pv);
        if (pmode) {
            double2 qn;
            qn.x = apply_diag(pd, pmode, 2 * i2, pv.x);
            qn.y = apply_diag(pd, pmode, 2 * i2 + 1, pv.y);
            st2(q, i2, qn);
        }
    }
};

struct BcgFinX {
    DevScalars *s;
    double     *hist;
    __device__ void operator()(const double *t) const
    {
        s->n_iter++;
        if (s->skip_half) {          // stays latched: it also gates the second SpMV
            s->done = 1;
            return;
        }
        const double resid = sqrt(t[0]);
        s->resid = resid;
        hist_push(s, hist, 1, resid, 0.0);                                   // :141
        if (resid <= s->threshold || s->n_matvec >= s->matvec_max) {         // :142
            s->done = 1;
            s->skip_half = 3;
        }
    }
};

// setup, bicgstab.py:62-83
struct R0SetupEpi {      // r0 = rhs - A x, fused rho = r0.r0, copies into the work vectors
    double       *r0, *a, *b;      // a, b: optional copies (r / p ...), may be null
    const double *rhs;
    __device__ void init() {}
    __device__ void operator()(int row, double ax, double *acc) const
    {
        const double v = __dsub_rn(rhs[row], ax);                            // :64
        r0[row] = v;
        if (a) a[row] = v;
        if (b) b[row] = v;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(v, v));                         // :68
    }
};

struct R0SetupBody {     // zero guess: r0 = rhs
    double       *r0, *a, *b;
    const double *rhs;
    __device__ void init() {}
    __device__ void operator()(int i, double *acc) const
    {
        const double v = rhs[i];
        r0[i] = v;
        if (a) a[i] = v;
        if (b) b[i] = v;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(v, v));
    }
};

struct BcgSetupFin {
    DevScalars *s;
    double     *hist;
    int         guess;
    __device__ void operator()(const double *t) const
    {
        const double rho_next = t[0];
        const double resid = fabs(sqrt(rho_next));                           // :69
        s->resid0 = s->resid = resid;
        s->threshold = fmax(s->abstol, s->reltol * resid);
        s->n_matvec = guess ? 1 : 0;                                         // :65
        hist_push(s, hist, 1, resid, 0.0);
        s->s[S_ALPHA] = 1.0;                                                 // :67
        s->s[S_OMEGA] = 1.0;
        s->s[S_RHO_NEXT] = rho_next;
        s->s[S_BETA] = rho_next / 1.0 * 1.0 / 1.0;                           // :87 first trip
        s->s[S_RHO] = rho_next;                                              // :88
        if (resid <= s->threshold || s->n_matvec >= s->matvec_max) {         // :72
            s->done = 1;
            s->skip_half = 3;
        }
    }
};

struct DiagApplyBody {   // q = M p
    double       *q;
    const double *p, *pd;
    int           pmode;
    __device__ void init() {}
    __device__ void operator()(int i) const { q[i] = apply_diag(pd, pmode, i, p[i]); }
};

static int bicgstab_setup(kry_solver *S, int guess)
{
    double *x = solver_vec(S, "x"), *r0 = solver_vec(S, "r0"), *r = solver_vec(S, "r");
    double *p = solver_vec(S, "p"), *rhs = solver_vec(S, "rhs"), *v = solver_vec(S, "v");
    BcgSetupFin fin{S->ds, S->hist, guess};
    KRY_CUDA(cudaMemsetAsync(v, 0, (size_t)S->n * sizeof(double), S->ctx->stream));   // :83
    // first trip: p = beta*0 - beta*omega*0 + r = r  (bicgstab.py:82,91-93)
    if (guess) {
        R0SetupEpi e{r0, r, p, rhs};
        KRY_TRY((solver_spmv<1>(S, GatherPlain{x}, e, fin, &S->ds->done, x)));
    } else {
        R0SetupBody b{r0, r, p, rhs};
        KRY_TRY((solver_pass<1>(S, b, fin, &S->ds->done)));
    }
    if (S->precon_mode) {
        DiagApplyBody q{solver_vec(S, "q"), p, S->dinv, S->precon_mode};
        KRY_TRY(vec_map_launch(S->ctx, S->n, q, &S->ds->done));
    }
    return KRY_OK;
}

static int bicgstab_iterate(kry_solver *S)
{
    double *x = solver_vec(S, "x"), *r0 = solver_vec(S, "r0"), *r = solver_vec(S, "r");
    double *p = solver_vec(S, "p"), *v = solver_vec(S, "v"), *sv = solver_vec(S, "s");
    double *t = solver_vec(S, "t"), *q = solver_vec(S, "q"), *z = solver_vec(S, "z");
    const int pm = S->precon_mode;
    const int *done = &S->ds->done;
    const int *skip = &S->ds->skip_half;
    double *qin = pm ? q : p;
    KRY_TRY((solver_spmv<1>(S, GatherPlain{qin}, BcgEpiV{v, r0}, BcgFinV{S->ds}, done, qin)));
    BcgBodyS bs{sv, z, r, v, S->dinv, pm, S->ds, 0.0};
    KRY_TRY((solver_pass<1>(S, bs, BcgFinS{S->ds, S->hist}, done)));
    double *zin = pm ? z : sv;
    // `skip` doubles as the early-out flag of the second half
    KRY_TRY((solver_spmv<3>(S, GatherPlain{zin}, BcgEpiT{t, sv, r0}, BcgFinT{S->ds}, skip, zin)));
    BcgBodyX bx{x, r, p, q, sv, z, t, v, S->dinv, pm, S->ds, 0, 0, 0, 0, 0};
    return solver_pass<1>(S, bx, BcgFinX{S->ds, S->hist}, done);
}

// ====================================================================== CGS
// cgs/cgs.py:76-117.  4 launches per iteration:
//   K1  v = A y ; sigma = r0.v ; alpha = rho/sigma                        [spmv]
//   K2  q = u - alpha v ; z = M (u + q) ; x += alpha z
//   K3  r -= alpha (A z) ; |r| ; rho' = r0.r ; stop test ; beta           [spmv]
//   K4  u = r + beta q ; p = beta (beta p + q) + u ; y = M p
struct CgsEpiV {
    double       *v;
    const double *r0;
    __device__ void init() {}
    __device__ void operator()(int row, double ax, double *acc) const
    {
        v[row] = ax;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(r0[row], ax));                  // cgs.py:85
    }
};

struct CgsFinV {
    DevScalars *s;
    __device__ void operator()(const double *t) const
    {
        s->n_matvec++;                                                       // :84
        s->s[S_SIGMA] = t[0];
        s->s[S_ALPHA] = s->s[S_RHO] / t[0];                                  // :86
    }
};

struct CgsBodyQ {
    double       *q, *z, *x;
    const double *u, *v, *pd;
    int           pmode;
    DevScalars   *s;
    double        alpha;
    __device__ void init() { alpha = s->s[S_ALPHA]; }
    __device__ void operator()(int i) const
    {
        const double qi = __dsub_rn(u[i], __dmul_rn(alpha, v[i]));           // :87
        q[i] = qi;
        const double zi = apply_diag(pd, pmode, i, __dadd_rn(u[i], qi));     // :89-92
        z[i] = zi;
        x[i] = __dadd_rn(x[i], __dmul_rn(alpha, zi));                        // :95
    }
};

struct CgsEpiR {
    static constexpr int kMinBlocks = 8;     // 32 registers: the row kernel needs full occupancy
    double       *r;
    const double *r0;
    DevScalars   *s;
    double        alpha;
    __device__ void init() { alpha = s->s[S_ALPHA]; }
    __device__ void operator()(int row, double ax, double *acc) const
    {
        const double ri = __dsub_rn(r[row], __dmul_rn(alpha, ax));           // :97
        r[row] = ri;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(ri, ri));                       // :100
        acc[1] = __dadd_rn(acc[1], __dmul_rn(r0[row], ri));                  // :106
    }
};

struct CgsFinR {
    DevScalars *s;
    double     *hist;
    __device__ void operator()(const double *t) const
    {
        s->n_matvec++;                                                       // :96
        s->n_iter++;
        const double resid = sqrt(t[0]);
        s->resid = resid;
        hist_push(s, hist, 1, resid, 0.0);
        if (resid <= s->threshold || s->n_matvec >= s->matvec_max) {         // :102-104
            s->done = 1;
            return;
        }
        const double rho_next = t[1];
        s->s[S_BETA] = rho_next / s->s[S_RHO];                               // :107
        s->s[S_RHO] = rho_next;                                              // :108
    }
};

struct CgsBodyP {
    double       *u, *p, *y;
    const double *r, *q, *pd;
    int           pmode;
    DevScalars   *s;
    double        beta;
    __device__ void init() { beta = s->s[S_BETA]; }
    __device__ void operator()(int i) const
    {
        const double ui = __dadd_rn(r[i], __dmul_rn(beta, q[i]));            // :109
        u[i] = ui;
        double pi = __dmul_rn(p[i], beta);                                   // :112-115
        pi = __dadd_rn(pi, q[i]);
        pi = __dmul_rn(pi, beta);
        pi = __dadd_rn(pi, ui);
        p[i] = pi;
        if (pmode) y[i] = apply_diag(pd, pmode, i, pi);                      // :79-82
    }
};

struct RhoSetupFin {     // CGS / TFQMR: rho = r0.r0, residNorm0, threshold (matvec NOT counted)
    DevScalars *s;
    double     *hist;
    long long   nmv0;
    __device__ void operator()(const double *t) const
    {
        s->s[S_RHO] = t[0];                                                  // cgs.py:62, tfqmr.py:61
        const double resid = fabs(sqrt(t[0]));
        s->resid0 = s->resid = resid;
        s->threshold = fmax(s->abstol, s->reltol * resid);
        s->n_matvec = nmv0;
        hist_push(s, hist, 1, resid, 0.0);
        if (resid <= s->threshold || s->n_matvec >= s->matvec_max) s->done = 1;
    }
};

static int cgs_setup(kry_solver *S, int guess)
{
    double *x = solver_vec(S, "x"), *r0 = solver_vec(S, "r0"), *r = solver_vec(S, "r");
    double *p = solver_vec(S, "p"), *u = solver_vec(S, "u"), *rhs = solver_vec(S, "rhs");
    RhoSetupFin fin{S->ds, S->hist, 0};
    if (guess) {
        R0SetupEpi e{r0, r, p, rhs};
        KRY_TRY((solver_spmv<1>(S, GatherPlain{x}, e, fin, &S->ds->done, x)));
    } else {
        R0SetupBody b{r0, r, p, rhs};
        KRY_TRY((solver_pass<1>(S, b, fin, &S->ds->done)));
    }
    KRY_CUDA(cudaMemcpyAsync(u, r0, (size_t)S->n * sizeof(double), cudaMemcpyDeviceToDevice,
                             S->ctx->stream));                               // cgs.py:73
    if (S->precon_mode) {
        DiagApplyBody y{solver_vec(S, "y"), p, S->dinv, S->precon_mode};
        KRY_TRY(vec_map_launch(S->ctx, S->n, y, &S->ds->done));
    }
    return KRY_OK;
}

static int cgs_iterate(kry_solver *S)
{
    double *x = solver_vec(S, "x"), *r0 = solver_vec(S, "r0"), *r = solver_vec(S, "r");
    double *p = solver_vec(S, "p"), *u = solver_vec(S, "u"), *q = solver_vec(S, "q");
    double *v = solver_vec(S, "v"), *z = solver_vec(S, "z"), *y = solver_vec(S, "y");
    const int pm = S->precon_mode;
    const int *done = &S->ds->done;
    double *yin = pm ? y : p;
    KRY_TRY((solver_spmv<1>(S, GatherPlain{yin}, CgsEpiV{v, r0}, CgsFinV{S->ds}, done, yin)));
    CgsBodyQ bq{q, z, x, u, v, S->dinv, pm, S->ds, 0.0};
    KRY_TRY(vec_map_launch(S->ctx, S->n, bq, done));
    CgsEpiR er{r, r0, S->ds, 0.0};
    KRY_TRY((solver_spmv<2>(S, GatherPlain{z}, er, CgsFinR{S->ds, S->hist}, done, z)));
    CgsBodyP bp{u, p, y, r, q, S->dinv, pm, S->ds, 0.0};
    return vec_map_launch(S->ctx, S->n, bp, done);
}

// ==================================================================== TFQMR
// tfqmr/tfqmr.py:85-153.  State: x, r0, y, w, d, u, v (+ z = M y).  The two
// half-steps share one kernel (`TfqHalfBody`); launches per iteration:
//   K1  sigma = r0.v -> alpha                (fused into the tail of the previous trip's
//                                             SpMV when possible; a plain dot on trip 1)
//   K2  w -= alpha u ; |w|                                       first half, :92,95
//   K3  d = coef d + z ; x += eta d ; y -= alpha v ; z = M y     :93-94,99,109  (needs theta, eta)
//   K4  u = A z ; w -= alpha u ; |w|                             :114-118       [spmv]
//   K5  d = coef d + z ; x += eta d ; (rho' = r0.w)              :117-123,128
//   K6  y = beta y + w ; v = beta (beta v + u) ; z = M y          :133-139
//   K7  u = A z ; v += u ; sigma = r0.v -> alpha of next trip     :146-149       [spmv]
struct TfqBodyW {             // w -= alpha u ; |w|^2
    double       *w;
    const double *u;
    DevScalars   *s;
    double        alpha;
    __device__ void init() { alpha = s->s[S_ALPHA]; }
    __device__ void operator()(int i, double *acc) const
    {
        const double wi = __dsub_rn(w[i], __dmul_rn(alpha, u[i]));           // :92 / :116
        w[i] = wi;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(wi, wi));
    }
};

// scalar part shared by both halves, tfqmr.py:93,95-98 / :117,119-122
__device__ __forceinline__ void tfq_half_scalars(DevScalars *s, double *hist, double wnorm2, double m)
{
    const double theta_old = s->s[S_THETA], eta_old = s->s[S_ETA], alpha = s->s[S_ALPHA];
    s->s[S_DCOEF] = theta_old * theta_old * eta_old / alpha;                 // :93 (uses OLD theta, eta)
    const double theta = sqrt(wnorm2) / s->resid;                            // :95
    const double c = 1.0 / sqrt(1 + theta * theta);                          // :96
    s->resid = s->resid * (theta * c);                                       // :97
    s->s[S_THETA] = theta;
    s->s[S_ETA] = c * c * alpha;                                             // :98
    s->s[S_M] = m;
    hist_push(s, hist, 1, s->resid, 0.0);
}

struct TfqFinW1 {
    DevScalars *s;
    double     *hist;
    __device__ void operator()(const double *t) const
    {
        const double m = 2.0 * s->s[S_K] - 1.0;                              // :101
        tfq_half_scalars(s, hist, t[0], m);
        // stop test is applied after x has been updated (K3 latches it): :103-105
        if (s->resid * sqrt(m + 1) < s->threshold || s->n_matvec >= s->matvec_max) s->skip_half = 1;
    }
};

struct TfqBodyD1 {            // d = coef d + z ; x += eta d ; then (if continuing) y -= alpha v ; z = M y
    double       *d, *x, *y, *z;
    const double *v, *pd;
    int           pmode;
    DevScalars   *s;
    double        coef, eta, alpha;
    int           stop;
    __device__ void init()
    {
        coef = s->s[S_DCOEF];
        eta = s->s[S_ETA];
        alpha = s->s[S_ALPHA];
        stop = s->skip_half;
    }
    __device__ void operator()(int i, double *acc) const
    {
        const double zi = pmode ? z[i] : y[i];
        const double di = __dadd_rn(__dmul_rn(d[i], coef), zi);              // :93-94
        d[i] = di;
        x[i] = __dadd_rn(x[i], __dmul_rn(eta, di));                          // :99
        if (stop) return;
        const double yi = __dsub_rn(y[i], __dmul_rn(alpha, v[i]));           // :109
        y[i] = yi;
        if (pmode) z[i] = apply_diag(pd, pmode, i, yi);                      // :110-113
    }
};

struct TfqFinD1 {
    DevScalars *s;
    __device__ void operator()(const double *) const
    {
        if (s->skip_half) {
            s->skip_half = 0;
            s->n_iter++;
            s->done = 1;
        }
    }
};

struct TfqEpiU2 {             // u = A z ; w -= alpha u ; |w|^2
    double     *u, *w;
    DevScalars *s;
    double      alpha;
    __device__ void init() { alpha = s->s[S_ALPHA]; }
    __device__ void operator()(int row, double ax, double *acc) const
    {
        u[row] = ax;                                                         // :114
        const double wi = __dsub_rn(w[row], __dmul_rn(alpha, ax));           // :116
        w[row] = wi;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(wi, wi));
    }
};

struct TfqFinW2 {
    DevScalars *s;
    double     *hist;
    __device__ void operator()(const double *t) const
    {
        s->n_matvec++;                                                       // :114
        const double m = 2.0 * s->s[S_K];                                    // :108
        tfq_half_scalars(s, hist, t[0], m);
        if (s->resid * sqrt(m + 1) < s->threshold || s->n_matvec >= s->matvec_max) s->skip_half = 1;
    }
};

struct TfqBodyD2 {            // d = coef d + z ; x += eta d ; rho' = r0.w
    double       *d, *x;
    const double *y, *z, *r0, *w;
    int           pmode;
    DevScalars   *s;
    double        coef, eta;
    __device__ void init()
    {
        coef = s->s[S_DCOEF];
        eta = s->s[S_ETA];
    }
    __device__ void operator()(int i, double *acc) const
    {
        const double zi = pmode ? z[i] : y[i];
        const double di = __dadd_rn(__dmul_rn(d[i], coef), zi);              // :117-118
        d[i] = di;
        x[i] = __dadd_rn(x[i], __dmul_rn(eta, di));                          // :123
        acc[0] = __dadd_rn(acc[0], __dmul_rn(r0[i], w[i]));                  // :128
    }
};

struct TfqFinD2 {
    DevScalars *s;
    __device__ void operator()(const double *t) const
    {
        s->n_iter++;
        if (s->skip_half) {                                                  // :125-127
            s->skip_half = 0;
            s->done = 1;
            return;
        }
        s->s[S_BETA] = t[0] / s->s[S_RHO];                                   // :129
        s->s[S_RHO] = t[0];                                                  // :130
    }
};

struct TfqBodyYV {            // y = beta y + w ; v = beta (beta v + u) ; z = M y
    double       *y, *v, *z;
    const double *w, *u, *pd;
    int           pmode;
    DevScalars   *s;
    double        beta;
    __device__ void init() { beta = s->s[S_BETA]; }
    __device__ void operator()(int i) const
    {
        const double yi = __dadd_rn(__dmul_rn(y[i], beta), w[i]);            // :133-134
        y[i] = yi;
        double vi = __dmul_rn(v[i], beta);                                   // :137-139
        vi = __dadd_rn(vi, u[i]);
        vi = __dmul_rn(vi, beta);
        v[i] = vi;
        if (pmode) z[i] = apply_diag(pd, pmode, i, yi);                      // :142-145
    }
};

struct TfqEpiU3 {             // u = A z ; v += u ; sigma = r0.v (next trip)
    double       *u, *v;
    const double *r0;
    __device__ void init() {}
    __device__ void operator()(int row, double ax, double *acc) const
    {
        u[row] = ax;                                                         // :146
        const double vi = __dadd_rn(v[row], ax);                             // :149
        v[row] = vi;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(r0[row], vi));                  // :88 of the next trip
    }
};

struct TfqFinU3 {
    DevScalars *s;
    __device__ void operator()(const double *t) const
    {
        s->n_matvec++;                                                       // :146
        s->s[S_K] += 1.0;                                                    // :87 next trip
        s->s[S_SIGMA] = t[0];
        s->s[S_ALPHA] = s->s[S_RHO] / t[0];                                  // :89
    }
};

struct TfqSetupEpiU {         // u = A z ; v = u  (tfqmr.py:84-85) ; sigma = r0.v for trip 1
    double       *u, *v;
    const double *r0;
    __device__ void init() {}
    __device__ void operator()(int row, double ax, double *acc) const
    {
        u[row] = ax;
        v[row] = ax;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(r0[row], ax));
    }
};

struct TfqSetupFinU {
    DevScalars *s;
    __device__ void operator()(const double *t) const
    {
        s->n_matvec++;                                                       // :84
        s->s[S_K] = 1.0;                                                     // :87 first trip
        s->s[S_SIGMA] = t[0];
        s->s[S_ALPHA] = s->s[S_RHO] / t[0];                                  // :89
        s->s[S_THETA] = 0.0;                                                 // :76-77
        s->s[S_ETA] = 0.0;
    }
};

static int tfqmr_setup(kry_solver *S, int guess)
{
    double *x = solver_vec(S, "x"), *r0 = solver_vec(S, "r0"), *y = solver_vec(S, "y");
    double *w = solver_vec(S, "w"), *d = solver_vec(S, "d"), *u = solver_vec(S, "u");
    double *v = solver_vec(S, "v"), *z = solver_vec(S, "z"), *rhs = solver_vec(S, "rhs");
    const int pm = S->precon_mode;
    RhoSetupFin fin{S->ds, S->hist, 0};
    if (guess) {
        R0SetupEpi e{r0, y, w, rhs};
        KRY_TRY((solver_spmv<1>(S, GatherPlain{x}, e, fin, &S->ds->done, x)));
    } else {
        R0SetupBody b{r0, y, w, rhs};
        KRY_TRY((solver_pass<1>(S, b, fin, &S->ds->done)));
    }
    KRY_CUDA(cudaMemsetAsync(d, 0, (size_t)S->n * sizeof(double), S->ctx->stream));     // :75
    if (pm) {
        DiagApplyBody zb{z, y, S->dinv, pm};
        KRY_TRY(vec_map_launch(S->ctx, S->n, zb, &S->ds->done));
    }
    double *zin = pm ? z : y;
    return solver_spmv<1>(S, GatherPlain{zin}, TfqSetupEpiU{u, v, r0}, TfqSetupFinU{S->ds},
                          &S->ds->done, zin);
}

static int tfqmr_iterate(kry_solver *S)
{
    double *x = solver_vec(S, "x"), *r0 = solver_vec(S, "r0"), *y = solver_vec(S, "y");
    double *w = solver_vec(S, "w"), *d = solver_vec(S, "d"), *u = solver_vec(S, "u");
    double *v = solver_vec(S, "v"), *z = solver_vec(S, "z");
    const int pm = S->precon_mode;
    const int *done = &S->ds->done;
    // first half
    TfqBodyW bw{w, u, S->ds, 0.0};
    KRY_TRY((solver_pass<1>(S, bw, TfqFinW1{S->ds, S->hist}, done)));
    TfqBodyD1 bd1{d, x, y, z, v, S->dinv, pm, S->ds, 0, 0, 0, 0};
    KRY_TRY((solver_pass<1>(S, bd1, TfqFinD1{S->ds}, done)));
    // second half
    double *zin = pm ? z : y;
    TfqEpiU2 eu2{u, w, S->ds, 0.0};
    KRY_TRY((solver_spmv<1>(S, GatherPlain{zin}, eu2, TfqFinW2{S->ds, S->hist}, done, zin)));
    TfqBodyD2 bd2{d, x, y, z, r0, w, pm, S->ds, 0, 0};
    KRY_TRY((solver_pass<1>(S, bd2, TfqFinD2{S->ds}, done)));
    // final updates
    TfqBodyYV byv{y, v, z, w, u, S->dinv, pm, S->ds, 0.0};
    KRY_TRY(vec_map_launch(S->ctx, S->n, byv, done));
    return solver_spmv<1>(S, GatherPlain{zin}, TfqEpiU3{u, v, r0}, TfqFinU3{S->ds}, done, zin);
}

// =================================================================== MINRES
// minres/minres.py:218-383.  3 launches per iteration:
//   K1  v = y/beta (on the fly) ; y' = A v - shift v - (beta/oldb) r1 ; alfa = v.y'   [spmv]
//   K2  y = y' - (alfa/beta) r2 ; r1 <- r2 ; r2 <- y ; (y = M r2) ; beta' = r2.y ;
//       whole scalar QR step, norms and the five stopping tests on device
//   K3  w = (v - oldeps w1 - delta w2)/gamma ; x += phi w      (w1 <- w2 <- w rotate)
// Vectors rotate by pointer on the host side (the launch sequence is static, so
// the rotation is a fixed period-3 / period-2 schedule, see minres_iterate()).
struct MinGather {            // x[c] * (1/beta): v = s*y with s = 1.0/beta, minres.py:236-237
    const double *y;
    DevScalars   *s;
    double        inv;
    __device__ void   init() { inv = 1.0 / s->s[M_BETA]; }
    __device__ double operator()(int c) const { return __dmul_rn(inv, __ldg(y + c)); }
    __device__ double boundary(int j) const { return y[j]; }
    __device__ double coherent(int c) const { return __dmul_rn(inv, y[c]); }
};

struct MinEpiY {
    double       *yn;         // y' (becomes r2 after K2)
    const double *y, *r1;
    DevScalars   *s;
    double        inv, shift, c_r1;
    int           first;
    __device__ void init()
    {
        inv = 1.0 / s->s[M_BETA];
        shift = s->shift;
        first = (s->n_iter == 0);
        c_r1 = first ? 0.0 : s->s[M_BETA] / s->s[M_OLDB];                    // :243
    }
    __device__ void operator()(int row, double ax, double *acc) const
    {
        const double vi = __dmul_rn(inv, y[row]);                            // :237
        double yi = __dsub_rn(ax, __dmul_rn(shift, vi));                     // :240
        if (!first) yi = __dsub_rn(yi, __dmul_rn(c_r1, r1[row]));            // :242-243
        yn[row] = yi;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(vi, yi));                       // :245
    }
};

struct MinFinAlfa {
    DevScalars *s;
    __device__ void operator()(const double *t) const
    {
        s->s[M_ALFA] = t[0];
        s->s[M_C_R2] = -t[0] / s->s[M_BETA];                                 // :246
    }
};

struct MinBodyR2 {            // y = (-alfa/beta) r2 + y' ; beta'^2 = r2new . (M r2new)
    double       *yn, *ypre;
    const double *r2, *pd;
    int           pmode;
    DevScalars   *s;
    double        c;
    __device__ void init() { c = s->s[M_C_R2]; }
    __device__ void operator()(int i, double *acc) const
    {
        const double yi = __dadd_rn(__dmul_rn(c, r2[i]), yn[i]);             // :246
        yn[i] = yi;                                                          // :248 (new r2)
        const double yp = apply_diag(pd, pmode, i, yi);                      // :249
        if (pmode) ypre[i] = yp;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(yi, yp));                       // :251
    }
    static constexpr bool kPair = true;
    __device__ void pair(int i2, double *acc) const      // (no preconditioner on this path)
    {
        const double2 rv = ld2(r2, i2);
        double2 yv = ld2(yn, i2);
        yv.x = __dadd_rn(__dmul_rn(c, rv.x), yv.x);
        yv.y = __dadd_rn(__dmul_rn(c, rv.y), yv.y);
        st2(yn, i2, yv);
        acc[0] = __dadd_rn(acc[0], __dmul_rn(yv.x, yv.x));
        acc[0] = __dadd_rn(acc[0], __dmul_rn(yv.y, yv.y));
    }
};

struct MinBodyR2Pre {         // preconditioned K2: r2' = (-alfa/beta) r2 + y' ; y = M r2' ; beta'^2 = r2'.y
    static constexpr int kMinBlocks = 4;       // the fp64 division of `r ./ d` spills under 40 registers
    double       *yn, *ypre;  // yn: y' in, the new r2 out; ypre: the new (preconditioned) y
    const double *r2, *pd;
    int           pmode;
    DevScalars   *s;
    double        c;
    __device__ void init() { c = s->s[M_C_R2]; }
    __device__ void operator()(int i, double *acc) const
    {
        const double yi = __dadd_rn(__dmul_rn(c, r2[i]), yn[i]);             // :246
        yn[i] = yi;                                                          // :248 (new r2)
        const double yp = apply_diag(pd, pmode, i, yi);                      // :249
        ypre[i] = yp;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(yi, yp));                       // :251
    }
};

struct MinFinQR {
    DevScalars *s;
    double     *hist;
    int         fuse;        // 2-launch plan: this launch is the last of the trip -> it latches `done`
                             // itself and records that the w / x update of the trip is still owed
    __device__ void operator()(const double *t) const
    {
        const double eps = 2.220446049250313e-16;
        double *v = s->s;
        if (fuse == 1) v[M_WPEND] = 0.0; // the body of this launch paid what the previous trip owed
        if (fuse == 2) v[M_STOPNOW] = 0.0;
        s->n_iter++;
        s->n_matvec++;
        const long long itn = s->n_iter;
        const double alfa = v[M_ALFA];
        const double oldb = v[M_BETA];                                       // :250
        v[M_OLDB] = oldb;
        double beta = t[0];                                                  // :251
        if (beta < 0) {                                                      // :252-254
            s->istop = 6;
            s->done = 1;
            if (fuse == 2) v[M_STOPNOW] = 1.0;
            return;
        }
        beta = sqrt(beta);                                                   // :255
        v[M_BETA] = beta;
        v[M_TNORM2] = v[M_TNORM2] + alfa * alfa + oldb * oldb + beta * beta; // :256
        if (itn == 1) {                                                      // :258-264
            if (beta / v[M_BETA1] <= 10 * eps) s->istop = -1;
            v[M_GMAX] = fabs(alfa);
            v[M_GMIN] = v[M_GMAX];
        }
        const double cs = v[M_CS], sn = v[M_SN], dbar = v[M_DBAR];
        const double oldeps = v[M_EPSLN];                                    // :270
        const double delta = cs * dbar + sn * alfa;                          // :271
        const double gbar = sn * dbar - cs * alfa;                           // :272
        const double epsln = sn * beta;                                      // :277
        v[M_DBAR] = -cs * beta;                                              // :278
        const double root = sqrt(gbar * gbar + v[M_DBAR] * v[M_DBAR]);       // :279
        v[M_ARNORM] = v[M_PHIBAR] * root;                                    // :280
        double gamma = sqrt(gbar * gbar + beta * beta);                      // :284
        gamma = fmax(gamma, eps);                                            // :285
        v[M_CS] = gbar / gamma;                                              // :286
        v[M_SN] = beta / gamma;                                              // :287
        const double phi = v[M_CS] * v[M_PHIBAR];                            // :288
        v[M_PHIBAR] = v[M_SN] * v[M_PHIBAR];                                 // :289
        v[M_OLDEPS] = oldeps;
        v[M_DELTA] = delta;
        v[M_GBAR] = gbar;
        v[M_EPSLN] = epsln;
        v[M_PHI] = phi;
        v[M_DENOM] = 1.0 / gamma;                                            // :293
        v[M_INVBETA] = 1.0 / oldb;     // s of this trip (v = s*y), for K3
        // energy norm / truncated direct error, :303-310
        v[M_XNRG2] += phi * phi;
        const int window = s->window;
        double direrr = nan("");
        s->derr[itn % window] = phi;
        if (itn > window) {
            double ss = 0.0;           // np.linalg.norm(dErr): sqrt of the ordered sum of squares
            for (int k = 0; k < window; ++k) ss += s->derr[k] * s->derr[k];
            const double trnc = sqrt(ss);
            v[M_TRNC] = trnc;
            const double xnrg = sqrt(v[M_XNRG2]);
            direrr = trnc / xnrg;                                            // :308
            if (trnc < s->etol * xnrg) s->istop = 10;
        }
        v[M_GMAX] = fmax(v[M_GMAX], gamma);                                  // :314-319
        v[M_GMIN] = fmin(v[M_GMIN], gamma);
        const double z = v[M_RHS1] / gamma;
        v[M_YNORM2] = z * z + v[M_YNORM2];
        v[M_RHS1] = v[M_RHS2] - delta * z;
        v[M_RHS2] = -epsln * z;
        const double Anorm = sqrt(v[M_TNORM2]);                              // :323-334
        const double ynorm = sqrt(v[M_YNORM2]);
        const double epsx = Anorm * ynorm * eps;
        const double rnorm = v[M_PHIBAR];
        const double test1 = rnorm / (Anorm * ynorm);
        const double test2 = root / Anorm;
        v[M_ANORM] = Anorm;
        v[M_YNORM] = ynorm;
        v[M_TEST1] = test1;
        v[M_TEST2] = test2;
        s->resid = rnorm;
        hist_push(s, hist, 2, rnorm, direrr);                                // :336, :308
        const double Acond = v[M_GMAX] / v[M_GMIN];                          // :344
        v[M_ACOND] = Acond;
        if (s->istop == 0) {                                                 // :349-361
            const double t1 = 1 + test1, t2 = 1 + test2;
            if (t2 <= 1) s->istop = 2;
            if (t1 <= 1) s->istop = 1;
            if (itn >= s->matvec_max) s->istop = 6;
            if (Acond >= 0.1 / eps) s->istop = 4;
            if (epsx >= v[M_BETA1]) s->istop = 3;
            if (test2 <= s->rtol) s->istop = 2;
            if (test1 <= s->rtol) s->istop = 1;
        }
        if (fuse == 1) {
            v[M_WPEND] = 1.0;                                                // minres.py:294-297 of this trip
            if (s->istop > 0 || itn >= s->matvec_max) s->done = 1;           // :381, :218
            return;
        }
        if (fuse == 2) {             // persistent kernel: every CTA reads `done` right after this barrier, runs
            if (s->istop > 0 || itn >= s->matvec_max) s->done = 1;   // the w / x update of the trip and leaves
            return;
        }
        // `done` is latched by K3 (the x update of this trip must still run)
        if (s->istop > 0 || itn >= s->matvec_max) s->skip_half = 1;          // :381, :218
    }
};

// 2-launch plan: K2 of this trip + the w / x update the previous trip still owes.  The owed
// update uses the previous trip's buffers (y_prev = this trip's r1, the two w buffers swapped)
// and the scalars its MinFinQR left (this launch's finalize overwrites them only after every
// CTA has finished its body).
template <bool PEND>
struct MinBodyR2W {
    static constexpr int kMinBlocks = 4;       // 7 vectors in flight: allow 64 registers
    double       *yn;                          // K2 part (as MinBodyR2, no preconditioner)
    const double *r2;
    double       *wnew, *x;                    // owed part (as MinBodyW, previous trip's roles)
    const double *w2, *yold;
    DevScalars   *s;
    double        c, inv, oldeps, delta, denom, phi;
    __device__ void init()
    {
        c = s->s[M_C_R2];
        inv = s->s[M_INVBETA];
        oldeps = s->s[M_OLDEPS];
        delta = s->s[M_DELTA];
        denom = s->s[M_DENOM];
        phi = s->s[M_PHI];
    }
    __device__ void operator()(int i, double *acc) const
    {
        const double yi = __dadd_rn(__dmul_rn(c, r2[i]), yn[i]);             // :246
        yn[i] = yi;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(yi, yi));                       // :251
        if constexpr (PEND) {
            const double vi = __dmul_rn(inv, yold[i]);                       // :237 of the previous trip
            double wi = __dsub_rn(vi, __dmul_rn(oldeps, wnew[i]));           // :296
            wi = __dsub_rn(wi, __dmul_rn(delta, w2[i]));
            wi = __dmul_rn(wi, denom);
            wnew[i] = wi;
            x[i] = __dadd_rn(x[i], __dmul_rn(phi, wi));                      // :297
        }
    }
    static constexpr bool kPair = true;
    __device__ void pair(int i2, double *acc) const
    {
        const double2 rv = ld2(r2, i2);
        double2 yv = ld2(yn, i2);
        yv.x = __dadd_rn(__dmul_rn(c, rv.x), yv.x);
        yv.y = __dadd_rn(__dmul_rn(c, rv.y), yv.y);
        st2(yn, i2, yv);
        acc[0] = __dadd_rn(acc[0], __dmul_rn(yv.x, yv.x));
        acc[0] = __dadd_rn(acc[0], __dmul_rn(yv.y, yv.y));
        if constexpr (PEND) {
            const double2 yo = ld2(yold, i2), w1 = ld2(wnew, i2), wv = ld2(w2, i2);
            double2 xv = ld2(x, i2), wn;
            wn.x = __dmul_rn(__dsub_rn(__dsub_rn(__dmul_rn(inv, yo.x), __dmul_rn(oldeps, w1.x)), __dmul_rn(delta, wv.x)), denom);
            wn.y = __dmul_rn(__dsub_rn(__dsub_rn(__dmul_rn(inv, yo.y), __dmul_rn(oldeps, w1.y)), __dmul_rn(delta, wv.y)), denom);
            st2(wnew, i2, wn);
            xv.x = __dadd_rn(xv.x, __dmul_rn(phi, wn.x));
            xv.y = __dadd_rn(xv.y, __dmul_rn(phi, wn.y));
            st2(x, i2, xv);
        }
    }
};

// Pay the owed w / x update (driven by the device flag and the device trip counter, so it
// is right whichever launch latched `done`); afterwards the state is that of the 3-launch plan.
struct MinSettleBody {
    double     *R[3], *W[2], *x;
    DevScalars *s;
    double     *wnew;
    const double *w2, *yold;
    double      inv, oldeps, delta, denom, phi;
    int         pend;
    __device__ void init()
    {
        pend = (s->s[M_WPEND] != 0.0);
        const long long t = s->n_iter - 1;              // the trip whose update is owed
        const int k = (int)(((t % 3) + 3) % 3), j = (int)(((t % 2) + 2) % 2);
        yold = R[k];
        wnew = W[1 - j];
        w2 = W[j];
        inv = s->s[M_INVBETA];
        oldeps = s->s[M_OLDEPS];
        delta = s->s[M_DELTA];
        denom = s->s[M_DENOM];
        phi = s->s[M_PHI];
    }
    __device__ void operator()(int i) const
    {
        if (!pend) return;
        const double vi = __dmul_rn(inv, yold[i]);
        double wi = __dsub_rn(vi, __dmul_rn(oldeps, wnew[i]));
        wi = __dsub_rn(wi, __dmul_rn(delta, w2[i]));
        wi = __dmul_rn(wi, denom);
        wnew[i] = wi;
        x[i] = __dadd_rn(x[i], __dmul_rn(phi, wi));
    }
};

struct MinBodyW {             // w = (v - oldeps w1 - delta w2) * denom ; x += phi w
    double       *wnew, *x;   // wnew overwrites the w1 buffer (w1 dies here)
    const double *w2, *yold;  // yold: the y this trip's v was built from
    DevScalars   *s;
    double        inv, oldeps, delta, denom, phi;
    __device__ void init()
    {
        inv = s->s[M_INVBETA];
        oldeps = s->s[M_OLDEPS];
        delta = s->s[M_DELTA];
        denom = s->s[M_DENOM];
        phi = s->s[M_PHI];
    }
    __device__ void operator()(int i, double *acc) const
    {
        const double vi = __dmul_rn(inv, yold[i]);                           // :237
        double wi = __dsub_rn(vi, __dmul_rn(oldeps, wnew[i]));               // :296 (wnew holds w1)
        wi = __dsub_rn(wi, __dmul_rn(delta, w2[i]));
        wi = __dmul_rn(wi, denom);
        wnew[i] = wi;
        x[i] = __dadd_rn(x[i], __dmul_rn(phi, wi));                          // :297
    }
    static constexpr bool kPair = true;
    __device__ void pair(int i2, double *) const
    {
        const double2 yv = ld2(yold, i2), w1 = ld2(wnew, i2), wv = ld2(w2, i2);
        double2 xv = ld2(x, i2), wn;
        wn.x = __dmul_rn(__dsub_rn(__dsub_rn(__dmul_rn(inv, yv.x), __dmul_rn(oldeps, w1.x)), __dmul_rn(delta, wv.x)), denom);
        wn.y = __dmul_rn(__dsub_rn(__dsub_rn(__dmul_rn(inv, yv.y), __dmul_rn(oldeps, w1.y)), __dmul_rn(delta, wv.y)), denom);
        st2(wnew, i2, wn);
        xv.x = __dadd_rn(xv.x, __dmul_rn(phi, wn.x));
        xv.y = __dadd_rn(xv.y, __dmul_rn(phi, wn.y));
        st2(x, i2, xv);
    }
};

struct MinFinW {
    DevScalars *s;
    __device__ void operator()(const double *) const
    {
        if (s->skip_half) {
            s->skip_half = 0;
            s->done = 1;
        }
    }
};

struct MinSetupBody {         // y = M b ; beta1^2 = b.y   (minres.py:160-166)
    double       *y;
    const double *b, *pd;
    int           pmode;
    __device__ void init() {}
    __device__ void operator()(int i, double *acc) const
    {
        const double yi = apply_diag(pd, pmode, i, b[i]);
        y[i] = yi;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(b[i], yi));
    }
};

struct MinSetupFin {
    DevScalars *s;
    __device__ void operator()(const double *t) const
    {
        double *v = s->s;
        double beta1 = t[0];
        if (beta1 < 0) {                                                     // :170-173
            s->istop = 9;
            s->done = 1;
        }
        if (beta1 == 0.0) s->done = 1;                                       // :175-177
        if (beta1 > 0) beta1 = sqrt(beta1);                                  // :179-180
        v[M_BETA1] = beta1;
        s->resid0 = beta1;
        s->resid = 0.0;
        v[M_OLDB] = 0.0;                                                     // :201-205
        v[M_BETA] = beta1;
        v[M_DBAR] = 0.0;
        v[M_EPSLN] = 0.0;
        v[M_PHIBAR] = beta1;
        v[M_RHS1] = beta1;
        v[M_RHS2] = 0.0;
        v[M_TNORM2] = 0.0;
        v[M_YNORM2] = 0.0;
        v[M_CS] = -1.0;
        v[M_SN] = 0.0;
        v[M_ARNORM] = 0.0;
        v[M_XNRG2] = 0.0;
        v[M_ANORM] = v[M_ACOND] = v[M_YNORM] = 0.0;
        if (s->matvec_max <= 0) s->done = 1;                                 // :218
    }
};

static int minres_setup(kry_solver *S, int)
{
    // y = b.copy() lives in R[0]; r2 == y value-wise (no preconditioner), r1 is
    // first read on trip 2, when it is the rotated r2 (minres.py:160-165, 208).
    const size_t bytes = (size_t)S->n * sizeof(double);
    if (S->precon_mode) {
        // with a preconditioner y = M r2 is a vector of its own (two buffers: a trip reads one and
        // writes the other); r2 = r1 = b                                      minres.py:160-165
        KRY_CUDA(cudaMemcpyAsync(solver_vec(S, "ra"), solver_vec(S, "rhs"), bytes, cudaMemcpyDeviceToDevice,
                                 S->ctx->stream));
        MinSetupBody sb{solver_vec(S, "ya"), solver_vec(S, "rhs"), S->dinv, S->precon_mode};
        KRY_TRY((solver_pass<1>(S, sb, MinSetupFin{S->ds}, &S->ds->done)));
    } else {
        MinSetupBody sb{solver_vec(S, "ra"), solver_vec(S, "rhs"), S->dinv, 0};
        KRY_TRY((solver_pass<1>(S, sb, MinSetupFin{S->ds}, &S->ds->done)));
    }
    KRY_CUDA(cudaMemsetAsync(solver_vec(S, "wa"), 0, bytes, S->ctx->stream));   // :206-207
    KRY_CUDA(cudaMemsetAsync(solver_vec(S, "wb"), 0, bytes, S->ctx->stream));
    S->rot = 0;
    S->minres_fuse = S->ctx->minres_fuse;
    S->minres_persistent = S->ctx->minres_persistent && !S->sharded && !S->A->halo.active && !S->precon_mode &&
                           (S->A->kind == KRY_SPMV_AUTO || S->A->kind == KRY_SPMV_ROW) && S->A->A.max_row <= 64;
#ifdef KRY_EMULATE
    if (!emu_fibers_on) S->minres_persistent = false;          // the host emulation plays it in SIMT mode only
#endif
    if (S->minres_persistent) S->minres_fuse = 0;                // its HBM state is that of the 3-launch plan
    S->fresh = true;
    return KRY_OK;
}

static int minres_iterate(kry_solver *S)
{
    // r buffers rotate with period 3: r2 (== y without precon) = R[k], r1 = R[k-1],
    // R[k+1] receives this trip's y.  w buffers rotate with period 2: w = W[j],
    // w2 = W[1-j]; the new w overwrites w2's buffer (w1 := w2 dies in the same pass).
    double *R[3] = {solver_vec(S, "ra"), solver_vec(S, "rb"), solver_vec(S, "rc")};
    double *W[2] = {solver_vec(S, "wa"), solver_vec(S, "wb")};
    double *x = solver_vec(S, "x");
    const int k = (int)(S->rot % 3), j = (int)(S->rot % 2);
    double *r2 = R[k], *r1 = R[(k + 2) % 3], *rn = R[(k + 1) % 3];
    const int *done = &S->ds->done;
    if (S->precon_mode) {
        // preconditioned: v = y/beta is built from y = M r2 (Y[j]); K2 writes the next y to Y[1-j]
        double *Y[2] = {solver_vec(S, "ya"), solver_vec(S, "yb")};
        double *y = Y[j], *ynext = Y[1 - j];
        MinGather g{y, S->ds, 0.0};
        MinEpiY e{rn, y, r1, S->ds, 0, 0, 0, 0};
        KRY_TRY((solver_spmv<1>(S, g, e, MinFinAlfa{S->ds}, done, y)));
        MinBodyR2Pre b2{rn, ynext, r2, S->dinv, S->precon_mode, S->ds, 0.0};
        KRY_TRY((solver_pass<1>(S, b2, MinFinQR{S->ds, S->hist}, done)));
        MinBodyW bw{W[1 - j], x, W[j], y, S->ds, 0, 0, 0, 0, 0};
        KRY_TRY((solver_pass<1>(S, bw, MinFinW{S->ds}, done)));
        S->rot++;
        return KRY_OK;
    }
    MinGather g{r2, S->ds, 0.0};
    MinEpiY e{rn, r2, r1, S->ds, 0, 0, 0, 0};
    KRY_TRY((solver_spmv<1>(S, g, e, MinFinAlfa{S->ds}, done, r2)));
    if (S->minres_fuse) {
        // 2 launches: the w / x update of the PREVIOUS trip rides in this trip's second launch.
        // Previous trip (rot-1): y_prev = R[(k+2)%3] = r1, it wrote W[1-j_prev] = W[j] from w2 = W[1-j].
        MinFinQR fin{S->ds, S->hist, 1};
        if (S->fresh) {
            MinBodyR2W<false> b{rn, r2, nullptr, x, nullptr, nullptr, S->ds, 0, 0, 0, 0, 0, 0};
            KRY_TRY((solver_pass<1>(S, b, fin, done)));
        } else {
            MinBodyR2W<true> b{rn, r2, W[j], x, W[1 - j], r1, S->ds, 0, 0, 0, 0, 0, 0};
            KRY_TRY((solver_pass<1>(S, b, fin, done)));
        }
        S->fresh = false;
        S->rot++;
        return KRY_OK;
    }
    MinBodyR2 b2{rn, nullptr, r2, S->dinv, 0, S->ds, 0.0};
    KRY_TRY((solver_pass<1>(S, b2, MinFinQR{S->ds, S->hist, 0}, done)));
    MinBodyW bw{W[1 - j], x, W[j], r2, S->ds, 0, 0, 0, 0, 0};
    KRY_TRY((solver_pass<1>(S, bw, MinFinW{S->ds}, done)));
    S->rot++;
    return KRY_OK;
}

// ---- MINRES as one cooperative persistent kernel (KRY_OPT_MINRES_PERSISTENT).  At N = 10^6
// (BASELINE config 2) a trip moves ~180 MB, i.e. ~30 us of HBM time, and three dependent launches
// cost about as much again in launch gaps, ramp-up, tails and last-CTA finalisation.  Here one
// CTA wave stays resident for the whole kry_solver_iterate call; the three phases of a trip are
// the same loops and functors as the three kernels (MinGather/MinEpiY, MinBodyR2, MinBodyW), and
// the two reductions ride in grid-wide barriers: the CTA that arrives last sums the partials in
// index order, runs the scalar step (MinFinAlfa / MinFinQR) and releases the others.  Nothing
// __shared__ is live across such a barrier.  Phase 3 needs no barrier before the next trip's
// phase 1 (disjoint buffers; the next barrier orders everything else).
template <int ND, class Fin>
__device__ __forceinline__ void grid_reduce_sync(double (&acc)[ND], const ReduceWs &ws, Fin &fin,
                                                 unsigned *gen_ptr, unsigned &my_gen)
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
    if (s_last) {
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
        if (threadIdx.x == 0) {
#pragma unroll
            for (int d = 0; d < ND; ++d) ws.sums[d] = tot[d];
            *ws.counter = 0u;
            fin(tot);
            __threadfence();
            atomicAdd(gen_ptr, 1u);                               // release
        }
    } else if (threadIdx.x == 0) {
        while (*reinterpret_cast<volatile unsigned *>(gen_ptr) == my_gen) __nanosleep(20);
        __threadfence();
    }
    my_gen++;
    __syncthreads();
}

__global__ void __launch_bounds__(256, 4)
minres_persistent_kernel(CsrView A, double *Ra, double *Rb, double *Rc, double *Wa, double *Wb, double *x,
                         DevScalars *s, double *hist, ReduceWs ws, unsigned *gen_ptr, long long n_iters)
{
    if (s->done) return;
    unsigned my_gen = *reinterpret_cast<volatile unsigned *>(gen_ptr);
    long long rot = s->n_iter;                       // trips done so far: fixes the buffer rotation
    double *R[3] = {Ra, Rb, Rc}, *W[2] = {Wa, Wb};
    const int stride = (int)(gridDim.x * blockDim.x), t0 = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int nn = A.nrows, half = nn >> 1;
    for (long long it = 0; it < n_iters; ++it, ++rot) {
        const int k = (int)(rot % 3), j = (int)(rot % 2);
        double *r2 = R[k], *r1 = R[(k + 2) % 3], *rn = R[(k + 1) % 3];
        // phase 1: y' = A v - shift v - (beta/oldb) r1 with v = y/beta ; alfa = v.y'      (K1)
        double acc[1] = {0.0};
        {
            MinGather g{r2, s, 0.0};
            MinEpiY   e{rn, r2, r1, s, 0, 0, 0, 0};
            g.init();
            e.init();
            for (int row = t0; row < nn; row += stride) {
                const int rs = __ldg(A.rowptr + row), re = __ldg(A.rowptr + row + 1);
                double sum = 0.0;
                for (int q = rs; q < re; ++q)
                    sum = __dadd_rn(sum, __dmul_rn(__ldg(A.val + q), g(__ldg(A.col + q))));
                e(row, sum, acc);
            }
            MinFinAlfa f{s};
            grid_reduce_sync<1>(acc, ws, f, gen_ptr, my_gen);
        }
        // phase 2: r2' = (-alfa/beta) r2 + y' ; beta'^2 = r2'.r2' ; QR step, norms, stopping tests   (K2)
        {
            MinBodyR2 b{rn, nullptr, r2, nullptr, 0, s, 0.0};
            b.init();
            acc[0] = 0.0;
            for (int i = t0; i < half; i += stride) b.pair(i, acc);
            if ((nn & 1) && t0 == 0) b(nn - 1, acc);
            MinFinQR f{s, hist, 2};
            grid_reduce_sync<1>(acc, ws, f, gen_ptr, my_gen);
        }
        // both written by the last CTA before it released the barrier, by nobody afterwards: uniform
        const int stop = s->done;
        if (stop && s->s[M_STOPNOW] != 0.0) break;     // beta < 0: the reference leaves before the w update
        // phase 3: w = (v - oldeps w1 - delta w2)/gamma ; x += phi w                      (K3)
        {
            MinBodyW b{W[1 - j], x, W[j], r2, s, 0, 0, 0, 0, 0};
            b.init();
            for (int i = t0; i < half; i += stride) b.pair(i, nullptr);
            if ((nn & 1) && t0 == 0) b(nn - 1, nullptr);
        }
        if (stop) break;
    }
}

static int minres_persistent_iterate(kry_solver *S, int64_t n_iters)
{
    kry_ctx *c = S->ctx;
    static int per_sm = 0;
    if (per_sm == 0) {
        int b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, minres_persistent_kernel, 256, 0) != cudaSuccess || b < 1) b = 1;
        per_sm = b;
    }
    int64_t need = (S->n + 255) / 256;
    if (need < 1) need = 1;
    const int64_t cap = (int64_t)c->sm_count * per_sm;             // one resident wave: co-residency is required
    const int grid = (int)(need < cap ? need : cap);
    KRY_TRY(kry_ctx_ensure_partials(c, grid));
    ReduceWs ws = kry_ws(c);
    CsrView A = csr_view(S->A->A);
    double *Ra = solver_vec(S, "ra"), *Rb = solver_vec(S, "rb"), *Rc = solver_vec(S, "rc");
    double *Wa = solver_vec(S, "wa"), *Wb = solver_vec(S, "wb"), *x = solver_vec(S, "x");
    DevScalars *ds = S->ds;
    double *hist = S->hist;
    unsigned *gen_ptr = c->counter + 16;                           // inside the zeroed 256-byte counter block
    long long n = (long long)n_iters;
#ifdef KRY_EMULATE
    KRY_REQUIRE(emu_fibers_on, KRY_ERR_UNSUPPORTED, "minres_persistent needs the SIMT mode of the emulation");
    auto body = [&] { minres_persistent_kernel(A, Ra, Rb, Rc, Wa, Wb, x, ds, hist, ws, gen_ptr, n); };
    emu_launch_fibers_mode(grid, 256, &body, [](const void *k) { (*static_cast<const decltype(body) *>(k))(); }, 1);
#else
    void *args[] = {&A, &Ra, &Rb, &Rc, &Wa, &Wb, &x, &ds, &hist, &ws, &gen_ptr, &n};
    KRY_CUDA(cudaLaunchCooperativeKernel((const void *)minres_persistent_kernel, dim3(grid), dim3(256), args, 0, c->stream));
#endif
    c->launches++;
    S->rot += n_iters;
    KRY_CUDA(cudaGetLastError());
    return KRY_OK;
}

static int minres_settle(kry_solver *S)
{
    if (S->method != KRY_MINRES || !S->minres_fuse || !S->ready) return KRY_OK;
    MinSettleBody b{{solver_vec(S, "ra"), solver_vec(S, "rb"), solver_vec(S, "rc")},
                    {solver_vec(S, "wa"), solver_vec(S, "wb")}, solver_vec(S, "x"), S->ds,
                    nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, 0};
    KRY_TRY(vec_map_launch(S->ctx, S->n, b, &S->ctx->never_done[0]));
    KRY_CUDA(cudaMemsetAsync(&S->ds->s[M_WPEND], 0, sizeof(double), S->ctx->stream));
    S->fresh = true;
    S->warm = false;      // the next trip must run un-captured (PEND = false variant)
    return KRY_OK;
}

// ============================================================ C ABI plumbing
struct VecSpec {
    const char *name;
    bool        gathered;    // SpMV input: needs the halo tail on sharded runs
};

static const VecSpec *method_vectors(kry_method m, int *count)
{
    static const VecSpec cg[] = {{"x", true}, {"p", true}, {"r", false}, {"Ap", false}, {"rhs", false},
                                 {"p2", true}};     // second p buffer of the fused forms (ping-pong)
    static const VecSpec bcg[] = {{"x", true}, {"p", true}, {"s", true}, {"q", true}, {"z", true},
                                  {"r0", false}, {"r", false}, {"v", false}, {"t", false}, {"rhs", false}};
    static const VecSpec cgs[] = {{"x", true}, {"p", true}, {"z", true}, {"y", true}, {"r0", false},
                                  {"r", false}, {"u", false}, {"q", false}, {"v", false}, {"rhs", false}};
    static const VecSpec tfq[] = {{"x", true}, {"y", true}, {"z", true}, {"r0", false}, {"w", false},
                                  {"d", false}, {"u", false}, {"v", false}, {"rhs", false}};
    static const VecSpec mr[] = {{"x", false}, {"ra", true}, {"rb", true}, {"rc", true},
                                 {"wa", false}, {"wb", false}, {"rhs", false},
                                 {"ya", true}, {"yb", true}};    // y = M r2 of the preconditioned form
    switch (m) {
        case KRY_CG: *count = 6; return cg;
        case KRY_BICGSTAB: *count = 10; return bcg;
        case KRY_CGS: *count = 10; return cgs;
        case KRY_TFQMR: *count = 9; return tfq;
        case KRY_MINRES: *count = 9; return mr;
    }
    *count = 0;
    return nullptr;
}

extern "C" int kry_solver_create(kry_ctx *c, kry_method method, kry_csr *A, kry_solver **out)
{
    KRY_REQUIRE(c && A && out, KRY_ERR_INVALID, "kry_solver_create: NULL argument");
    *out = nullptr;
    KRY_REQUIRE(A->ctx == c, KRY_ERR_INVALID, "kry_solver_create: operator belongs to another context");
    int nv = 0;
    const VecSpec *specs = method_vectors(method, &nv);
    KRY_REQUIRE(specs, KRY_ERR_INVALID, "kry_solver_create: unknown method %d", (int)method);
    const int64_t n = A->A.nrows;
    const int64_t n_in = A->halo.active ? A->A.nrows : A->A.ncols;
    KRY_REQUIRE(n == n_in, KRY_ERR_SHAPE, "kry_solver_create: operator is %lld x %lld, not square",
                (long long)n, (long long)n_in);
    KRY_CUDA(cudaSetDevice(c->device));
    kry_solver *S = new (std::nothrow) kry_solver();
    KRY_REQUIRE(S, KRY_ERR_NOMEM, "kry_solver_create: host allocation failed");
    memset(S, 0, sizeof(*S));
    S->ctx = c;
    S->A = A;
    S->method = method;
    S->n = n;
    S->ncap = A->halo.active ? A->A.ncols : n;
    S->sharded = A->halo.active && c->nranks > 1;
    S->hist_width = (method == KRY_CG || method == KRY_MINRES) ? 2 : 1;
    auto pad = [](int64_t v) { return (v + 31) & ~(int64_t)31; };
    int64_t total = 0;
    for (int i = 0; i < nv; ++i) total += pad(specs[i].gathered ? S->ncap : n);
    total += pad(n);   // preconditioner diagonal
    S->slab_doubles = total;
    // shards: the slab is mapped into the peers (kry_halo_link) -- it gets a driver allocation of its
    // own (>= 4 MiB, never a sub-allocation) and a magic word behind the vectors to verify the mapping
    size_t slab_bytes = (size_t)(total + 32) * sizeof(double);
    if (S->sharded && slab_bytes < ((size_t)4 << 20)) slab_bytes = (size_t)4 << 20;
    int rc = kry_alloc((void **)&S->slab, slab_bytes);
    if (rc == KRY_OK) rc = kry_alloc((void **)&S->ds, sizeof(DevScalars));
    if (rc == KRY_OK)
        rc = kry_alloc((void **)&S->hist, (size_t)KRY_HIST_CAP * S->hist_width * sizeof(double));
    if (rc != KRY_OK) {
        cudaFree(S->slab);
        cudaFree(S->ds);
        cudaFree(S->hist);
        delete S;
        return rc;
    }
    cudaMemsetAsync(S->slab, 0, (size_t)total * sizeof(double), c->stream);
    int64_t off = 0;
    for (int i = 0; i < nv; ++i) {
        S->vecs[i].name = specs[i].name;
        S->vecs[i].d = S->slab + off;
        S->vecs[i].n = n;
        off += pad(specs[i].gathered ? S->ncap : n);
    }
    S->nvecs = nv;
    S->dinv = S->slab + off;
    S->precon_mode = 0;
    kry_ctx_retain(c);
    if (S->sharded && c->p2p_inbox) {
        // best effort and collective: without the mapping the pack + ncclAllGather exchange stays in use
        rc = kry_halo_link(S);
        if (rc != KRY_OK) {
            kry_solver_destroy(S);
            return rc;
        }
    }
    *out = S;
    return KRY_OK;
}

extern "C" int kry_solver_destroy(kry_solver *S)
{
    if (!S) return KRY_OK;
    if (!S->ctx->closed) cudaStreamSynchronize(S->ctx->stream);
    solver_drop_graph(S);
    for (int k = 0; k < 2; ++k) {
        if (S->snap_host[k]) {
            cudaEventDestroy(S->snap_ev[k]);
            cudaFreeHost(S->snap_host[k]);
        }
    }
    kry_halo_unlink(S);
    cudaFree(S->slab);
    cudaFree(S->ds);
    cudaFree(S->hist);
    kry_ctx_release(S->ctx);
    delete S;
    return KRY_OK;
}

extern "C" int kry_solver_set_precon_diag(kry_solver *S, const double *diag_host, int mode)
{
    KRY_REQUIRE(S, KRY_ERR_INVALID, "kry_solver_set_precon_diag: NULL solver");
    KRY_CTX_LIVE(S->ctx, "kry_solver_set_precon_diag");
    if (!diag_host || mode == 0) {
        S->precon_mode = 0;
        return KRY_OK;
    }
    KRY_REQUIRE(mode == 1 || mode == 2, KRY_ERR_INVALID, "kry_solver_set_precon_diag: mode %d", mode);
    KRY_CUDA(cudaMemcpyAsync(S->dinv, diag_host, (size_t)S->n * sizeof(double),
                             cudaMemcpyHostToDevice, S->ctx->stream));
    KRY_CUDA(cudaStreamSynchronize(S->ctx->stream));
    S->precon_mode = mode;
    return KRY_OK;
}

static int solver_setup_common(kry_solver *S, int guess, const kry_solver_params *p)
{
    KRY_REQUIRE(p, KRY_ERR_INVALID, "kry_solver_setup: NULL params");
    KRY_REQUIRE(S->method != KRY_MINRES || (p->window >= 1 && p->window <= 16), KRY_ERR_INVALID,
                "kry_solver_setup: MINRES window %d not in [1,16]", p->window);
    S->warm = false;      // (a captured graph survives: see solver_graph_key)
    S->snap_pending[0] = S->snap_pending[1] = false;
    S->params = *p;
    DevScalars h;
    memset(&h, 0, sizeof(h));
    h.definite = 1;
    h.matvec_max = p->matvec_max;
    h.abstol = p->abstol;
    h.reltol = p->reltol;
    h.check_curvature = p->check_curvature;
    h.window = p->window;
    h.shift = p->shift;
    h.rtol = p->rtol;
    h.etol = p->etol;
    h.s[S_RHO] = 1.0;
    cudaStream_t st = S->ctx->stream;
    KRY_CUDA(cudaMemcpyAsync(S->ds, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    KRY_CUDA(cudaStreamSynchronize(st));       // `h` is on the stack
    int rc = KRY_OK;
    switch (S->method) {
        case KRY_CG: rc = cg_setup(S, guess); break;
        case KRY_BICGSTAB: rc = bicgstab_setup(S, guess); break;
        case KRY_CGS: rc = cgs_setup(S, guess); break;
        case KRY_TFQMR: rc = tfqmr_setup(S, guess); break;
        case KRY_MINRES: rc = minres_setup(S, guess); break;
    }
    S->ready = (rc == KRY_OK);
    return rc;
}

extern "C" int kry_solver_setup(kry_solver *S, const double *rhs_host, const double *guess_host,
                                const kry_solver_params *params)
{
    KRY_REQUIRE(S && rhs_host, KRY_ERR_INVALID, "kry_solver_setup: NULL argument");
    KRY_CTX_LIVE(S->ctx, "kry_solver_setup");
    KRY_CUDA(cudaSetDevice(S->ctx->device));
    cudaStream_t st = S->ctx->stream;
    const size_t bytes = (size_t)S->n * sizeof(double);
    KRY_CUDA(cudaMemcpyAsync(solver_vec(S, "rhs"), rhs_host, bytes, cudaMemcpyHostToDevice, st));
    if (guess_host)
        KRY_CUDA(cudaMemcpyAsync(solver_vec(S, "x"), guess_host, bytes, cudaMemcpyHostToDevice, st));
    else
        KRY_CUDA(cudaMemsetAsync(solver_vec(S, "x"), 0, bytes, st));
    KRY_CUDA(cudaStreamSynchronize(st));
    return solver_setup_common(S, guess_host != nullptr, params);
}

extern "C" int kry_solver_setup_dev(kry_solver *S, const kry_vec *rhs, const kry_vec *guess,
                                    const kry_solver_params *params)
{
    KRY_REQUIRE(S && rhs, KRY_ERR_INVALID, "kry_solver_setup_dev: NULL argument");
    KRY_CTX_LIVE(S->ctx, "kry_solver_setup_dev");
    KRY_REQUIRE(rhs->n == S->n && (!guess || guess->n == S->n), KRY_ERR_SHAPE,
                "kry_solver_setup_dev: rhs/guess size does not match the operator (%lld rows)",
                (long long)S->n);
    KRY_CUDA(cudaSetDevice(S->ctx->device));
    cudaStream_t st = S->ctx->stream;
    const size_t bytes = (size_t)S->n * sizeof(double);
    KRY_CUDA(cudaMemcpyAsync(solver_vec(S, "rhs"), rhs->d, bytes, cudaMemcpyDeviceToDevice, st));
    if (guess)
        KRY_CUDA(cudaMemcpyAsync(solver_vec(S, "x"), guess->d, bytes, cudaMemcpyDeviceToDevice, st));
    else
        KRY_CUDA(cudaMemsetAsync(solver_vec(S, "x"), 0, bytes, st));
    return solver_setup_common(S, guess != nullptr, params);
}

static int iterate_once(kry_solver *S)
{
    switch (S->method) {
        case KRY_CG: return cg_iterate(S);
        case KRY_BICGSTAB: return bicgstab_iterate(S);
        case KRY_CGS: return cgs_iterate(S);
        case KRY_TFQMR: return tfqmr_iterate(S);
        case KRY_MINRES: return minres_iterate(S);
    }
    return KRY_ERR_INVALID;
}

static void solver_drop_graph(kry_solver *S)
{
    if (S->graph_exec) cudaGraphExecDestroy(S->graph_exec);
    S->graph_exec = nullptr;
    S->graph_launches = 0;
}

// Everything besides the solver's own (fixed) buffers that a captured launch sequence has baked
// into its kernel arguments and grid sizes.  A graph is kept across kry_solver_setup calls --
// re-instantiating it costs more than a small solve -- and dropped only when this key changes.
static uint64_t solver_graph_key(const kry_solver *S)
{
    const kry_ctx *c = S->ctx;
    uint64_t k = 1469598103934665603ull;
    auto mix = [&k](uint64_t v) { k = (k ^ v) * 1099511628211ull; };
    mix((uint64_t)c->partials_gen);
    mix((uint64_t)c->l2_hints);
    mix((uint64_t)S->precon_mode);
    mix((uint64_t)S->cg_fuse);
    mix((uint64_t)S->minres_fuse);
    mix((uint64_t)S->A->kind);
    mix((uint64_t)S->A->tile_nnz);
    mix((uint64_t)S->A->threads);
    mix((uint64_t)(uintptr_t)S->A->A.rowblk);
    return k;
}

// Capture KRY_GRAPH_ITERS iterations of the (static) launch sequence into a graph.
static int solver_capture_graph(kry_solver *S)
{
    kry_ctx *c = S->ctx;
    const int64_t l0 = c->launches;
    const long long rot0 = S->rot;
    cudaGraph_t graph = nullptr;
    KRY_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = KRY_OK;
    for (int i = 0; i < KRY_GRAPH_ITERS && rc == KRY_OK; ++i) rc = iterate_once(S);
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    S->graph_launches = c->launches - l0;
    c->launches = l0;                      // nothing has executed yet
    S->rot = rot0;
    if (rc != KRY_OK || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        KRY_REQUIRE(rc == KRY_OK, rc, "graph capture: a launch failed");
        KRY_CUDA(e);
        return KRY_ERR_CUDA;
    }
    S->graph_key = solver_graph_key(S);
    e = cudaGraphInstantiate(&S->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    KRY_CUDA(e);
    return KRY_OK;
}

extern "C" int kry_solver_iterate(kry_solver *S, int64_t n_iters)
{
    KRY_REQUIRE(S, KRY_ERR_INVALID, "kry_solver_iterate: NULL solver");
    KRY_CTX_LIVE(S->ctx, "kry_solver_iterate");
    KRY_REQUIRE(S->ready, KRY_ERR_STATE, "kry_solver_iterate: call kry_solver_setup first");
    KRY_REQUIRE(n_iters >= 0 && n_iters < KRY_HIST_CAP / 2, KRY_ERR_INVALID,
                "kry_solver_iterate: n_iters=%lld not in [0,%d)", (long long)n_iters, KRY_HIST_CAP / 2);
    kry_ctx *c = S->ctx;
    KRY_CUDA(cudaSetDevice(c->device));
    if (S->method == KRY_CG && S->one_cta) return n_iters > 0 ? cg_one_cta_iterate(S, n_iters) : KRY_OK;
    if (S->method == KRY_MINRES && S->minres_persistent)
        return n_iters > 0 ? minres_persistent_iterate(S, n_iters) : KRY_OK;
    int64_t left = n_iters;
    // Graph replay: not on sharded runs (NCCL in the sequence), not while per-launch
    // profiling events are being recorded, and only once the sequence ran un-captured
    // (first-use allocations / attribute calls must not happen inside a capture).
    const bool graphs = c->use_graphs && !S->sharded && c->prof_cap == 0;
    if (graphs && left >= KRY_GRAPH_ITERS + 6) {
        while (left > 0 && (!S->warm || S->rot % 6 != 0)) {
            KRY_TRY(iterate_once(S));
            S->warm = true;
            --left;
        }
        if (S->graph_exec && S->graph_key != solver_graph_key(S)) solver_drop_graph(S);
        if (!S->graph_exec && left >= KRY_GRAPH_ITERS) KRY_TRY(solver_capture_graph(S));
        while (S->graph_exec && left >= KRY_GRAPH_ITERS) {
            KRY_CUDA(cudaGraphLaunch(S->graph_exec, c->stream));
            c->launches += S->graph_launches;
            S->rot += KRY_GRAPH_ITERS;
            left -= KRY_GRAPH_ITERS;
        }
    }
    for (; left > 0; --left) {
        KRY_TRY(iterate_once(S));
        S->warm = true;
    }
    return KRY_OK;
}

static void status_decode(const kry_solver *S, const DevScalars &h, kry_solver_status *out);

extern "C" int kry_solver_status_read(kry_solver *S, kry_solver_status *out)
{
    KRY_REQUIRE(S && out, KRY_ERR_INVALID, "kry_solver_status_read: NULL argument");
    KRY_CTX_LIVE(S->ctx, "kry_solver_status_read");
    DevScalars h;
    KRY_CUDA(cudaMemcpyAsync(&h, S->ds, sizeof(h), cudaMemcpyDeviceToHost, S->ctx->stream));
    KRY_CUDA(cudaStreamSynchronize(S->ctx->stream));
    KRY_CUDA(cudaGetLastError());
    status_decode(S, h, out);
    return KRY_OK;
}

extern "C" int kry_solver_status_enqueue(kry_solver *S, int slot)
{
    KRY_REQUIRE(S && (slot == 0 || slot == 1), KRY_ERR_INVALID, "kry_solver_status_enqueue: bad argument");
    KRY_CTX_LIVE(S->ctx, "kry_solver_status_enqueue");
    KRY_REQUIRE(S->ready, KRY_ERR_STATE, "kry_solver_status_enqueue: call kry_solver_setup first");
    if (!S->snap_host[slot]) {
        KRY_CUDA(cudaMallocHost((void **)&S->snap_host[slot], sizeof(DevScalars)));
        KRY_CUDA(cudaEventCreateWithFlags(&S->snap_ev[slot], cudaEventDisableTiming));
    }
    KRY_CUDA(cudaMemcpyAsync(S->snap_host[slot], S->ds, sizeof(DevScalars), cudaMemcpyDeviceToHost,
                             S->ctx->stream));
    KRY_CUDA(cudaEventRecord(S->snap_ev[slot], S->ctx->stream));
    S->snap_pending[slot] = true;
    return KRY_OK;
}

extern "C" int kry_solver_status_wait(kry_solver *S, int slot, kry_solver_status *out)
{
    KRY_REQUIRE(S && out && (slot == 0 || slot == 1), KRY_ERR_INVALID, "kry_solver_status_wait: bad argument");
    KRY_CTX_LIVE(S->ctx, "kry_solver_status_wait");
    KRY_REQUIRE(S->snap_pending[slot], KRY_ERR_STATE, "kry_solver_status_wait: nothing enqueued in slot %d", slot);
    KRY_CUDA(cudaEventSynchronize(S->snap_ev[slot]));
    S->snap_pending[slot] = false;
    status_decode(S, *S->snap_host[slot], out);
    return KRY_OK;
}

static void status_decode(const kry_solver *S, const DevScalars &h, kry_solver_status *out)
{
    memset(out, 0, sizeof(*out));
    out->done = h.done;
    out->definite = h.definite;
    out->istop = h.istop;
    out->n_matvec = h.n_matvec;
    out->n_iter = h.n_iter;
    out->hist_count = h.hist_count;
    out->resid_norm0 = h.resid0;
    out->resid_norm = h.resid;
    out->threshold = h.threshold;
    switch (S->method) {
        case KRY_MINRES:
            out->converged = (h.istop == 1 || h.istop == 2 || h.istop == 3 || h.istop == 4 ||
                              h.istop == 10);                                // minres.py:395
            out->aux[0] = h.s[M_ANORM];
            out->aux[1] = h.s[M_ACOND];
            out->aux[2] = h.s[M_YNORM];
            out->aux[3] = h.s[M_ARNORM];
            out->aux[4] = h.s[M_BETA1];
            out->aux[5] = h.s[M_ALFA];
            out->aux[6] = h.s[M_BETA];
            out->aux[7] = h.s[M_PHI];
            out->aux[8] = h.s[M_TEST1];
            out->aux[9] = h.s[M_TEST2];
            out->aux[10] = h.s[M_TRNC];
            out->aux[11] = h.s[M_XNRG2];
            out->aux[12] = h.s[M_GBAR];
            break;
        case KRY_TFQMR:
            out->converged = (h.n_iter > 0 || h.hist_count > 1)
                                 ? (h.resid * sqrt(h.s[S_M] + 1) < h.threshold)   // tfqmr.py:156
                                 : 0;
            out->aux[0] = h.s[S_RHO];
            out->aux[1] = h.s[S_SIGMA];
            out->aux[2] = h.s[S_ALPHA];
            out->aux[3] = h.s[S_BETA];
            out->aux[4] = h.s[S_THETA];
            out->aux[5] = h.s[S_ETA];
            out->aux[6] = h.s[S_M];
            break;
        default:
            out->converged = (h.resid <= h.threshold);                       // cg.py:161 etc.
            out->aux[0] = h.s[S_RY];
            out->aux[1] = h.s[S_PAP];
            out->aux[2] = h.s[S_ALPHA];
            out->aux[3] = h.s[S_BETA];
            out->aux[4] = h.s[S_RHO];
            out->aux[5] = h.s[S_OMEGA];
            out->aux[6] = h.s[S_SIGMA];
            out->aux[7] = h.s[S_RHO_NEXT];
            break;
    }
}

static int solver_history_on(kry_solver *S, cudaStream_t st, int64_t first, int64_t count, double *host,
                             int32_t *width);

extern "C" int kry_solver_history(kry_solver *S, int64_t first, int64_t count, double *host,
                                  int32_t *width)
{
    KRY_REQUIRE(S && (host || count == 0), KRY_ERR_INVALID, "kry_solver_history: NULL argument");
    KRY_CTX_LIVE(S->ctx, "kry_solver_history");
    return solver_history_on(S, S->ctx->stream, first, count, host, width);
}

extern "C" int kry_solver_history_nowait(kry_solver *S, int64_t first, int64_t count, double *host,
                                         int32_t *width)
{
    KRY_REQUIRE(S && (host || count == 0), KRY_ERR_INVALID, "kry_solver_history_nowait: NULL argument");
    KRY_CTX_LIVE(S->ctx, "kry_solver_history_nowait");
    kry_ctx *c = S->ctx;
    if (!c->copy_stream) KRY_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    return solver_history_on(S, c->copy_stream, first, count, host, width);
}

static int solver_history_on(kry_solver *S, cudaStream_t st, int64_t first, int64_t count, double *host,
                             int32_t *width)
{
    if (width) *width = S->hist_width;
    KRY_REQUIRE(first >= 0 && count >= 0 && count <= KRY_HIST_CAP, KRY_ERR_INVALID,
                "kry_solver_history: range [%lld,+%lld) invalid", (long long)first, (long long)count);
    const int w = S->hist_width;
    int64_t done = 0;
    while (done < count) {      // the ring may wrap
        const int64_t slot = (first + done) % KRY_HIST_CAP;
        int64_t run = KRY_HIST_CAP - slot;
        if (run > count - done) run = count - done;
        KRY_CUDA(cudaMemcpyAsync(host + done * w, S->hist + slot * w, (size_t)run * w * sizeof(double),
                                 cudaMemcpyDeviceToHost, st));
        done += run;
    }
    KRY_CUDA(cudaStreamSynchronize(st));
    return KRY_OK;
}

extern "C" int kry_solver_solution(kry_solver *S, double *x_host)
{
    KRY_REQUIRE(S && x_host, KRY_ERR_INVALID, "kry_solver_solution: NULL argument");
    KRY_CTX_LIVE(S->ctx, "kry_solver_solution");
    KRY_TRY(cg_settle(S));
    KRY_TRY(minres_settle(S));
    KRY_CUDA(cudaMemcpyAsync(x_host, solver_vec(S, "x"), (size_t)S->n * sizeof(double),
                             cudaMemcpyDeviceToHost, S->ctx->stream));
    KRY_CUDA(cudaStreamSynchronize(S->ctx->stream));
    return KRY_OK;
}

// MINRES rotates its buffers; map the reference's names onto the rotation that
// the device actually reached (trips executed, not trips enqueued).
static double *solver_vec_logical(kry_solver *S, const char *name)
{
    if (S->method == KRY_MINRES) {
        long long itn = 0;
        cudaMemcpyAsync(&itn, &S->ds->n_iter, sizeof(itn), cudaMemcpyDeviceToHost, S->ctx->stream);
        cudaStreamSynchronize(S->ctx->stream);
        const char *R[3] = {"ra", "rb", "rc"}, *W[2] = {"wa", "wb"};
        const int k = (int)(itn % 3), j = (int)(itn % 2);
        if (!strcmp(name, "y") && S->precon_mode) return solver_vec(S, (itn % 2) ? "yb" : "ya");
        if (!strcmp(name, "r2") || !strcmp(name, "y")) return solver_vec(S, R[k]);
        if (!strcmp(name, "r1")) return solver_vec(S, R[(k + 2) % 3]);
        if (!strcmp(name, "w")) return solver_vec(S, W[j]);
        if (!strcmp(name, "w2")) return solver_vec(S, W[1 - j]);
    }
    if (S->method == KRY_CG && S->cg_fuse && !strcmp(name, "p")) {
        // after cg_settle(): p lives in P[(n_iter+1)&1] with P[1] = "p", P[0] = "p2"
        long long itn = 0;
        cudaMemcpyAsync(&itn, &S->ds->n_iter, sizeof(itn), cudaMemcpyDeviceToHost, S->ctx->stream);
        cudaStreamSynchronize(S->ctx->stream);
        return solver_vec(S, ((itn + 1) & 1) ? "p" : "p2");
    }
    return solver_vec(S, name);
}

extern "C" int kry_solver_get_vector(kry_solver *S, const char *name, double *host)
{
    KRY_REQUIRE(S && name && host, KRY_ERR_INVALID, "kry_solver_get_vector: NULL argument");
    KRY_CTX_LIVE(S->ctx, "kry_solver_get_vector");
    KRY_TRY(cg_settle(S));
    KRY_TRY(minres_settle(S));
    double *d = solver_vec_logical(S, name);
    KRY_REQUIRE(d, KRY_ERR_INVALID, "kry_solver_get_vector: no vector named '%s'", name);
    KRY_CUDA(cudaMemcpyAsync(host, d, (size_t)S->n * sizeof(double), cudaMemcpyDeviceToHost,
                             S->ctx->stream));
    KRY_CUDA(cudaStreamSynchronize(S->ctx->stream));
    return KRY_OK;
}

extern "C" int kry_solver_set_vector(kry_solver *S, const char *name, const double *host)
{
    KRY_REQUIRE(S && name && host, KRY_ERR_INVALID, "kry_solver_set_vector: NULL argument");
    KRY_CTX_LIVE(S->ctx, "kry_solver_set_vector");
    KRY_TRY(cg_settle(S));
    KRY_TRY(minres_settle(S));
    double *d = solver_vec_logical(S, name);
    KRY_REQUIRE(d, KRY_ERR_INVALID, "kry_solver_set_vector: no vector named '%s'", name);
    KRY_CUDA(cudaMemcpyAsync(d, host, (size_t)S->n * sizeof(double), cudaMemcpyHostToDevice,
                             S->ctx->stream));
    KRY_CUDA(cudaStreamSynchronize(S->ctx->stream));
    return KRY_OK;
}

struct ScalarName {
    const char *name;
    int         idx;
};
static const ScalarName kScalarNames[] = {
    {"ry", S_RY}, {"pAp", S_PAP}, {"alpha", S_ALPHA}, {"beta", S_BETA}, {"rho", S_RHO},
    {"rho_next", S_RHO_NEXT}, {"omega", S_OMEGA}, {"sigma", S_SIGMA}, {"theta", S_THETA},
    {"eta", S_ETA}, {"k", S_K}, {"m", S_M}, {"alfa", M_ALFA}, {"mbeta", M_BETA}, {"oldb", M_OLDB},
    {"beta1", M_BETA1}, {"dbar", M_DBAR}, {"epsln", M_EPSLN}, {"phibar", M_PHIBAR},
    {"cs", M_CS}, {"sn", M_SN}, {"phi", M_PHI}, {"tnorm2", M_TNORM2}, {"ynorm2", M_YNORM2},
    {"rhs1", M_RHS1}, {"rhs2", M_RHS2}, {"gmax", M_GMAX}, {"gmin", M_GMIN}};

static int scalar_index(const char *name)
{
    for (size_t i = 0; i < sizeof(kScalarNames) / sizeof(kScalarNames[0]); ++i)
        if (!strcmp(kScalarNames[i].name, name)) return kScalarNames[i].idx;
    return -1;
}

// Address of a named scalar inside the device block.  Besides the recurrence
// scalars this covers the counters and the MINRES error window so that the
// single-step parity tests can transplant a complete state.
static int scalar_locate(kry_solver *S, const char *name, void **addr, int *kind)
{
    *kind = 0;                                    // 0: double, 1: long long, 2: int
    const int idx = scalar_index(name);
    if (idx >= 0) { *addr = &S->ds->s[idx]; return KRY_OK; }
    if (!strcmp(name, "resid")) { *addr = &S->ds->resid; return KRY_OK; }
    if (!strcmp(name, "resid0")) { *addr = &S->ds->resid0; return KRY_OK; }
    if (!strcmp(name, "threshold")) { *addr = &S->ds->threshold; return KRY_OK; }
    if (!strcmp(name, "xnrg2")) { *addr = &S->ds->s[M_XNRG2]; return KRY_OK; }
    if (!strcmp(name, "n_iter")) { *addr = &S->ds->n_iter; *kind = 1; return KRY_OK; }
    if (!strcmp(name, "n_matvec")) { *addr = &S->ds->n_matvec; *kind = 1; return KRY_OK; }
    if (!strcmp(name, "istop")) { *addr = &S->ds->istop; *kind = 2; return KRY_OK; }
    if (!strcmp(name, "done")) { *addr = &S->ds->done; *kind = 2; return KRY_OK; }
    if (!strncmp(name, "derr", 4)) {
        const int k = atoi(name + 4);
        if (k >= 0 && k < 16) { *addr = &S->ds->derr[k]; return KRY_OK; }
    }
    kry_set_error("no scalar named '%s'", name);
    return KRY_ERR_INVALID;
}

extern "C" int kry_solver_get_scalar(kry_solver *S, const char *name, double *value)
{
    KRY_REQUIRE(S && name && value, KRY_ERR_INVALID, "kry_solver_get_scalar: NULL argument");
    KRY_CTX_LIVE(S->ctx, "kry_solver_get_scalar");
    void *addr = nullptr;
    int kind = 0;
    KRY_TRY(scalar_locate(S, name, &addr, &kind));
    double d = 0.0;
    long long ll = 0;
    int i = 0;
    void *dst = kind == 0 ? (void *)&d : (kind == 1 ? (void *)&ll : (void *)&i);
    const size_t bytes = kind == 0 ? sizeof(d) : (kind == 1 ? sizeof(ll) : sizeof(i));
    KRY_CUDA(cudaMemcpyAsync(dst, addr, bytes, cudaMemcpyDeviceToHost, S->ctx->stream));
    KRY_CUDA(cudaStreamSynchronize(S->ctx->stream));
    *value = kind == 0 ? d : (kind == 1 ? (double)ll : (double)i);
    return KRY_OK;
}

extern "C" int kry_solver_set_scalar(kry_solver *S, const char *name, double value)
{
    KRY_REQUIRE(S && name, KRY_ERR_INVALID, "kry_solver_set_scalar: NULL argument");
    KRY_CTX_LIVE(S->ctx, "kry_solver_set_scalar");
    void *addr = nullptr;
    int kind = 0;
    KRY_TRY(scalar_locate(S, name, &addr, &kind));
    double d = value;
    long long ll = (long long)value;
    int i = (int)value;
    const void *src = kind == 0 ? (void *)&d : (kind == 1 ? (void *)&ll : (void *)&i);
    const size_t bytes = kind == 0 ? sizeof(d) : (kind == 1 ? sizeof(ll) : sizeof(i));
    KRY_CUDA(cudaMemcpyAsync(addr, src, bytes, cudaMemcpyHostToDevice, S->ctx->stream));
    KRY_CUDA(cudaStreamSynchronize(S->ctx->stream));
    if (!strcmp(name, "n_iter")) S->rot = ll;      // keep the MINRES buffer rotation in step
    return KRY_OK;
}
