// vecops.cuh -- the fused multi-AXPY body and the SpMV epilogue  z = a (A x) + b w  (+ dot), shared
// by the generic entry points (ops.cu: results into the scalar slots) and the lls / SYMMLQ planes
// (lls.cu: the same launches with the scalar recurrence of the next phase in their finalize).
#pragma once
#include "launch.cuh"

// ------------------------------------------------------------- multi-AXPY
struct AxpbyDev {
    double       *z;
    const double *u, *w;
    double        a, b;
    int           a_slot, b_slot, a_neg, b_neg;
};

template <int ND>
struct MultiAxpyBody {
    AxpbyDev      op[4];
    int           n_ops;
    const double *du[ND > 0 ? ND : 1], *dw[ND > 0 ? ND : 1];
    const double *slots;
    double        ca[4], cb[4];      // resolved coefficients (registers: all loops are unrolled)
    __device__ void init()
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ca[k] = op[k].a;
            cb[k] = op[k].b;
            if (k < n_ops) {
                if (op[k].a_slot >= 0) ca[k] = slots[op[k].a_slot];
                if (op[k].b_slot >= 0) cb[k] = slots[op[k].b_slot];
                if (op[k].a_neg & 1) ca[k] = -ca[k];
                if (op[k].b_neg & 1) cb[k] = -cb[k];
            }
        }
    }
    // coefficient applied as a multiplier, or as a divisor when bit 1 of the flag is set
    // (`u /= beta` in the reference is a true division, not a multiply by 1/beta)
    __device__ static double term(double c, double v, int flag)
    {
        return (flag & 2) ? __ddiv_rn(v, c) : __dmul_rn(c, v);
    }
    __device__ void update(int i) const
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= n_ops) break;
            const double *u = op[k].u, *w = op[k].w;
            double r;
            if (u && w)
                r = __dadd_rn(term(ca[k], u[i], op[k].a_neg), term(cb[k], w[i], op[k].b_neg));
            else if (u)
                r = term(ca[k], u[i], op[k].a_neg);
            else if (w)
                r = term(cb[k], w[i], op[k].b_neg);
            else
                r = 0.0;
            op[k].z[i] = r;
        }
    }
    __device__ void operator()(int i, double *acc) const
    {
        update(i);
#pragma unroll
        for (int d = 0; d < ND; ++d) acc[d] = __dadd_rn(acc[d], __dmul_rn(du[d][i], dw[d][i]));
    }
    __device__ void operator()(int i) const { update(i); }

    // Elements 2*i2 and 2*i2+1 with 16-byte accesses (the vector kernels run this form: twice the bytes
    // in flight per thread).  Ops run in order on each element, later ops see the outputs of earlier
    // ones (same thread, same addresses), exactly like update().
    static constexpr bool kPair = true;
    static constexpr int  kMinBlocks = 4;      // up to 64 registers: the 8 resolved coefficients stay in registers
    __device__ void update2(int i2) const
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= n_ops) break;
            const double *u = op[k].u, *w = op[k].w;
            double2 r = make_double2(0.0, 0.0);
            if (u && w) {
                const double2 a = ld2(u, i2), b = ld2(w, i2);
                r.x = __dadd_rn(term(ca[k], a.x, op[k].a_neg), term(cb[k], b.x, op[k].b_neg));
                r.y = __dadd_rn(term(ca[k], a.y, op[k].a_neg), term(cb[k], b.y, op[k].b_neg));
            } else if (u) {
                const double2 a = ld2(u, i2);
                r.x = term(ca[k], a.x, op[k].a_neg);
                r.y = term(ca[k], a.y, op[k].a_neg);
            } else if (w) {
                const double2 b = ld2(w, i2);
                r.x = term(cb[k], b.x, op[k].b_neg);
                r.y = term(cb[k], b.y, op[k].b_neg);
            }
            st2(op[k].z, i2, r);
        }
    }
    __device__ void pair(int i2, double *acc) const
    {
        update2(i2);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const double2 a = ld2(du[d], i2), b = ld2(dw[d], i2);
            acc[d] = __dadd_rn(acc[d], __dmul_rn(a.x, b.x));
            acc[d] = __dadd_rn(acc[d], __dmul_rn(a.y, b.y));
        }
    }
    __device__ void pair(int i2) const { update2(i2); }
};


// SpMV epilogue: z[row] = a * (A x)[row] + b * w[row], then (ND == 1) accumulates dw[row] * z[row]
// (dw == nullptr: z . z).  Same coefficient rules and the same un-fused arithmetic as an op of
// MultiAxpyBody whose `u` is the product: fusing it changes no bit of z.
template <int ND, int MINB = 8, bool DIV = true>
struct EpiAxpbyDot {
    // Resident CTAs per SM the row kernels are compiled for: 8 (32 registers) like the plain row kernel.
    // With the lls recurrence in the finalize (LlsFin, a call) the kernel would otherwise be allocated
    // the callee's 62 registers and the latency-bound row loop would lose half its warps; lls.cu asks
    // for 6 (40 registers: no spill in the row loop, the callee spills instead, once per launch).
    static constexpr int kMinBlocks = MINB;
    double       *z;
    const double *w, *dw;
    const double *slots;
    double        a, b;
    int           a_slot, b_slot, a_neg, b_neg;
    double        ca, cb;
    __device__ void init()
    {
        ca = a_slot >= 0 ? slots[a_slot] : a;
        cb = b_slot >= 0 ? slots[b_slot] : b;
        if (a_neg & 1) ca = -ca;
        if (b_neg & 1) cb = -cb;
    }
    __device__ void operator()(int row, double ax, double *acc) const
    {
        double r;
        if constexpr (DIV) {
            r = MultiAxpyBody<0>::term(ca, ax, a_neg);
            if (w) r = __dadd_rn(r, MultiAxpyBody<0>::term(cb, w[row], b_neg));
        } else {                                  // the caller checked that no coefficient divides
            r = __dmul_rn(ca, ax);
            if (w) r = __dadd_rn(r, __dmul_rn(cb, w[row]));
        }
        z[row] = r;
        if constexpr (ND > 0) acc[0] = __dadd_rn(acc[0], __dmul_rn(dw ? dw[row] : r, r));
    }
};

// argument checks shared by kry_multi_axpy_dot and kry_lls_multi_axpy_dot; n = common length
int multi_axpy_check(const char *who, kry_ctx *c, int n_ops, const kry_axpby *ops, int n_dots,
                     const kry_dotspec *dots, int slot0, int64_t *n_out);
// argument checks of the fused SpMV epilogue
int spmv_axpby_check(const char *who, kry_csr *A, int trans, const kry_vec *x, const kry_axpby *op,
                     int n_dots, const kry_vec *dot_with, int slot0);

template <int ND>
static void multi_axpy_fill(MultiAxpyBody<ND> &b, kry_ctx *c, int n_ops, const kry_axpby *ops, const kry_dotspec *dots)
{
    b.n_ops = n_ops;
    b.slots = c->scalars;
    for (int k = 0; k < n_ops; ++k) {
        b.op[k].z = ops[k].z->d;
        b.op[k].u = ops[k].u ? ops[k].u->d : nullptr;
        b.op[k].w = ops[k].w ? ops[k].w->d : nullptr;
        b.op[k].a = ops[k].a;
        b.op[k].b = ops[k].b;
        b.op[k].a_slot = ops[k].a_slot;
        b.op[k].b_slot = ops[k].b_slot;
        b.op[k].a_neg = ops[k].a_neg;
        b.op[k].b_neg = ops[k].b_neg;
    }
    b.du[0] = b.dw[0] = nullptr;
    for (int d = 0; d < ND; ++d) {
        b.du[d] = dots[d].u->d;
        b.dw[d] = dots[d].w->d;
    }
}

// one fused vector pass; `fin` receives the ND totals (all-reduced first on sharded contexts)
template <int ND, class Fin>
static int multi_axpy_run(kry_ctx *c, int64_t n, int n_ops, const kry_axpby *ops, const kry_dotspec *dots, Fin fin)
{
    MultiAxpyBody<ND> b;
    multi_axpy_fill<ND>(b, c, n_ops, ops, dots);
    if constexpr (ND == 0) {
        return vec_map_launch(c, n, b, kry_gate(c));
    } else {
        if (c->nranks > 1) {
            extern int kry_allreduce_sums(kry_ctx * c, int n);
            KRY_TRY((vec_pass_launch<ND>(c, n, b, fin, kry_gate(c), 1)));
            KRY_TRY(kry_allreduce_sums(c, ND));
            return finalize_launch(c, fin, kry_gate(c));
        }
        return vec_pass_launch<ND>(c, n, b, fin, kry_gate(c), 0);
    }
}

// y-side fused product  z = a (A x) + b w  with ND (0 or 1) inner products into `fin`
template <int ND, int MINB = 8, bool DIV = true, class Fin>
static int spmv_axpby_run(kry_csr *A, int trans, const kry_vec *x, const kry_axpby *op, const kry_vec *dot_with, Fin fin)
{
    kry_ctx *c = A->ctx;
    GatherPlain g{x->d};
    EpiAxpbyDot<ND, MINB, DIV> e;
    e.z = op->z->d;
    e.w = op->w ? op->w->d : nullptr;
    e.dw = dot_with ? dot_with->d : nullptr;
    e.slots = c->scalars;
    e.a = op->a;
    e.b = op->b;
    e.a_slot = op->a_slot;
    e.b_slot = op->b_slot;
    e.a_neg = op->a_neg;
    e.b_neg = op->b_neg;
    e.ca = e.cb = 0.0;
    if constexpr (ND > 0) {
        if (c->nranks > 1 && A->halo.active) {
            extern int kry_allreduce_sums(kry_ctx * c, int n);
            KRY_TRY(spmv_launch<ND>(A, trans != 0, g, e, fin, kry_gate(c), 1));
            KRY_TRY(kry_allreduce_sums(c, ND));
            return finalize_launch(c, fin, kry_gate(c));
        }
    }
    return spmv_launch<ND>(A, trans != 0, g, e, fin, kry_gate(c), 0);
}
