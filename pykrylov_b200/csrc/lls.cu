// lls.cu -- device-resident scalar planes of LSQR, LSMR, CRAIG, CRAIG-MR and SYMMLQ.
//
// The vector work of these solvers already runs as CUDA kernels (CSR SpMV with A / A^T and
// the fused multi-AXPY + dot launches of ops.cu).  What the reference does *between* them --
// the plane rotations, norm estimates and stopping tests of pykrylov/lls/lsqr.py:277-390,
// lls/lsmr.py:337-475, lls/craig.py:314-455, lls/craigmr.py:159-215 and
// symmlq/symmlq.py:235-355 -- is a few dozen flops on ~40 scalars.  Here that scalar plane runs
// on the device too: one single-thread step kernel per phase reads the inner products the
// previous launch left in the context's scalar slots, advances the recurrence (same
// expressions, same operation order, no FMA: the library is built with -fmad=false) and
// writes the coefficients of the next vector launches back into slots, which those launches
// read on the device (kry_axpby::a_slot / b_slot).  The host enqueues whole iterations without
// ever needing a scalar; it reads one status block per check interval.  Once the stopping test
// of the reference's loop fires, `done` is latched and every later launch of the sequence --
// vector kernels included, through the context's gate -- is a no-op, so results do not depend
// on the check interval.
#include <math.h>
#include <string.h>

#include <new>

#include "vecops.cuh"

constexpr int LLS_HIST_CAP = 1 << 15;
constexpr int LLS_HIST_W = 4;
constexpr int LLS_NV = 64;
#ifndef LLS_SPMV_MINB
#define LLS_SPMV_MINB 6
#endif

// coefficient slots in the context's scalar block (0..3 receive the fused inner products)
enum {
    SL_D0 = 0, SL_D1 = 1, SL_D2 = 2,
    SL_ALPHA = 8,      // alpha of the Golub-Kahan process (used as -alpha by  Mu = A v - alpha Mu)
    SL_U_DIV,          // u /= beta        (1 when beta == 0: the reference skips the scaling)
    SL_NV_A, SL_NV_B,  // Nv = NV_A * (A^T u) + NV_B * Nv    ((1, -beta), or (0, 1) when beta == 0)
    SL_V_DIV,          // v /= alpha       (1 when the reference skips it)
    SL_C0, SL_C1, SL_C2, SL_C3, SL_C4, SL_C5, SL_C6, SL_C7      // method specific
};

struct LlsDev {
    int       done, istop, method, window;
    long long itn, nmatvec, hist_count, itnlim;
    double    damp, atol, btol, ctol, etol, rtol, shift, eps;
    double    v[LLS_NV];
    double    derr[16];
};

struct kry_lls {
    kry_ctx *ctx;
    LlsDev  *dev;
    double  *hist;
    int      method;
};

// ---------------------------------------------------------------- scalar names (setup / status)
#define LSQR_NAMES(X) X(alpha) X(beta) X(rhobar) X(phibar) X(bnorm) X(Anorm) X(Acond) X(ddnorm) X(res2) X(xnorm) \
    X(xxnorm) X(z) X(cs2) X(sn2) X(rnorm) X(r1norm) X(r2norm) X(Arnorm) X(xNrgNorm2) X(trncDirErr) X(psi) X(phi) \
    X(rho) X(theta) X(tau)
#define LSMR_NAMES(X) X(alpha) X(beta) X(zetabar) X(alphabar) X(rho) X(rhobar) X(cbar) X(sbar) X(betadd) X(betad) \
    X(rhodold) X(tautildeold) X(thetatilde) X(zeta) X(d) X(normA2) X(maxrbar) X(minrbar) X(normA) X(condA) X(normx) \
    X(xNrgNorm2) X(trncDirErr) X(normb) X(normr) X(normar) X(chat) X(shat) X(c) X(s) X(rhoold) X(rhobarold) \
    X(zetaold) X(thetabar) X(rhotemp) X(thetanew)
#define CRAIG_NAMES(X) X(alpha) X(beta) X(c) X(s) X(tau) X(zeta) X(eta) X(xi) X(rnorm) X(xnorm) X(r1norm) X(r2norm) \
    X(Arnorm) X(rNrgNorm2) X(xNrgNorm2) X(trncDirErr) X(bnorm) X(delta) X(beta_hat) X(alpha_hat) X(s2)
#define CRAIGMR_NAMES(X) X(alpha) X(beta) X(c) X(s) X(zeta_hat) X(alpha_tilde) X(theta) X(xNrgNorm2) X(trncDirErr) \
    X(zeta) X(beta_hat) X(alpha_hat) X(rho) X(theta_old)
#define SYMMLQ_NAMES(X) X(beta1) X(beta) X(oldb) X(alfa) X(tnorm) X(ynorm2) X(gbar) X(dbar) X(rhs1) X(rhs2) X(snprod) \
    X(bstep) X(gmax) X(gmin) X(anorm) X(ynorm) X(acond) X(cgnorm) X(qrnorm) X(lqnorm) X(diag) X(zbar) X(z) X(epsa) \
    X(epsx) X(epsr) X(cs) X(sn)

#define AS_ENUM_Q(n) Q_##n,
#define AS_ENUM_M(n) M_##n,
#define AS_ENUM_C(n) C_##n,
#define AS_ENUM_R(n) R_##n,
#define AS_ENUM_Y(n) Y_##n,
#define AS_STR(n) #n,
enum { LSQR_NAMES(AS_ENUM_Q) Q_COUNT };
enum { LSMR_NAMES(AS_ENUM_M) M_COUNT };
enum { CRAIG_NAMES(AS_ENUM_C) C_COUNT };
enum { CRAIGMR_NAMES(AS_ENUM_R) R_COUNT };
enum { SYMMLQ_NAMES(AS_ENUM_Y) Y_COUNT };
static_assert(Q_COUNT <= LLS_NV && M_COUNT <= LLS_NV && C_COUNT <= LLS_NV && R_COUNT <= LLS_NV && Y_COUNT <= LLS_NV,
              "LlsDev::v too small");
static const char *const kLsqrNames[] = {LSQR_NAMES(AS_STR) nullptr};
static const char *const kLsmrNames[] = {LSMR_NAMES(AS_STR) nullptr};
static const char *const kCraigNames[] = {CRAIG_NAMES(AS_STR) nullptr};
static const char *const kCraigmrNames[] = {CRAIGMR_NAMES(AS_STR) nullptr};
static const char *const kSymmlqNames[] = {SYMMLQ_NAMES(AS_STR) nullptr};

static const char *const *lls_names(int method)
{
    switch (method) {
        case KRY_LLS_LSQR: return kLsqrNames;
        case KRY_LLS_LSMR: return kLsmrNames;
        case KRY_LLS_CRAIG: return kCraigNames;
        case KRY_LLS_CRAIGMR: return kCraigmrNames;
        case KRY_LLS_SYMMLQ: return kSymmlqNames;
    }
    return nullptr;
}

#if defined(__CUDACC__) || defined(KRY_EMULATE)
// ---------------------------------------------------------------- helpers
__device__ static inline double normof2(double x, double y) { return sqrt(x * x + y * y); }
__device__ static inline double normof4(double a, double b, double c, double d)
{
    return sqrt(a * a + b * b + c * c + d * d);
}
__device__ static inline double sgn(double a) { return a > 0 ? 1.0 : (a < 0 ? -1.0 : 0.0); }   // np.sign

// Stable Givens rotation, lsmr.py:500-518
__device__ static inline void sym_ortho(double a, double b, double &c, double &s, double &r)
{
    if (b == 0) {
        c = sgn(a); s = 0; r = fabs(a);
    } else if (a == 0) {
        c = 0; s = sgn(b); r = fabs(b);
    } else if (fabs(b) > fabs(a)) {
        const double tau = a / b;
        s = sgn(b) / sqrt(1 + tau * tau);
        c = s * tau;
        r = b / s;
    } else {
        const double tau = b / a;
        c = sgn(a) / sqrt(1 + tau * tau);
        s = c * tau;
        r = a / c;
    }
}

__device__ static inline void lls_hist(LlsDev *st, double *hist, double a, double b, double c, double d)
{
    double *h = hist + (st->hist_count % LLS_HIST_CAP) * LLS_HIST_W;
    h[0] = a; h[1] = b; h[2] = c; h[3] = d;
    st->hist_count++;
}

// np.linalg.norm(dErr): sqrt of the ordered sum of squares
__device__ static inline double window_norm(const LlsDev *st)
{
    double ss = 0.0;
    for (int k = 0; k < st->window; ++k) ss += st->derr[k] * st->derr[k];
    return sqrt(ss);
}

// first half of a Golub-Kahan step, shared by LSQR / LSMR / CRAIG / CRAIG-MR:
//   beta M u = A v - alpha M u   (lsqr.py:243-256): beta = |u| arrives in slot D0
__device__ static inline void gk_beta(double *sl, double &beta)
{
    beta = sqrt(sl[SL_D0]);
    if (beta > 0) {
        sl[SL_U_DIV] = beta;
        sl[SL_NV_A] = 1.0;
        sl[SL_NV_B] = -beta;
    } else {                      // the reference skips the scaling and the A^T product
        sl[SL_U_DIV] = 1.0;
        sl[SL_NV_A] = 0.0;
        sl[SL_NV_B] = 1.0;
    }
}
// second half: alpha N v = A'u - beta N v  (lsqr.py:262-275): alpha = |v| arrives in slot D0
__device__ static inline bool gk_alpha(double *sl, double beta, double &alpha)
{
    if (beta > 0) alpha = sqrt(sl[SL_D0]);
    const bool scale_v = beta > 0 && alpha > 0;
    sl[SL_V_DIV] = scale_v ? alpha : 1.0;
    sl[SL_ALPHA] = alpha;
    return scale_v;
}

// ================================================================= LSQR (lls/lsqr.py:243-413)
__device__ static void lsqr_step(LlsDev *st, double *sl, double *hist, int phase)
{
    double *v = st->v;
    if (phase == 1) {
        st->itn++;
        gk_beta(sl, v[Q_beta]);
        if (v[Q_beta] > 0) v[Q_Anorm] = normof4(v[Q_Anorm], v[Q_alpha], v[Q_beta], st->damp);      // :260
        return;
    }
    if (phase == 2) {
        gk_alpha(sl, v[Q_beta], v[Q_alpha]);
        const double alpha = v[Q_alpha], beta = v[Q_beta], damp = st->damp;
        const double rhobar1 = normof2(v[Q_rhobar], damp);                   // :277-296
        const double cs1 = v[Q_rhobar] / rhobar1;
        const double sn1 = damp / rhobar1;
        v[Q_psi] = sn1 * v[Q_phibar];
        double phibar = cs1 * v[Q_phibar];
        const double rho = normof2(rhobar1, beta);
        const double cs = rhobar1 / rho;
        const double sn = beta / rho;
        v[Q_theta] = sn * alpha;
        v[Q_rhobar] = -cs * alpha;
        v[Q_phi] = cs * phibar;
        phibar = sn * phibar;
        v[Q_phibar] = phibar;
        v[Q_tau] = sn * v[Q_phi];
        v[Q_rho] = rho;
        sl[SL_C0] = 1.0 / rho;               // dk = (1/rho) w
        sl[SL_C1] = v[Q_phi] / rho;          // x += t1 w
        sl[SL_C2] = -v[Q_theta] / rho;       // w = t2 w + v
        return;
    }
    // phase 3: |dk|^2 in D0
    const double dk2 = sl[SL_D0];
    const double sdk = sqrt(dk2);
    v[Q_ddnorm] = v[Q_ddnorm] + sdk * sdk;                                   // :306
    const double phi = v[Q_phi], rho = v[Q_rho], theta = v[Q_theta];
    v[Q_xNrgNorm2] += phi * phi;                                             // :311-322
    const long long itn = st->itn;
    st->derr[itn % st->window] = phi;
    double direrr = nan("");
    if (itn > st->window) {
        const double trnc = window_norm(st);
        v[Q_trncDirErr] = trnc;
        const double xnrg = sqrt(v[Q_xNrgNorm2]);
        direrr = trnc / xnrg;
        if (trnc < st->etol * xnrg) st->istop = 8;
    }
    const double delta = v[Q_sn2] * rho;                                     // :324-332
    const double gambar = -v[Q_cs2] * rho;
    const double rhs_ = phi - delta * v[Q_z];
    const double zbar = rhs_ / gambar;
    v[Q_xnorm] = sqrt(v[Q_xxnorm] + zbar * zbar);
    const double gamma = normof2(gambar, theta);
    v[Q_cs2] = gambar / gamma;
    v[Q_sn2] = theta / gamma;
    v[Q_z] = rhs_ / gamma;
    v[Q_xxnorm] += v[Q_z] * v[Q_z];
    const double Anorm = v[Q_Anorm], bnorm = v[Q_bnorm];                     // :338-390
    v[Q_Acond] = Anorm * sqrt(v[Q_ddnorm]);
    const double res1 = v[Q_phibar] * v[Q_phibar];
    v[Q_res2] = v[Q_res2] + v[Q_psi] * v[Q_psi];
    const double rnorm = sqrt(res1 + v[Q_res2]);
    v[Q_rnorm] = rnorm;
    v[Q_Arnorm] = v[Q_alpha] * fabs(v[Q_tau]);
    const double r1sq = rnorm * rnorm - st->damp * st->damp * v[Q_xxnorm];
    double r1norm = sqrt(fabs(r1sq));
    if (r1sq < 0) r1norm = -r1norm;
    v[Q_r1norm] = r1norm;
    v[Q_r2norm] = rnorm;
    const double inf = 1.0 / 0.0;
    const double test1 = rnorm / bnorm;
    const double test2 = (Anorm == 0. || rnorm == 0.) ? inf : v[Q_Arnorm] / (Anorm * rnorm);
    const double test3 = (v[Q_Acond] == 0.0) ? inf : 1.0 / v[Q_Acond];
    const double t1 = test1 / (1 + Anorm * v[Q_xnorm] / bnorm);
    const double rtol = st->btol + st->atol * Anorm * v[Q_xnorm] / bnorm;
    lls_hist(st, hist, rnorm, v[Q_Arnorm], direrr, 0.0);
    if (itn >= st->itnlim) st->istop = 7;
    if (1 + test3 <= 1) st->istop = 6;
    if (1 + test2 <= 1) st->istop = 5;
    if (1 + t1 <= 1) st->istop = 4;
    if (test3 <= st->ctol) st->istop = 3;
    if (test2 <= st->atol) st->istop = 2;
    if (test1 <= rtol) st->istop = 1;
    if (st->istop > 0 || itn >= st->itnlim) st->done = 1;
}

// ================================================================= LSMR (lls/lsmr.py:303-475)
__device__ static void lsmr_step(LlsDev *st, double *sl, double *hist, int phase)
{
    double *v = st->v;
    if (phase == 1) {
        st->itn++;
        gk_beta(sl, v[M_beta]);
        return;
    }
    if (phase == 2) {
        gk_alpha(sl, v[M_beta], v[M_alpha]);
        const double alpha = v[M_alpha], beta = v[M_beta];
        double alphahat;
        sym_ortho(v[M_alphabar], st->damp, v[M_chat], v[M_shat], alphahat);          // :337-353
        v[M_rhoold] = v[M_rho];
        sym_ortho(alphahat, beta, v[M_c], v[M_s], v[M_rho]);
        v[M_thetanew] = v[M_s] * alpha;
        v[M_alphabar] = v[M_c] * alpha;
        v[M_rhobarold] = v[M_rhobar];
        v[M_zetaold] = v[M_zeta];
        v[M_thetabar] = v[M_sbar] * v[M_rho];
        v[M_rhotemp] = v[M_cbar] * v[M_rho];
        double cbar, sbar, rhobar;
        sym_ortho(v[M_cbar] * v[M_rho], v[M_thetanew], cbar, sbar, rhobar);
        v[M_cbar] = cbar; v[M_sbar] = sbar; v[M_rhobar] = rhobar;
        v[M_zeta] = cbar * v[M_zetabar];
        v[M_zetabar] = -sbar * v[M_zetabar];
        sl[SL_C0] = -(v[M_thetabar] * v[M_rho] / (v[M_rhoold] * v[M_rhobarold]));   // hbar = h + C0 hbar
        sl[SL_C1] = v[M_zeta] / (v[M_rho] * v[M_rhobar]);                            // x += C1 hbar
        sl[SL_C2] = -(v[M_thetanew] / v[M_rho]);                                     // h = v + C2 h
        return;
    }
    // phase 3: x.x in D0
    const double xx = sl[SL_D0];
    const long long itn = st->itn;
    const double zeta = v[M_zeta];
    v[M_xNrgNorm2] += zeta * zeta;                                           // :358-366
    st->derr[itn % st->window] = zeta;
    double direrr = nan("");
    if (itn > st->window) {
        const double trnc = window_norm(st);
        v[M_trncDirErr] = trnc;
        const double xnrg = sqrt(v[M_xNrgNorm2]);
        direrr = trnc / xnrg;
        if (trnc < st->etol * xnrg) st->istop = 8;
    }
    const double betaacute = v[M_chat] * v[M_betadd];                        // :368-392
    const double betacheck = -v[M_shat] * v[M_betadd];
    const double betahat = v[M_c] * betaacute;
    v[M_betadd] = -v[M_s] * betaacute;
    const double thetatildeold = v[M_thetatilde];
    double ctildeold, stildeold, rhotildeold;
    sym_ortho(v[M_rhodold], v[M_thetabar], ctildeold, stildeold, rhotildeold);
    v[M_thetatilde] = stildeold * v[M_rhobar];
    v[M_rhodold] = ctildeold * v[M_rhobar];
    v[M_betad] = -stildeold * v[M_betad] + ctildeold * betahat;
    v[M_tautildeold] = (v[M_zetaold] - thetatildeold * v[M_tautildeold]) / rhotildeold;
    const double taud = (zeta - v[M_thetatilde] * v[M_tautildeold]) / v[M_rhodold];
    v[M_d] = v[M_d] + betacheck * betacheck;
    const double dd = v[M_betad] - taud;
    const double normr = sqrt(v[M_d] + dd * dd + v[M_betadd] * v[M_betadd]);
    v[M_normr] = normr;
    v[M_normA2] = v[M_normA2] + v[M_beta] * v[M_beta];                       // :395-403
    const double normA = sqrt(v[M_normA2]);
    v[M_normA] = normA;
    v[M_normA2] = v[M_normA2] + v[M_alpha] * v[M_alpha];
    v[M_maxrbar] = fmax(v[M_maxrbar], v[M_rhobarold]);
    if (itn > 1) v[M_minrbar] = fmin(v[M_minrbar], v[M_rhobarold]);
    const double condA = fmax(v[M_maxrbar], v[M_rhotemp]) / fmin(v[M_minrbar], v[M_rhotemp]);
    v[M_condA] = condA;
    const double normar = fabs(v[M_zetabar]);
    v[M_normar] = normar;
    const double normx = sqrt(xx);
    v[M_normx] = normx;
    const double normb = v[M_normb];
    const double test1 = normr / normb;
    const double test2 = normar / (normA * normr);
    const double test3 = 1 / condA;
    const double t1 = test1 / (1 + normA * normx / normb);
    const double rtol = st->btol + st->atol * normA * normx / normb;
    lls_hist(st, hist, normr, normar, v[M_xNrgNorm2], direrr);
    if (itn >= st->itnlim) st->istop = 7;
    if (1 + test3 <= 1) st->istop = 6;
    if (1 + test2 <= 1) st->istop = 5;
    if (1 + t1 <= 1) st->istop = 4;
    if (test3 <= st->ctol) st->istop = 3;
    if (test2 <= st->atol) st->istop = 2;
    if (test1 <= rtol) st->istop = 1;
    if (st->istop > 0 || itn >= st->itnlim) st->done = 1;
}

// ================================================================= CRAIG (lls/craig.py:293-455)
__device__ static void craig_step(LlsDev *st, double *sl, double *hist, int phase)
{
    double *v = st->v;
    if (phase == 1) {
        st->itn++;
        const double alpha_old = v[C_alpha];
        gk_beta(sl, v[C_beta]);
        v[C_Arnorm] = fabs(alpha_old * v[C_beta] * v[C_s] * v[C_zeta]);      // :314
        return;
    }
    // phase 2: alpha, rotations, norms and stopping tests (no further inner product is needed)
    gk_alpha(sl, v[C_beta], v[C_alpha]);
    const double alpha = v[C_alpha], beta = v[C_beta];
    const double beta_hat = v[C_c] * beta;                                   // :336-347
    const double gamma = v[C_s] * beta;
    const double delta = normof2(gamma, 1);
    const double s2 = gamma / delta;
    const double alpha_hat = normof2(alpha, delta);
    v[C_c] = alpha / alpha_hat;
    v[C_s] = delta / alpha_hat;
    v[C_tau] = -beta_hat * v[C_tau] / alpha_hat;
    v[C_zeta] = -beta_hat * v[C_zeta] / alpha_hat;
    v[C_eta] = v[C_c] * v[C_zeta];
    v[C_xi] = v[C_s] * v[C_zeta];
    v[C_delta] = delta; v[C_beta_hat] = beta_hat; v[C_alpha_hat] = alpha_hat; v[C_s2] = s2;
    sl[SL_C0] = beta_hat;            // d = u - beta_hat d
    sl[SL_C1] = alpha_hat;           // d /= alpha_hat
    sl[SL_C2] = v[C_tau];            // r += tau d
    sl[SL_C3] = s2;                  // wbar *= s2
    sl[SL_C4] = v[C_c];              // w = c v + s wbar ; wbar = -c wbar + s v
    sl[SL_C5] = v[C_s];
    sl[SL_C6] = v[C_zeta];           // x += zeta w
    const long long itn = st->itn;
    const double tau = v[C_tau];
    v[C_rNrgNorm2] += tau * tau;                                             // :366-378
    v[C_xNrgNorm2] += v[C_zeta] * v[C_zeta];
    st->derr[itn % st->window] = tau;
    double direrr = nan("");
    if (itn > st->window) {
        const double trnc = window_norm(st);
        v[C_trncDirErr] = trnc;
        const double rnrg = sqrt(v[C_rNrgNorm2]);
        direrr = trnc / rnrg;
        if (trnc < st->etol * rnrg) st->istop = 8;
    }
    v[C_rnorm] += tau * tau;                                                 // :380-395
    v[C_xnorm] += v[C_eta] * v[C_eta];
    v[C_r1norm] += v[C_xi] * v[C_xi];
    v[C_r2norm] = v[C_rnorm];
    const double test1 = sqrt(v[C_rnorm]) / v[C_bnorm];
    const double t1 = test1;
    const double rtol = st->btol;
    lls_hist(st, hist, v[C_r2norm], v[C_Arnorm], v[C_xNrgNorm2], direrr);
    if (itn >= st->itnlim) st->istop = 7;
    if (1 + t1 <= 1) st->istop = 4;
    if (test1 <= rtol) st->istop = 1;
    // the vector updates of this trip still run (they are enqueued behind this step and must not
    // be gated away): `done` is latched by phase 3, which carries no arithmetic
    if (st->istop > 0 || itn >= st->itnlim) st->window = -st->window;        // marks "stop after the updates"
}

// ================================================================= CRAIG-MR (lls/craigmr.py:128-215)
__device__ static void craigmr_step(LlsDev *st, double *sl, double *hist, int phase)
{
    double *v = st->v;
    if (phase == 1) {
        st->itn++;
        gk_beta(sl, v[R_beta]);
        return;
    }
    gk_alpha(sl, v[R_beta], v[R_alpha]);
    const double alpha = v[R_alpha], beta = v[R_beta];
    const double beta_hat = v[R_c] * beta;                                   // :159-176
    const double gamma = v[R_s] * beta;
    const double delta = sqrt(gamma * gamma + 1);
    const double alpha_hat = sqrt(alpha * alpha + delta * delta);
    v[R_c] = alpha / alpha_hat;
    v[R_s] = delta / alpha_hat;
    const double rho = sqrt(v[R_alpha_tilde] * v[R_alpha_tilde] + beta_hat * beta_hat);
    const double c_hat = v[R_alpha_tilde] / rho;
    const double s_hat = beta_hat / rho;
    v[R_theta_old] = v[R_theta];
    v[R_theta] = s_hat * alpha_hat;
    v[R_alpha_tilde] = -c_hat * alpha_hat;
    const double zeta = c_hat * v[R_zeta_hat];
    v[R_zeta_hat] = s_hat * v[R_zeta_hat];
    v[R_zeta] = zeta; v[R_beta_hat] = beta_hat; v[R_alpha_hat] = alpha_hat; v[R_rho] = rho;
    v[R_xNrgNorm2] += zeta * zeta;
    sl[SL_C0] = v[R_theta_old];      // dbar = d - theta_old dbar
    sl[SL_C1] = rho;                 // dbar /= rho
    sl[SL_C2] = beta_hat;            // d = u - beta_hat d
    sl[SL_C3] = alpha_hat;           // d /= alpha_hat
    sl[SL_C4] = zeta;                // x += zeta dbar
    const long long itn = st->itn;
    st->derr[itn % st->window] = zeta;
    double direrr = nan("");
    if (itn > st->window) {
        const double trnc = window_norm(st);
        v[R_trncDirErr] = trnc;
        const double xnrg = sqrt(v[R_xNrgNorm2]);
        direrr = trnc / xnrg;
        if (trnc < st->etol * xnrg) st->istop = 8;
    }
    lls_hist(st, hist, v[R_xNrgNorm2], fabs(zeta), direrr, 0.0);
    if (itn >= st->itnlim) st->istop = 7;
    if (st->istop > 0 || itn >= st->itnlim) st->window = -st->window;        // stop after this trip's updates
}

// ================================================================= SYMMLQ (symmlq/symmlq.py:235-355)
__device__ static void symmlq_step(LlsDev *st, double *sl, double *hist, int phase)
{
    double *v = st->v;
    const double eps = st->eps;
    if (phase == 1) {
        // top of the loop: `while nMatvec < matvec_max` (the reference leaves with istop == 0 here)
        if (st->nmatvec >= st->itnlim) {
            st->done = 1;
            return;
        }
        st->itn++;
        v[Y_anorm] = sqrt(v[Y_tnorm]);                                       // :236-262
        v[Y_ynorm] = sqrt(v[Y_ynorm2]);
        v[Y_epsa] = v[Y_anorm] * eps;
        v[Y_epsx] = v[Y_anorm] * v[Y_ynorm] * eps;
        v[Y_epsr] = v[Y_anorm] * v[Y_ynorm] * st->rtol;
        double diag = v[Y_gbar];
        if (diag == 0) diag = v[Y_epsa];
        v[Y_diag] = diag;
        v[Y_lqnorm] = sqrt(v[Y_rhs1] * v[Y_rhs1] + v[Y_rhs2] * v[Y_rhs2]);
        v[Y_qrnorm] = v[Y_snprod] * v[Y_beta1];
        v[Y_cgnorm] = v[Y_qrnorm] * v[Y_beta] / fabs(diag);
        if (v[Y_lqnorm] < v[Y_cgnorm]) v[Y_acond] = v[Y_gmax] / v[Y_gmin];
        else v[Y_acond] = v[Y_gmax] / fmin(v[Y_gmin], fabs(diag));
        v[Y_zbar] = v[Y_rhs1] / diag;
        v[Y_z] = (v[Y_snprod] * v[Y_zbar] + v[Y_bstep]) / v[Y_beta1];
        if (st->istop == 0) {                                                // :271-276
            if (st->nmatvec >= st->itnlim) st->istop = 5;
            if (v[Y_acond] >= 0.1 / eps) st->istop = 4;
            if (v[Y_epsx] >= v[Y_beta1]) st->istop = 3;
            if (v[Y_cgnorm] <= v[Y_epsx]) st->istop = 2;
            if (v[Y_cgnorm] <= v[Y_epsr]) st->istop = 1;
        }
        lls_hist(st, hist, v[Y_cgnorm], v[Y_qrnorm], v[Y_anorm], v[Y_acond]);
        if (st->istop != 0) {
            st->done = 1;
            return;
        }
        sl[SL_C0] = 1 / v[Y_beta];                   // v = s*y, s = 1/beta             :300-301
        sl[SL_C1] = -(v[Y_beta] / v[Y_oldb]);        // y -= (beta/oldb) r1              :307
        st->nmatvec++;
        return;
    }
    if (phase == 2) {                                // alfa = v.y in D0                 :308
        v[Y_alfa] = sl[SL_D0];
        sl[SL_C2] = -(v[Y_alfa] / v[Y_beta]);        // y -= (alfa/beta) r2              :309
        return;
    }
    // phase 3: beta^2 = r2.y in D0
    v[Y_oldb] = v[Y_beta];                                                   // :316-349
    double beta = sl[SL_D0];
    if (beta < 0) {
        v[Y_beta] = beta;
        st->istop = 6;
        st->done = 1;
        return;
    }
    beta = sqrt(beta);
    v[Y_beta] = beta;
    const double alfa = v[Y_alfa], oldb = v[Y_oldb];
    v[Y_tnorm] = v[Y_tnorm] + alfa * alfa + oldb * oldb + beta * beta;
    const double gamma = sqrt(v[Y_gbar] * v[Y_gbar] + oldb * oldb);
    const double cs = v[Y_gbar] / gamma;
    const double sn = oldb / gamma;
    const double delta = cs * v[Y_dbar] + sn * alfa;
    v[Y_gbar] = sn * v[Y_dbar] - cs * alfa;
    const double epsln = sn * beta;
    v[Y_dbar] = -cs * beta;
    const double z = v[Y_rhs1] / gamma;
    sl[SL_C3] = z * cs;                              // tmp = s w + t v ; x += tmp       :332-338
    sl[SL_C4] = z * sn;
    sl[SL_C5] = sn;                                  // w = sn w - cs v
    sl[SL_C6] = cs;
    v[Y_cs] = cs; v[Y_sn] = sn;
    v[Y_bstep] = v[Y_snprod] * cs * z + v[Y_bstep];
    v[Y_snprod] = v[Y_snprod] * sn;
    v[Y_gmax] = fmax(v[Y_gmax], gamma);
    v[Y_gmin] = fmin(v[Y_gmin], gamma);
    v[Y_ynorm2] = z * z + v[Y_ynorm2];
    v[Y_rhs1] = v[Y_rhs2] - delta * z;
    v[Y_rhs2] = -epsln * z;
}

// One phase of the scalar recurrence (one thread).  Not inlined: it is also the cold tail of the
// fused vector / SpMV launches (LlsFin), whose hot loops must not pay for its registers.
#if defined(__CUDACC__)
__device__ __noinline__
#endif
static void lls_step_phase(LlsDev *st, double *slots, double *hist, int phase)
{
    if (st->done) return;
    if (phase == 9) {                // CRAIG / CRAIG-MR: latch `done` behind the trip's vector updates
        if (st->window < 0) {
            st->window = -st->window;
            st->done = 1;
        }
        return;
    }
    switch (st->method) {
        case KRY_LLS_LSQR: lsqr_step(st, slots, hist, phase); break;
        case KRY_LLS_LSMR: lsmr_step(st, slots, hist, phase); break;
        case KRY_LLS_CRAIG: craig_step(st, slots, hist, phase); break;
        case KRY_LLS_CRAIGMR: craigmr_step(st, slots, hist, phase); break;
        case KRY_LLS_SYMMLQ: symmlq_step(st, slots, hist, phase); break;
    }
}

__global__ void lls_step_kernel(LlsDev *st, double *slots, double *hist, int phase)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    lls_step_phase(st, slots, hist, phase);
}

// Finalize of a fused launch: publish the inner products into slots 0.., then run the phase of the
// recurrence that consumes them -- what a separate lls_step_kernel launch would do next.
struct LlsFin {
    LlsDev *st;
    double *slots, *hist;
    int     n, phase;
    __device__ void operator()(const double *t) const
    {
        for (int d = 0; d < n; ++d) slots[d] = t[d];
        lls_step_phase(st, slots, hist, phase);
    }
};
#endif  // __CUDACC__ || KRY_EMULATE

// ---------------------------------------------------------------- C ABI
extern "C" int kry_lls_create(kry_ctx *c, int method, kry_lls **out)
{
    KRY_REQUIRE(c && out, KRY_ERR_INVALID, "kry_lls_create: NULL argument");
    *out = nullptr;
    KRY_REQUIRE(lls_names(method), KRY_ERR_INVALID, "kry_lls_create: unknown method %d", method);
    KRY_REQUIRE(!c->closed, KRY_ERR_STATE, "kry_lls_create: the context was destroyed");
    KRY_CUDA(cudaSetDevice(c->device));
    kry_lls *L = new (std::nothrow) kry_lls();
    KRY_REQUIRE(L, KRY_ERR_NOMEM, "kry_lls_create: host allocation failed");
    L->ctx = c;
    L->method = method;
    L->dev = nullptr;
    L->hist = nullptr;
    int rc = kry_alloc((void **)&L->dev, sizeof(LlsDev));
    if (rc == KRY_OK) rc = kry_alloc((void **)&L->hist, (size_t)LLS_HIST_CAP * LLS_HIST_W * sizeof(double));
    if (rc != KRY_OK) {
        cudaFree(L->dev);
        cudaFree(L->hist);
        delete L;
        return rc;
    }
    kry_ctx_retain(c);
    *out = L;
    return KRY_OK;
}

extern "C" int kry_lls_destroy(kry_lls *L)
{
    if (!L) return KRY_OK;
    if (!L->ctx->closed) cudaStreamSynchronize(L->ctx->stream);
    if (L->ctx->gate == &L->dev->done) L->ctx->gate = nullptr;
    cudaFree(L->dev);
    cudaFree(L->hist);
    kry_ctx_release(L->ctx);
    delete L;
    return KRY_OK;
}

extern "C" const char *kry_lls_scalar_name(int method, int index)
{
    const char *const *names = lls_names(method);
    if (!names || index < 0) return nullptr;
    for (int i = 0; names[i]; ++i)
        if (i == index) return names[i];
    return nullptr;
}

extern "C" int kry_lls_setup(kry_lls *L, const kry_lls_params *p, const double *scalars, int n_scalars)
{
    KRY_REQUIRE(L && p && (scalars || n_scalars == 0), KRY_ERR_INVALID, "kry_lls_setup: NULL argument");
    KRY_CTX_LIVE(L->ctx, "kry_lls_setup");
    KRY_REQUIRE(n_scalars >= 0 && n_scalars <= LLS_NV, KRY_ERR_INVALID, "kry_lls_setup: %d scalars", n_scalars);
    KRY_REQUIRE(p->window >= 1 && p->window <= 16, KRY_ERR_INVALID, "kry_lls_setup: window %d not in [1,16]", p->window);
    kry_ctx *c = L->ctx;
    LlsDev h;
    memset(&h, 0, sizeof(h));
    h.method = L->method;
    h.window = p->window;
    h.itn = p->itn;
    h.nmatvec = p->nmatvec;
    h.itnlim = p->itnlim;
    h.istop = p->istop;
    h.damp = p->damp; h.atol = p->atol; h.btol = p->btol; h.ctol = p->ctol; h.etol = p->etol;
    h.rtol = p->rtol; h.shift = p->shift; h.eps = p->eps;
    for (int i = 0; i < n_scalars; ++i) h.v[i] = scalars[i];
    KRY_CUDA(cudaSetDevice(c->device));
    KRY_CUDA(cudaMemcpyAsync(L->dev, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
    // the Golub-Kahan methods read alpha from its slot from the first trip on
    double alpha0 = n_scalars > 0 ? scalars[0] : 0.0;
    if (L->method != KRY_LLS_SYMMLQ)
        KRY_CUDA(cudaMemcpyAsync(c->scalars + SL_ALPHA, &alpha0, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    KRY_CUDA(cudaStreamSynchronize(c->stream));
    c->gate = &L->dev->done;         // the stand-alone vector launches of this context are gated from now on
    return KRY_OK;
}

extern "C" int kry_lls_release_gate(kry_lls *L)
{
    KRY_REQUIRE(L, KRY_ERR_INVALID, "kry_lls_release_gate: NULL argument");
    if (L->ctx->gate == &L->dev->done) L->ctx->gate = nullptr;
    return KRY_OK;
}

extern "C" int kry_lls_step(kry_lls *L, int phase)
{
    KRY_REQUIRE(L, KRY_ERR_INVALID, "kry_lls_step: NULL argument");
    KRY_CTX_LIVE(L->ctx, "kry_lls_step");
    kry_ctx *c = L->ctx;
#ifdef KRY_EMULATE
    emu_launch<0>(1, 1, ReduceWs(), NoFin(), [&] { lls_step_kernel(L->dev, c->scalars, L->hist, phase); });
#else
    lls_step_kernel<<<1, 32, 0, c->stream>>>(L->dev, c->scalars, L->hist, phase);
#endif
    c->launches++;
    KRY_CUDA(cudaGetLastError());
    return KRY_OK;
}

// The two fused forms of a trip's launches: the vector pass / the SpMV with its y-side update, the
// inner products, and phase `phase` of the recurrence in the launch's finalize.
extern "C" int kry_lls_multi_axpy_dot(kry_lls *L, int phase, int n_ops, const kry_axpby *ops, int n_dots,
                                      const kry_dotspec *dots)
{
    KRY_REQUIRE(L, KRY_ERR_INVALID, "kry_lls_multi_axpy_dot: NULL argument");
    KRY_CTX_LIVE(L->ctx, "kry_lls_multi_axpy_dot");
    kry_ctx *c = L->ctx;
    int64_t n = -1;
    KRY_TRY(multi_axpy_check("kry_lls_multi_axpy_dot", c, n_ops, ops, n_dots, dots, 0, &n));
    KRY_REQUIRE(n_dots >= 1 && n > 0, KRY_ERR_INVALID,
                "kry_lls_multi_axpy_dot: the phase rides on a reduction (n_dots=%d, n=%lld)", n_dots, (long long)n);
    KRY_CUDA(cudaSetDevice(c->device));
    LlsFin fin{L->dev, c->scalars, L->hist, n_dots, phase};
    switch (n_dots) {
        case 1: return multi_axpy_run<1>(c, n, n_ops, ops, dots, fin);
        case 2: return multi_axpy_run<2>(c, n, n_ops, ops, dots, fin);
        default: return multi_axpy_run<3>(c, n, n_ops, ops, dots, fin);
    }
}

extern "C" int kry_lls_spmv_axpby_dot(kry_lls *L, int phase, kry_csr *A, int trans, const kry_vec *x,
                                      const kry_axpby *op, const kry_vec *dot_with)
{
    KRY_REQUIRE(L, KRY_ERR_INVALID, "kry_lls_spmv_axpby_dot: NULL argument");
    KRY_TRY(spmv_axpby_check("kry_lls_spmv_axpby_dot", A, trans, x, op, 1, dot_with, 0));
    KRY_REQUIRE(A->ctx == L->ctx, KRY_ERR_INVALID, "kry_lls_spmv_axpby_dot: operator and plane on different contexts");
    kry_ctx *c = L->ctx;
    KRY_CUDA(cudaSetDevice(c->device));
    KRY_TRY(kry_halo_exchange(A, x->d));
    LlsFin fin{L->dev, c->scalars, L->hist, 1, phase};
    if (((op->a_neg | op->b_neg) & 2) == 0)          // what the trips use: plain multiplies, leaner kernel
        return spmv_axpby_run<1, LLS_SPMV_MINB, false>(A, trans, x, op, dot_with, fin);
    return spmv_axpby_run<1, 6, true>(A, trans, x, op, dot_with, fin);
}

extern "C" int kry_lls_status(kry_lls *L, kry_lls_status_t *out, double *scalars, int n_scalars)
{
    KRY_REQUIRE(L && out, KRY_ERR_INVALID, "kry_lls_status: NULL argument");
    KRY_CTX_LIVE(L->ctx, "kry_lls_status");
    KRY_REQUIRE(n_scalars >= 0 && n_scalars <= LLS_NV, KRY_ERR_INVALID, "kry_lls_status: %d scalars", n_scalars);
    LlsDev h;
    KRY_CUDA(cudaMemcpyAsync(&h, L->dev, sizeof(h), cudaMemcpyDeviceToHost, L->ctx->stream));
    KRY_CUDA(cudaStreamSynchronize(L->ctx->stream));
    KRY_CUDA(cudaGetLastError());
    out->done = h.done;
    out->istop = h.istop;
    out->itn = h.itn;
    out->nmatvec = h.nmatvec;
    out->hist_count = h.hist_count;
    for (int i = 0; i < n_scalars; ++i) scalars[i] = h.v[i];
    return KRY_OK;
}

extern "C" int kry_lls_history(kry_lls *L, int64_t first, int64_t count, double *host)
{
    KRY_REQUIRE(L && (host || count == 0), KRY_ERR_INVALID, "kry_lls_history: NULL argument");
    KRY_CTX_LIVE(L->ctx, "kry_lls_history");
    KRY_REQUIRE(first >= 0 && count >= 0 && count <= LLS_HIST_CAP, KRY_ERR_INVALID, "kry_lls_history: bad range");
    cudaStream_t st = L->ctx->stream;
    int64_t done = 0;
    while (done < count) {           // the ring may wrap
        const int64_t slot = (first + done) % LLS_HIST_CAP;
        int64_t run = LLS_HIST_CAP - slot;
        if (run > count - done) run = count - done;
        KRY_CUDA(cudaMemcpyAsync(host + done * LLS_HIST_W, L->hist + slot * LLS_HIST_W,
                                 (size_t)run * LLS_HIST_W * sizeof(double), cudaMemcpyDeviceToHost, st));
        done += run;
    }
    KRY_CUDA(cudaStreamSynchronize(st));
    return KRY_OK;
}
