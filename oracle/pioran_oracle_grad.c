/*
 * pioran_oracle_grad.c — CPU restatement of the gradient of Pioran.jl's likelihood hot path.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as pioran_oracle.c: only tests/, smoke() and bench.py's CPU legs may use it).
 *
 * What the reference does: `ForwardDiff.gradient(loglike, p)` with p = [α₁, f₁, α₂, variance, ν, μ] pushed through
 * `approx` → `ScalableGP` → `logpdf` (test/test_likelihood.jl:24-43,55; the Turing scripts examples/turing_distributed/{single,double}_pl.jl
 * sample with NUTS on the same function).  ForwardDiff is forward-mode automatic differentiation with dual numbers: every
 * arithmetic operation of the code path carries the partial derivatives along.  This file restates that: the SAME
 * operation sequence as pioran_oracle.c (approx: src/psd.jl:214-289; logl: src/celerite_solver.jl:12-158,312-334)
 * evaluated on dual numbers, one tangent direction per sweep.
 *
 * Parity status: the reference pins gradients only through `all(isfinite.(grad))` (test/test_likelihood.jl:60), so the
 * VALUES are pinned indirectly: tests/test_oracle.py checks this file (i) value part == pioran_oracle.c bit for bit and
 * (ii) derivative part against central differences of the pinned oracle log-likelihood.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PSD_DBPL 1
#define ORC_BASIS_SHO 0

void orc_build_approx(int J, double f0, double fM, int basis, double *fj, double *B);

/* The arithmetic type.  The FP64 build is the oracle proper (the reference's precision); compiled a second time with
 * ORC_GRAD_LONG_DOUBLE (pioran_oracle_grad_ld.c) the same code runs in x87 80-bit arithmetic and exports
 * orc_approx_logl_grad_batch_ld — used only to tell conditioning from error when a gradient comparison at an
 * ill-conditioned parameter vector exceeds the tolerance (the role orc_celerite_logl_ld plays for the likelihood). */
#ifdef ORC_GRAD_LONG_DOUBLE
typedef long double real;
#define RM(f) f##l
#define R_PI 3.14159265358979323846264338327950288L
#define ORC_GRAD_ENTRY orc_approx_logl_grad_batch_ld
#define ORC_GRAD_LOGSHIFT_ENTRY orc_approx_logl_logshift_grad_batch_ld
#else
typedef double real;
#define RM(f) f
#define R_PI M_PI
#define ORC_GRAD_ENTRY orc_approx_logl_grad_batch
#define ORC_GRAD_LOGSHIFT_ENTRY orc_approx_logl_logshift_grad_batch
#endif


typedef struct { real v, d; } dual;

static inline dual dk(real v) { dual r = {v, 0.0}; return r; }
static inline dual dmk(real v, real d) { dual r = {v, d}; return r; }
static inline dual dadd(dual a, dual b) { return dmk(a.v + b.v, a.d + b.d); }
static inline dual dsub(dual a, dual b) { return dmk(a.v - b.v, a.d - b.d); }
static inline dual dneg(dual a) { return dmk(-a.v, -a.d); }
static inline dual dmul(dual a, dual b) { return dmk(a.v * b.v, a.d * b.v + a.v * b.d); }
static inline dual dmulc(dual a, real c) { return dmk(a.v * c, a.d * c); }
static inline dual ddivc(dual a, real c) { return dmk(a.v / c, a.d / c); }
static inline dual ddiv(dual a, dual b) { real q = a.v / b.v; return dmk(q, (a.d - q * b.d) / b.v); }
static inline dual dlog(dual a) { return dmk(RM(log)(a.v), a.d / a.v); }
static inline dual dlogabs(dual a) { return dmk(RM(log)(RM(fabs)(a.v)), a.d / a.v); }
/* x^e with both dual: d = x^e (e' ln x + e x'/x) */
static inline dual dpow(dual x, dual e)
{
    real p = RM(pow)(x.v, e.v);
    return dmk(p, p * (e.d * RM(log)(x.v) + e.v * x.d / x.v));
}

/* test/test_psd.jl:6,12 (Tonari closed forms) on duals */
static dual psd_eval_dual(int model, const dual *p, real f)
{
    dual x = ddiv(dk(f), p[1]);
    dual v = ddiv(dpow(x, dneg(p[0])), dadd(dk(1.0), dpow(x, dsub(p[2], p[0]))));
    if (model == ORC_PSD_DBPL)
        v = ddiv(v, dadd(dk(1.0), dpow(ddiv(dk(f), p[3]), dsub(p[4], p[2]))));
    return v;
}

/* LU with partial pivoting on a constant matrix, dual right-hand side (src/psd.jl:109-112) */
static int lu_solve_dual(int n, real *A, dual *x)
{
    for (int k = 0; k < n; k++) {
        int piv = k;
        real best = RM(fabs)(A[k + (size_t)k * n]);
        for (int i = k + 1; i < n; i++) {
            real v = RM(fabs)(A[i + (size_t)k * n]);
            if (v > best) { best = v; piv = i; }
        }
        if (best == 0.0) return -1;
        if (piv != k) {
            for (int j = 0; j < n; j++) {
                real tmp = A[k + (size_t)j * n];
                A[k + (size_t)j * n] = A[piv + (size_t)j * n];
                A[piv + (size_t)j * n] = tmp;
            }
            dual tmp = x[k]; x[k] = x[piv]; x[piv] = tmp;
        }
        real inv = 1.0 / A[k + (size_t)k * n];
        for (int i = k + 1; i < n; i++) A[i + (size_t)k * n] *= inv;
        for (int j = k + 1; j < n; j++) {
            real akj = A[k + (size_t)j * n];
            for (int i = k + 1; i < n; i++) A[i + (size_t)j * n] -= A[i + (size_t)k * n] * akj;
        }
    }
    for (int k = 0; k < n; k++)
        for (int i = k + 1; i < n; i++) x[i] = dsub(x[i], dmulc(x[k], A[i + (size_t)k * n]));
    for (int k = n - 1; k >= 0; k--) {
        x[k] = ddivc(x[k], A[k + (size_t)k * n]);
        for (int i = 0; i < k; i++) x[i] = dsub(x[i], dmulc(x[k], A[i + (size_t)k * n]));
    }
    return 0;
}

/* src/psd.jl:301-305, 318-324: the antiderivatives are linear in the amplitudes; the bracket is a constant */
static dual integral_basis_dual(int J, const dual *a, const real *c, real x, int basis)
{
    dual acc = dk(0.0);
    if (basis == ORC_BASIS_SHO) {
        const real s2 = RM(sqrt)(2.0);
        for (int j = 0; j < J; j++) {
            real poly = (x * x + s2 * c[j] * x + c[j] * c[j]) / (x * x - s2 * c[j] * x + c[j] * c[j]);
            real br = RM(log)(poly) + 2.0 * RM(atan2)(c[j] * s2 * x, c[j] * c[j] - x * x);
            acc = dadd(acc, dmulc(ddivc(dmulc(a[j], c[j]), 4.0 * s2), br));
        }
    } else {
        const real s3 = RM(sqrt)(3.0);
        for (int j = 0; j < J; j++) {
            real drw = RM(atan)(x / c[j]);
            real poly = (x * x + s3 * c[j] * x + c[j] * c[j]) / (x * x - s3 * c[j] * x + c[j] * c[j]);
            real cel = 0.5 * RM(atan2)(x * x - c[j] * c[j], c[j] * x) + s3 / 4.0 * RM(log)(poly);
            acc = dadd(acc, dmulc(ddivc(dmulc(a[j], c[j]), 3.0), drw + cel));
        }
    }
    return acc;
}

/* src/psd.jl:214-289 approx on duals; c, d are θ-independent.  Returns Jt. */
static int approx_dual(int model, const dual *psd_par, real f_min, real f_max, int J, dual norm, real S_low,
                       real S_high, int is_integrated_power, int basis, dual *a, dual *b, real *c, real *d)
{
    real f0 = f_min / S_low, fM = f_max * S_high;
    real *B = (real *)malloc(sizeof(real) * (size_t)J * J);
    real *fj = (real *)malloc(sizeof(real) * J);
    dual *amp = (dual *)malloc(sizeof(dual) * J);
    double *Bd = (double *)malloc(sizeof(double) * (size_t)J * J);      /* the grid and its matrix come from pioran_oracle.c */
    double *fd = (double *)malloc(sizeof(double) * J);
    if (!B || !fj || !amp || !Bd || !fd) { free(B); free(fj); free(amp); free(Bd); free(fd); return -2; }
    orc_build_approx(J, (double)f0, (double)fM, basis, fd, Bd);
    for (int j = 0; j < J; j++) fj[j] = fd[j];
    for (size_t q = 0; q < (size_t)J * J; q++) B[q] = Bd[q];
    free(Bd); free(fd);
    dual p0 = psd_eval_dual(model, psd_par, fj[0]);
    for (int j = 0; j < J; j++) amp[j] = ddiv(psd_eval_dual(model, psd_par, fj[j]), p0);
    int rc = lu_solve_dual(J, B, amp);
    if (rc) { free(B); free(fj); free(amp); return rc; }
    dual integ;
    if (is_integrated_power) {
        integ = dsub(integral_basis_dual(J, amp, fj, f_max, basis), integral_basis_dual(J, amp, fj, f_min, basis));
    } else {
        dual s = dk(0.0);
        for (int j = 0; j < J; j++) s = dadd(s, dmulc(amp[j], fj[j]));
        integ = (basis == ORC_BASIS_SHO) ? ddivc(dmulc(s, R_PI), RM(sqrt)(2.0)) : ddivc(dmulc(s, 2.0 * R_PI), 3.0);
    }
    dual scale = ddiv(norm, integ);
    for (int j = 0; j < J; j++) amp[j] = dmul(amp[j], scale);
    int Jt;
    if (basis == ORC_BASIS_SHO) {
        for (int j = 0; j < J; j++) {
            a[j] = ddivc(dmulc(dmulc(amp[j], fj[j]), R_PI), RM(sqrt)(2.0));
            b[j] = a[j];
            c[j] = RM(sqrt)(2.0) * R_PI * fj[j];
            d[j] = c[j];
        }
        Jt = J;
    } else {
        for (int j = 0; j < J; j++) {
            dual aj = ddivc(dmulc(dmulc(amp[j], fj[j]), R_PI), 3.0);
            real cj = R_PI * fj[j];
            a[j] = aj;     b[j] = dmulc(aj, RM(sqrt)(3.0)); c[j] = cj;           d[j] = RM(sqrt)(3.0) * cj;
            a[J + j] = aj; b[J + j] = dk(0.0);          c[J + j] = 2.0 * cj; d[J + j] = 0.0;
        }
        Jt = 2 * J;
    }
    free(B); free(fj); free(amp);
    return Jt;
}

/* src/celerite_solver.jl:312-334 on duals: a, b, y, σ² carry tangents; c, d, t do not. */
static dual celerite_logl_dual(int Jt, const dual *a, const dual *b, const real *c, const real *d, int64_t N,
                               const double *t, const dual *y, const dual *s2)
{
    const int R = 2 * Jt;
    dual *S = (dual *)calloc((size_t)R * R, sizeof(dual));
    real *phi = (real *)malloc(sizeof(real) * (size_t)R * (N > 1 ? N - 1 : 1));
    dual *U = (dual *)malloc(sizeof(dual) * (size_t)R * N);
    dual *V = (dual *)malloc(sizeof(dual) * (size_t)R * N);
    dual *D = (dual *)malloc(sizeof(dual) * N);
    dual *z = (dual *)malloc(sizeof(dual) * N);
    dual *f = (dual *)calloc(R, sizeof(dual));
    dual *g = (dual *)calloc(R, sizeof(dual));
    dual result = dmk(NAN, NAN);
    if (!S || !phi || !U || !V || !D || !z || !f || !g) goto done;

    dual suma = dk(0.0);
    for (int j = 0; j < Jt; j++) suma = dadd(suma, a[j]);
    D[0] = dadd(suma, s2[0]);
    {
        dual buff = ddiv(dk(1.0), D[0]);
        for (int j = 0; j < Jt; j++) {
            real co = RM(cos)(d[j] * t[0]), si = RM(sin)(d[j] * t[0]);
            V[2 * j + 1] = dmulc(buff, si);
            V[2 * j] = dmulc(buff, co);
            U[2 * j + 1] = dsub(dmulc(a[j], si), dmulc(b[j], co));
            U[2 * j] = dadd(dmulc(a[j], co), dmulc(b[j], si));
        }
    }
    for (int64_t n = 1; n < N; n++) {
        dual s = dk(0.0);
        real tn = t[n], dt = tn - (real)t[n - 1];
        dual *Un = U + (size_t)R * n, *Vn = V + (size_t)R * n, *Vp = V + (size_t)R * (n - 1);
        real *ph = phi + (size_t)R * (n - 1);
        for (int j = 0; j < Jt; j++) {
            real co = RM(cos)(d[j] * tn), si = RM(sin)(d[j] * tn), ec = RM(exp)(-c[j] * dt);
            ph[2 * j + 1] = ec; ph[2 * j] = ec;
            Un[2 * j + 1] = dsub(dmulc(a[j], si), dmulc(b[j], co));
            Un[2 * j] = dadd(dmulc(a[j], co), dmulc(b[j], si));
            Vn[2 * j + 1] = dk(si); Vn[2 * j] = dk(co);
        }
        for (int j = 0; j < R; j++) {
            dual uj = Un[j], vn = Vp[j];
            real phj = ph[j];
            dual dn = dmul(D[n - 1], vn);
            dual vnj = Vn[j];
            for (int k = 0; k < j; k++) {
                dual uk = Un[k];
                dual r = dmulc(dadd(S[j + (size_t)k * R], dmul(dn, Vp[k])), phj * ph[k]);
                S[j + (size_t)k * R] = r;
                dual v = dmul(uj, r);
                Vn[k] = dsub(Vn[k], v);
                vnj = dsub(vnj, dmul(uk, r));
                s = dadd(s, dmul(dmulc(v, 2.0), uk));
            }
            S[j + (size_t)j * R] = dmulc(dadd(S[j + (size_t)j * R], dmul(dn, vn)), phj * phj);
            dual r = dmul(S[j + (size_t)j * R], uj);
            s = dadd(s, dmul(r, uj));
            Vn[j] = dsub(vnj, r);
        }
        dual dn = dsub(dadd(suma, s2[n]), s);
        D[n] = dn;
        for (int j = 0; j < R; j++) Vn[j] = ddiv(Vn[j], dn);
    }
    {
        dual logdetD = dlog(D[0]);
        z[0] = y[0];
        for (int64_t n = 1; n < N; n++) {
            dual s = dk(0.0), zp = z[n - 1];
            const dual *Wp = V + (size_t)R * (n - 1), *Un = U + (size_t)R * n;
            const real *ph = phi + (size_t)R * (n - 1);
            for (int j = 0; j < R; j++) {
                f[j] = dmulc(dadd(f[j], dmul(Wp[j], zp)), ph[j]);
                s = dadd(s, dmul(Un[j], f[j]));
            }
            logdetD = dadd(logdetD, dlogabs(D[n]));
            z[n] = dsub(y[n], s);
        }
        z[N - 1] = ddiv(z[N - 1], D[N - 1]);
        for (int64_t n = N - 2; n >= 0; n--) {
            dual s = dk(0.0), zn = z[n + 1];
            const dual *Un1 = U + (size_t)R * (n + 1), *Wn = V + (size_t)R * n;
            const real *ph = phi + (size_t)R * n;
            for (int j = 0; j < R; j++) {
                g[j] = dmulc(dadd(g[j], dmul(Un1[j], zn)), ph[j]);
                s = dadd(s, dmul(Wn[j], g[j]));
            }
            z[n] = dsub(ddiv(z[n], D[n]), s);
        }
        dual yz = dk(0.0);
        for (int64_t n = 0; n < N; n++) yz = dadd(yz, dmul(y[n], z[n]));
        result = dsub(dsub(ddivc(dneg(logdetD), 2.0), dk((real)N * RM(log)(2 * R_PI) / 2)), ddivc(yz, 2.0));
    }
done:
    free(S); free(phi); free(U); free(V); free(D); free(z); free(f); free(g);
    return result;
}

/* Batched driver: θ row = [psd params…, norm, ν, μ] (as orc_approx_logl_batch).  logl_out [B] (may be NULL),
 * grad_out [B × (n_psd_par + 3)] = ∂logL/∂θ, one forward sweep per direction. */
void ORC_GRAD_ENTRY(int model, int n_psd_par, int B, const double *theta, double f_min, double f_max, int J,
                                double S_low, double S_high, int is_integrated_power, int basis, int64_t N,
                                const double *t, const double *y, const double *s2_base, double *logl_out,
                                double *grad_out, int nthreads)
{
    const int P = n_psd_par + 3;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) collapse(2)
#endif
    for (int i = 0; i < B; i++) {
        for (int k = 0; k < P; k++) {
            const double *th = theta + (size_t)i * P;
            dual par[8];
            for (int q = 0; q < P; q++) par[q] = dmk(th[q], q == k ? 1.0 : 0.0);
            dual norm = par[n_psd_par], nu = par[n_psd_par + 1], mu = par[n_psd_par + 2];
            int Jt_max = 2 * J;
            dual *ab = (dual *)malloc(sizeof(dual) * 2 * Jt_max);
            real *cd = (real *)malloc(sizeof(real) * 2 * Jt_max);
            dual *yy = (dual *)malloc(sizeof(dual) * N);
            dual *ss = (dual *)malloc(sizeof(dual) * N);
            for (int64_t n = 0; n < N; n++) { yy[n] = dsub(dk(y[n]), mu); ss[n] = dmulc(nu, s2_base[n]); }
            int Jt = approx_dual(model, par, f_min, f_max, J, norm, S_low, S_high, is_integrated_power, basis, ab,
                                 ab + Jt_max, cd, cd + Jt_max);
            dual r = dmk(NAN, NAN);
            if (Jt > 0) r = celerite_logl_dual(Jt, ab, ab + Jt_max, cd, cd + Jt_max, N, t, yy, ss);
            if (logl_out && k == 0) logl_out[i] = (double)r.v;
            grad_out[(size_t)i * P + k] = (double)r.d;
            free(ab); free(cd); free(yy); free(ss);
        }
    }
}


/* Log-normal model (docs/src/ultranest.md:197-217; docs/src/timeseries.md:16-21): θ row = [psd params…, norm, ν, μ, c],
 * yn = log(y − c), σ² = ν σ²/(y − c)², logpdf(ScalableGP(μ, 𝓡)(t, σ²), yn) — the same dual-number sweep with y and σ² carrying
 * the tangent of c.  grad_out [B × (n_psd_par + 4)]. */
void ORC_GRAD_LOGSHIFT_ENTRY(int model, int n_psd_par, int B, const double *theta, double f_min, double f_max, int J,
                             double S_low, double S_high, int is_integrated_power, int basis, int64_t N,
                             const double *t, const double *y, const double *s2_base, double *logl_out,
                             double *grad_out, int nthreads)
{
    const int P = n_psd_par + 4;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) collapse(2)
#endif
    for (int i = 0; i < B; i++) {
        for (int k = 0; k < P; k++) {
            const double *th = theta + (size_t)i * P;
            dual par[9];
            for (int q = 0; q < P; q++) par[q] = dmk(th[q], q == k ? 1.0 : 0.0);
            dual norm = par[n_psd_par], nu = par[n_psd_par + 1], mu = par[n_psd_par + 2], cs = par[n_psd_par + 3];
            int Jt_max = 2 * J;
            dual *ab = (dual *)malloc(sizeof(dual) * 2 * Jt_max);
            real *cd = (real *)malloc(sizeof(real) * 2 * Jt_max);
            dual *yy = (dual *)malloc(sizeof(dual) * N);
            dual *ss = (dual *)malloc(sizeof(dual) * N);
            for (int64_t n = 0; n < N; n++) {
                dual dl = dsub(dk(y[n]), cs);                         /* y − c */
                yy[n] = dsub(dlog(dl), mu);                           /* log(y − c) − μ */
                ss[n] = ddiv(dmulc(nu, s2_base[n]), dmul(dl, dl));    /* ν σ²/(y − c)² */
            }
            int Jt = approx_dual(model, par, f_min, f_max, J, norm, S_low, S_high, is_integrated_power, basis, ab,
                                 ab + Jt_max, cd, cd + Jt_max);
            dual r = dmk(NAN, NAN);
            if (Jt > 0) r = celerite_logl_dual(Jt, ab, ab + Jt_max, cd, cd + Jt_max, N, t, yy, ss);
            if (logl_out && k == 0) logl_out[i] = (double)r.v;
            grad_out[(size_t)i * P + k] = (double)r.d;
            free(ab); free(cd); free(yy); free(ss);
        }
    }
}
