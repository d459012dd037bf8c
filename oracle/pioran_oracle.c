/*
 * pioran_oracle.c — CPU restatement of Pioran.jl's likelihood hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pioran.jl_b200/ (the product) may include, link or call this
 * file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker or as the reported CPU baseline.
 *
 * Parity status: PINNED.  This restatement is checked (tests/test_oracle.py) against
 *   - the 20 golden SHO amplitudes of reference test/test_psd.jl:38,
 *   - the PSD closed forms of test/test_psd.jl:6,12 and the grid formula of test/test_psd.jl:28,
 *   - Σa = variance (test/test_psd.jl:114,141) and the integral normalisation (test/test_psd.jl:171,196),
 *   - celerite ≡ −dense on test/data/simu_log.txt (test/test_likelihood.jl:58-59) and on the literal N=6
 *     inputs of test/test_scalablegp.jl:110-128,
 *   - the ≈21 k (θ, logL) pairs the reference itself produced and ships under
 *     examples/ultranest/inference/{simu_single,simu_double,simu_periodic_rednoise_123_factor}/chains/weighted_post.txt.
 * The reference is Julia; no Julia toolchain exists in this image, so oracle/_ref cannot be built.
 *
 * Third-party arithmetic on the path: Tonari.jl (Project.toml:46, compat "^0.2", no Manifest → unpinned)
 * supplies SingleBendingPowerLaw / DoubleBendingPowerLaw; their formulas are pinned with `==` by the
 * reference's own tests (test/test_psd.jl:6,12) and restated in orc_psd_eval below.
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).
 * Arithmetic is IEEE double, evaluated in the reference's operation order; compile with -ffp-contract=off
 * (Julia does not contract a*b+c into fma unless asked).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PSD_SBPL 0 /* SingleBendingPowerLaw(α₁,f₁,α₂)            */
#define ORC_PSD_DBPL 1 /* DoubleBendingPowerLaw(α₁,f₁,α₂,f₂,α₃)      */
#define ORC_BASIS_SHO 0
#define ORC_BASIS_DRWCELERITE 1

/* test/test_psd.jl:6  : SBPL(f) = (f/f₁)^(−α₁) / (1 + (f/f₁)^(α₂−α₁))
 * test/test_psd.jl:12 : DBPL(f) = (f/f₁)^(−α₁) / (1 + (f/f₁)^(α₂−α₁)) / (1 + (f/f₂)^(α₃−α₂)) */
double orc_psd_eval(int model, const double *p, double f)
{
    double x = f / p[1];
    double v = pow(x, -p[0]) / (1.0 + pow(x, p[2] - p[0]));
    if (model == ORC_PSD_DBPL)
        v = v / (1.0 + pow(f / p[3], p[4] - p[2]));
    return v;
}

/* src/psd.jl:73-102 (build_approx / init_psd_decomp!): log-spaced grid and spectral matrix.
 * B is column-major J×J with B[j + k*J] = 1/(1 + (f_j/f_k)^p), p = 4 (SHO) | 6 (DRWCelerite). */
void orc_build_approx(int J, double f0, double fM, int basis, double *fj, double *B)
{
    for (int j = 0; j < J; j++)
        fj[j] = f0 * pow(fM / f0, (double)j / (double)(J - 1));
    for (int j = 0; j < J; j++)
        for (int k = 0; k < J; k++) {
            double r = fj[j] / fj[k];
            double r2 = r * r;
            double rp = (basis == ORC_BASIS_SHO) ? r2 * r2 : r2 * r2 * r2;
            B[j + (size_t)k * J] = 1.0 / (1.0 + rp);
        }
}

/* src/psd.jl:109-112 (psd_decomp: `spectral_matrix \ psd_normalised`).  Julia's `\` on a dense square,
 * non-triangular matrix is LU with partial pivoting (LAPACK getrf + getrs).  Unblocked right-looking
 * elimination; A (column-major n×n) is overwritten, x holds the right-hand side on entry, solution on exit. */
int orc_lu_solve(int n, double *A, double *x)
{
    for (int k = 0; k < n; k++) {
        int piv = k;
        double best = fabs(A[k + (size_t)k * n]);
        for (int i = k + 1; i < n; i++) {
            double v = fabs(A[i + (size_t)k * n]);
            if (v > best) { best = v; piv = i; }
        }
        if (best == 0.0) return -1;
        if (piv != k) {
            for (int j = 0; j < n; j++) {
                double tmp = A[k + (size_t)j * n];
                A[k + (size_t)j * n] = A[piv + (size_t)j * n];
                A[piv + (size_t)j * n] = tmp;
            }
            double tmp = x[k]; x[k] = x[piv]; x[piv] = tmp;
        }
        double inv = 1.0 / A[k + (size_t)k * n];
        for (int i = k + 1; i < n; i++) A[i + (size_t)k * n] *= inv;
        for (int j = k + 1; j < n; j++) {
            double akj = A[k + (size_t)j * n];
            for (int i = k + 1; i < n; i++) A[i + (size_t)j * n] -= A[i + (size_t)k * n] * akj;
        }
    }
    /* forward (unit lower) then backward (upper) substitution */
    for (int k = 0; k < n; k++)
        for (int i = k + 1; i < n; i++) x[i] -= A[i + (size_t)k * n] * x[k];
    for (int k = n - 1; k >= 0; k--) {
        x[k] /= A[k + (size_t)k * n];
        for (int i = 0; i < k; i++) x[i] -= A[i + (size_t)k * n] * x[k];
    }
    return 0;
}

/* src/psd.jl:301-305 integral_sho(a, c, x): Σ_j c_j a_j/(4√2) · [ log((x²+√2 c x+c²)/(x²−√2 c x+c²)) + 2·atan2(c√2 x, c²−x²) ]
 * (Julia reads `√2c * x` as (√2·c)·x and `4√2` as 4·√2.) */
double orc_integral_sho(int J, const double *a, const double *c, double x)
{
    const double s2 = sqrt(2.0);
    double acc = 0.0;
    for (int j = 0; j < J; j++) {
        double nrm = c[j] * a[j] / (4.0 * s2);
        double poly = (x * x + s2 * c[j] * x + c[j] * c[j]) / (x * x - s2 * c[j] * x + c[j] * c[j]);
        acc += nrm * (log(poly) + 2.0 * atan2(c[j] * s2 * x, c[j] * c[j] - x * x));
    }
    return acc;
}

/* src/psd.jl:318-324 integral_drwcelerite(a, c, x):
 * Σ_j a_j c_j/3 · [ atan(x/c) + ½·atan2(x²−c², c x) + (√3/4)·log((x²+√3 c x+c²)/(x²−√3 c x+c²)) ] */
double orc_integral_drwcelerite(int J, const double *a, const double *c, double x)
{
    const double s3 = sqrt(3.0);
    double acc = 0.0;
    for (int j = 0; j < J; j++) {
        double nrm = a[j] * c[j] / 3.0;
        double drw = atan(x / c[j]);
        double poly = (x * x + s3 * c[j] * x + c[j] * c[j]) / (x * x - s3 * c[j] * x + c[j] * c[j]);
        double cel = 0.5 * atan2(x * x - c[j] * c[j], c[j] * x) + s3 / 4.0 * log(poly);
        acc += nrm * (drw + cel);
    }
    return acc;
}

/* src/psd.jl:341-349 integrate_basis_function */
double orc_integrate_basis(int J, const double *a, const double *c, double x1, double x2, int basis)
{
    if (basis == ORC_BASIS_SHO) return orc_integral_sho(J, a, c, x2) - orc_integral_sho(J, a, c, x1);
    return orc_integral_drwcelerite(J, a, c, x2) - orc_integral_drwcelerite(J, a, c, x1);
}

/* src/psd.jl:122-128 get_approx_coefficients: raw (un-normalised) amplitudes on the grid [f0,fM]. */
int orc_get_approx_coefficients(int model, const double *psd_par, double f0, double fM, int J, int basis,
                                double *amp, double *fj_out)
{
    double *B = (double *)malloc(sizeof(double) * (size_t)J * J);
    double *fj = (double *)malloc(sizeof(double) * J);
    if (!B || !fj) { free(B); free(fj); return -2; }
    orc_build_approx(J, f0, fM, basis, fj, B);
    /* src/psd.jl:52-56 get_normalised_psd: divide by the PSD at the FIRST grid point */
    double p0 = orc_psd_eval(model, psd_par, fj[0]);
    for (int j = 0; j < J; j++) amp[j] = orc_psd_eval(model, psd_par, fj[j]) / p0;
    int rc = orc_lu_solve(J, B, amp);
    if (fj_out) memcpy(fj_out, fj, sizeof(double) * J);
    free(B); free(fj);
    return rc;
}

/* src/psd.jl:214-289 approx, continuum only (orc_approx_features below adds the QPO feature branch).
 * Writes Jt = J (SHO) or 2J (DRWCelerite) celerite terms; returns Jt or a negative error. */
int orc_approx(int model, const double *psd_par, double f_min, double f_max, int J, double norm,
               double S_low, double S_high, int is_integrated_power, int basis,
               double *a, double *b, double *c, double *d)
{
    double f0 = f_min / S_low;  /* :216 */
    double fM = f_max * S_high; /* :217 */
    double *amp = (double *)malloc(sizeof(double) * J);
    double *fj = (double *)malloc(sizeof(double) * J);
    if (!amp || !fj) { free(amp); free(fj); return -2; }
    int rc = orc_get_approx_coefficients(model, psd_par, f0, fM, J, basis, amp, fj); /* :218-226 */
    if (rc) { free(amp); free(fj); return rc; }
    /* :236 + :375-395 get_norm_psd */
    double integ;
    if (is_integrated_power) {
        integ = orc_integrate_basis(J, amp, fj, f_min, f_max, basis);
    } else {
        double s = 0.0;
        for (int j = 0; j < J; j++) s += amp[j] * fj[j];
        integ = (basis == ORC_BASIS_SHO) ? s * M_PI / sqrt(2.0) : s * 2.0 * M_PI / 3.0;
    }
    double scale = norm / integ; /* :238 amplitudes *= norm / integ */
    for (int j = 0; j < J; j++) amp[j] *= scale;
    int Jt;
    if (basis == ORC_BASIS_SHO) { /* :247-252 */
        for (int j = 0; j < J; j++) {
            a[j] = amp[j] * fj[j] * M_PI / sqrt(2.0);
            b[j] = a[j];
            c[j] = sqrt(2.0) * M_PI * fj[j];
            d[j] = c[j];
        }
        Jt = J;
    } else { /* :261-275: celerite part (a, √3a, πf, √3πf) followed by DRW part (a, 0, 2πf, 0) */
        for (int j = 0; j < J; j++) {
            double aj = amp[j] * fj[j] * M_PI / 3.0;
            double cj = M_PI * fj[j];
            a[j] = aj;         b[j] = sqrt(3.0) * aj; c[j] = cj;           d[j] = sqrt(3.0) * cj;
            a[J + j] = aj;     b[J + j] = 0.0;        c[J + j] = 2.0 * cj; d[J + j] = 0.0;
        }
        Jt = 2 * J;
    }
    free(amp); free(fj);
    return Jt;
}

/* src/psd.jl:330-334 integral_celerite: antiderivative of the celerite PSD with coefficients (a, b, c, d) at x. */
double orc_integral_celerite(double a, double b, double c, double d, double x)
{
    double num = c * c + (d + 2.0 * M_PI * x) * (d + 2.0 * M_PI * x);
    double den = c * c + (d - 2.0 * M_PI * x) * (d - 2.0 * M_PI * x);
    return (2.0 * a * (atan2(c, d - 2.0 * M_PI * x) - atan2(c, d + 2.0 * M_PI * x)) + b * log(num / den)) / (2.0 * M_PI);
}

/* src/psd.jl:214-289 approx with PSD features: continuum as in orc_approx plus nfeat QPO components feat[3k..] = (S0, f0, Q)
 * (convert_feature :15-28; amplitudes divided by the continuum's normalisation :230-233; get_norm_psd with features :380-388;
 * terms appended doubled :254-259, :277-282).  Writes Jt + nfeat terms; returns that count or a negative error. */
int orc_approx_features(int model, const double *psd_par, double f_min, double f_max, int J, double norm,
                        double S_low, double S_high, int is_integrated_power, int basis, int nfeat, const double *feat,
                        double *a, double *b, double *c, double *d)
{
    double f0 = f_min / S_low, fM = f_max * S_high;
    double *amp = (double *)malloc(sizeof(double) * J);
    double *fj = (double *)malloc(sizeof(double) * J);
    double *cf = (double *)malloc(sizeof(double) * 4 * (nfeat > 0 ? nfeat : 1));
    if (!amp || !fj || !cf) { free(amp); free(fj); free(cf); return -2; }
    int rc = orc_get_approx_coefficients(model, psd_par, f0, fM, J, basis, amp, fj);
    if (rc) { free(amp); free(fj); free(cf); return rc; }
    double psd_norm = orc_psd_eval(model, psd_par, fj[0]); /* :52-56 psd_zero */
    for (int k = 0; k < nfeat; k++) { /* :15-28 */
        double S0 = feat[3 * k], fq = feat[3 * k + 1], Q = feat[3 * k + 2];
        double Delta = sqrt(4.0 * Q * Q - 1.0);
        double w0 = 2.0 * M_PI * fq;
        double ak = S0 * w0 * Q / 4.0;
        double bk = ak / Delta;
        double ck = w0 / Q / 2.0;
        double dk = ck * Delta;
        cf[4 * k] = ak / psd_norm; cf[4 * k + 1] = bk / psd_norm; cf[4 * k + 2] = ck; cf[4 * k + 3] = dk; /* :231-232 */
    }
    double integ;
    if (is_integrated_power) { /* :376-388 */
        integ = orc_integrate_basis(J, amp, fj, f_min, f_max, basis);
        for (int k = 0; k < nfeat; k++)
            integ += orc_integral_celerite(cf[4 * k], cf[4 * k + 1], cf[4 * k + 2], cf[4 * k + 3], f_max) -
                     orc_integral_celerite(cf[4 * k], cf[4 * k + 1], cf[4 * k + 2], cf[4 * k + 3], f_min);
    } else {
        double s = 0.0;
        for (int j = 0; j < J; j++) s += amp[j] * fj[j];
        integ = (basis == ORC_BASIS_SHO) ? s * M_PI / sqrt(2.0) : s * 2.0 * M_PI / 3.0;
    }
    double scale = norm / integ;
    for (int j = 0; j < J; j++) amp[j] *= scale;
    for (int k = 0; k < nfeat; k++) { cf[4 * k] *= scale; cf[4 * k + 1] *= scale; } /* :241-242 */
    int Jt;
    if (basis == ORC_BASIS_SHO) {
        for (int j = 0; j < J; j++) {
            a[j] = amp[j] * fj[j] * M_PI / sqrt(2.0); b[j] = a[j];
            c[j] = sqrt(2.0) * M_PI * fj[j]; d[j] = c[j];
        }
        Jt = J;
    } else {
        for (int j = 0; j < J; j++) {
            double aj = amp[j] * fj[j] * M_PI / 3.0, cj = M_PI * fj[j];
            a[j] = aj;     b[j] = sqrt(3.0) * aj; c[j] = cj;           d[j] = sqrt(3.0) * cj;
            a[J + j] = aj; b[J + j] = 0.0;        c[J + j] = 2.0 * cj; d[J + j] = 0.0;
        }
        Jt = 2 * J;
    }
    for (int k = 0; k < nfeat; k++) { /* :254-259, :277-282 */
        a[Jt + k] = 2.0 * cf[4 * k]; b[Jt + k] = 2.0 * cf[4 * k + 1]; c[Jt + k] = cf[4 * k + 2]; d[Jt + k] = cf[4 * k + 3];
    }
    free(amp); free(fj); free(cf);
    return Jt + nfeat;
}

/* src/celerite_solver.jl:312-334 logl = :12-100 init_semi_separable! + :115-158 solve_prec!.
 * Same storage (U, V→W, ϕ materialised, column-major R×N; S lower triangle), same loop order, same
 * expression grouping.  Returns logL. */
double orc_celerite_logl(int Jt, const double *a, const double *b, const double *c, const double *d,
                         int64_t N, const double *t, const double *y, const double *s2)
{
    const int R = 2 * Jt;
    double *S = (double *)calloc((size_t)R * R, sizeof(double));
    double *phi = (double *)malloc(sizeof(double) * (size_t)R * (N > 1 ? N - 1 : 1));
    double *U = (double *)malloc(sizeof(double) * (size_t)R * N);
    double *V = (double *)malloc(sizeof(double) * (size_t)R * N);
    double *D = (double *)malloc(sizeof(double) * N);
    double *z = (double *)malloc(sizeof(double) * N);
    double *f = (double *)calloc(R, sizeof(double));
    double *g = (double *)calloc(R, sizeof(double));
    double result = NAN;
    if (!S || !phi || !U || !V || !D || !z || !f || !g) goto done;

    double suma = 0.0; /* :21 */
    for (int j = 0; j < Jt; j++) suma += a[j];
    D[0] = suma + s2[0]; /* :27 */
    {
        double buff = 1.0 / D[0];
        double t1 = t[0];
        for (int j = 0; j < Jt; j++) { /* :33-42 */
            double co = cos(d[j] * t1), si = sin(d[j] * t1);
            V[2 * j + 1] = si * buff;
            V[2 * j] = co * buff;
            U[2 * j + 1] = a[j] * si - b[j] * co;
            U[2 * j] = a[j] * co + b[j] * si;
        }
    }
    for (int64_t n = 1; n < N; n++) { /* :44-99 */
        double s = 0.0;
        double tn = t[n], dt = tn - t[n - 1];
        double *Un = U + (size_t)R * n, *Vn = V + (size_t)R * n, *Vp = V + (size_t)R * (n - 1);
        double *ph = phi + (size_t)R * (n - 1);
        for (int j = 0; j < Jt; j++) { /* :51-64 */
            double co = cos(d[j] * tn), si = sin(d[j] * tn), ec = exp(-c[j] * dt);
            ph[2 * j + 1] = ec; ph[2 * j] = ec;
            Un[2 * j + 1] = a[j] * si - b[j] * co;
            Un[2 * j] = a[j] * co + b[j] * si;
            Vn[2 * j + 1] = si; Vn[2 * j] = co;
        }
        for (int j = 0; j < R; j++) { /* :69-90 */
            double uj = Un[j], phj = ph[j], vn = Vp[j];
            double dn = D[n - 1] * vn;
            double vnj = Vn[j];
            for (int k = 0; k < j; k++) {
                double uk = Un[k];
                double r = phj * ph[k] * (S[j + (size_t)k * R] + dn * Vp[k]);
                S[j + (size_t)k * R] = r;
                double v = uj * r;
                Vn[k] -= v;
                vnj -= uk * r;
                s += 2 * v * uk;
            }
            S[j + (size_t)j * R] = phj * phj * (S[j + (size_t)j * R] + dn * vn);
            double r = S[j + (size_t)j * R] * uj;
            s += r * uj;
            Vn[j] = vnj - r;
        }
        double dn = suma + s2[n] - s; /* :92 */
        D[n] = dn;
        for (int j = 0; j < R; j++) Vn[j] /= dn; /* :95-98 */
    }
    /* solve_prec! :115-158 */
    {
        double logdetD = log(D[0]); /* :126 (no abs on the first pivot) */
        z[0] = y[0];
        for (int64_t n = 1; n < N; n++) { /* :132-142 */
            double s = 0.0, zp = z[n - 1];
            const double *Wp = V + (size_t)R * (n - 1), *ph = phi + (size_t)R * (n - 1), *Un = U + (size_t)R * n;
            for (int j = 0; j < R; j++) {
                f[j] = (f[j] + Wp[j] * zp) * ph[j];
                s += Un[j] * f[j];
            }
            logdetD += log(fabs(D[n]));
            z[n] = y[n] - s;
        }
        z[N - 1] /= D[N - 1]; /* :145 */
        for (int64_t n = N - 2; n >= 0; n--) { /* :146-155 */
            double s = 0.0, zn = z[n + 1];
            const double *Un1 = U + (size_t)R * (n + 1), *ph = phi + (size_t)R * n, *Wn = V + (size_t)R * n;
            for (int j = 0; j < R; j++) {
                g[j] = (g[j] + Un1[j] * zn) * ph[j];
                s += Wn[j] * g[j];
            }
            z[n] = z[n] / D[n] - s;
        }
        double yz = 0.0;
        for (int64_t n = 0; n < N; n++) yz += y[n] * z[n];
        result = -logdetD / 2 - (double)N * log(2 * M_PI) / 2 - yz / 2; /* :333 */
    }
done:
    free(S); free(phi); free(U); free(V); free(D); free(z); free(f); free(g);
    return result;
}

/* Extended-precision (x87 80-bit long double) evaluation of the same likelihood, forward-only form
 * (y'K⁻¹y = Σ z_n²/D_n).  Used only to triage ill-conditioned parameter vectors in the
 * parity reports: it tells which of two FP64 answers is nearer the exact value. */
long double orc_celerite_logl_ld(int Jt, const double *a, const double *b, const double *c, const double *d,
                                 int64_t N, const double *t, const double *y, const double *s2)
{
    const int R = 2 * Jt;
    long double *S = (long double *)calloc((size_t)R * R, sizeof(long double));
    long double *u = (long double *)malloc(sizeof(long double) * R);
    long double *v = (long double *)malloc(sizeof(long double) * R);
    long double *w = (long double *)calloc(R, sizeof(long double));
    long double *ph = (long double *)malloc(sizeof(long double) * R);
    long double *f = (long double *)calloc(R, sizeof(long double));
    long double *p = (long double *)malloc(sizeof(long double) * R);
    long double suma = 0, logdet = 0, chi2 = 0, Dprev = 0, zprev = 0;
    for (int j = 0; j < Jt; j++) suma += a[j];
    for (int64_t n = 0; n < N; n++) {
        long double tn = t[n];
        for (int j = 0; j < Jt; j++) {
            long double co = cosl((long double)d[j] * tn), si = sinl((long double)d[j] * tn);
            u[2 * j] = a[j] * co + b[j] * si; u[2 * j + 1] = a[j] * si - b[j] * co;
            v[2 * j] = co; v[2 * j + 1] = si;
            if (n > 0) ph[2 * j] = ph[2 * j + 1] = expl(-(long double)c[j] * (tn - (long double)t[n - 1]));
        }
        long double Dn, zn;
        if (n == 0) {
            Dn = suma + s2[0];
            for (int j = 0; j < R; j++) w[j] = v[j] / Dn;
            zn = y[0];
        } else {
            for (int j = 0; j < R; j++)
                for (int k = 0; k <= j; k++) {
                    long double r = ph[j] * ph[k] * (S[j + (size_t)k * R] + Dprev * w[j] * w[k]);
                    S[j + (size_t)k * R] = r; S[k + (size_t)j * R] = r;
                }
            long double s = 0, uf = 0;
            for (int j = 0; j < R; j++) {
                long double acc = 0;
                for (int k = 0; k < R; k++) acc += S[j + (size_t)k * R] * u[k];
                p[j] = acc; s += u[j] * acc;
                f[j] = ph[j] * (f[j] + w[j] * zprev);
                uf += u[j] * f[j];
            }
            Dn = suma + s2[n] - s;
            for (int j = 0; j < R; j++) w[j] = (v[j] - p[j]) / Dn;
            zn = y[n] - uf;
        }
        logdet += logl(fabsl(Dn));
        chi2 += zn * zn / Dn;
        Dprev = Dn; zprev = zn;
    }
    free(S); free(u); free(v); free(w); free(ph); free(f); free(p);
    return -logdet / 2 - (long double)N * logl(2 * 3.141592653589793238462643383279502884L) / 2 - chi2 / 2;
}

/* src/direct_solver.jl:6-21 log_likelihood_direct with the kernel of src/Celerite.jl:42-44 summed over terms
 * (src/acvf.jl:138-140), Euclidean metric τ = |t_i − t_j|.  Returns +NLL like the reference; *info = 1 and
 * NaN when the matrix is not positive definite (the reference throws PosDefException). */
double orc_direct_nll(int Jt, const double *a, const double *b, const double *c, const double *d,
                      int64_t N, const double *t, const double *y, const double *s2, int *info)
{
    double *K = (double *)malloc(sizeof(double) * (size_t)N * N);
    double *z = (double *)malloc(sizeof(double) * N);
    double res = NAN;
    if (info) *info = 0;
    if (!K || !z) goto done;
    for (int64_t i = 0; i < N; i++)
        for (int64_t j = 0; j <= i; j++) {
            double tau = fabs(t[i] - t[j]);
            double k = 0.0;
            for (int m = 0; m < Jt; m++)
                k += exp(-c[m] * tau) * (a[m] * cos(d[m] * tau) + b[m] * sin(d[m] * tau));
            if (i == j) k += s2[i];
            K[i + (size_t)j * N] = k;
        }
    /* lower Cholesky, column by column */
    for (int64_t j = 0; j < N; j++) {
        double djj = K[j + (size_t)j * N];
        for (int64_t k = 0; k < j; k++) djj -= K[j + (size_t)k * N] * K[j + (size_t)k * N];
        if (!(djj > 0.0)) { if (info) *info = 1; goto done; }
        djj = sqrt(djj);
        K[j + (size_t)j * N] = djj;
        for (int64_t i = j + 1; i < N; i++) {
            double v = K[i + (size_t)j * N];
            for (int64_t k = 0; k < j; k++) v -= K[i + (size_t)k * N] * K[j + (size_t)k * N];
            K[i + (size_t)j * N] = v / djj;
        }
    }
    {
        double logdet = 0.0, zz = 0.0;
        for (int64_t i = 0; i < N; i++) {
            double v = y[i];
            for (int64_t k = 0; k < i; k++) v -= K[i + (size_t)k * N] * z[k];
            z[i] = v / K[i + (size_t)i * N];
            logdet += log(K[i + (size_t)i * N]);
            zz += z[i] * z[i];
        }
        res = logdet + 0.5 * zz + 0.5 * (double)N * log(2 * M_PI);
    }
done:
    free(K); free(z);
    return res;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Widening rows: posterior mean `pred` and GP draw `sim`.
 * Both start from the factorisation of init_semi_separable! (:12-100); ss_factor below repeats that part of
 * orc_celerite_logl verbatim (U, W = V/D, ϕ, D materialised) and ss_solve repeats solve_prec! (:115-158). */
static void ss_factor(int Jt, const double *a, const double *b, const double *c, const double *d, int64_t N,
                      const double *t, const double *s2, double *U, double *V, double *phi, double *D, double *S)
{
    const int R = 2 * Jt;
    double suma = 0.0;
    for (int j = 0; j < Jt; j++) suma += a[j];
    memset(S, 0, sizeof(double) * (size_t)R * R);
    D[0] = suma + s2[0];
    {
        double buff = 1.0 / D[0], t1 = t[0];
        for (int j = 0; j < Jt; j++) {
            double co = cos(d[j] * t1), si = sin(d[j] * t1);
            V[2 * j + 1] = si * buff; V[2 * j] = co * buff;
            U[2 * j + 1] = a[j] * si - b[j] * co; U[2 * j] = a[j] * co + b[j] * si;
        }
    }
    for (int64_t n = 1; n < N; n++) {
        double s = 0.0, tn = t[n], dt = tn - t[n - 1];
        double *Un = U + (size_t)R * n, *Vn = V + (size_t)R * n, *Vp = V + (size_t)R * (n - 1);
        double *ph = phi + (size_t)R * (n - 1);
        for (int j = 0; j < Jt; j++) {
            double co = cos(d[j] * tn), si = sin(d[j] * tn), ec = exp(-c[j] * dt);
            ph[2 * j + 1] = ec; ph[2 * j] = ec;
            Un[2 * j + 1] = a[j] * si - b[j] * co; Un[2 * j] = a[j] * co + b[j] * si;
            Vn[2 * j + 1] = si; Vn[2 * j] = co;
        }
        for (int j = 0; j < R; j++) {
            double uj = Un[j], phj = ph[j], vn = Vp[j], dn = D[n - 1] * vn, vnj = Vn[j];
            for (int k = 0; k < j; k++) {
                double uk = Un[k];
                double r = phj * ph[k] * (S[j + (size_t)k * R] + dn * Vp[k]);
                S[j + (size_t)k * R] = r;
                double v = uj * r;
                Vn[k] -= v; vnj -= uk * r; s += 2 * v * uk;
            }
            S[j + (size_t)j * R] = phj * phj * (S[j + (size_t)j * R] + dn * vn);
            double r = S[j + (size_t)j * R] * uj;
            s += r * uj;
            Vn[j] = vnj - r;
        }
        double dn = suma + s2[n] - s;
        D[n] = dn;
        for (int j = 0; j < R; j++) Vn[j] /= dn;
    }
}
/* solve_prec! (:115-158): z ← K⁻¹ y */
static void ss_solve(int R, int64_t N, const double *y, const double *U, const double *W, const double *D,
                     const double *phi, double *z, double *f, double *g)
{
    memset(f, 0, sizeof(double) * R); memset(g, 0, sizeof(double) * R);
    z[0] = y[0];
    for (int64_t n = 1; n < N; n++) {
        double s = 0.0, zp = z[n - 1];
        const double *Wp = W + (size_t)R * (n - 1), *ph = phi + (size_t)R * (n - 1), *Un = U + (size_t)R * n;
        for (int j = 0; j < R; j++) { f[j] = (f[j] + Wp[j] * zp) * ph[j]; s += Un[j] * f[j]; }
        z[n] = y[n] - s;
    }
    z[N - 1] /= D[N - 1];
    for (int64_t n = N - 2; n >= 0; n--) {
        double s = 0.0, zn = z[n + 1];
        const double *Un1 = U + (size_t)R * (n + 1), *ph = phi + (size_t)R * n, *Wn = W + (size_t)R * n;
        for (int j = 0; j < R; j++) { g[j] = (g[j] + Un1[j] * zn) * ph[j]; s += Wn[j] * g[j]; }
        z[n] = z[n] / D[n] - s;
    }
}

/* src/celerite_solver.jl:376-483 `pred`: posterior mean at the (ascending) points tau given (t, y, σ²), with the
 * reference's own index bookkeeping (start / stop / n₀ = searchsortedfirst(t, τ) − 1).  Indices below are 1-based
 * like the Julia source; arrays are read with [n − 1].  Returns 0, or −1 on allocation failure. */
int orc_celerite_predict(int Jt, const double *a, const double *b, const double *c, const double *d, int64_t M,
                         const double *tau, int64_t N, const double *t, const double *y, const double *s2, double *mu_out)
{
    const int R = 2 * Jt;
    double *S = (double *)malloc(sizeof(double) * (size_t)R * R);
    double *phi = (double *)malloc(sizeof(double) * (size_t)R * (N > 1 ? N - 1 : 1));
    double *U = (double *)malloc(sizeof(double) * (size_t)R * N);
    double *V = (double *)malloc(sizeof(double) * (size_t)R * N);
    double *D = (double *)malloc(sizeof(double) * N);
    double *z = (double *)malloc(sizeof(double) * N);
    double *f = (double *)malloc(sizeof(double) * R), *g = (double *)malloc(sizeof(double) * R);
    double *Q = (double *)calloc(R, sizeof(double)), *Sv = (double *)calloc(R, sizeof(double));
    int64_t *n0L = (int64_t *)malloc(sizeof(int64_t) * (M > 0 ? M : 1));
    int rc = -1;
    if (!S || !phi || !U || !V || !D || !z || !f || !g || !Q || !Sv || !n0L) goto done;
    ss_factor(Jt, a, b, c, d, N, t, s2, U, V, phi, D, S);
    ss_solve(R, N, y, U, V, D, phi, z, f, g);
    for (int64_t m = 0; m < M; m++) { /* :395 searchsortedfirst(t, τ) − 1 = number of t_n < τ */
        int64_t lo = 0, hi = N;
        while (lo < hi) { int64_t mid = (lo + hi) / 2; if (t[mid] < tau[m]) lo = mid + 1; else hi = mid; }
        n0L[m] = lo;
        mu_out[m] = 0.0;
    }
    /* forward pass :400-435 */
    {
        int64_t start = 1;
        for (int64_t m = 0; m < M; m++) {
            const int64_t n0 = n0L[m];
            const double tm = tau[m];
            for (int64_t n = start; n <= n0 - 1; n++) {
                start += 1;
                double tn = t[n - 1], tn1 = t[n], zn = z[n - 1];
                for (int j = 0; j < Jt; j++) {
                    double e = exp(-c[j] * (tn1 - tn));
                    Q[2 * j] = (Q[2 * j] + zn * cos(d[j] * tn)) * e;
                    Q[2 * j + 1] = (Q[2 * j + 1] + zn * sin(d[j] * tn)) * e;
                }
            }
            if (start >= n0 && n0 != 0) {
                double tn = t[n0 - 1], zn = z[n0 - 1];
                double tn1 = (n0 == N) ? t[n0 - 1] : t[n0];
                for (int j = 0; j < Jt; j++) {
                    double e = exp(-c[j] * (tm - tn));
                    Sv[2 * j + 1] = (Q[2 * j + 1] + zn * sin(d[j] * tn)) * e * (a[j] * sin(d[j] * tm) - b[j] * cos(d[j] * tm));
                    Sv[2 * j] = (Q[2 * j] + zn * cos(d[j] * tn)) * e * (a[j] * cos(d[j] * tm) + b[j] * sin(d[j] * tm));
                }
                if (start + 1 == n0) {
                    start += 1;
                    for (int j = 0; j < Jt; j++) {
                        double e = exp(-c[j] * (tn1 - tn));
                        Q[2 * j] = (Q[2 * j] + zn * cos(d[j] * tn)) * e;
                        Q[2 * j + 1] = (Q[2 * j + 1] + zn * sin(d[j] * tn)) * e;
                    }
                }
            }
            double sum = 0.0;
            for (int j = 0; j < R; j++) sum += Sv[j];
            mu_out[m] = sum;
        }
    }
    memset(Q, 0, sizeof(double) * R);
    /* backward pass :439-480 */
    {
        int64_t stop = N;
        for (int64_t m = M - 1; m >= 0; m--) {
            const int64_t n0 = n0L[m];
            if (n0 == N) continue;
            const double tm = tau[m];
            const int64_t stop_cur = stop;
            for (int64_t n = stop_cur; n >= n0 + 2; n--) {
                stop -= 1;
                double tn = t[n - 1], tnm = t[n - 2], zn = z[n - 1];
                for (int j = 0; j < Jt; j++) {
                    double e = exp(-c[j] * (tn - tnm));
                    Q[2 * j] = (Q[2 * j] + zn * (a[j] * cos(d[j] * tn) + b[j] * sin(d[j] * tn))) * e;
                    Q[2 * j + 1] = (Q[2 * j + 1] + zn * (a[j] * sin(d[j] * tn) - b[j] * cos(d[j] * tn))) * e;
                }
            }
            const int64_t n = n0 + 1;
            double zn = z[n - 1], tn = t[n - 1];
            double tnm = (n == 1) ? t[0] : t[n - 2];
            for (int j = 0; j < Jt; j++) {
                double e = exp(-c[j] * (tn - tm));
                Sv[2 * j] = (Q[2 * j] + zn * (a[j] * cos(d[j] * tn) + b[j] * sin(d[j] * tn))) * e * cos(d[j] * tm);
                Sv[2 * j + 1] = (Q[2 * j + 1] + zn * (a[j] * sin(d[j] * tn) - b[j] * cos(d[j] * tn))) * e * sin(d[j] * tm);
            }
            const int64_t k = (m != 0) ? m : 1;   /* 0-based: the Julia k = m (m ≠ 1) else 2 */
            if (stop_cur == n0 + 1 && M > 1 && n0L[k] != n0L[k - 1]) {
                stop -= 1;
                for (int j = 0; j < Jt; j++) {
                    double e = exp(-c[j] * (tn - tnm));
                    Q[2 * j] = (Q[2 * j] + zn * (a[j] * cos(d[j] * tn) + b[j] * sin(d[j] * tn))) * e;
                    Q[2 * j + 1] = (Q[2 * j + 1] + zn * (a[j] * sin(d[j] * tn) - b[j] * cos(d[j] * tn))) * e;
                }
            }
            double sum = 0.0;
            for (int j = 0; j < R; j++) sum += Sv[j];
            mu_out[m] += sum;
        }
    }
    rc = 0;
done:
    free(S); free(phi); free(U); free(V); free(D); free(z); free(f); free(g); free(Q); free(Sv); free(n0L);
    return rc;
}

/* src/celerite_solver.jl:515-549 `sim`: y = L q for standard-normal q (the caller supplies q; Julia's randn stream
 * cannot be reproduced).  NaN where a pivot is negative (the reference raises DomainError in sqrt). */
int orc_celerite_simulate(int Jt, const double *a, const double *b, const double *c, const double *d, int64_t N,
                          const double *t, const double *s2, const double *q, double *y_sim)
{
    const int R = 2 * Jt;
    double *S = (double *)malloc(sizeof(double) * (size_t)R * R);
    double *phi = (double *)malloc(sizeof(double) * (size_t)R * (N > 1 ? N - 1 : 1));
    double *U = (double *)malloc(sizeof(double) * (size_t)R * N);
    double *V = (double *)malloc(sizeof(double) * (size_t)R * N);
    double *D = (double *)malloc(sizeof(double) * N);
    double *f = (double *)calloc(R, sizeof(double)), *g = (double *)calloc(R, sizeof(double));
    int rc = -1;
    if (!S || !phi || !U || !V || !D || !f || !g) goto done;
    ss_factor(Jt, a, b, c, d, N, t, s2, U, V, phi, D, S);
    for (int64_t n = 0; n < N; n++) y_sim[n] = 0.0;
    y_sim[0] = sqrt(D[0]) * q[0];
    for (int64_t n = 1; n < N; n++) { /* :538-546 */
        const double *ph = phi + (size_t)R * (n - 1), *Wp = V + (size_t)R * (n - 1), *Un = U + (size_t)R * n;
        for (int j = 0; j < R; j++) {
            f[j] = ph[j] * (g[j] + Wp[j] * sqrt(D[n - 1]) * q[n - 1]);
            y_sim[n] += Un[j] * f[j];
        }
        for (int j = 0; j < R; j++) g[j] = f[j];
        y_sim[n] += sqrt(D[n]) * q[n];
    }
    rc = 0;
done:
    free(S); free(phi); free(U); free(V); free(D); free(f); free(g);
    return rc;
}

/* src/direct_solver.jl:74-119 predict_direct (mean only): K_τ0 (K0 + diag σ²)⁻¹ y by Cholesky.  The reference's own
 * tests pin pred against it (test/test_prediction.jl:49-58, test/test_predict_celerite.jl). */
int orc_direct_predict(int Jt, const double *a, const double *b, const double *c, const double *d, int64_t M,
                       const double *tau, int64_t N, const double *t, const double *y, const double *s2, double *mu_out)
{
    double *K = (double *)malloc(sizeof(double) * (size_t)N * N);
    double *z = (double *)malloc(sizeof(double) * N);
    int rc = -1;
    if (!K || !z) goto done;
    for (int64_t i = 0; i < N; i++)
        for (int64_t j = 0; j <= i; j++) {
            double dt = fabs(t[i] - t[j]), k = 0.0;
            for (int m = 0; m < Jt; m++) k += exp(-c[m] * dt) * (a[m] * cos(d[m] * dt) + b[m] * sin(d[m] * dt));
            if (i == j) k += s2[i];
            K[i + (size_t)j * N] = k;
        }
    for (int64_t j = 0; j < N; j++) {
        double djj = K[j + (size_t)j * N];
        for (int64_t k = 0; k < j; k++) djj -= K[j + (size_t)k * N] * K[j + (size_t)k * N];
        if (!(djj > 0.0)) { rc = 1; goto done; }
        djj = sqrt(djj);
        K[j + (size_t)j * N] = djj;
        for (int64_t i = j + 1; i < N; i++) {
            double v = K[i + (size_t)j * N];
            for (int64_t k = 0; k < j; k++) v -= K[i + (size_t)k * N] * K[j + (size_t)k * N];
            K[i + (size_t)j * N] = v / djj;
        }
    }
    for (int64_t i = 0; i < N; i++) {             /* L z' = y */
        double v = y[i];
        for (int64_t k = 0; k < i; k++) v -= K[i + (size_t)k * N] * z[k];
        z[i] = v / K[i + (size_t)i * N];
    }
    for (int64_t i = N - 1; i >= 0; i--) {        /* Lᵀ z = z' */
        double v = z[i];
        for (int64_t k = i + 1; k < N; k++) v -= K[k + (size_t)i * N] * z[k];
        z[i] = v / K[i + (size_t)i * N];
    }
    for (int64_t m = 0; m < M; m++) {
        double acc = 0.0;
        for (int64_t n = 0; n < N; n++) {
            double dt = fabs(tau[m] - t[n]), k = 0.0;
            for (int jj = 0; jj < Jt; jj++) k += exp(-c[jj] * dt) * (a[jj] * cos(d[jj] * dt) + b[jj] * sin(d[jj] * dt));
            acc += k * z[n];
        }
        mu_out[m] = acc;
    }
    rc = 0;
done:
    free(K); free(z);
    return rc;
}

/* Batched driver = what a sampler does per parameter vector (examples/ultranest/single_pl.jl:65-93):
 * θ row = [psd params…, norm, ν, μ]; σ² = ν·s2_base; y' = y − μ (src/scalable_GP.jl:162-166); approx; logl.
 * OpenMP over θ when nthreads > 1 (the reference itself is one Julia thread per process; its users run one
 * process per core, examples/ultranest/single_pl.jl:19-21). */
void orc_approx_logl_batch(int model, int n_psd_par, int B, const double *theta, double f_min, double f_max,
                           int J, double S_low, double S_high, int is_integrated_power, int basis,
                           int64_t N, const double *t, const double *y, const double *s2_base,
                           double *logl_out, int nthreads)
{
    const int stride = n_psd_par + 3;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (int i = 0; i < B; i++) {
        const double *th = theta + (size_t)i * stride;
        double norm = th[n_psd_par], nu = th[n_psd_par + 1], mu = th[n_psd_par + 2];
        int Jt_max = 2 * J;
        double *co = (double *)malloc(sizeof(double) * 4 * Jt_max);
        double *yy = (double *)malloc(sizeof(double) * N);
        double *ss = (double *)malloc(sizeof(double) * N);
        for (int64_t n = 0; n < N; n++) { yy[n] = y[n] - mu; ss[n] = nu * s2_base[n]; }
        int Jt = orc_approx(model, th, f_min, f_max, J, norm, S_low, S_high, is_integrated_power, basis,
                            co, co + Jt_max, co + 2 * Jt_max, co + 3 * Jt_max);
        logl_out[i] = (Jt > 0) ? orc_celerite_logl(Jt, co, co + Jt_max, co + 2 * Jt_max, co + 3 * Jt_max, N, t, yy, ss) : NAN;
        free(co); free(yy); free(ss);
    }
}

/* Batched generic-coefficient driver: B coefficient sets (row-major [B×Jt]) on one series. */
void orc_celerite_logl_batch(int B, int Jt, const double *a, const double *b, const double *c, const double *d,
                             const double *mu, const double *nu,
                             int64_t N, const double *t, const double *y, const double *s2_base,
                             double *logl_out, int nthreads)
{
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (int i = 0; i < B; i++) {
        double *yy = (double *)malloc(sizeof(double) * N);
        double *ss = (double *)malloc(sizeof(double) * N);
        double m = mu ? mu[i] : 0.0, v = nu ? nu[i] : 1.0;
        for (int64_t n = 0; n < N; n++) { yy[n] = y[n] - m; ss[n] = v * s2_base[n]; }
        logl_out[i] = orc_celerite_logl(Jt, a + (size_t)i * Jt, b + (size_t)i * Jt, c + (size_t)i * Jt, d + (size_t)i * Jt, N, t, yy, ss);
        free(yy); free(ss);
    }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
