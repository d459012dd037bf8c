// approx.cuh — K1: fused PSD → celerite-coefficient kernel (src/psd.jl:214-289).
//
// Per parameter vector θ = (psd parameters…, norm): evaluate the PSD on the J grid points, solve the J×J
// spectral system with the host-factored LU of the θ-independent spectral matrix (src/psd.jl:86-97 depends on
// J, f0, fM and the basis only — the reference refactors it on every call), normalise by the analytic
// integral (src/psd.jl:301-324,375-395) and emit the coefficients.  One thread per θ; J ≤ MAXJ.
#pragma once
#include "common.cuh"

namespace pioran {

constexpr int MAXJ = 64;

struct ApproxPlan {           // device-resident, built once per (J, f0, fM, basis)
    double fj[MAXJ];          // spectral points  (src/psd.jl:81-83)
    double lu[MAXJ * MAXJ];   // column-major LU factors of the spectral matrix (unit lower + upper)
    int piv[MAXJ];            // row interchanges (getrf convention: row k swapped with piv[k] at step k)
    int J, basis, model, n_psd_par, is_integrated_power;
    double f_min, f_max;
};

// Tonari.jl closed forms pinned by the reference's tests (test/test_psd.jl:6,12).
__device__ __forceinline__ double psd_eval(int model, const double* p, double f) {
    const double x = f / p[1];
    double v = pow(x, -p[0]) / (1.0 + pow(x, p[2] - p[0]));
    if (model == 1) v = v / (1.0 + pow(f / p[3], p[4] - p[2]));
    return v;
}

// Σ_j of the basis-function antiderivatives at x (src/psd.jl:301-305, 318-324)
__device__ __forceinline__ double basis_integral(int basis, int J, const double* amp, const double* fj, double x) {
    double acc = 0.0;
    if (basis == 0) {
        const double s2 = 1.4142135623730951;
        for (int j = 0; j < J; j++) {
            const double c = fj[j];
            const double nrm = c * amp[j] / (4.0 * s2);
            const double poly = (x * x + s2 * c * x + c * c) / (x * x - s2 * c * x + c * c);
            acc += nrm * (log(poly) + 2.0 * atan2(c * s2 * x, c * c - x * x));
        }
    } else {
        const double s3 = 1.7320508075688772;
        for (int j = 0; j < J; j++) {
            const double c = fj[j];
            const double nrm = amp[j] * c / 3.0;
            const double drw = atan(x / c);
            const double poly = (x * x + s3 * c * x + c * c) / (x * x - s3 * c * x + c * c);
            const double cel = 0.5 * atan2(x * x - c * c, c * x) + s3 / 4.0 * log(poly);
            acc += nrm * (drw + cel);
        }
    }
    return acc;
}

// theta: [B × tstride], first n_psd_par entries = PSD parameters, entry n_psd_par = norm.
// Outputs (any may be nullptr):
//   a,b,c,d   : [B × Jt] celerite coefficients in the reference's order (src/psd.jl:247-275)
//   amp_rows  : [B × RP] row amplitudes for the shared-table kernel, rows as built by make_rows()
//   suma      : [B] Σ_terms a  (celerite_solver.jl:21)
// PSD features (src/psd.jl:15-44, 229-243): nfeat QPO components (S₀, f₀, Q) per θ, read from theta[feat_off + 3k …]; each
// becomes one more celerite term after the Jt continuum terms (output rows then have Jt + nfeat entries).  nfeat = 0: continuum only.
constexpr int MAXFEAT = 8;

// ∫ of the celerite PSD with coefficients (a, b, c, d) (src/psd.jl:330-334)
__device__ __forceinline__ double integral_celerite(double a, double b, double c, double d, double x) {
    const double TWO_PI = 6.283185307179586;
    const double num = c * c + (d + TWO_PI * x) * (d + TWO_PI * x);
    const double den = c * c + (d - TWO_PI * x) * (d - TWO_PI * x);
    return (2.0 * a * (atan2(c, d - TWO_PI * x) - atan2(c, d + TWO_PI * x)) + b * log(num / den)) / TWO_PI;
}

__global__ void approx_kernel(const ApproxPlan* __restrict__ plan, int B, const double* __restrict__ theta, int tstride,
                              double* __restrict__ a, double* __restrict__ b, double* __restrict__ c,
                              double* __restrict__ d, double* __restrict__ amp_rows, int RP,
                              double* __restrict__ suma, int nfeat = 0, int feat_off = 0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const ApproxPlan& P = *plan;
    const int J = P.J;
    const double* th = theta + (size_t)i * tstride;
    double x[MAXJ];
    // get_normalised_psd: divide by the PSD at the first grid point (src/psd.jl:52-56)
    const double p0 = psd_eval(P.model, th, P.fj[0]);
    for (int j = 0; j < J; j++) x[j] = psd_eval(P.model, th, P.fj[j]) / p0;
    // amplitudes = B \ p  (src/psd.jl:109-112): apply the row interchanges, then L (unit) and U solves
    for (int k = 0; k < J; k++) {
        const int pk = P.piv[k];
        if (pk != k) { const double tmp = x[k]; x[k] = x[pk]; x[pk] = tmp; }
    }
    for (int k = 0; k < J; k++) {
        const double xk = x[k];
        for (int r = k + 1; r < J; r++) x[r] -= P.lu[r + k * J] * xk;
    }
    for (int k = J - 1; k >= 0; k--) {
        x[k] /= P.lu[k + k * J];
        const double xk = x[k];
        for (int r = 0; r < k; r++) x[r] -= P.lu[r + k * J] * xk;
    }
    // normalisation (src/psd.jl:236-238, 375-395)
    double integ;
    if (P.is_integrated_power) {
        integ = basis_integral(P.basis, J, x, P.fj, P.f_max) - basis_integral(P.basis, J, x, P.fj, P.f_min);
    } else {
        double s = 0.0;
        for (int j = 0; j < J; j++) s += x[j] * P.fj[j];
        integ = (P.basis == 0) ? s * 3.141592653589793 / 1.4142135623730951 : s * 2.0 * 3.141592653589793 / 3.0;
    }
    // features: convert_feature (src/psd.jl:15-28), amplitudes divided by the continuum's normalisation (:230-233), their
    // integrals added to the norm (:380-388; only when the norm is the integrated power)
    double fa[MAXFEAT], fb[MAXFEAT], fc[MAXFEAT], fd[MAXFEAT];
    for (int k = 0; k < nfeat; k++) {
        const double S0 = th[feat_off + 3 * k], f0q = th[feat_off + 3 * k + 1], Q = th[feat_off + 3 * k + 2];
        const double dl = sqrt(4.0 * Q * Q - 1.0);
        const double w0 = 2.0 * 3.141592653589793 * f0q;
        const double ak = S0 * w0 * Q / 4.0;
        fa[k] = ak / p0; fb[k] = (ak / dl) / p0;
        fc[k] = w0 / Q / 2.0; fd[k] = fc[k] * dl;
        if (P.is_integrated_power)
            integ += integral_celerite(fa[k], fb[k], fc[k], fd[k], P.f_max) - integral_celerite(fa[k], fb[k], fc[k], fd[k], P.f_min);
    }
    const double scale = th[P.n_psd_par] / integ;
    const int Jout = (P.basis == 0 ? J : 2 * J) + nfeat;     // entries per output row
    if (a)
        for (int k = 0; k < nfeat; k++) {                     // src/psd.jl:254-259, 277-282
            const size_t q = (size_t)i * Jout + (Jout - nfeat) + k;
            a[q] = 2.0 * (fa[k] * scale); b[q] = 2.0 * (fb[k] * scale); c[q] = fc[k]; d[q] = fd[k];
        }
    const double PI = 3.141592653589793, S2 = 1.4142135623730951, S3 = 1.7320508075688772;
    double sa = 0.0;
    if (P.basis == 0) {  // SHO: a = b = A f π/√2, c = d = √2 π f   (src/psd.jl:249-252)
        for (int j = 0; j < J; j++) {
            const double aj = (x[j] * scale) * P.fj[j] * PI / S2;
            sa += aj;
            if (a) {
                const size_t k = (size_t)i * Jout + j;
                a[k] = aj; b[k] = aj; c[k] = S2 * PI * P.fj[j]; d[k] = S2 * PI * P.fj[j];
            }
            if (amp_rows) { amp_rows[(size_t)i * RP + 2 * j] = aj; amp_rows[(size_t)i * RP + 2 * j + 1] = aj; }
        }
        if (amp_rows) for (int r = 2 * J; r < RP; r++) amp_rows[(size_t)i * RP + r] = 0.0;
    } else {             // DRWCelerite: (a, √3a, πf, √3πf) ++ (a, 0, 2πf, 0)   (src/psd.jl:264-275)
        for (int j = 0; j < J; j++) {
            const double aj = (x[j] * scale) * P.fj[j] * PI / 3.0;
            const double cj = PI * P.fj[j];
            if (a) {
                const size_t k = (size_t)i * Jout + j, k2 = k + J;
                a[k] = aj;  b[k] = S3 * aj; c[k] = cj;        d[k] = S3 * cj;
                a[k2] = aj; b[k2] = 0.0;    c[k2] = 2.0 * cj; d[k2] = 0.0;
            }
            if (amp_rows) {
                amp_rows[(size_t)i * RP + 2 * j] = aj; amp_rows[(size_t)i * RP + 2 * j + 1] = aj;
                amp_rows[(size_t)i * RP + 2 * J + j] = aj;
            }
        }
        // Σ over the 2J terms in the reference's order: first the celerite parts, then the DRW parts
        for (int pass = 0; pass < 2; pass++)
            for (int j = 0; j < J; j++) sa += (x[j] * scale) * P.fj[j] * PI / 3.0;
        if (amp_rows) for (int r = 3 * J; r < RP; r++) amp_rows[(size_t)i * RP + r] = 0.0;
    }
    if (suma) suma[i] = sa;
}

}  // namespace pioran
