// prior.cuh — device-side prior transform of unit-cube points (missing #5).
//
// The reference's nested-sampling scripts map a unit-cube point to a parameter vector with Distributions.jl quantiles, one
// column at a time (examples/ultranest/single_pl.jl:96-104: α₁ ~ Uniform(0, 1.5), f₁ ~ LogUniform(f0·4, fM/4), α₂ ~ Uniform(α₁, 4),
// variance ~ LogNormal(μₙ, σₙ), ν ~ Gamma(2, 0.5), μ ~ Normal(x̄, 5√va); docs/src/ultranest.md:165-190 adds c ~ LogUniform).
// One thread per point; columns left to right, so a prior may take its lower edge from an earlier column of the same point.
#pragma once
#include "common.cuh"

namespace pioran {

// kinds of pioran_prior_spec (include/pioran_b200.h)
enum PriorKind : int { PRIOR_UNIFORM = 0, PRIOR_UNIFORM_FROM = 1, PRIOR_LOGUNIFORM = 2, PRIOR_NORMAL = 3, PRIOR_LOGNORMAL = 4, PRIOR_GAMMA = 5 };
struct PriorSpec { int kind, ref_col; double p0, p1; };
constexpr int PRIOR_MAX_COLS = 16;

// quantile of Gamma(shape k (integer), scale 1): P(k, x) = 1 − e^{−x} Σ_{i<k} x^i/i! = u, by safeguarded Newton steps on x
// (the CDF is increasing and concave beyond its mode; bisection bounds keep every step inside [lo, hi])
__device__ inline double gamma_int_quantile(const int k, const double u) {
    if (!(u > 0.0)) return 0.0;
    if (!(u < 1.0)) return INFINITY;
    const double uc = 1.0 - u;                         // exact for u ≥ 1/2 (Sterbenz)
    // residual P(k, x) − u without cancellation: the lower series e^{−x} Σ_{i≥k} x^i/i! below the mode region, (1 − u) − Q above it
    auto resid = [&](double x, double& pdf) {
        const double ex = exp(-x);
        double term = 1.0, sum = 1.0;                  // Σ_{i<k} x^i / i!, term = x^{k−1}/(k−1)! at the end
        for (int i = 1; i < k; i++) { term *= x / i; sum += term; }
        pdf = term * ex;                               // x^{k−1} e^{−x} / (k−1)!
        if (x <= (double)k + 1.0) {
            double tl = term * x / k, sl = tl;         // x^k/k!, then x^{k+j}/(k+j)!
            for (int j = 1; j < 200; j++) { tl *= x / (k + j); sl += tl; if (tl <= 1e-17 * sl) break; }
            return ex * sl - u;
        }
        return uc - ex * sum;
    };
    double lo = 0.0, hi = (double)k + 10.0, pdf;
    while (resid(hi, pdf) < 0.0) { lo = hi; hi *= 2.0; if (hi > 1e6) break; }
    double fac = 1.0;
    for (int i = 2; i <= k; i++) fac *= i;
    double x = fmin(fmax(pow(u * fac, 1.0 / k), 1e-300), hi);        // small-u behaviour P ≈ x^k / k!
    for (int it = 0; it < 200; it++) {
        const double f = resid(x, pdf);
        if (f > 0.0) hi = x; else lo = x;
        double xn = (pdf > 0.0) ? x - f / pdf : 0.5 * (lo + hi);
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        if (fabs(xn - x) <= 4e-16 * fabs(xn)) { x = xn; break; }
        x = xn;
    }
    return x;
}

// cube [B × ncol] → theta [B × tstride] (columns 0 … ncol−1).  grid = ceil(B / 128), block = 128.
__global__ void prior_transform_kernel(const PriorSpec* __restrict__ priors, int ncol, int B, const double* __restrict__ cube,
                                       double* __restrict__ theta, int tstride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double row[PRIOR_MAX_COLS];
    for (int k = 0; k < ncol; k++) {
        const PriorSpec p = priors[k];
        const double u = cube[(size_t)i * ncol + k];
        double v;
        switch (p.kind) {
            case PRIOR_UNIFORM:      v = p.p0 + u * (p.p1 - p.p0); break;
            case PRIOR_UNIFORM_FROM: { const double lo = row[p.ref_col]; v = lo + u * (p.p1 - lo); } break;
            case PRIOR_LOGUNIFORM:   { const double la = log(p.p0), lb = log(p.p1); v = exp(la + u * (lb - la)); } break;
            case PRIOR_NORMAL:       v = fma(p.p1, normcdfinv(u), p.p0); break;
            case PRIOR_LOGNORMAL:    v = exp(fma(p.p1, normcdfinv(u), p.p0)); break;
            default:                 v = p.p1 * gamma_int_quantile((int)p.p0, u); break;      // Gamma(shape p0, scale p1)
        }
        row[k] = v;
        theta[(size_t)i * tstride + k] = v;
    }
}

}  // namespace pioran
