// wide_grad.cuh — K5w: gradient of the fused approx + celerite log-likelihood at ranks 65 … 96 (FP64, sm_100a).
//
// The reference differentiates logpdf with ForwardDiff whatever the number of basis functions (test/test_likelihood.jl:55); its
// benchmark grid goes to DRWCelerite J = 30 … 50 (benchmark/benchmarks.jl:16-18).  The warp-per-evaluation gradient kernels
// (blocked_grad.cuh, grad.cuh, grad_pipe.cuh) hold ranks up to 64.  Here: the register-file CTA kernel of wide.cuh on (value,
// tangent) pairs — one CTA per (parameter vector, θ-direction), thread (ty, tx) of a 16×16 grid owning the interleaved TS×TS
// tile of the full square state AND of its tangent (TS = 5, 6: 2·TS² ≤ 72 doubles per thread).  Same forward-mode rule as
// grad.cuh, on the amplitude-scaled state of celerite.cuh (T = diag(amp) S diag(amp), rows Ũ amplitude-free: cos + ρ sin,
// sin − ρ cos, or 1; ρ = b/a of the basis, src/psd.jl:249-252, 264-275): on the approx path the decay rates and frequencies belong
// to the spectral grid, so only the row amplitudes (through q = amp∘V − p), Σa, ν and μ carry tangents.  Directions k < ND: PSD parameters; k = ND: ν, which also reports ∂/∂norm through the
// homogeneity identity norm ∂/∂norm + ν ∂/∂ν = ½ yᵀK⁻¹y − N/2; ∂/∂μ moves only the right-hand side and rides along in every CTA
// as one more vector (g_μ), reported by the k = 0 CTA.
#pragma once
#include "wide.cuh"
#include "grad.cuh"

namespace pioran {

struct WideGradArgs {
    const WorkItem* work;       // one item per parameter vector (series pointers; theta_begin = θ index)
    const double* amp;          // [nθ × RP] row amplitudes (approx_grad_kernel, logical rows)
    const double* damp;         // [nθ × ND × RP]
    const double* suma;         // [nθ]
    const double* dsuma;        // [nθ × ND]
    const double* c; const double* d;   // [Jt] decay rates and frequencies of the terms (θ-independent on the approx path)
    const double* rho;          // [Jt] b/a of the term (0 for a real term)
    const int* term_row;        // as in the generic kernels
    int Jt, R, RP, ND;
    const double* theta;        // [nθ × pstride]; norm at column ND, ν at ND + 1, μ at ND + 2
    int pstride;
    double* logl;               // [nθ] or nullptr
    double* grad;               // [nθ × pstride]
};

// grid = nθ · (ND + 1); block = 256.
template <int TS>
__global__ void __launch_bounds__(WIDE_THREADS, 1) celerite_wide_grad_kernel(const WideGradArgs args) {
    constexpr int LD = 16 * TS, NWARP = WIDE_THREADS / 32;
    __shared__ __align__(16) double tabU[WIDE_CH][LD], tabV[WIDE_CH][LD], tabP[WIDE_CH][LD];     // Ũ (amplitude-free), V, φ
    __shared__ __align__(16) double amp_s[LD], damp_s[LD];
    __shared__ __align__(16) double qphi[LD], wphi[LD], dqphi[LD], dwphi[LD], p_s[LD], dp_s[LD];
    __shared__ double red[5 * NWARP];

    const int ND = args.ND;
    const int th = blockIdx.x / (ND + 1), k = blockIdx.x - th * (ND + 1);
    const WorkItem wk = args.work[th];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int Jt = args.Jt;
    const int64_t N = wk.N;
    const bool amp_dir = k < ND;
    const double* trow = args.theta + (size_t)th * args.pstride;
    const double nu = trow[ND + 1], mu = trow[ND + 2];
    const double suma = args.suma[th], dsuma = amp_dir ? args.dsuma[(size_t)th * ND + k] : 0.0;

    double Tm[TS][TS], dT[TS][TS];
#pragma unroll
    for (int i = 0; i < TS; i++)
#pragma unroll
        for (int j = 0; j < TS; j++) { Tm[i][j] = 0.0; dT[i][j] = 0.0; }
    for (int q = tid; q < WIDE_CH * LD; q += WIDE_THREADS) { (&tabU[0][0])[q] = 0.0; (&tabV[0][0])[q] = 0.0; (&tabP[0][0])[q] = 0.0; }
    if (tid < LD) {
        const bool live = tid < args.R;
        amp_s[tid] = live ? args.amp[(size_t)th * args.RP + tid] : 0.0;
        damp_s[tid] = (live && amp_dir) ? args.damp[((size_t)th * ND + k) * args.RP + tid] : 0.0;
    }
    const bool owner = tid < LD;
    double g = 0.0, dg = 0.0, gmu = 0.0, q = 0.0, dq = 0.0, w = 0.0, dw = 0.0;
    double zprev = 0.0, dzprev = 0.0, zmuprev = 0.0;
    double chi2 = 0.0, dchi2 = 0.0, chimu = 0.0, dlog = 0.0;
    double logacc = 0.0, dkeep = 1.0, dfirst = 1.0;
    const bool b3 = (tx & 8) != 0, b2 = (tx & 4) != 0, b1 = (tx & 2) != 0;
    __syncthreads();

    for (int64_t nb = 0; nb < N; nb += WIDE_CH) {
        const int ns = (int)((N - nb) < WIDE_CH ? (N - nb) : WIDE_CH);
        for (int idx = tid; idx < ns * Jt; idx += WIDE_THREADS) {           // celerite_solver.jl:51-64, amplitude-free rows
            const int s = idx / Jt, m = idx - s * Jt;
            const int64_t n = nb + s;
            const double tn = wk.t[n];
            const double ph = (n >= 1) ? exp(-args.c[m] * (tn - wk.t[n - 1])) : 0.0;
            const int tr = args.term_row[m];
            if (tr < 0) {
                const int r0 = -tr - 1;
                tabU[s][r0] = 1.0; tabV[s][r0] = 1.0; tabP[s][r0] = ph;
            } else {
                double si, co;
                sincos_large(args.d[m] * tn, &si, &co);
                const double rho = args.rho[m];
                tabU[s][tr] = fma(rho, si, co);          // (a cos + b sin) / a
                tabU[s][tr + 1] = fma(-rho, co, si);     // (a sin − b cos) / a
                tabV[s][tr] = co; tabV[s][tr + 1] = si;
                tabP[s][tr] = ph; tabP[s][tr + 1] = ph;
            }
        }
        __syncthreads();
        for (int s = 0; s < ns; s++) {
            const int64_t n = nb + s;
            const double* Un = tabU[s];
            const double* Vn = tabV[s];
            const double* Pn = tabP[s];
            // ---- phase 0: owners publish (q, w)φ and their tangents, advance g, ġ, g_μ
            if (owner) {
                const double ph = Pn[tid];
                qphi[tid] = q * ph; wphi[tid] = w * ph;
                dqphi[tid] = dq * ph; dwphi[tid] = dw * ph;
                dg = ph * fma(dw, zprev, fma(w, dzprev, dg));
                gmu = ph * fma(w, zmuprev, gmu);
                g = ph * fma(w, zprev, g);                 // celerite_solver.jl:137
            }
            __syncthreads();
            // ---- phase 1: tile update and row sums, value and tangent
            double phr[TS], qr[TS], dqr[TS], rs[8], drs[8];
#pragma unroll
            for (int i = 0; i < TS; i++) { phr[i] = Pn[ty + 16 * i]; qr[i] = qphi[ty + 16 * i]; dqr[i] = dqphi[ty + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 8; i++) { rs[i] = 0.0; drs[i] = 0.0; }
#pragma unroll
            for (int j = 0; j < TS; j++) {
                const int cidx = tx + 16 * j;
                const double pc = Pn[cidx], wc = wphi[cidx], dwc = dwphi[cidx], uc = Un[cidx];
#pragma unroll
                for (int i = 0; i < TS; i++) {
                    const double t = fma(phr[i], pc * Tm[i][j], qr[i] * wc);                       // celerite_solver.jl:76
                    const double dt = fma(phr[i], pc * dT[i][j], fma(dqr[i], wc, qr[i] * dwc));
                    Tm[i][j] = t; dT[i][j] = dt;
                    rs[i] = fma(t, uc, rs[i]);
                    drs[i] = fma(dt, uc, drs[i]);
                }
            }
            {   // reduce-scatter of the (padded) 8 row sums over the 16 lanes that share ty, both components
                double e4[4], e2[2], f4[4], f2[2];
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const double r0 = __shfl_xor_sync(FULL, b3 ? rs[m] : rs[m + 4], 8);
                    const double r1 = __shfl_xor_sync(FULL, b3 ? drs[m] : drs[m + 4], 8);
                    e4[m] = (b3 ? rs[m + 4] : rs[m]) + r0;
                    f4[m] = (b3 ? drs[m + 4] : drs[m]) + r1;
                }
#pragma unroll
                for (int m = 0; m < 2; m++) {
                    const double r0 = __shfl_xor_sync(FULL, b2 ? e4[m] : e4[m + 2], 4);
                    const double r1 = __shfl_xor_sync(FULL, b2 ? f4[m] : f4[m + 2], 4);
                    e2[m] = (b2 ? e4[m + 2] : e4[m]) + r0;
                    f2[m] = (b2 ? f4[m + 2] : f4[m]) + r1;
                }
                const double r0 = __shfl_xor_sync(FULL, b1 ? e2[0] : e2[1], 2);
                const double r1 = __shfl_xor_sync(FULL, b1 ? f2[0] : f2[1], 2);
                double tot = (b1 ? e2[1] : e2[0]) + r0;
                double dtot = (b1 ? f2[1] : f2[0]) + r1;
                tot += __shfl_xor_sync(FULL, tot, 1);
                dtot += __shfl_xor_sync(FULL, dtot, 1);
                const int isel = 4 * (b3 ? 1 : 0) + 2 * (b2 ? 1 : 0) + (b1 ? 1 : 0);
                if ((tx & 1) == 0 && isel < TS) { p_s[ty + 16 * isel] = tot; dp_s[ty + 16 * isel] = dtot; }
            }
            __syncthreads();
            // ---- phase 2: UᵀTU, Uᵀg and their tangents, Uᵀg_μ
            double p = 0.0, dp = 0.0, v5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            if (owner) {
                p = p_s[tid]; dp = dp_s[tid];
                const double u = Un[tid];
                v5[0] = u * p;
                v5[1] = u * dp;
                v5[2] = u * g;
                v5[3] = u * dg;
                v5[4] = u * gmu;
            }
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1)
#pragma unroll
                for (int m = 0; m < 5; m++) v5[m] += __shfl_xor_sync(FULL, v5[m], sft);
            if (lane == 0)
#pragma unroll
                for (int m = 0; m < 5; m++) red[5 * warp + m] = v5[m];
            __syncthreads();
            double t5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int wq = 0; wq < NWARP; wq++)
#pragma unroll
                for (int m = 0; m < 5; m++) t5[m] += red[5 * wq + m];
            // ---- phase 3: pivot, innovation, next q and w — on pairs
            const double s2n = wk.s2[n];
            const double D = fma(nu, s2n, suma) - t5[0];                 // celerite_solver.jl:92
            const double dD = (amp_dir ? dsuma : s2n) - t5[1];
            const double z = (wk.y[n] - mu) - t5[2];                     // celerite_solver.jl:141
            const double dz = -t5[3];
            const double zmu = -1.0 - t5[4];
            const double rD = fast_rcp(D);
            if (owner) {
                q = Vn[tid] * amp_s[tid] - p;                            // amplitude-scaled state: q = amp∘V − p
                dq = Vn[tid] * damp_s[tid] - dp;
                w = q * rD;                                               // celerite_solver.jl:95-98
                dw = (dq - w * dD) * rD;
            }
            zprev = z; dzprev = dz; zmuprev = zmu;
            chi2 = fma(z * z, rD, chi2);
            dchi2 = fma(fma(2.0 * z, dz, -(z * z) * (dD * rD)), rD, dchi2);
            chimu = fma(2.0 * z * zmu, rD, chimu);
            dlog = fma(dD, rD, dlog);
            if (warp == NWARP - 1) {                                     // celerite_solver.jl:126,140
                if (n == 0) dfirst = D;
                else if ((int)(n & 31) == lane) dkeep = D;
                if ((n & 31) == 31) { logacc += log(fabs(dkeep)); dkeep = 1.0; }
            }
        }
        __syncthreads();
    }
    if (warp == NWARP - 1) {
        double la = logacc + log(fabs(dkeep));
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) la += __shfl_xor_sync(FULL, la, sft);
        const double logdet = log(dfirst) + la;
        if (lane == 0) {
            const int P = args.pstride;
            if (k == 0 && args.logl) args.logl[th] = -logdet / 2 - (double)N * 1.8378770664093453 / 2 - chi2 / 2;   // celerite_solver.jl:333
            const double gk = -dlog / 2 - dchi2 / 2;
            if (k < ND) {
                args.grad[(size_t)th * P + k] = gk;
            } else {      // the ν CTA also reports ∂/∂norm
                args.grad[(size_t)th * P + ND + 1] = gk;
                args.grad[(size_t)th * P + ND] = (0.5 * chi2 - 0.5 * (double)N - nu * gk) / trow[ND];
            }
            if (k == 0) args.grad[(size_t)th * P + ND + 2] = -chimu / 2;
        }
    }
}

}  // namespace pioran
