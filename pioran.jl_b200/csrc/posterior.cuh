// posterior.cuh — posterior mean of the celerite GP (widening row SURVEY §8f #2): batched `pred` of the reference
// (src/celerite_solver.jl:376-483) = factor + solve_prec! + two O(N + M) sweeps, for B parameter vectors at once.
//
//   1. the generic K2 kernel in STEP_STORE mode (celerite.cuh) leaves D_n, W_n and the forward-substitution z_n in HBM;
//   2. celerite_backsolve_kernel finishes solve_prec! (celerite_solver.jl:145-155):
//          z_N ← z_N / D_N;   g ← φ_n ∘ (g + U_{n+1} z_{n+1});   z_n ← z_n / D_n − W_nᵀ g          → z = K⁻¹ (y − μ)
//   3. celerite_predict_kernel evaluates  mean(τ_m) = μ + Σ_n k(|τ_m − t_n|) z_n  with the two sweeps of
//      celerite_solver.jl:400-480: a forward recursion over the data points with t_n < τ_m and a backward one over the
//      points with t_n ≥ τ_m (n₀ = searchsortedfirst(t, τ) − 1, computed once per call on the host — it does not depend
//      on the parameter vector).
// One warp per parameter vector; lane l owns the celerite terms l, l + 32, … (PT per lane: 2 up to 64 terms, 4 up to 128).  cos/sin are taken at absolute
// times like the reference (sincos_large, common.cuh).
#pragma once
#include "common.cuh"

namespace pioran {

struct PostArgs {
    const double* t; const double* tau;     // data times [N], prediction times [M] (ascending)
    const int* n0;                          // [M] number of data times strictly below τ_m
    int64_t N; int64_t M;
    int B, Jt, RPL;                         // RPL = logical row count of the stored factor (8·BS)
    const double* a; const double* b; const double* c; const double* d;   // [B × Jt]
    const int* term_row;                    // as in the generic K2 kernel
    const double* mu;                       // [B] or nullptr
    const double* W; const double* D;       // stored factor [B × N × RPL], [B × N]
    double* z;                              // in: forward z [B × N]; out: K⁻¹(y − μ)
    double* mean;                           // [B × M]
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
    return v;
}

// grid = ceil(B / 4), block = 128 (4 warps, one parameter vector each)
template <int PT>
__global__ void __launch_bounds__(128) celerite_backsolve_kernel(const PostArgs pa) {
    const int th = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (th >= pa.B) return;
    const int Jt = pa.Jt;
    const int64_t N = pa.N;
    double ca[PT], cb[PT], cc[PT], cd[PT], gc[PT], gs[PT];
    int r0[PT], r1[PT];
#pragma unroll
    for (int k = 0; k < PT; k++) {
        const int m = lane + 32 * k;
        const bool on = m < Jt;
        gc[k] = 0.0; gs[k] = 0.0;
        ca[k] = on ? pa.a[(size_t)th * Jt + m] : 0.0; cb[k] = on ? pa.b[(size_t)th * Jt + m] : 0.0;
        cc[k] = on ? pa.c[(size_t)th * Jt + m] : 0.0; cd[k] = on ? pa.d[(size_t)th * Jt + m] : 0.0;
        const int tr = on ? pa.term_row[m] : -1;
        r0[k] = !on ? -1 : (tr < 0 ? -tr - 1 : tr);
        r1[k] = (!on || tr < 0) ? -1 : tr + 1;          // real terms have no sin-row
    }
    const double* W = pa.W + (size_t)th * N * pa.RPL;
    const double* D = pa.D + (size_t)th * N;
    double* z = pa.z + (size_t)th * N;
    double zn1 = z[N - 1] / D[N - 1];                   // celerite_solver.jl:145
    __syncwarp();
    if (lane == 0) z[N - 1] = zn1;
    for (int64_t n = N - 2; n >= 0; n--) {              // celerite_solver.jl:146-155
        const double tn1 = pa.t[n + 1], dt = tn1 - pa.t[n];
        double part = 0.0;
#pragma unroll
        for (int k = 0; k < PT; k++) {
            if (r0[k] >= 0) {
                const double ph = exp(-cc[k] * dt);
                if (r1[k] >= 0) {
                    double si, co;
                    sincos_large(cd[k] * tn1, &si, &co);
                    gc[k] = (gc[k] + (ca[k] * co + cb[k] * si) * zn1) * ph;
                    gs[k] = (gs[k] + (ca[k] * si - cb[k] * co) * zn1) * ph;
                    part = fma(W[n * pa.RPL + r0[k]], gc[k], part);
                    part = fma(W[n * pa.RPL + r1[k]], gs[k], part);
                } else {
                    gc[k] = (gc[k] + ca[k] * zn1) * ph;
                    part = fma(W[n * pa.RPL + r0[k]], gc[k], part);
                }
            }
        }
        const double s = warp_sum(part);
        const double zf = z[n];
        zn1 = zf / D[n] - s;
        __syncwarp();
        if (lane == 0) z[n] = zn1;
    }
}

// grid = ceil(B / 4), block = 128.  mean[th][m] = μ_th + Σ_n k(|τ_m − t_n|) z_n.
template <int PT>
__global__ void __launch_bounds__(128) celerite_predict_kernel(const PostArgs pa) {
    const int th = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (th >= pa.B) return;
    const int Jt = pa.Jt;
    const int64_t N = pa.N, M = pa.M;
    double ca[PT], cb[PT], cc[PT], cd[PT];
    bool on[PT];
#pragma unroll
    for (int k = 0; k < PT; k++) {
        const int m = lane + 32 * k;
        on[k] = m < Jt;
        ca[k] = on[k] ? pa.a[(size_t)th * Jt + m] : 0.0; cb[k] = on[k] ? pa.b[(size_t)th * Jt + m] : 0.0;
        cc[k] = on[k] ? pa.c[(size_t)th * Jt + m] : 0.0; cd[k] = on[k] ? pa.d[(size_t)th * Jt + m] : 0.0;
    }
    const double* z = pa.z + (size_t)th * N;
    double* mean = pa.mean + (size_t)th * M;
    const double mu = pa.mu ? pa.mu[th] : 0.0;

    // ---- forward sweep (celerite_solver.jl:400-435): Q_j = Σ_{n ≤ n_abs} z_n (cos, sin)(d_j t_n) e^{−c_j (t_{n_abs} − t_n)}
    double qc[PT], qs[PT];
#pragma unroll
    for (int k = 0; k < PT; k++) { qc[k] = 0.0; qs[k] = 0.0; }
    int64_t nabs = 0;                                   // data points absorbed so far; Q refers to time t[nabs − 1]
    for (int64_t m = 0; m < M; m++) {
        const int64_t n0 = pa.n0[m];
        while (nabs < n0) {
            const double tn = pa.t[nabs], zn = z[nabs];
            const double dt = nabs > 0 ? tn - pa.t[nabs - 1] : 0.0;
#pragma unroll
            for (int k = 0; k < PT; k++)
                if (on[k]) {
                    double si, co;
                    sincos_large(cd[k] * tn, &si, &co);
                    const double ph = nabs > 0 ? exp(-cc[k] * dt) : 0.0;
                    qc[k] = fma(qc[k], ph, zn * co);
                    qs[k] = fma(qs[k], ph, zn * si);
                }
            nabs++;
        }
        double part = 0.0;
        if (n0 > 0) {
            const double tm = pa.tau[m], dt = tm - pa.t[n0 - 1];
#pragma unroll
            for (int k = 0; k < PT; k++)
                if (on[k]) {
                    double si, co;
                    sincos_large(cd[k] * tm, &si, &co);
                    const double e = exp(-cc[k] * dt);
                    part += e * (qc[k] * (ca[k] * co + cb[k] * si) + qs[k] * (ca[k] * si - cb[k] * co));
                }
        }
        const double tot = warp_sum(part);
        if (lane == 0) mean[m] = mu + tot;
    }
    __syncwarp();
    // ---- backward sweep (celerite_solver.jl:439-480): P_j = Σ_{n ≥ n_abs} z_n U_j(t_n) e^{−c_j (t_n − t_{n_abs})}
    double pc[PT], ps[PT];
#pragma unroll
    for (int k = 0; k < PT; k++) { pc[k] = 0.0; ps[k] = 0.0; }
    nabs = N;                                           // points nabs … N−1 absorbed; P refers to time t[nabs]
    for (int64_t m = M - 1; m >= 0; m--) {
        const int64_t n0 = pa.n0[m];
        while (nabs > n0) {
            const int64_t n = nabs - 1;
            const double tn = pa.t[n], zn = z[n];
            const double dt = nabs < N ? pa.t[nabs] - tn : 0.0;
#pragma unroll
            for (int k = 0; k < PT; k++)
                if (on[k]) {
                    double si, co;
                    sincos_large(cd[k] * tn, &si, &co);
                    const double ph = nabs < N ? exp(-cc[k] * dt) : 0.0;
                    pc[k] = fma(pc[k], ph, zn * (ca[k] * co + cb[k] * si));
                    ps[k] = fma(ps[k], ph, zn * (ca[k] * si - cb[k] * co));
                }
            nabs--;
        }
        double part = 0.0;
        if (n0 < N) {
            const double tm = pa.tau[m], dt = pa.t[n0] - tm;
#pragma unroll
            for (int k = 0; k < PT; k++)
                if (on[k]) {
                    double si, co;
                    sincos_large(cd[k] * tm, &si, &co);
                    part += exp(-cc[k] * dt) * (pc[k] * co + ps[k] * si);
                }
        }
        const double tot = warp_sum(part);
        if (lane == 0) mean[m] += tot;
    }
}

}  // namespace pioran
