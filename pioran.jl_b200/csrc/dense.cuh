// dense.cuh — K4: batched dense-Cholesky log-likelihood, the drop-in for  log_likelihood_direct(cov, t, y, σ²)
// (src/direct_solver.jl:6-21) with the kernel  k(τ) = Σ_m exp(−c_m τ)(a_m cos d_m τ + b_m sin d_m τ)
// (src/Celerite.jl:42-44 summed over terms, src/acvf.jl:138-140).  Used as the small-N cross-check of K2.
//
// Per parameter vector the (N+1)×(N+1) augmented matrix
//        [ K + diag(ν σ²)   · ]
//        [ (y − μ)ᵀ         0 ]
// is factorised in place (lower triangle, row-major, leading dimension Npad = multiple of 64; rows beyond N are
// identity padding).  Row N of the factor is zᵀ = (L⁻¹(y−μ))ᵀ and the Schur complement left on its diagonal is
// −zᵀz, so the forward substitution of direct_solver.jl:16 comes out of the same blocked sweep:
//        +NLL = Σ log L_ii + ½ zᵀz + ½ N log 2π            (direct_solver.jl:20)
// Blocked right-looking Cholesky, NB = 64, three kernels per panel, all batched over θ (blockIdx.y):
//   dense_potrf_kernel : 64×64 diagonal block in shared memory; log-pivots, first non-positive pivot → info
//   dense_trsm_kernel  : row blocks below the diagonal block, one thread per row, L_kk broadcast from shared memory
//   dense_syrk_kernel  : trailing update C_ij −= A_ik A_jkᵀ (i ≥ j > k), 64×64 tiles, 4×4 register blocking
// Matrices stay in HBM/L2 (32 MB each at N = 2 000; the 126 MB L2 holds the working set of a few of them).
#pragma once
#include "common.cuh"

namespace pioran {

constexpr int DNB = 64;

__device__ __forceinline__ void tri_index(int p, int& ti, int& tj) {   // p → (ti ≥ tj) of a packed lower triangle
    ti = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= p) ti++;
    while (ti * (ti + 1) / 2 > p) ti--;
    tj = p - ti * (ti + 1) / 2;
}

// A[θ] lower tiles ← covariance.  grid = (ntile·(ntile+1)/2, B), block = 256.
__global__ void __launch_bounds__(256) dense_fill_kernel(double* __restrict__ A, int64_t ld, int64_t N,
                                                         const double* __restrict__ t, const double* __restrict__ y,
                                                         const double* __restrict__ s2, int Jt,
                                                         const double* __restrict__ a, const double* __restrict__ b,
                                                         const double* __restrict__ c, const double* __restrict__ d,
                                                         const double* __restrict__ mu, const double* __restrict__ nu,
                                                         int theta0) {
    extern __shared__ double sm[];
    double* ca = sm; double* cb = ca + Jt; double* cc = cb + Jt; double* cd = cc + Jt;
    double* ti_s = cd + Jt; double* tj_s = ti_s + DNB;
    const int th = blockIdx.y;
    int bi, bj;
    tri_index(blockIdx.x, bi, bj);
    for (int m = threadIdx.x; m < Jt; m += blockDim.x) {
        const size_t k = (size_t)(theta0 + th) * Jt + m;
        ca[m] = a[k]; cb[m] = b[k]; cc[m] = c[k]; cd[m] = d[k];
    }
    if (threadIdx.x < DNB) {
        const int64_t gi = (int64_t)bi * DNB + threadIdx.x;
        ti_s[threadIdx.x] = gi < N ? t[gi] : 0.0;
    } else if (threadIdx.x < 2 * DNB) {
        const int64_t gj = (int64_t)bj * DNB + threadIdx.x - DNB;
        tj_s[threadIdx.x - DNB] = gj < N ? t[gj] : 0.0;
    }
    __syncthreads();
    const double m_ = mu ? mu[theta0 + th] : 0.0, v_ = nu ? nu[theta0 + th] : 1.0;
    double* At = A + (size_t)th * ld * ld;
    for (int e = threadIdx.x; e < DNB * DNB; e += blockDim.x) {
        const int r = e / DNB, q = e - r * DNB;
        const int64_t gi = (int64_t)bi * DNB + r, gj = (int64_t)bj * DNB + q;
        if (gj > gi) continue;
        double val;
        if (gi < N) {                                  // covariance entry (src/Celerite.jl:42-44), τ = |t_i − t_j|
            const double tau = fabs(ti_s[r] - tj_s[q]);
            double k = 0.0;
            for (int m = 0; m < Jt; m++) {
                double si, co;
                sincos(cd[m] * tau, &si, &co);
                k += exp(-cc[m] * tau) * (ca[m] * co + cb[m] * si);
            }
            if (gi == gj) k += v_ * s2[gi];            // K + Diagonal(σ²)   (direct_solver.jl:12)
            val = k;
        } else if (gi == N) {
            val = gj < N ? y[gj] - m_ : 0.0;           // augmented row: (y − μ)ᵀ, corner 0
        } else {
            val = gi == gj ? 1.0 : 0.0;                // identity padding
        }
        At[gi * ld + gj] = val;
    }
}

// Diagonal block kb of every matrix.  grid = (1, B), block = 256.  acc[θ] = {Σ log L_ii, zᵀz}; info[θ] = first bad minor.
__global__ void __launch_bounds__(256) dense_potrf_kernel(double* __restrict__ A, int64_t ld, int64_t N, int kb,
                                                          double* __restrict__ acc, int* __restrict__ info) {
    __shared__ double L[DNB][DNB + 1];
    __shared__ double piv_s;
    const int th = blockIdx.y;
    double* At = A + (size_t)th * ld * ld + ((size_t)kb * DNB) * ld + (size_t)kb * DNB;
    for (int e = threadIdx.x; e < DNB * DNB; e += blockDim.x) {
        const int r = e / DNB, q = e - r * DNB;
        L[r][q] = q <= r ? At[(size_t)r * ld + q] : 0.0;
    }
    __syncthreads();
    double logsum = 0.0;
    for (int k = 0; k < DNB; k++) {
        const int64_t g = (int64_t)kb * DNB + k;
        if (threadIdx.x == 0) {
            double p = L[k][k];
            if (g < N) {
                if (!(p > 0.0)) {                      // PosDefException in the reference (direct_solver.jl:14)
                    if (info[th] == 0) info[th] = (int)(g + 1);
                    p = 1.0;
                }
                p = sqrt(p);
                logsum += log(p);
            } else if (g == N) {
                acc[2 * th + 1] = -p;                  // Schur complement of the augmented corner = −zᵀz
                p = 1.0;
            } else {
                p = 1.0;
            }
            L[k][k] = p;
            piv_s = p;
        }
        __syncthreads();
        const double ip = 1.0 / piv_s;
        for (int r = k + 1 + threadIdx.x; r < DNB; r += blockDim.x) L[r][k] *= ip;
        __syncthreads();
        const int rem = DNB - k - 1;
        for (int e = threadIdx.x; e < rem * rem; e += blockDim.x) {
            const int r = k + 1 + e / rem, q = k + 1 + e % rem;
            if (q <= r) L[r][q] -= L[r][k] * L[q][k];
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < DNB * DNB; e += blockDim.x) {
        const int r = e / DNB, q = e - r * DNB;
        if (q <= r) At[(size_t)r * ld + q] = L[r][q];
    }
    if (threadIdx.x == 0) acc[2 * th] += logsum;
}

// Row blocks i > kb: A_ik ← A_ik L_kk^{-T}.  grid = (nblk − kb − 1, B), block = 64 (one thread per row).
__global__ void __launch_bounds__(DNB) dense_trsm_kernel(double* __restrict__ A, int64_t ld, int kb) {
    __shared__ double L[DNB][DNB + 1];
    const int th = blockIdx.y;
    const int ib = kb + 1 + blockIdx.x;
    double* At = A + (size_t)th * ld * ld;
    const double* Lk = At + ((size_t)kb * DNB) * ld + (size_t)kb * DNB;
    double* Ai = At + ((size_t)ib * DNB) * ld + (size_t)kb * DNB;
    for (int e = threadIdx.x; e < DNB * DNB; e += DNB) {
        const int r = e / DNB, q = e - r * DNB;      // coalesced along q
        L[r][q] = Lk[(size_t)r * ld + q];
    }
    __syncthreads();
    double* row = Ai + (size_t)threadIdx.x * ld;     // 64 consecutive doubles, 16-byte aligned (ld, kb·64 even)
    double x[DNB];
#pragma unroll
    for (int j = 0; j < DNB; j += 2) {
        const double2 v = *reinterpret_cast<const double2*>(row + j);
        x[j] = v.x; x[j + 1] = v.y;
    }
#pragma unroll
    for (int j = 0; j < DNB; j++) {
        double v = x[j];
#pragma unroll
        for (int l = 0; l < j; l++) v = fma(-x[l], L[j][l], v);
        x[j] = v / L[j][j];
    }
#pragma unroll
    for (int j = 0; j < DNB; j += 2) *reinterpret_cast<double2*>(row + j) = make_double2(x[j], x[j + 1]);
}

// Trailing update: C_ij −= A_ik A_jkᵀ for kb < j ≤ i.  grid = (m(m+1)/2, B) with m = nblk − kb − 1, block = 256.
__global__ void __launch_bounds__(256) dense_syrk_kernel(double* __restrict__ A, int64_t ld, int kb) {
    constexpr int KH = DNB / 2;
    __shared__ __align__(16) double As[KH][DNB + 2];   // [k][row], one half of the panel width at a time
    __shared__ __align__(16) double Bs[KH][DNB + 2];
    const int th = blockIdx.y;
    int pi, pj;
    tri_index(blockIdx.x, pi, pj);
    const int ib = kb + 1 + pi, jb = kb + 1 + pj;
    double* At = A + (size_t)th * ld * ld;
    const double* Ai = At + ((size_t)ib * DNB) * ld + (size_t)kb * DNB;
    const double* Aj = At + ((size_t)jb * DNB) * ld + (size_t)kb * DNB;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++) acc[u][v] = 0.0;
    for (int half = 0; half < 2; half++) {
        if (half) __syncthreads();
        for (int e = threadIdx.x; e < DNB * KH; e += 256) {
            const int r = e / KH, k = e - r * KH;        // coalesced along k
            As[k][r] = Ai[(size_t)r * ld + half * KH + k];
            Bs[k][r] = Aj[(size_t)r * ld + half * KH + k];
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < KH; k++) {
            const double2 a01 = *reinterpret_cast<const double2*>(&As[k][ty * 4]);
            const double2 a23 = *reinterpret_cast<const double2*>(&As[k][ty * 4 + 2]);
            const double2 b01 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4]);
            const double2 b23 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4 + 2]);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int v = 0; v < 4; v++) acc[u][v] = fma(av[u], bv[v], acc[u][v]);
        }
    }
    double* C = At + ((size_t)ib * DNB) * ld + (size_t)jb * DNB;
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const int r = ty * 4 + u, q = tx * 4 + v;
            if (ib != jb || q <= r) C[(size_t)r * ld + q] -= acc[u][v];
        }
}

// acc → +NLL (direct_solver.jl:20); NaN where the matrix was not positive definite.
__global__ void dense_finish_kernel(const double* __restrict__ acc, const int* __restrict__ info, int64_t N, int B,
                                    double* __restrict__ nll) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const double v = acc[2 * i] + 0.5 * acc[2 * i + 1] + 0.5 * (double)N * 1.8378770664093453;
    nll[i] = info[i] ? __longlong_as_double(0x7ff8000000000000LL) : v;
}

}  // namespace pioran
