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
// Blocked right-looking Cholesky, NB = 64, all kernels batched over θ (blockIdx.y):
//   dense_potrf_kernel : 64×64 diagonal block in registers, one barrier per pivot; log-pivots, first non-positive pivot → info
//   dense_trsm_kernel  : row blocks below the diagonal block, blocked substitution on the FP64 tensor cores
//   dense_syrk_kernel  : trailing update C_ij −= A_ik A_jkᵀ (i ≥ j > k), 64×64 tiles on the FP64 tensor cores (DMMA m8n8k4),
//                        panels applied in groups of four (api.cu: pioran_direct_logl)
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

// Separable factors of the kernel per FULL 64-block, for dense_fill_kernel (round 2).  With α = t_k − t_first(b) ≥ 0 (row role)
// and γ = t_last(b) − t_k ≥ 0 (column role):
//     P = e^{−cα}(a cos dα + b sin dα),  Q = e^{−cα}(b cos dα − a sin dα),  Gc = e^{−cγ} cos dγ,  Gs = e^{−cγ} sin dγ
// tab[θ][b][P|Q|Gc|Gs][Jt][64].  grid = (nfull, B), block = 256.  2·64 transcendental triples per term and BLOCK — the tiles of
// a block row/column share them (they were recomputed per tile before: 1.9 of C5's 12 ms went into the fill).
__global__ void __launch_bounds__(256) dense_fill_tables_kernel(double* __restrict__ tab, int nfull, const double* __restrict__ t,
                                                                int Jt, const double* __restrict__ a, const double* __restrict__ b,
                                                                const double* __restrict__ c, const double* __restrict__ d,
                                                                int theta0) {
    const int th = blockIdx.y, blk = blockIdx.x;
    const size_t JD = (size_t)Jt * DNB;
    double* T = tab + ((size_t)th * nfull + blk) * 4 * JD;
    const double t0 = t[(int64_t)blk * DNB], t1 = t[(int64_t)blk * DNB + DNB - 1];
    for (int e = threadIdx.x; e < Jt * DNB; e += blockDim.x) {
        const int m = e / DNB, k = e - m * DNB;
        const size_t idx = (size_t)(theta0 + th) * Jt + m;
        const double am = a[idx], bm = b[idx], cm = c[idx], dm = d[idx];
        const double tk = t[(int64_t)blk * DNB + k];
        double si, co;
        const double al = tk - t0;
        sincos(dm * al, &si, &co);
        double ex = exp(-cm * al);
        T[e] = ex * (am * co + bm * si);
        T[JD + e] = ex * (bm * co - am * si);
        const double ga = t1 - tk;
        sincos(dm * ga, &si, &co);
        ex = exp(-cm * ga);
        T[2 * JD + e] = ex * co;
        T[3 * JD + e] = ex * si;
    }
}

constexpr int DFILL_DIAG = 8 * 36;      // entries of the 8 diagonal 8×8 sub-tiles of a diagonal tile
__host__ __device__ inline size_t dense_fill_smem_doubles(int Jt) {
    return 4 * (size_t)Jt + 2 * DNB + 4 * (size_t)Jt * DNB + 24 * (size_t)Jt + 4 * DFILL_DIAG;
}

// A[θ] lower tiles ← covariance.  grid = (ntile·(ntile+1)/2, B), block = 256; big tiles first.
__global__ void __launch_bounds__(256) dense_fill_kernel(double* __restrict__ A, int64_t ld, int64_t N,
                                                         const double* __restrict__ t, const double* __restrict__ y,
                                                         const double* __restrict__ s2, int Jt,
                                                         const double* __restrict__ a, const double* __restrict__ b,
                                                         const double* __restrict__ c, const double* __restrict__ d,
                                                         const double* __restrict__ mu, const double* __restrict__ nu,
                                                         int theta0, const double* __restrict__ tab, int nfull, int direct, int fpart, int G) {
    extern __shared__ double sm[];
    double* ca = sm; double* cb = ca + Jt; double* cc = cb + Jt; double* cd = cc + Jt;
    double* ti_s = cd + Jt; double* tj_s = ti_s + DNB;
    double* Ps = tj_s + DNB;            // [Jt][DNB] each
    double* Qs = Ps + (size_t)Jt * DNB;
    double* Xs = Qs + (size_t)Jt * DNB;
    double* Ys = Xs + (size_t)Jt * DNB;
    double* rot = Ys + (size_t)Jt * DNB;    // [8 warps][3][Jt]: e^{−cβ}, cos dβ, sin dβ of a (row block, column block) pair
    double* part = rot + 24 * (size_t)Jt;   // [4][DFILL_DIAG]
    const int th = blockIdx.y;
    int bi, bj;
    // fpart 0: the whole lower triangle; 1: the first G block columns (grid.x = nblk·G); 2: the triangle behind them
    if (fpart == 1) {
        bi = blockIdx.x / G; bj = blockIdx.x - bi * G;
        if (bj > bi) return;
    } else {
        tri_index(gridDim.x - 1 - blockIdx.x, bi, bj);
        if (fpart == 2) { bi += G; bj += G; }
    }
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
    const int JD = Jt * DNB;
    // Off-diagonal tiles: every row time ≥ every column time, and the kernel separates.  With α = t_i − t_r0 (r0 = first row of
    // the tile), β = t_r0 − t_c1 (c1 = last column), γ = t_c1 − t_j, all ≥ 0 and τ = α+β+γ:
    //     e^{−cτ}(a cos dτ + b sin dτ) = P_i X_j + Q_i Y_j,    X + iY = e^{−cβ} e^{i dβ} (Gc_j + i Gs_j)
    // P, Q, Gc, Gs come from the block tables, the rotation by β costs one transcendental triple per term and TILE, then two FMAs
    // per entry and term.  No exponent is positive, so nothing overflows (the celerite instability of separating
    // e^{−c t_i} e^{+c t_j} globally does not arise per tile).  A partial last row block computes its few valid P, Q rows here.
    if (bi != bj && !direct) {
        const bool full = bi < nfull;
        for (int m = threadIdx.x; m < Jt; m += blockDim.x) {
            const double beta = fmax(ti_s[0] - tj_s[DNB - 1], 0.0);      // (an all-padding row block has no times; its rows are overwritten below)
            double si, co;
            sincos(cd[m] * beta, &si, &co);
            rot[m] = exp(-cc[m] * beta); rot[Jt + m] = co; rot[2 * Jt + m] = si;
        }
        __syncthreads();
        const double* Tj = tab + ((size_t)th * nfull + bj) * 4 * JD;
        for (int e = threadIdx.x; e < JD; e += blockDim.x) {
            const int m = e / DNB;
            const double gc = Tj[2 * (size_t)JD + e], gs = Tj[3 * (size_t)JD + e], ex = rot[m], co = rot[Jt + m], si = rot[2 * Jt + m];
            Xs[e] = ex * (co * gc - si * gs);
            Ys[e] = ex * (si * gc + co * gs);
        }
        if (full) {
            const double* Ti = tab + ((size_t)th * nfull + bi) * 4 * JD;
            for (int e = threadIdx.x; e < JD; e += blockDim.x) { Ps[e] = Ti[e]; Qs[e] = Ti[(size_t)JD + e]; }
        } else {
            const double tr0 = ti_s[0];
            for (int e = threadIdx.x; e < JD; e += blockDim.x) {
                const int m = e / DNB, k = e - m * DNB;
                double pv = 0.0, qv = 0.0;
                if ((int64_t)bi * DNB + k < N) {
                    const double al = ti_s[k] - tr0;
                    double si, co;
                    sincos(cd[m] * al, &si, &co);
                    const double ex = exp(-cc[m] * al);
                    pv = ex * (ca[m] * co + cb[m] * si);
                    qv = ex * (cb[m] * co - ca[m] * si);
                }
                Ps[e] = pv; Qs[e] = qv;
            }
        }
        __syncthreads();
        // 4×4 entries per thread: rows 4 ty + u, columns tx + 16 w — the 16 lanes of a half-warp read 16 consecutive doubles of X / Y
        // (one wavefront; columns 4 tx + w cost two per load and made the loop shared-memory bound: ncu LSU 80 %, FP64 34 %)
        const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
        double v[4][4];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int w = 0; w < 4; w++) v[u][w] = 0.0;
        for (int m = 0; m < Jt; m++) {
            double pr[4], qr[4], xc[4], yc[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { pr[u] = Ps[m * DNB + 4 * ty + u]; qr[u] = Qs[m * DNB + 4 * ty + u]; }
#pragma unroll
            for (int w = 0; w < 4; w++) { xc[w] = Xs[m * DNB + tx + 16 * w]; yc[w] = Ys[m * DNB + tx + 16 * w]; }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int w = 0; w < 4; w++) v[u][w] = fma(pr[u], xc[w], fma(qr[u], yc[w], v[u][w]));
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t gi = (int64_t)bi * DNB + 4 * ty + u;
            if (!full && gi >= N) {                    // augmented row (y − μ)ᵀ, then identity padding (zeros off the diagonal)
#pragma unroll
                for (int w = 0; w < 4; w++) v[u][w] = gi == N ? y[(int64_t)bj * DNB + tx + 16 * w] - m_ : 0.0;
            }
            double* dst = At + gi * ld + (int64_t)bj * DNB + tx;
#pragma unroll
            for (int w = 0; w < 4; w++) dst[16 * w] = v[u][w];
        }
        return;
    }
    // Full diagonal tiles: the same separation one level down, on the 8×8 grid of 8×8 sub-tiles.  The 28 sub-tiles below the
    // diagonal use P, Q relative to the first row of their 8-row group and Gc, Gs relative to the last column of their 8-column
    // group, rotated by the gap β between the two groups; only the 8 diagonal sub-tiles (288 entries) are evaluated directly.
    if (bi == bj && bi < nfull && !direct) {
        for (int e = threadIdx.x; e < 2 * JD; e += blockDim.x) {
            const int side = e / JD, f = e - side * JD, m = f / DNB, k = f - m * DNB;
            double si, co;
            if (side == 0) {
                const double al = ti_s[k] - ti_s[k & ~7];
                sincos(cd[m] * al, &si, &co);
                const double ex = exp(-cc[m] * al);
                Ps[f] = ex * (ca[m] * co + cb[m] * si);
                Qs[f] = ex * (cb[m] * co - ca[m] * si);
            } else {
                const double ga = ti_s[k | 7] - ti_s[k];
                sincos(cd[m] * ga, &si, &co);
                const double ex = exp(-cc[m] * ga);
                Xs[f] = ex * co;
                Ys[f] = ex * si;
            }
        }
        __syncthreads();
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        double* wr = rot + (size_t)warp * 3 * Jt;
        for (int p = warp; p < 28; p += 8) {
            int I, J;
            tri_index(p, I, J);
            I += 1;                                     // strictly below the diagonal: I > J
            const double beta = ti_s[8 * I] - ti_s[8 * J + 7];
            for (int m = lane; m < Jt; m += 32) {
                double si, co;
                sincos(cd[m] * beta, &si, &co);
                wr[m] = exp(-cc[m] * beta); wr[Jt + m] = co; wr[2 * Jt + m] = si;
            }
            __syncwarp();
            const int r = 8 * I + (lane >> 2), q = 8 * J + 2 * (lane & 3);
            double v0 = 0.0, v1 = 0.0;
            for (int m = 0; m < Jt; m++) {
                const double ex = wr[m], co = wr[Jt + m], si = wr[2 * Jt + m];
                const double pr = Ps[m * DNB + r] * ex, qr = Qs[m * DNB + r] * ex;
                const double2 gc = *reinterpret_cast<const double2*>(Xs + m * DNB + q);
                const double2 gs = *reinterpret_cast<const double2*>(Ys + m * DNB + q);
                v0 = fma(pr, co * gc.x - si * gs.x, fma(qr, si * gc.x + co * gs.x, v0));
                v1 = fma(pr, co * gc.y - si * gs.y, fma(qr, si * gc.y + co * gs.y, v1));
            }
            *reinterpret_cast<double2*>(At + ((int64_t)bi * DNB + r) * ld + (int64_t)bj * DNB + q) = make_double2(v0, v1);
            __syncwarp();
        }
        for (int it = threadIdx.x; it < 4 * DFILL_DIAG; it += blockDim.x) {     // (entry, quarter of the terms)
            const int qt = it / DFILL_DIAG, ent = it - qt * DFILL_DIAG, I = ent / 36;
            int r, q;
            tri_index(ent - 36 * I, r, q);
            const double tau = ti_s[8 * I + r] - ti_s[8 * I + q];
            double k = 0.0;
            for (int m = qt; m < Jt; m += 4) {
                double si, co;
                sincos(cd[m] * tau, &si, &co);
                k += exp(-cc[m] * tau) * (ca[m] * co + cb[m] * si);
            }
            part[it] = k;
        }
        __syncthreads();
        for (int ent = threadIdx.x; ent < DFILL_DIAG; ent += blockDim.x) {
            const int I = ent / 36;
            int r, q;
            tri_index(ent - 36 * I, r, q);
            const int64_t gi = (int64_t)bi * DNB + 8 * I + r, gj = (int64_t)bj * DNB + 8 * I + q;
            double k = (part[ent] + part[DFILL_DIAG + ent]) + (part[2 * DFILL_DIAG + ent] + part[3 * DFILL_DIAG + ent]);
            if (gi == gj) k += v_ * s2[gi];            // K + Diagonal(σ²)   (direct_solver.jl:12)
            At[gi * ld + gj] = k;
        }
        return;
    }
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
    // One barrier per pivot, block in REGISTERS (round 2; three barriers and a shared-memory read-modify-write of every entry
    // before): thread (ty, tx) of a 16×16 grid owns the entries (ty + 16 i, tx + 16 j).  Column k is never written after step k,
    // so its owners publish it UNSCALED through a double-buffered shared column, every thread reads the pivot p_k itself and
    // updates e_rq −= L_rk L_qk / p_k; the scaling by 1/√p_k happens once, when the block is written back.
    __shared__ double col_s[2][DNB];
    __shared__ double sp_s[DNB];                       // √pivot of every column
    const int th = blockIdx.y, ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double* At = A + (size_t)th * ld * ld + ((size_t)kb * DNB) * ld + (size_t)kb * DNB;
    double e[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = ty + 16 * i, q = tx + 16 * j;
            e[i][j] = q <= r ? At[(size_t)r * ld + q] : 0.0;
        }
#pragma unroll
    for (int j0 = 0; j0 < 4; j0++) {
        for (int kk = 0; kk < 16; kk++) {
            const int k = 16 * j0 + kk;
            double* col = col_s[k & 1];
            if (tx == kk) {
#pragma unroll
                for (int i = 0; i < 4; i++) col[ty + 16 * i] = e[i][j0];
            }
            __syncthreads();
            const int64_t g = (int64_t)kb * DNB + k;
            double p = col[k];
            if (g < N) {
                if (!(p > 0.0)) {                      // PosDefException in the reference (direct_solver.jl:14)
                    if (threadIdx.x == 0 && info[th] == 0) info[th] = (int)(g + 1);
                    p = 1.0;
                }
            } else {
                if (g == N && threadIdx.x == 0) acc[2 * th + 1] = -p;   // Schur complement of the augmented corner = −zᵀz
                p = 1.0;
            }
            if (threadIdx.x == 0) sp_s[k] = p;         // pivot now; √ and log after the loop, 64 at a time, off the critical path
            const double ip = 1.0 / p;
            double lr[4], lq[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { lr[i] = col[ty + 16 * i] * ip; lq[i] = col[tx + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int r = ty + 16 * i, q = tx + 16 * j;
                    if (q > k && q <= r) e[i][j] = fma(-lr[i], lq[j], e[i][j]);
                }
        }
    }
    __syncthreads();
    if (threadIdx.x < DNB) {
        const double sp = sqrt(sp_s[threadIdx.x]);
        sp_s[threadIdx.x] = sp;
        double lg = log(sp);                           // Σ log L_ii of the block (padding and the augmented corner carry p = 1)
#pragma unroll
        for (int o = 16; o; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
        if ((threadIdx.x & 31) == 0) col_s[0][threadIdx.x >> 5] = lg;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = ty + 16 * i, q = tx + 16 * j;
            if (q < r) At[(size_t)r * ld + q] = e[i][j] / sp_s[q];
            else if (q == r) At[(size_t)r * ld + q] = sp_s[q];
        }
    if (threadIdx.x == 0) acc[2 * th] += col_s[0][0] + col_s[0][1];
}

// FP64 tensor-core tile product: D(8×8) += A(8×4, row-major) · B(4×8, column-major).  Fragments (PTX ISA, m8n8k4.f64):
// lane l holds A[l>>2][l&3], B[l&3][l>>2] and D[l>>2][2·(l&3) + {0,1}].
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Row blocks i > kb: A_ik ← A_ik L_kk^{-T}.  grid = (nblk − kb − 1, B), block = 64 = 2 warps of 32 rows.
// Blocked substitution on the tensor pipe (round 2; one thread per row with L_kk broadcast from shared memory before — bound by
// the shared-memory pipe at 14 % of the FP64 rate): a warp holds its 32 × 64 slice as 4 × 8 DMMA accumulator tiles.  Per 8-column
// group g: the 8 × 8 triangle L_gg is substituted inside the 4-lane groups that share a row (x_j travels by shuffle), −X_g goes
// through a warp-private shared slice into A-fragment order, and the later groups get X_g' −= X_g L_g'gᵀ as 4 × 2 DMMAs each.
__global__ void __launch_bounds__(DNB) dense_trsm_kernel(double* __restrict__ A, int64_t ld, int kb) {
    constexpr int LS = DNB + 4, XS = 12;             // row strides (doubles): fragment loads fall on 32 distinct banks
    __shared__ __align__(16) double L[DNB][LS];
    __shared__ double invd[DNB];
    __shared__ __align__(16) double Xs[2][32][XS];
    const int th = blockIdx.y;
    const int ib = kb + 1 + blockIdx.x;
    double* At = A + (size_t)th * ld * ld;
    const double* Lk = At + ((size_t)kb * DNB) * ld + (size_t)kb * DNB;
    double* Ai = At + ((size_t)ib * DNB) * ld + (size_t)kb * DNB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gr = lane >> 2, gc = lane & 3;
    double* rows = Ai + (size_t)(32 * warp) * ld;
    double x[4][8][2];                               // the slice's loads are in flight while L_kk is staged
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int g = 0; g < 8; g++) {
            const double2 v = *reinterpret_cast<const double2*>(rows + (size_t)(8 * u + gr) * ld + 8 * g + 2 * gc);
            x[u][g][0] = v.x; x[u][g][1] = v.y;
        }
    const double dg = Lk[(size_t)threadIdx.x * ld + threadIdx.x];
#pragma unroll 16
    for (int e = threadIdx.x; e < DNB * DNB / 2; e += DNB) {      // 16-byte loads, 16 in flight per thread
        const int r = e / (DNB / 2), q = 2 * (e - r * (DNB / 2));
        *reinterpret_cast<double2*>(&L[r][q]) = *reinterpret_cast<const double2*>(Lk + (size_t)r * ld + q);
    }
    invd[threadIdx.x] = 1.0 / dg;
    __syncthreads();
    double (*xs)[XS] = Xs[warp];
#pragma unroll
    for (int g = 0; g < 8; g++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {                // x_j is final once the earlier columns are eliminated
            const int src = (lane & ~3) | (j >> 1);
            const double idj = invd[8 * g + j];
            const double l0 = L[8 * g + 2 * gc][8 * g + j], l1 = L[8 * g + 2 * gc + 1][8 * g + j];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const double xj = __shfl_sync(0xffffffffu, x[u][g][j & 1] * idj, src);
                if (gc == (j >> 1)) x[u][g][j & 1] = xj;
                if (2 * gc > j) x[u][g][0] = fma(-xj, l0, x[u][g][0]);
                if (2 * gc + 1 > j) x[u][g][1] = fma(-xj, l1, x[u][g][1]);
            }
        }
        if (g < 7) {
#pragma unroll
            for (int u = 0; u < 4; u++)
                *reinterpret_cast<double2*>(&xs[8 * u + gr][2 * gc]) = make_double2(-x[u][g][0], -x[u][g][1]);
            __syncwarp();
            double af[4][2];
#pragma unroll
            for (int u = 0; u < 4; u++) { af[u][0] = xs[8 * u + gr][gc]; af[u][1] = xs[8 * u + gr][4 + gc]; }
#pragma unroll
            for (int g2 = g + 1; g2 < 8; g2++)
#pragma unroll
                for (int ks = 0; ks < 2; ks++) {
                    const double bf = L[8 * g2 + gr][8 * g + 4 * ks + gc];
#pragma unroll
                    for (int u = 0; u < 4; u++) dmma_8x8x4(x[u][g2][0], x[u][g2][1], af[u][ks], bf);
                }
            __syncwarp();
        }
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int g = 0; g < 8; g++)
            *reinterpret_cast<double2*>(rows + (size_t)(8 * u + gr) * ld + 8 * g + 2 * gc) = make_double2(x[u][g][0], x[u][g][1]);
}

// Trailing update: C_ij −= A_ik A_jkᵀ for kb < j ≤ i — the one dense contraction of the path, on the FP64 tensor cores
// (DMMA m8n8k4).  grid = (m(m+1)/2, B) with m = nblk − j0 (or (m, B) when narrow), block = 256 = 8 warps; warp (wy, wx) owns the 32×16 sub-tile
// rows 32·wy…, columns 16·wx… as 4×2 DMMA tiles.  The two 64×32 panel halves sit in shared memory row-major with a row
// stride of 36 doubles: the 8 rows × 4 k of a fragment load then fall on 32 distinct banks per half-warp.  Per k-step of 4
// a warp issues 6 shared loads for 8 DMMAs (2 048 FMAs) — the scalar version needed 4 128-bit loads per 512 FMAs and was
// bound by the shared-memory pipe (profiles/r01_launches_k3_k4.csv: 20.5 of K4's 36 ms).
// Several panels per trailing update: the panels kb … kb+kw−1 (64·kw contiguous columns) are applied at once to the tiles
// (i ≥ j ≥ j0), so a trailing tile is read and written once per GROUP of panels; inside a group the next block column alone gets
// the group's panels so far (narrow = 1: tiles (i, j0), i ≥ j0) so that it can be factorised.
__global__ void __launch_bounds__(256, 2) dense_syrk_kernel(double* __restrict__ A, int64_t ld, int kb, int kw, int j0, int narrow) {
    constexpr int KH = DNB / 2, LDSM = KH + 4;
    __shared__ __align__(16) double As[DNB][LDSM];   // [row][k], one half of a panel's width at a time
    __shared__ __align__(16) double Bs[DNB][LDSM];
    const int th = blockIdx.y;
    int pi, pj;
    if (narrow) { pi = blockIdx.x; pj = 0; }
    else tri_index(blockIdx.x, pi, pj);
    const int ib = j0 + pi, jb = j0 + pj;
    double* At = A + (size_t)th * ld * ld;
    const double* Ai = At + ((size_t)ib * DNB) * ld + (size_t)kb * DNB;
    const double* Aj = At + ((size_t)jb * DNB) * ld + (size_t)kb * DNB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy = warp >> 2, wx = warp & 3, gr = lane >> 2, gc = lane & 3;
    double acc[4][2][2];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 2; v++) acc[u][v][0] = acc[u][v][1] = 0.0;
    // software pipeline: the next half-panel travels global → registers while the tensor cores work on the current one, and
    // the output tile is fetched up front (the update is a read-modify-write)
    double* C = At + ((size_t)ib * DNB) * ld + (size_t)jb * DNB;
    double2 cv[4][2];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 2; v++)
            cv[u][v] = *reinterpret_cast<const double2*>(C + (size_t)(32 * wy + 8 * u + gr) * ld + 16 * wx + 8 * v + 2 * gc);
    constexpr int PER = DNB * KH / 256;      // 8 elements of each panel block per thread and half
    double pa[PER], pb[PER];
    const int nhalf = 2 * kw;
#pragma unroll
    for (int q = 0; q < PER; q++) {
        const int e = threadIdx.x + q * 256, r = e / KH, k = e - r * KH;        // coalesced along k
        pa[q] = Ai[(size_t)r * ld + k];
        pb[q] = Aj[(size_t)r * ld + k];
    }
    for (int half = 0; half < nhalf; half++) {
        if (half) __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int e = threadIdx.x + q * 256, r = e / KH, k = e - r * KH;
            As[r][k] = pa[q];
            Bs[r][k] = pb[q];
        }
        __syncthreads();
        if (half + 1 < nhalf) {
#pragma unroll
            for (int q = 0; q < PER; q++) {
                const int e = threadIdx.x + q * 256, r = e / KH, k = e - r * KH;
                pa[q] = Ai[(size_t)r * ld + (half + 1) * KH + k];
                pb[q] = Aj[(size_t)r * ld + (half + 1) * KH + k];
            }
        }
#pragma unroll
        for (int k4 = 0; k4 < KH / 4; k4++) {
            double a[4], b[2];
#pragma unroll
            for (int u = 0; u < 4; u++) a[u] = As[32 * wy + 8 * u + gr][4 * k4 + gc];
#pragma unroll
            for (int v = 0; v < 2; v++) b[v] = Bs[16 * wx + 8 * v + gr][4 * k4 + gc];
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int v = 0; v < 2; v++) dmma_8x8x4(acc[u][v][0], acc[u][v][1], a[u], b[v]);
        }
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 2; v++) {
            const int r = 32 * wy + 8 * u + gr, q = 16 * wx + 8 * v + 2 * gc;
            double* dst = C + (size_t)r * ld + q;
            if (ib != jb || q + 1 <= r) {
                *reinterpret_cast<double2*>(dst) = make_double2(cv[u][v].x - acc[u][v][0], cv[u][v].y - acc[u][v][1]);
            } else if (q <= r) {
                dst[0] = cv[u][v].x - acc[u][v][0];
            }
        }
}

// cp.async variant of the trailing update (round 2): the panel halves travel global → shared memory without passing through
// registers (16-byte cp.async.cg, two stages), and the output tile is staged into the free stage during the last half instead of
// being held in registers, so the kernel fits 80 registers and THREE CTAs share an SM — one CTA's prologue / barriers / epilogue
// are covered by the other two.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gmem) : "memory");
}
constexpr size_t DENSE_SYRK_ASYNC_SMEM = sizeof(double) * 4 * DNB * (DNB / 2 + 4);
__global__ void __launch_bounds__(256, 3) dense_syrk_async_kernel(double* __restrict__ A, int64_t ld, int kb, int kw, int j0, int narrow) {
    constexpr int KH = DNB / 2, LDSM = KH + 4;
    extern __shared__ __align__(16) double syrk_sm[];     // As[2][64][36] | Bs[2][64][36]
    double (*As)[DNB][LDSM] = reinterpret_cast<double (*)[DNB][LDSM]>(syrk_sm);
    double (*Bs)[DNB][LDSM] = reinterpret_cast<double (*)[DNB][LDSM]>(syrk_sm + 2 * DNB * LDSM);
    const int th = blockIdx.y;
    int pi, pj;
    if (narrow) { pi = blockIdx.x; pj = 0; }
    else tri_index(blockIdx.x, pi, pj);
    const int ib = j0 + pi, jb = j0 + pj;
    double* At = A + (size_t)th * ld * ld;
    const double* Ai = At + ((size_t)ib * DNB) * ld + (size_t)kb * DNB;
    const double* Aj = At + ((size_t)jb * DNB) * ld + (size_t)kb * DNB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy = warp >> 2, wx = warp & 3, gr = lane >> 2, gc = lane & 3;
    const int nhalf = 2 * kw;
    auto issue = [&](int half) {      // 64 rows × 32 doubles of both panels: 4 + 4 chunks of 16 bytes per thread
        const int st = half & 1;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int e = threadIdx.x + q * 256, r = e >> 4, k = 2 * (e & 15);
            cp_async16(&As[st][r][k], Ai + (size_t)r * ld + half * KH + k);
            cp_async16(&Bs[st][r][k], Aj + (size_t)r * ld + half * KH + k);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0);
    // the output tile is needed at the end only (no registers to hold it meanwhile): ask for its lines now, so that the copy into
    // shared memory during the last half finds them in the L2 (ncu before the staging: 29 % of the stall samples sat on the
    // epilogue's first DADD, waiting on the tile; the prefetch alone moved the trailing updates from 6.92 to 6.89 ms, the staging
    // to 6.63)
    double* C = At + ((size_t)ib * DNB) * ld + (size_t)jb * DNB;
    if (gc == 0) {
#pragma unroll
        for (int u = 0; u < 4; u++)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(C + (size_t)(32 * wy + 8 * u + gr) * ld + 16 * wx));
    }
    double acc[4][2][2];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 2; v++) acc[u][v][0] = acc[u][v][1] = 0.0;
    const bool dead = ib == jb && 16 * wx > 32 * wy + 31;     // a warp wholly above the diagonal of a diagonal tile does no math
    constexpr int CS = DNB + 2;       // row stride (doubles) of the output tile staged in the free stage during the last half
    for (int half = 0; half < nhalf; half++) {
        const int cur = half & 1;
        const bool last = half + 1 == nhalf;
        if (!last) {
            issue(half + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (last) {
            // the other stage is free now: the output tile travels into it (rows 0–31 into the A part, 32–63 into the B part) while
            // the tensor cores work on the last half, so the epilogue's read-modify-write does not wait on global memory
            double* c0 = &As[cur ^ 1][0][0];
            double* c1 = &Bs[cur ^ 1][0][0];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int e = threadIdx.x + q * 256, r = e >> 5, k = 2 * (e & 31);
                cp_async16((r < 32 ? c0 + r * CS : c1 + (r - 32) * CS) + k, C + (size_t)r * ld + k);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        if (!dead) {
#pragma unroll
            for (int k4 = 0; k4 < KH / 4; k4++) {
                double a[4], b[2];
#pragma unroll
                for (int u = 0; u < 4; u++) a[u] = As[cur][32 * wy + 8 * u + gr][4 * k4 + gc];
#pragma unroll
                for (int v = 0; v < 2; v++) b[v] = Bs[cur][16 * wx + 8 * v + gr][4 * k4 + gc];
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int v = 0; v < 2; v++) dmma_8x8x4(acc[u][v][0], acc[u][v][1], a[u], b[v]);
            }
        }
        if (last) asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();              // the stage is refilled by the issue of the iteration after next / the staged output tile is complete
    }
    if (dead) return;
    const int fst = ((nhalf - 1) & 1) ^ 1;                                   // the stage that holds the output tile
    const double* cs = wy == 0 ? &As[fst][0][0] : &Bs[fst][0][0];            // this warp's 32 rows of it
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 2; v++) {
            const int r = 32 * wy + 8 * u + gr, q = 16 * wx + 8 * v + 2 * gc;
            double* dst = C + (size_t)r * ld + q;
            const double2 cvv = *reinterpret_cast<const double2*>(cs + (8 * u + gr) * CS + q);
            if (ib != jb || q + 1 <= r) {
                *reinterpret_cast<double2*>(dst) = make_double2(cvv.x - acc[u][v][0], cvv.y - acc[u][v][1]);
            } else if (q <= r) {
                dst[0] = cvv.x - acc[u][v][0];
            }
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
