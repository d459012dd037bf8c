// scan.cuh — K3: parallel-in-time celerite log-likelihood for very long single series (N ~ 1e6).
//
// The reference recursion (src/celerite_solver.jl:12-158) is sequential in n.  Written on the state (S, g) entering
// step n — S the R×R matrix of :69-90, g the forward-substitution vector of :132-142 — one step is
//     D = A_n − UᵀSU,  w = (V − SU)/D,  z = y_n − Uᵀg,   S ← Φ(S + D wwᵀ)Φ,  g ← Φ(g + w z)        (Φ = diag φ_{n+1})
// which is a linear-fractional map of the same family as the Kalman-filter elements of Särkkä & García-Fernández
// (2021): with  𝒜 = Φ(I − VUᵀ/A_n), b = ΦV y/A_n, C = ΦVVᵀΦ/A_n, η = −U y/A_n, J = −UUᵀ/A_n  the step is
//     S ← 𝒜 (I + S J)⁻¹ S 𝒜ᵀ + C,      g ← 𝒜 (I + S J)⁻¹ (g + S η) + b,
// and such maps compose associatively (combine below).  Three passes over P chunks of the time axis:
//   1. scan_fold_kernel   — one CTA per chunk folds its steps into one composite (𝒜, b, C, η, J).  Folding a single
//      step into a composite is O(R²): (C, b) follow the celerite recursion itself started from zero, and
//      𝒜 ← Φ(𝒜 − w (𝒜ᵀu)ᵀ),  J ← J − (𝒜ᵀu)(𝒜ᵀu)ᵀ/D̂,  η ← η − (𝒜ᵀu) ẑ/D̂   (rank-1 updates).
//      The three 64×64 matrices live in REGISTERS (4×4 tile of each per thread, 256 threads); shared memory only
//      carries the per-step vectors.
//   2. scan_ks_kernel / scan_gather_kernel / scan_group_states_kernel / scan_states_kernel / scan_substates_kernel — two-level
//      scan over the P composites, each level a Kogge–Stone sweep of independent generic O(R³) combines (64×64 LU with partial
//      pivoting + matrix products in shared memory), then one apply per chunk and per inner boundary: the exact state entering
//      every quarter chunk.
//   3. the generic K2 kernel (celerite.cuh) re-sweeps every chunk from its incoming state, one warp per chunk, and
//      returns (Σ log|D_n|, Σ z_n²/D_n); the host adds them up.
// Algorithmic HBM traffic: 2 passes over (t, y, σ²) = 48 N bytes plus P·(3·64² + 2·64)·8 B of composites written and
// read a few times (≈ 0.1 GB at P = 296) — the path is FP64/latency-bound, not HBM-bound.
// Accuracy: rounding-level agreement with the sequential recursion on well-conditioned covariances (median 2–5e-15 over prior
// draws); on ill-conditioned ones (steep DRWCelerite slopes) the composites lose digits (1e-8 … 1e-6 seen), so pass 3 carries
// a self-check — every warp sweeps 8 steps beyond its sub-chunk and the next warp sums its first 8 steps separately; the two
// must agree (scan_check_estimate) — and the host re-evaluates what fails (api.cu: scan_logl_locked; tests/test_gpu_scan.py).
#pragma once
#include "common.cuh"

namespace pioran {

constexpr int SR = 64;                        // padded rank of the scan path (R ≤ 64)
constexpr int SEL = 3 * SR * SR + 2 * SR;     // doubles per composite: 𝒜 | C | J | b | η
constexpr int SSTATE = SR * SR + SR;          // doubles per state: S | g
constexpr int SB = 8;                         // steps whose U, V, φ vectors are prepared at once in pass 1

struct ScanArgs {
    const double* t; const double* y; const double* s2;
    int64_t N;
    int P;                      // chunks per parameter vector
    const int64_t* bounds;      // [P + 1] chunk boundaries
    const double* a; const double* b; const double* c; const double* d;  // [B × Jt]
    int Jt;
    const int* term_row;        // as in the generic K2 kernel
    const double* mu; const double* nu;   // [B] or nullptr
    double* elems;              // [B × P × SEL] chunk composites
    int SUB;                    // sub-chunks per chunk (≥ 1): the fold also stores its running composite at the SUB − 1 inner
    double* subel;              // boundaries, [B × P × (SUB−1) × SEL] — the prefix composites of the chunk's own steps
};
// Inner boundary j (0 … SUB) of the chunk [n0, n1): a multiple of SB steps from n0 (the fold prepares SB steps at a time).
__host__ __device__ inline int64_t scan_sub_bound(int64_t n0, int64_t n1, int j, int SUB) {
    if (j >= SUB) return n1;
    return n0 + ((n1 - n0) * j / SUB) / 8 * 8;
}

// Rows / columns of the 4×4 register tile of fold thread (ty, tx): two adjacent pairs 32 apart — {2t, 2t+1, 32+2t, 33+2t}.
// The 128-bit shared-memory accesses of a half-warp are then contiguous (16 lanes × 16 B); with four adjacent entries per
// thread they were 32 B apart and 4-way bank-conflicted, which made the fold bound by the shared-memory pipe.
__device__ __forceinline__ int fold_idx(int t, int e) { return (e < 2) ? 2 * t + e : 32 + 2 * t + (e - 2); }

// ------------------------------------------------------------------------------------------------ pass 1
// grid = (P, B), block = 256.  Thread (ty, tx) = (tid >> 4, tid & 15) owns rows fold_idx(ty, ·) × columns fold_idx(tx, ·).
__global__ void __launch_bounds__(256, 2) scan_fold_kernel(const ScanArgs args) {
    __shared__ __align__(16) double Us[SB][SR], Vs[SB][SR], Ps[SB][SR];   // U_n, V_n, φ_{n+1}
    __shared__ double An_s[SB], yn_s[SB];
    __shared__ __align__(16) double part[16][SR];     // partial column sums of 𝒜ᵀu per thread row
    __shared__ __align__(16) double cu_s[SR], w_s[SR], dw_s[SR], atu_s[SR], atus_s[SR], b_s[SR], eta_s[SR];
    __shared__ double red_s[4];

    const int th = blockIdx.y, ch = blockIdx.x;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, lane = tid & 31;
    const int Jt = args.Jt;
    const int64_t n0 = args.bounds[ch], n1 = args.bounds[ch + 1], N = args.N;
    const double* ca = args.a + (size_t)th * Jt;
    const double* cb = args.b + (size_t)th * Jt;
    const double* cc = args.c + (size_t)th * Jt;
    const double* cd = args.d + (size_t)th * Jt;
    const double mu = args.mu ? args.mu[th] : 0.0, nu = args.nu ? args.nu[th] : 1.0;

    double suma = 0.0;
    for (int m = 0; m < Jt; m++) suma += ca[m];       // celerite_solver.jl:21

    double A[4][4], C[4][4], Jm[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) { A[r][c] = (ty == tx && r == c) ? 1.0 : 0.0; C[r][c] = 0.0; Jm[r][c] = 0.0; }
    for (int k = tid; k < SB * SR; k += 256) { (&Us[0][0])[k] = 0.0; (&Vs[0][0])[k] = 0.0; (&Ps[0][0])[k] = 0.0; }
    if (tid < SR) { b_s[tid] = 0.0; eta_s[tid] = 0.0; }
    __syncthreads();

    int next_sub = 1;
    int64_t next_bound = scan_sub_bound(n0, n1, 1, args.SUB);
    for (int64_t nb = n0; nb < n1; nb += SB) {
        const int ns = (int)((n1 - nb) < SB ? (n1 - nb) : SB);
        // ---- running composite at an inner boundary: the prefix of the chunk's steps [n0, nb)
        while (next_sub < args.SUB && nb >= next_bound) {
            double* E = args.subel + (((size_t)th * args.P + ch) * (args.SUB - 1) + (next_sub - 1)) * SEL;
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const size_t k = (size_t)fold_idx(ty, r) * SR + fold_idx(tx, c);
                    E[k] = A[r][c]; E[SR * SR + k] = C[r][c]; E[2 * SR * SR + k] = Jm[r][c];
                }
            if (tid < SR) { E[3 * SR * SR + tid] = b_s[tid]; E[3 * SR * SR + SR + tid] = eta_s[tid]; }
            next_sub++;
            next_bound = scan_sub_bound(n0, n1, next_sub, args.SUB);
        }
        // ---- U, V, φ of the next ns steps (celerite_solver.jl:51-64), one thread per (step, term)
        for (int idx = tid; idx < ns * Jt; idx += 256) {
            const int s = idx / Jt, m = idx - s * Jt;
            const int64_t n = nb + s;
            const double tn = args.t[n];
            const double ph = (n + 1 < N) ? exp(-cc[m] * (args.t[n + 1] - tn)) : 0.0;
            const int tr = args.term_row[m];
            if (tr < 0) {
                const int r0 = -tr - 1;
                Us[s][r0] = ca[m]; Vs[s][r0] = 1.0; Ps[s][r0] = ph;
            } else {
                double si, co;
                sincos_large(cd[m] * tn, &si, &co);
                Us[s][tr] = ca[m] * co + cb[m] * si;  Us[s][tr + 1] = ca[m] * si - cb[m] * co;
                Vs[s][tr] = co;                       Vs[s][tr + 1] = si;
                Ps[s][tr] = ph;                       Ps[s][tr + 1] = ph;
            }
        }
        if (tid < ns) { An_s[tid] = fma(nu, args.s2[nb + tid], suma); yn_s[tid] = args.y[nb + tid] - mu; }
        __syncthreads();

        for (int s = 0; s < ns; s++) {
            // ---- phase A: partial products  C u (rows),  𝒜ᵀ u (columns),  uᵀ b
            double ucol[4], urow[4];
            {
                const double2 c01 = *reinterpret_cast<const double2*>(&Us[s][2 * tx]), c23 = *reinterpret_cast<const double2*>(&Us[s][32 + 2 * tx]);
                const double2 r01 = *reinterpret_cast<const double2*>(&Us[s][2 * ty]), r23 = *reinterpret_cast<const double2*>(&Us[s][32 + 2 * ty]);
                ucol[0] = c01.x; ucol[1] = c01.y; ucol[2] = c23.x; ucol[3] = c23.y;
                urow[0] = r01.x; urow[1] = r01.y; urow[2] = r23.x; urow[3] = r23.y;
            }
            double cup[4], atp[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < 4; c++) acc = fma(C[r][c], ucol[c], acc);
                cup[r] = acc;
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < 4; r++) acc = fma(A[r][c], urow[r], acc);
                atp[c] = acc;
            }
            // rows: the 16 threads of one ty are 16 consecutive lanes → reduce-scatter (4 values → 2 → 1) + 2 rounds
            {
                const bool h8 = (tx & 8) != 0, h4 = (tx & 4) != 0;
                const double r0 = __shfl_xor_sync(0xffffffffu, h8 ? cup[0] : cup[2], 8);
                const double r1 = __shfl_xor_sync(0xffffffffu, h8 ? cup[1] : cup[3], 8);
                const double e0 = (h8 ? cup[2] : cup[0]) + r0;     // h8 = 0 keeps rows 0,1; h8 = 1 keeps rows 2,3
                const double e1 = (h8 ? cup[3] : cup[1]) + r1;
                const double r2 = __shfl_xor_sync(0xffffffffu, h4 ? e0 : e1, 4);
                double f = (h4 ? e1 : e0) + r2;                    // row 2·h8 + h4
                f += __shfl_xor_sync(0xffffffffu, f, 2);
                f += __shfl_xor_sync(0xffffffffu, f, 1);
                if ((tx & 3) == 0) cu_s[fold_idx(ty, 2 * (h8 ? 1 : 0) + (h4 ? 1 : 0))] = f;
            }
            *reinterpret_cast<double2*>(&part[ty][2 * tx]) = make_double2(atp[0], atp[1]);
            *reinterpret_cast<double2*>(&part[ty][32 + 2 * tx]) = make_double2(atp[2], atp[3]);
            __syncthreads();

            // ---- phase B (warps 0-1): D̂, ẑ, w, 𝒜ᵀu;  b, η updates
            if (tid < SR) {
                const int j = tid;
                double at0 = 0.0, at1 = 0.0, at2 = 0.0, at3 = 0.0;
#pragma unroll
                for (int q = 0; q < 16; q += 4) { at0 += part[q][j]; at1 += part[q + 1][j]; at2 += part[q + 2][j]; at3 += part[q + 3][j]; }
                const double at = (at0 + at1) + (at2 + at3);
                const double u = Us[s][j], v = Vs[s][j], cu = cu_s[j], bj = b_s[j];
                double s1 = u * cu, s2v = u * bj;
#pragma unroll
                for (int sft = 16; sft >= 1; sft >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, sft);
                    s2v += __shfl_xor_sync(0xffffffffu, s2v, sft);
                }
                if (lane == 0) { red_s[2 * (tid >> 5)] = s1; red_s[2 * (tid >> 5) + 1] = s2v; }
                // the two warps exchange their halves through shared memory
                asm volatile("bar.sync 1, 64;");
                const double den = An_s[s] - (red_s[0] + red_s[2]);       // celerite_solver.jl:92 on the chunk-local C
                const double z = yn_s[s] - (red_s[1] + red_s[3]);         // celerite_solver.jl:141 on the chunk-local b
                const double rden = 1.0 / den;
                const double w = (v - cu) * rden;
                const double ph = Ps[s][j];
                const double ats = at * rden;
                w_s[j] = w; dw_s[j] = v - cu; atu_s[j] = at; atus_s[j] = ats;
                b_s[j] = ph * fma(w, z, bj);
                eta_s[j] = fma(-ats, z, eta_s[j]);
            }
            __syncthreads();

            // ---- phase C: rank-1 updates and decay of the three tiles
            double phr[4], phc[4], wr[4], wc[4], dwr[4], atc[4], atsr[4];
            {
                const double2 p0 = *reinterpret_cast<const double2*>(&Ps[s][2 * ty]), p1 = *reinterpret_cast<const double2*>(&Ps[s][32 + 2 * ty]);
                const double2 p2 = *reinterpret_cast<const double2*>(&Ps[s][2 * tx]), p3 = *reinterpret_cast<const double2*>(&Ps[s][32 + 2 * tx]);
                const double2 w0 = *reinterpret_cast<const double2*>(&w_s[2 * ty]), w1 = *reinterpret_cast<const double2*>(&w_s[32 + 2 * ty]);
                const double2 w2 = *reinterpret_cast<const double2*>(&w_s[2 * tx]), w3 = *reinterpret_cast<const double2*>(&w_s[32 + 2 * tx]);
                const double2 d0 = *reinterpret_cast<const double2*>(&dw_s[2 * ty]), d1 = *reinterpret_cast<const double2*>(&dw_s[32 + 2 * ty]);
                const double2 a0 = *reinterpret_cast<const double2*>(&atu_s[2 * tx]), a1 = *reinterpret_cast<const double2*>(&atu_s[32 + 2 * tx]);
                const double2 s0 = *reinterpret_cast<const double2*>(&atus_s[2 * ty]), s1 = *reinterpret_cast<const double2*>(&atus_s[32 + 2 * ty]);
                phr[0] = p0.x; phr[1] = p0.y; phr[2] = p1.x; phr[3] = p1.y;
                phc[0] = p2.x; phc[1] = p2.y; phc[2] = p3.x; phc[3] = p3.y;
                wr[0] = w0.x; wr[1] = w0.y; wr[2] = w1.x; wr[3] = w1.y;
                wc[0] = w2.x; wc[1] = w2.y; wc[2] = w3.x; wc[3] = w3.y;
                dwr[0] = d0.x; dwr[1] = d0.y; dwr[2] = d1.x; dwr[3] = d1.y;
                atc[0] = a0.x; atc[1] = a0.y; atc[2] = a1.x; atc[3] = a1.y;
                atsr[0] = s0.x; atsr[1] = s0.y; atsr[2] = s1.x; atsr[3] = s1.y;
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    C[r][c] = (phr[r] * phc[c]) * fma(dwr[r], wc[c], C[r][c]);   // celerite_solver.jl:76,85
                    A[r][c] = phr[r] * fma(-wr[r], atc[c], A[r][c]);
                    Jm[r][c] = fma(-atsr[r], atc[c], Jm[r][c]);
                }
            // the next phase A reads only registers and Us; phase B of the next step is behind its own barrier
        }
        __syncthreads();   // Us/Vs/Ps are rebuilt for the next SB steps
    }

    // ---- composite → global
    double* E = args.elems + ((size_t)th * args.P + ch) * SEL;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const size_t k = (size_t)fold_idx(ty, r) * SR + fold_idx(tx, c);
            E[k] = A[r][c]; E[SR * SR + k] = C[r][c]; E[2 * SR * SR + k] = Jm[r][c];
        }
    if (tid < SR) { E[3 * SR * SR + tid] = b_s[tid]; E[3 * SR * SR + SR + tid] = eta_s[tid]; }
}

// ------------------------------------------------------------------------------------------------ 64×64 helpers
// All matrices are SR×SR row-major; shared-memory copies use a padded leading dimension.
constexpr int SLD = SR + 2;
constexpr int SMAT = SR * SLD;

__device__ __forceinline__ void sm_load(double* dst, const double* src) {   // global → shared
    for (int k = threadIdx.x; k < SR * SR; k += blockDim.x) dst[(k / SR) * SLD + (k % SR)] = src[k];
}
__device__ __forceinline__ void sm_store(double* __restrict__ dst, const double* src) {  // shared → global
    for (int k = threadIdx.x; k < SR * SR; k += blockDim.x) dst[k] = src[(k / SR) * SLD + (k % SR)];
}
// D(8×8) += A(8×4)·B(4×8) on the FP64 tensor pipe (mma.sync.m8n8k4.f64).  Fragments: A lane(g,t) = A[g][t]; B lane(g,t) = B[t][g];
// C/D lane(g,t) = [g][2t], [g][2t+1]  (g = lane >> 2, t = lane & 3).
__device__ __forceinline__ void scan_dmma(double& c0, double& c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// D = op(X)·op(Y) [+ addend (global, row-major)] → dst_s (shared) and/or dst_g (global).  256 threads = 8 warps; warp w forms the
// row tile w of D (8 rows × 64 columns) with DMMAs, operands read from shared memory as fragments (round 2; round 1 ran scalar FMAs
// on 4×4 register tiles: ≈ 15 µs per product against ≈ 1.5 µs).  The contraction stops at the live rank Rr (a multiple of 4).
// dst_s must not alias X or Y.
template <bool TX, bool TY>
__device__ __forceinline__ void sm_matmul(double* dst_s, double* dst_g, const double* X, const double* Y,
                                       const double* addend, bool symmetrise, int Rr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double acc[8][2];
#pragma unroll
    for (int K = 0; K < 8; K++) acc[K][0] = acc[K][1] = 0.0;
    const int arow = 8 * warp + g;
    for (int k0 = 0; k0 < Rr; k0 += 4) {
        const double a = TX ? X[(k0 + t) * SLD + arow] : X[arow * SLD + k0 + t];
#pragma unroll
        for (int K = 0; K < 8; K++) {
            const double b = TY ? Y[(8 * K + g) * SLD + k0 + t] : Y[(k0 + t) * SLD + 8 * K + g];
            scan_dmma(acc[K][0], acc[K][1], a, b);
        }
    }
    if (symmetrise) {
        // (D + Dᵀ)/2 through shared memory: needs dst_s
#pragma unroll
        for (int K = 0; K < 8; K++) { dst_s[arow * SLD + 8 * K + 2 * t] = acc[K][0]; dst_s[arow * SLD + 8 * K + 2 * t + 1] = acc[K][1]; }
        __syncthreads();
#pragma unroll
        for (int K = 0; K < 8; K++) {
            acc[K][0] = 0.5 * (acc[K][0] + dst_s[(8 * K + 2 * t) * SLD + arow]);
            acc[K][1] = 0.5 * (acc[K][1] + dst_s[(8 * K + 2 * t + 1) * SLD + arow]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int K = 0; K < 8; K++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int col = 8 * K + 2 * t + e;
            double v = acc[K][e];
            if (addend) v += addend[arow * SR + col];
            if (dst_s) dst_s[arow * SLD + col] = v;
            if (dst_g) dst_g[arow * SR + col] = v;
        }
    __syncthreads();
}
// y = op(X)·x (+ y0) for a shared matrix and shared vectors; threads 0..63 each one row.  No trailing barrier.
template <bool TX>
__device__ __forceinline__ double sm_matvec_row(const double* X, const double* x, int row, int Rr) {
    double acc = 0.0;
    for (int k = 0; k < Rr; k++) acc = fma(TX ? X[k * SLD + row] : X[row * SLD + k], x[k], acc);
    return acc;
}
// X ← M⁻¹ X for up to two shared SR×SR right-hand sides (R1, R2; R2 may be null) and nvec shared vectors (vecs + v·SR), by
// Gauss–Jordan elimination with partial pivoting on the augmented system [M | R1 | R2 | vecs]; M is destroyed.  One pass of
// Rr pivot steps, two CTA barriers each (pivot search by warp 0; rank-1 elimination of column k from every other row with a
// 16×16 thread grid over each 64-wide matrix) — one routine instead of an LU factorisation plus two substitution sweeps per
// right-hand side (the combine is bound by its barriers, not by this arithmetic).  Rows/columns ≥ Rr are the identity.
__device__ __forceinline__ void sm_gj_solve(double* M, double* R1, double* R2, double* vecs, int nvec, int Rr) {
    // Round 2: TWO pivots per pair of barriers.  No physical row interchange — the pivot row of column k is remembered and the
    // rows are put in order once at the end; no separate factor column (columns k, k+1 of M are never written again, so every
    // thread forms its factors from them in place); no per-pivot scaling (the rows are divided by their pivots in the final pass).
    // Warp 0 picks the pivot row p1 of column k (largest |M[r][k]| among the unused rows, as partial pivoting would), forms
    // column k+1 as it stands after that elimination, m'_r = M[r][k+1] − f1_r M[p1][k+1], and picks p2 from it; then everybody
    // applies both eliminations at once:  row_r ← row_r − f1_r row_p1 − f2_r row'_p2  with  row'_p2 = row_p2 − f1_p2 row_p1.
    __shared__ int piv_s[2];
    __shared__ int used_s[SR], prow_s[SR];
    __shared__ double pinv_s[SR];
    __shared__ double rowbuf_s[2][3 * SR];       // the two pivot rows (the second after the first elimination), snapshot by warp 0
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    if (tid < SR) used_s[tid] = 0;
    __syncthreads();
    for (int k = 0; k < Rr; k += 2) {            // Rr is a multiple of 4
        if (tid < 32) {
            // ---- pivot of column k
            double best = -1.0; int bi = -1;
            for (int r = tid; r < Rr; r += 32) {
                const double v = used_s[r] ? -1.0 : fabs(M[r * SLD + k]);
                if (v > best || (v == best && bi < 0)) { best = v; bi = r; }
            }
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, sft);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, sft);
                if (ob > best || (ob == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ob; bi = oi; }
            }
            if (bi < 0) {                         // a column of NaNs: any unused row keeps the bookkeeping valid, the result is NaN
                for (int r = 0; r < Rr; r++) if (!used_s[r]) { bi = r; break; }
            }
            const int p1 = bi;
            const double inv1 = 1.0 / M[p1 * SLD + k];
            const double mp = M[p1 * SLD + k + 1];
            // ---- column k+1 after that elimination, and its pivot among the rows still unused
            best = -1.0; bi = -1;
            for (int r = tid; r < Rr; r += 32) {
                double v = -1.0;
                if (!used_s[r] && r != p1) v = fabs(fma(-(M[r * SLD + k] * inv1), mp, M[r * SLD + k + 1]));
                if (v > best || (v == best && bi < 0 && v >= 0.0)) { best = v; bi = r; }
            }
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, sft);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, sft);
                if (ob > best || (ob == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ob; bi = oi; }
            }
            if (bi < 0) {
                for (int r = 0; r < Rr; r++) if (!used_s[r] && r != p1) { bi = r; break; }
            }
            const int p2 = bi;
            const double f1p2 = M[p2 * SLD + k] * inv1;
            // snapshot of the pivot rows over the augmented width: the elimination overwrites them while others still need them
            for (int q = tid; q < 3 * SR; q += 32) {
                const int which = q / SR, c = q - which * SR;
                const double* X = which == 0 ? M : (which == 1 ? R1 : R2);
                double a1 = 0.0, a2 = 0.0;
                if (X) { a1 = X[p1 * SLD + c]; a2 = fma(-f1p2, a1, X[p2 * SLD + c]); }
                rowbuf_s[0][q] = a1; rowbuf_s[1][q] = a2;
            }
            if (tid == 0) {
                const double m2 = fma(-f1p2, mp, M[p2 * SLD + k + 1]);
                piv_s[0] = p1; piv_s[1] = p2;
                used_s[p1] = 1; used_s[p2] = 1;
                prow_s[k] = p1; prow_s[k + 1] = p2;
                pinv_s[k] = inv1; pinv_s[k + 1] = 1.0 / m2;
            }
        }
        __syncthreads();
        const int p1 = piv_s[0], p2 = piv_s[1];
        const double inv1 = pinv_s[k], inv2 = pinv_s[k + 1];
        const double mp = M[p1 * SLD + k + 1];
        // factors of my rows: f1_r (0 for p1), f2_r from the updated column k+1 (0 for p2)
        double f1[4], f2[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int r = ty + 16 * q;
            f1[q] = 0.0; f2[q] = 0.0;
            if (r < Rr) {
                if (r != p1) f1[q] = M[r * SLD + k] * inv1;
                if (r != p2) f2[q] = fma(-f1[q], mp, M[r * SLD + k + 1]) * inv2;
            }
        }
        // my 4×4 entries of each matrix, fully unrolled (compile-time offsets: the loop overhead of the generic version was most of
        // the combine's 40 k instructions per warp)
        bool rok[4], isp2[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const int r = ty + 16 * q; rok[q] = r < Rr; isp2[q] = r == p2; }
#pragma unroll
        for (int which = 0; which < 3; which++) {
            double* X = which == 0 ? M : (which == 1 ? R1 : R2);
            if (!X) continue;
            double* Xt = X + ty * SLD + tx;
            const double* rb0 = rowbuf_s[0] + which * SR + tx;
            const double* rb1 = rowbuf_s[1] + which * SR + tx;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (which == 0 && tx + 16 * j < k + 2) continue;      // columns ≤ k+1 of M are spent
                const double a1 = rb0[16 * j];
                const double a2 = rb1[16 * j];                         // row p2 after the first elimination
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (rok[q]) {
                        double* e = Xt + 16 * q * SLD + 16 * j;
                        const double v = fma(-f2[q], a2, fma(-f1[q], a1, *e));      // r = p1: f1 = 0
                        *e = isp2[q] ? a2 : v;
                    }
                }
            }
        }
        if (tid < nvec * 32) {      // one warp per vector
            double* v = vecs + (size_t)(tid >> 5) * SR;
            const double a1 = v[p1];
            const double a2 = fma(-(M[p2 * SLD + k] * inv1), a1, v[p2]);
            __syncwarp();
            for (int r = (tid & 31); r < Rr; r += 32) {
                const double g1 = (r != p1) ? M[r * SLD + k] * inv1 : 0.0;
                const double g2 = (r != p2) ? fma(-g1, mp, M[r * SLD + k + 1]) * inv2 : 0.0;
                v[r] = (r == p2) ? a2 : fma(-g2, a2, fma(-g1, a1, v[r]));
            }
        }
        __syncthreads();
    }
    // rows into pivot order, divided by their pivots: X[k] ← X[prow[k]] / pivot_k (through the spent M as scratch)
    for (int which = 1; which < 3; which++) {
        double* X = which == 1 ? R1 : R2;
        if (!X) continue;
        for (int q = tid; q < Rr * SR; q += blockDim.x) { const int k = q / SR, c = q % SR; M[k * SLD + c] = X[prow_s[k] * SLD + c] * pinv_s[k]; }
        __syncthreads();
        for (int q = tid; q < Rr * SR; q += blockDim.x) { const int k = q / SR, c = q % SR; X[k * SLD + c] = M[k * SLD + c]; }
        __syncthreads();
    }
    for (int v = 0; v < nvec; v++) {
        double* x = vecs + (size_t)v * SR;
        double val = 0.0;
        if (tid < Rr) val = x[prow_s[tid]] * pinv_s[tid];
        __syncthreads();
        if (tid < Rr) x[tid] = val;
    }
    __syncthreads();
}

// Shared-memory workspace of the pass-2 kernels: 5 matrices + 8 vectors + permutation.
// Rr: live rank rounded up to a multiple of 4 (≤ SR).  Rows and columns ≥ Rr of every composite and state are exactly zero
// (the fold never touches them), so (I + C J) is the identity there: the sequential parts of combine / apply — pivot steps,
// substitution rows, the inner index of the products — stop at Rr, which makes a rank-8 combine ~8× shorter than a rank-64 one.
struct ScanSmem {
    double* m[5];
    double* v[8];
    int* perm;      // (unused since the Gauss–Jordan solve; kept for the workspace layout)
    int Rr;
};
__device__ __forceinline__ ScanSmem scan_smem(unsigned char* raw, int Rr) {
    ScanSmem w;
    w.Rr = Rr;
    double* p = reinterpret_cast<double*>(raw);
    for (int k = 0; k < 5; k++) { w.m[k] = p; p += SMAT; }
    for (int k = 0; k < 8; k++) { w.v[k] = p; p += SR; }
    w.perm = reinterpret_cast<int*>(p);
    return w;
}
constexpr size_t SCAN_SMEM_BYTES = sizeof(double) * (5 * (size_t)SMAT + 8 * SR) + sizeof(int) * SR;

// out = ei ⊗ ej (ei earlier).  Composite layout: 𝒜 | C | J | b | η.  out may alias neither input.
//   M = (I + C_i J_j)⁻¹;  𝒜 = 𝒜_j M 𝒜_i;  b = 𝒜_j M (b_i + C_i η_j) + b_j;  C = 𝒜_j M C_i 𝒜_jᵀ + C_j;
//   η = 𝒜_iᵀ (r − J_j M C_i r) + η_i,  r = η_j − J_j b_i;   J = 𝒜_iᵀ J_j M 𝒜_i + J_i
__device__ __forceinline__ void scan_combine(const ScanSmem& w, const double* ei, const double* ej, double* out) {
    const int tid = threadIdx.x;
    const double *Ai = ei, *Ci = ei + SR * SR, *Ji = ei + 2 * SR * SR, *bi = ei + 3 * SR * SR, *eti = bi + SR;
    const double *Aj = ej, *Cj = ej + SR * SR, *Jj = ej + 2 * SR * SR, *bj = ej + 3 * SR * SR, *etj = bj + SR;
    double *Ao = out, *Co = out + SR * SR, *Jo = out + 2 * SR * SR, *bo = out + 3 * SR * SR, *eto = bo + SR;
    sm_load(w.m[0], Ci);
    sm_load(w.m[1], Jj);
    if (tid < SR) { w.v[0][tid] = bi[tid]; w.v[1][tid] = etj[tid]; }
    __syncthreads();
    sm_matmul<false, false>(w.m[2], nullptr, w.m[0], w.m[1], nullptr, false, w.Rr);        // C_i J_j
    if (tid < SR) {
        w.m[2][tid * SLD + tid] += 1.0;
        w.v[2][tid] = w.v[1][tid] - sm_matvec_row<false>(w.m[1], w.v[0], tid, w.Rr);       // r = η_j − J_j b_i
        w.v[3][tid] = w.v[0][tid] + sm_matvec_row<false>(w.m[0], w.v[1], tid, w.Rr);       // b_i + C_i η_j
    }
    __syncthreads();
    if (tid < SR) w.v[4][tid] = sm_matvec_row<false>(w.m[0], w.v[2], tid, w.Rr);           // C_i r
    __syncthreads();
    sm_load(w.m[3], Ai);
    __syncthreads();
    sm_gj_solve(w.m[2], w.m[0], w.m[3], w.v[3], 2, w.Rr);   // m0 = M C_i;  m3 = M 𝒜_i;  v3 = M (b_i + C_i η_j);  v4 = M C_i r
    sm_load(w.m[4], Aj);
    __syncthreads();
    sm_matmul<false, false>(nullptr, Ao, w.m[4], w.m[3], nullptr, false, w.Rr);            // 𝒜 = 𝒜_j (M 𝒜_i)
    sm_matmul<false, false>(w.m[2], nullptr, w.m[4], w.m[0], nullptr, false, w.Rr);        // 𝒜_j (M C_i)
    if (tid < SR) bo[tid] = sm_matvec_row<false>(w.m[4], w.v[3], tid, w.Rr) + bj[tid];     // b
    if (tid < SR) w.v[5][tid] = w.v[2][tid] - sm_matvec_row<false>(w.m[1], w.v[4], tid, w.Rr);   // r − J_j M C_i r
    __syncthreads();
    sm_matmul<false, true>(w.m[0], Co, w.m[2], w.m[4], Cj, true, w.Rr);                    // C = (…) 𝒜_jᵀ + C_j
    sm_matmul<false, false>(w.m[2], nullptr, w.m[1], w.m[3], nullptr, false, w.Rr);        // J_j (M 𝒜_i)
    sm_load(w.m[4], Ai);
    __syncthreads();
    if (tid < SR) eto[tid] = sm_matvec_row<true>(w.m[4], w.v[5], tid, w.Rr) + eti[tid];    // η
    sm_matmul<true, false>(w.m[0], Jo, w.m[4], w.m[2], Ji, true, w.Rr);                    // J = 𝒜_iᵀ (…) + J_i
}

// (S', g') = el applied to (S, g):  S' = 𝒜 (I + S J)⁻¹ S 𝒜ᵀ + C,  g' = 𝒜 (I + S J)⁻¹ (g + S η) + b.
// in == nullptr means the zero state.  out may alias in.
__device__ __forceinline__ void scan_apply(const ScanSmem& w, const double* el, const double* in, double* out) {
    const int tid = threadIdx.x;
    const double *Ae = el, *Ce = el + SR * SR, *Je = el + 2 * SR * SR, *be = el + 3 * SR * SR, *ete = be + SR;
    if (!in) {
        for (int k = tid; k < SR * SR; k += blockDim.x) out[k] = Ce[k];
        if (tid < SR) out[SR * SR + tid] = be[tid];
        __syncthreads();
        return;
    }
    sm_load(w.m[0], in);
    sm_load(w.m[1], Je);
    if (tid < SR) { w.v[0][tid] = in[SR * SR + tid]; w.v[1][tid] = ete[tid]; }
    __syncthreads();
    sm_matmul<false, false>(w.m[2], nullptr, w.m[0], w.m[1], nullptr, false, w.Rr);        // S J
    if (tid < SR) {
        w.m[2][tid * SLD + tid] += 1.0;
        w.v[3][tid] = w.v[0][tid] + sm_matvec_row<false>(w.m[0], w.v[1], tid, w.Rr);       // g + S η
    }
    __syncthreads();
    sm_gj_solve(w.m[2], w.m[0], nullptr, w.v[3], 1, w.Rr);   // m0 = (I + S J)⁻¹ S;  v3 = (I + S J)⁻¹ (g + S η)
    sm_load(w.m[4], Ae);
    __syncthreads();
    sm_matmul<false, false>(w.m[2], nullptr, w.m[4], w.m[0], nullptr, false, w.Rr);        // 𝒜 W
    if (tid < SR) w.v[5][tid] = sm_matvec_row<false>(w.m[4], w.v[3], tid, w.Rr) + be[tid];
    __syncthreads();
    sm_matmul<false, true>(w.m[0], out, w.m[2], w.m[4], Ce, true, w.Rr);                   // S' = (𝒜 W) 𝒜ᵀ + C
    if (tid < SR) out[SR * SR + tid] = w.v[5][tid];
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ pass 2
// P chunks are split into G1 groups of G2 consecutive chunks (the last group may be shorter).  Both levels of the scan are
// Kogge–Stone sweeps — ⌈log2⌉ kernel launches of independent combines instead of a chain of G2 (then G1) dependent ones:
// (a) scan_ks_kernel over segments of G2 chunks: after ⌈log2 G2⌉ levels element i holds the product of its group's elements
//     up to i (the prefix composites inside the groups);
// (b) scan_gather_kernel copies the group totals, scan_ks_kernel (one segment of G1) turns them into their running products;
//     the last one is the composite of the whole range (what the multi-GPU path exchanges);
// (c) scan_group_states_kernel: state entering group g = running product g−1 applied to the incoming state (one apply each).
// One level: out[i] = in[i−d] ⊗ in[i] if i is at least d into its segment, else in[i].  grid = (n, B).
__global__ void __launch_bounds__(256, 1) scan_ks_kernel(const double* in, double* out, int n, int seg, int d, int Rr) {
    extern __shared__ __align__(16) unsigned char raw[];
    const ScanSmem w = scan_smem(raw, Rr);
    const int th = blockIdx.y, i = blockIdx.x;
    const size_t base = (size_t)th * n;
    const double* src = in + (base + i) * SEL;
    double* dst = out + (base + i) * SEL;
    if (i % seg >= d) {
        scan_combine(w, in + (base + i - d) * SEL, src, dst);
    } else {
        for (int k = threadIdx.x; k < SEL; k += blockDim.x) dst[k] = src[k];
    }
}
// tot[g] = pref[last chunk of group g].  grid = (G1, B).
__global__ void scan_gather_kernel(const double* __restrict__ pref, double* __restrict__ tot, int P, int G2, int G1) {
    const int th = blockIdx.y, g = blockIdx.x;
    const int last = min(P, (g + 1) * G2) - 1;
    const double* src = pref + ((size_t)th * P + last) * SEL;
    double* dst = tot + ((size_t)th * G1 + g) * SEL;
    for (int k = threadIdx.x; k < SEL; k += blockDim.x) dst[k] = src[k];
}
// gstate[0] = incoming state (init, or zero at the start of the series); gstate[g] = tp[g−1] applied to it.  grid = (G1, B).
//     init (nullable, [B × SSTATE]): the state entering the first chunk when the chunks cover only the tail of a series
//     (time axis split across GPUs).
__global__ void __launch_bounds__(256, 1) scan_group_states_kernel(const double* tp, double* gstate, int G1, const double* init,
                                                                   int Rr) {
    extern __shared__ __align__(16) unsigned char raw[];
    const ScanSmem w = scan_smem(raw, Rr);
    const int th = blockIdx.y, g = blockIdx.x;
    double* S = gstate + ((size_t)th * G1 + g) * SSTATE;
    const double* in = init ? init + (size_t)th * SSTATE : nullptr;
    if (g == 0) {
        for (int k = threadIdx.x; k < SSTATE; k += blockDim.x) S[k] = in ? in[k] : 0.0;
        return;
    }
    scan_apply(w, tp + ((size_t)th * G1 + g - 1) * SEL, in, S);
}
// (c) grid = (P, B): state entering chunk ch = pref[g][ch−1−g·G2] applied to the group state (or the group state itself).
__global__ void __launch_bounds__(256, 1) scan_states_kernel(const double* pref, const double* gstate, double* cstate,
                                                             int P, int G2, int G1, int has_init, int Rr) {
    extern __shared__ __align__(16) unsigned char raw[];
    const ScanSmem w = scan_smem(raw, Rr);
    const int th = blockIdx.y, ch = blockIdx.x;
    const int g = ch / G2;
    const double* Q = pref + (size_t)th * P * SEL;
    const double* Sg = gstate + ((size_t)th * G1 + g) * SSTATE;
    double* out = cstate + ((size_t)th * P + ch) * SSTATE;
    if (ch == g * G2) {
        for (int k = threadIdx.x; k < SSTATE; k += blockDim.x) out[k] = Sg[k];
        return;
    }
    scan_apply(w, Q + (size_t)(ch - 1) * SEL, (g == 0 && !has_init) ? nullptr : Sg, out);
}

// (c') grid = (P·(SUB−1), B): state entering inner boundary j of chunk ch = the chunk's own prefix composite (stored by the
//      fold) applied to the state entering the chunk.  Four times as many, four times shorter re-filter sweeps in pass 3.
__global__ void __launch_bounds__(256, 1) scan_substates_kernel(const double* subel, const double* cstate, double* substate,
                                                                int P, int SUB, int has_init, int Rr) {
    extern __shared__ __align__(16) unsigned char raw[];
    const ScanSmem w = scan_smem(raw, Rr);
    const int th = blockIdx.y, ch = blockIdx.x / (SUB - 1), j = blockIdx.x % (SUB - 1);
    const size_t q = (size_t)th * P + ch;
    const double* el = subel + (q * (SUB - 1) + j) * SEL;
    double* out = substate + (q * (SUB - 1) + j) * SSTATE;
    scan_apply(w, el, (ch == 0 && !has_init) ? nullptr : cstate + q * SSTATE, out);
}

// ------------------------------------------------------------------------------------------------ Newton refinement
// The composites lose digits in their combines on ill-conditioned covariances, so the chunk states S̃_k the scan hands to
// pass 3 can be ~1e-8 off — and the filter forgets such an error only algebraically on slowly decaying terms, so run-ups do
// not remove it.  One Newton step on the boundary states does, with the EXACT recursion as the residual: pass 3, run chunk by
// chunk, also returns the state it leaves behind, E_k = F_k(S̃_k) (WorkItem::exit); the correction δ_k of S̃_k obeys
//     δS_{k+1} = T_k δS_k T_kᵀ + (E_k − S̃_{k+1}),      δg_{k+1} = T_k (δg_k + δS_k m_k) + (Eg_k − g̃_{k+1}),
// T_k = 𝒜_k (I + S̃_k J_k)⁻¹ the closed-loop transition of chunk k and m_k = η_k − J_k ĝ_k, ĝ_k = (I + S̃_k J_k)⁻¹(g̃_k + S̃_k η_k)
// (the derivative of the linear-fractional map of scan_apply).  T_k and m_k come from the composites and are needed to first
// order only: their own 1e-8 error enters at second order.  Derivation and a numpy check: tests/tools/proto/scan_newton_math.py.
constexpr int SNEWT = SR * SR + SR;           // doubles per chunk: T | m
// grid = (P, B).  With X = (I + S J)⁻¹ S (the solve scan_apply does):  (I + S J)⁻¹ = I − X J,  so  T = 𝒜 − (𝒜 X) J.
__global__ void __launch_bounds__(256, 1) scan_newton_T_kernel(const double* elems, const double* cstate, double* tm, int P,
                                                               int Rr) {
    extern __shared__ __align__(16) unsigned char raw[];
    const ScanSmem w = scan_smem(raw, Rr);
    const int th = blockIdx.y, ch = blockIdx.x, tid = threadIdx.x;
    const size_t q = (size_t)th * P + ch;
    const double* el = elems + q * SEL;
    const double *Ae = el, *Je = el + 2 * SR * SR, *ete = el + 3 * SR * SR + SR;
    double* To = tm + q * SNEWT;
    double* mo = To + SR * SR;
    if (ch == 0) {      // zero state: T = 𝒜, m = η (δ_0 = 0, so neither is used)
        for (int k = tid; k < SR * SR; k += blockDim.x) To[k] = Ae[k];
        if (tid < SR) mo[tid] = ete[tid];
        return;
    }
    const double* in = cstate + q * SSTATE;
    sm_load(w.m[0], in);
    sm_load(w.m[1], Je);
    if (tid < SR) { w.v[0][tid] = in[SR * SR + tid]; w.v[1][tid] = ete[tid]; }
    __syncthreads();
    sm_matmul<false, false>(w.m[2], nullptr, w.m[0], w.m[1], nullptr, false, w.Rr);        // S J
    if (tid < SR) {
        w.m[2][tid * SLD + tid] += 1.0;
        w.v[3][tid] = w.v[0][tid] + sm_matvec_row<false>(w.m[0], w.v[1], tid, w.Rr);       // g + S η
    }
    __syncthreads();
    sm_gj_solve(w.m[2], w.m[0], nullptr, w.v[3], 1, w.Rr);   // m0 = X;  v3 = ĝ
    sm_load(w.m[4], Ae);
    __syncthreads();
    sm_matmul<false, false>(w.m[2], nullptr, w.m[4], w.m[0], nullptr, false, w.Rr);        // 𝒜 X
    sm_matmul<false, false>(w.m[3], nullptr, w.m[2], w.m[1], nullptr, false, w.Rr);        // (𝒜 X) J
    for (int k = tid; k < SR * SR; k += blockDim.x) To[k] = Ae[k] - w.m[3][(k / SR) * SLD + (k % SR)];
    if (tid < SR) mo[tid] = w.v[1][tid] - sm_matvec_row<false>(w.m[1], w.v[3], tid, w.Rr);
}
// The recurrence is a scan over affine maps: boundary k carries e_k = (T_{k−1}, r_k, m_{k−1}, r^g_k), and two consecutive maps
// (i earlier, j later) compose into  T = T_j T_i,  r = T_j r_i T_jᵀ + r_j,  m = m_i + T_iᵀ m_j,  r^g = T_j (r^g_i + r_i m_j) + r^g_j.
// A Kogge–Stone sweep over k = 1 … P−1 (⌈log2⌉ launches of independent compositions — three 64³ products each, no solve) leaves
// at boundary k the composition of the maps 1 … k; applied to δ_0 = 0 its (r, r^g) ARE the corrections (δS_k, δg_k).  A chain
// of P−1 dependent steps in one CTA took 8.3 ms at P = 296; the sweep takes 9 levels of ≈ 60 µs.
constexpr int SNEL = 2 * SR * SR + 2 * SR;    // doubles per affine element: T | r | m | r^g
// grid = (P, B): element of boundary k from (T | m) of chunk k−1, the exit state of chunk k−1 and the current state of chunk k.
__global__ void scan_newton_prep_kernel(const double* __restrict__ tm, const double* __restrict__ exits,
                                        const double* __restrict__ cstate, double* __restrict__ nel, int P) {
    const int th = blockIdx.y, k = blockIdx.x;
    if (k == 0) return;
    const size_t q = (size_t)th * P + k;
    const double* T = tm + (q - 1) * SNEWT;
    const double* E = exits + (q - 1) * SSTATE;
    const double* S = cstate + q * SSTATE;
    double* el = nel + q * SNEL;
    for (int i = threadIdx.x; i < SR * SR; i += blockDim.x) { el[i] = T[i]; el[SR * SR + i] = E[i] - S[i]; }
    for (int i = threadIdx.x; i < SR; i += blockDim.x) {
        el[2 * SR * SR + i] = T[SR * SR + i];
        el[2 * SR * SR + SR + i] = E[SR * SR + i] - S[SR * SR + i];
    }
}
// One Kogge–Stone level: out[k] = in[k−d] ∘ in[k] when k − d ≥ 1, else in[k].  grid = (P, B).
__global__ void __launch_bounds__(256, 1) scan_newton_ks_kernel(const double* in, double* out, int P, int d, int Rr) {
    extern __shared__ __align__(16) unsigned char raw[];
    const ScanSmem w = scan_smem(raw, Rr);
    const int th = blockIdx.y, k = blockIdx.x, tid = threadIdx.x;
    if (k == 0) return;
    const size_t q = (size_t)th * P + k;
    const double* ej = in + q * SNEL;
    double* eo = out + q * SNEL;
    if (k - d < 1) {
        for (int i = tid; i < SNEL; i += blockDim.x) eo[i] = ej[i];
        return;
    }
    const double* ei = in + (q - d) * SNEL;
    constexpr int MM = SR * SR;
    sm_load(w.m[0], ej);                 // T_j
    sm_load(w.m[1], ei);                 // T_i
    sm_load(w.m[2], ei + MM);            // r_i
    if (tid < SR) { w.v[0][tid] = ej[2 * MM + tid]; w.v[1][tid] = ei[2 * MM + SR + tid]; }       // m_j, r^g_i
    __syncthreads();
    if (tid < SR) {
        eo[2 * MM + tid] = ei[2 * MM + tid] + sm_matvec_row<true>(w.m[1], w.v[0], tid, w.Rr);     // m = m_i + T_iᵀ m_j
        w.v[3][tid] = w.v[1][tid] + sm_matvec_row<false>(w.m[2], w.v[0], tid, w.Rr);              // r^g_i + r_i m_j
    }
    __syncthreads();
    if (tid < SR) eo[2 * MM + SR + tid] = sm_matvec_row<false>(w.m[0], w.v[3], tid, w.Rr) + ej[2 * MM + SR + tid];   // r^g
    sm_matmul<false, false>(nullptr, eo, w.m[0], w.m[1], nullptr, false, w.Rr);                   // T = T_j T_i
    sm_matmul<false, false>(w.m[3], nullptr, w.m[0], w.m[2], nullptr, false, w.Rr);               // T_j r_i
    sm_matmul<false, true>(w.m[4], eo + MM, w.m[3], w.m[0], ej + MM, true, w.Rr);                 // r = (T_j r_i) T_jᵀ + r_j
}
// S̃_k ← S̃_k + δ_k (k ≥ 1).  grid = (P, B).
__global__ void scan_newton_apply_kernel(const double* __restrict__ nel, double* __restrict__ cstate, int P) {
    const int th = blockIdx.y, k = blockIdx.x;
    if (k == 0) return;
    const size_t q = (size_t)th * P + k;
    const double* el = nel + q * SNEL;
    double* S = cstate + q * SSTATE;
    for (int i = threadIdx.x; i < SR * SR; i += blockDim.x) S[i] += el[SR * SR + i];
    for (int i = threadIdx.x; i < SR; i += blockDim.x) S[SR * SR + i] += el[2 * SR * SR + SR + i];
}

// (e) state entering this range = the composites of the nprev earlier ranges applied, in order, to the zero state.
//     elems_prev is [nprev × B × SEL] (rank-major, as gathered); grid = (1, B).
__global__ void __launch_bounds__(256, 1) scan_chain_kernel(const double* elems_prev, int nprev, int B, double* state,
                                                            int Rr) {
    extern __shared__ __align__(16) unsigned char raw[];
    const ScanSmem w = scan_smem(raw, Rr);
    const int th = blockIdx.y;
    double* S = state + (size_t)th * SSTATE;
    for (int r = 0; r < nprev; r++) {
        __threadfence_block();
        scan_apply(w, elems_prev + ((size_t)r * B + th) * SEL, r == 0 ? nullptr : S, S);
        __syncthreads();
    }
}
// Σ over chunks of the pass-3 partial sums of a range → (Σ log|D_n|, Σ z_n²/D_n), one thread per parameter vector.
// Deviation estimate of one parameter vector from the self-check sums (WorkItem::chk): at every sub-chunk boundary the
// steps right after it were swept twice — by the previous warp, continuing from its own state, and by the next warp from the
// state the scan handed it.  The two contributions to log L differ by the effect of the scan's state error on those steps;
// `scale` (sub-chunk length / check length, ≥ 1) extends it to the whole sub-chunk — an upper bound while the effect of a state
// perturbation does not grow along the sweep.  NaN when any of the sums is not finite.
__device__ inline double scan_check_estimate(const double* __restrict__ chk, int P, double scale) {
    double est = 0.0;
    for (int k = 0; k + 1 < P; k++) {
        const double* a = chk + 4 * (size_t)k;
        est += 0.5 * (fabs(a[2] - a[4]) + fabs(a[3] - a[5]));
    }
    return est * scale;
}

__global__ void scan_partial_kernel(const double* __restrict__ parts, const double* __restrict__ chk, double scale, int P, int B,
                                    double* __restrict__ sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double ld = 0.0, chi = 0.0;
    for (int k = 0; k < P; k++) { ld += parts[2 * ((size_t)i * P + k)]; chi += parts[2 * ((size_t)i * P + k) + 1]; }
    sums[3 * i] = ld; sums[3 * i + 1] = chi; sums[3 * i + 2] = scan_check_estimate(chk + 4 * (size_t)i * P, P, scale);
}

// Σ over chunks of the pass-3 partial sums → logL (celerite_solver.jl:333).  One thread per parameter vector.
__global__ void scan_finish_kernel(const double* __restrict__ parts, const double* __restrict__ chk, double scale, int P, int B,
                                   int64_t N, double* __restrict__ out, double* __restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double ld = 0.0, chi = 0.0;
    for (int k = 0; k < P; k++) { ld += parts[2 * ((size_t)i * P + k)]; chi += parts[2 * ((size_t)i * P + k) + 1]; }
    out[i] = -ld / 2 - (double)N * 1.8378770664093453 / 2 - chi / 2;
    err[i] = scan_check_estimate(chk + 4 * (size_t)i * P, P, scale);
}

}  // namespace pioran
