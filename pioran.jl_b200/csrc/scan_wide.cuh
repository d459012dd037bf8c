// scan_wide.cuh — K3 at ranks 65 … 96: the parallel-in-time path of scan.cuh for the upper part of the reference's benchmark
// grid (DRWCelerite J = 30: rank 90, SHO J = 40: rank 80 — benchmark/benchmarks.jl:16-18).
//
// Same three passes and the same algebra as scan.cuh (composites (𝒜, b, C, η, J) of src/celerite_solver.jl:12-158 written as
// linear-fractional maps, Kogge–Stone levels of generic combines, re-filter of every sub-chunk from its incoming state, Newton
// refinement of the chunk states), re-laid for matrices that no longer fit the rank-64 shapes:
//   * fold: 512 threads, 𝒜 and C as 6×3 register tiles (rows fold pairs 32 apart, columns tx + 32 j), J — which the fold only
//     accumulates, J = −Σ (𝒜ᵀu)(𝒜ᵀu)ᵀ/D̂ — in shared memory, brought up to date once per 8 steps as a rank-8 update;
//   * combine / apply / Newton kernels: THREE shared-memory matrices of the live rank (Rr × (Rr | 1) doubles each: 205 KB at
//     rank 92) instead of five, operands reloaded from L2 when needed, the output's J slot used as scratch inside the combine;
//   * pass 3: the register-file CTA kernel of wide.cuh in chunked form (celerite_wide_chunk_kernel: injected state, the three
//     self-check segments, exit state for the Newton step).
// Composite layout: 𝒜 | C | J | b | η with leading dimension SRW = 96; states S | g likewise.
#pragma once
#include "scan.cuh"
#include "wide.cuh"

namespace pioran {

constexpr int SRW = 96;
constexpr int SELW = 3 * SRW * SRW + 2 * SRW;
constexpr int SSTATEW = SRW * SRW + SRW;
constexpr int SNEWTW = SRW * SRW + SRW;
constexpr int FOLDW_THREADS = 512;

// rows of fold thread row-group ty (0 … 15): three adjacent pairs 32 apart
__device__ __forceinline__ int foldw_row(int ty, int e) { return 32 * (e >> 1) + 2 * ty + (e & 1); }

// ------------------------------------------------------------------------------------------------ pass 1
// grid = (P, B), block = 512.  Thread (ty, tx) = (tid >> 5, tid & 31) owns rows foldw_row(ty, ·) × columns {tx, tx + 32, tx + 64}.
// dynamic smem: J [96 × 96] | Us, Vs, Ps [8 × 96 each] | part [16 × 96] | atu history [8 × 96], atus history [8 × 96] | vectors
constexpr size_t FOLDW_SMEM_BYTES = sizeof(double) * ((size_t)SRW * SRW + 3 * SB * SRW + 16 * SRW + 2 * SB * SRW + 8 * SRW + 32);
__global__ void __launch_bounds__(FOLDW_THREADS, 1) scanw_fold_kernel(const ScanArgs args) {
    extern __shared__ __align__(16) unsigned char raw[];
    double* Js = reinterpret_cast<double*>(raw);                // [SRW][SRW]
    double* Us = Js + SRW * SRW;                                // [SB][SRW]
    double* Vs = Us + SB * SRW;
    double* Ps = Vs + SB * SRW;
    double* part = Ps + SB * SRW;                               // [16][SRW]
    double* atu_h = part + 16 * SRW;                            // [SB][SRW]  𝒜ᵀu of the steps since the last J update
    double* atus_h = atu_h + SB * SRW;                          // [SB][SRW]  the same divided by D̂
    double* cu_s = atus_h + SB * SRW;
    double* w_s = cu_s + SRW;
    double* dw_s = w_s + SRW;
    double* b_s = dw_s + SRW;
    double* eta_s = b_s + SRW;
    double* An_s = eta_s + SRW;                                 // [SB] (+ yn_s [SB])
    double* yn_s = An_s + SB;
    double* red_s = yn_s + SB;                                  // [6]

    const int th = blockIdx.y, ch = blockIdx.x;
    const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31, lane = tx;
    const int Jt = args.Jt;
    const int64_t n0 = args.bounds[ch], n1 = args.bounds[ch + 1], N = args.N;
    const double* ca = args.a + (size_t)th * Jt;
    const double* cb = args.b + (size_t)th * Jt;
    const double* cc = args.c + (size_t)th * Jt;
    const double* cd = args.d + (size_t)th * Jt;
    const double mu = args.mu ? args.mu[th] : 0.0, nu = args.nu ? args.nu[th] : 1.0;

    double suma = 0.0;
    for (int m = 0; m < Jt; m++) suma += ca[m];       // celerite_solver.jl:21

    double A[6][3], C[6][3];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) { A[r][c] = (foldw_row(ty, r) == tx + 32 * c) ? 1.0 : 0.0; C[r][c] = 0.0; }
    for (int k = tid; k < SRW * SRW; k += FOLDW_THREADS) Js[k] = 0.0;
    for (int k = tid; k < 3 * SB * SRW; k += FOLDW_THREADS) Us[k] = 0.0;
    if (tid < SRW) { b_s[tid] = 0.0; eta_s[tid] = 0.0; }
    __syncthreads();

    // J tile of this thread ← J − Σ_s atus_s[row] atu_s[col] over the ns buffered steps (shared memory, rank-ns update)
    auto flush_J = [&](int ns) {
#pragma unroll
        for (int r = 0; r < 6; r++) {
            const int row = foldw_row(ty, r);
            double acc[3] = {Js[row * SRW + tx], Js[row * SRW + tx + 32], Js[row * SRW + tx + 64]};
            for (int s = 0; s < ns; s++) {
                const double ar = atus_h[s * SRW + row];
#pragma unroll
                for (int c = 0; c < 3; c++) acc[c] = fma(-ar, atu_h[s * SRW + tx + 32 * c], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < 3; c++) Js[row * SRW + tx + 32 * c] = acc[c];
        }
    };
    auto store_composite = [&](double* E) {
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const size_t k = (size_t)foldw_row(ty, r) * SRW + tx + 32 * c;
                E[k] = A[r][c]; E[SRW * SRW + k] = C[r][c]; E[2 * SRW * SRW + k] = Js[k];
            }
        if (tid < SRW) { E[3 * SRW * SRW + tid] = b_s[tid]; E[3 * SRW * SRW + SRW + tid] = eta_s[tid]; }
    };

    int next_sub = 1;
    int64_t next_bound = scan_sub_bound(n0, n1, 1, args.SUB);
    for (int64_t nb = n0; nb < n1; nb += SB) {
        const int ns = (int)((n1 - nb) < SB ? (n1 - nb) : SB);
        // ---- running composite at an inner boundary (J is up to date: flushed at the end of every group of SB steps)
        while (next_sub < args.SUB && nb >= next_bound) {
            store_composite(args.subel + (((size_t)th * args.P + ch) * (args.SUB - 1) + (next_sub - 1)) * SELW);
            next_sub++;
            next_bound = scan_sub_bound(n0, n1, next_sub, args.SUB);
        }
        // ---- U, V, φ of the next ns steps (celerite_solver.jl:51-64)
        for (int idx = tid; idx < ns * Jt; idx += FOLDW_THREADS) {
            const int s = idx / Jt, m = idx - s * Jt;
            const int64_t n = nb + s;
            const double tn = args.t[n];
            const double ph = (n + 1 < N) ? exp(-cc[m] * (args.t[n + 1] - tn)) : 0.0;
            const int tr = args.term_row[m];
            if (tr < 0) {
                const int r0 = -tr - 1;
                Us[s * SRW + r0] = ca[m]; Vs[s * SRW + r0] = 1.0; Ps[s * SRW + r0] = ph;
            } else {
                double si, co;
                sincos_large(cd[m] * tn, &si, &co);
                Us[s * SRW + tr] = ca[m] * co + cb[m] * si;  Us[s * SRW + tr + 1] = ca[m] * si - cb[m] * co;
                Vs[s * SRW + tr] = co;                       Vs[s * SRW + tr + 1] = si;
                Ps[s * SRW + tr] = ph;                       Ps[s * SRW + tr + 1] = ph;
            }
        }
        if (tid < ns) { An_s[tid] = fma(nu, args.s2[nb + tid], suma); yn_s[tid] = args.y[nb + tid] - mu; }
        __syncthreads();

        for (int s = 0; s < ns; s++) {
            const double* U = Us + s * SRW;
            // ---- phase A: partial products  C u (rows),  𝒜ᵀ u (columns)
            double ucol[3], urow[6];
#pragma unroll
            for (int c = 0; c < 3; c++) ucol[c] = U[tx + 32 * c];
#pragma unroll
            for (int e = 0; e < 6; e += 2) {
                const double2 v = *reinterpret_cast<const double2*>(&U[32 * (e >> 1) + 2 * ty]);
                urow[e] = v.x; urow[e + 1] = v.y;
            }
            double cup[8], atp[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int r = 0; r < 6; r++) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < 3; c++) { acc = fma(C[r][c], ucol[c], acc); atp[c] = fma(A[r][c], urow[r], atp[c]); }
                cup[r] = acc;
            }
            cup[6] = 0.0; cup[7] = 0.0;
            {   // the 32 lanes of a warp share ty: reduce-scatter of the (padded) 8 row sums, then two plain rounds
                const bool h16 = (tx & 16) != 0, h8 = (tx & 8) != 0, h4 = (tx & 4) != 0;
                double e4[4], e2[2];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double recv = __shfl_xor_sync(0xffffffffu, h16 ? cup[k] : cup[k + 4], 16);
                    e4[k] = (h16 ? cup[k + 4] : cup[k]) + recv;
                }
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const double recv = __shfl_xor_sync(0xffffffffu, h8 ? e4[k] : e4[k + 2], 8);
                    e2[k] = (h8 ? e4[k + 2] : e4[k]) + recv;
                }
                const double recv = __shfl_xor_sync(0xffffffffu, h4 ? e2[0] : e2[1], 4);
                double f = (h4 ? e2[1] : e2[0]) + recv;
                f += __shfl_xor_sync(0xffffffffu, f, 2);
                f += __shfl_xor_sync(0xffffffffu, f, 1);
                const int isel = 4 * (h16 ? 1 : 0) + 2 * (h8 ? 1 : 0) + (h4 ? 1 : 0);
                if ((tx & 3) == 0 && isel < 6) cu_s[foldw_row(ty, isel)] = f;
            }
#pragma unroll
            for (int c = 0; c < 3; c++) part[ty * SRW + tx + 32 * c] = atp[c];
            __syncthreads();

            // ---- phase B (warps 0-2): D̂, ẑ, w, 𝒜ᵀu;  b, η updates
            if (tid < SRW) {
                const int j = tid;
                double at0 = 0.0, at1 = 0.0, at2 = 0.0, at3 = 0.0;
#pragma unroll
                for (int q = 0; q < 16; q += 4) {
                    at0 += part[q * SRW + j]; at1 += part[(q + 1) * SRW + j]; at2 += part[(q + 2) * SRW + j]; at3 += part[(q + 3) * SRW + j];
                }
                const double at = (at0 + at1) + (at2 + at3);
                const double u = U[j], v = Vs[s * SRW + j], cu = cu_s[j], bj = b_s[j];
                double s1 = u * cu, s2v = u * bj;
#pragma unroll
                for (int sft = 16; sft >= 1; sft >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, sft);
                    s2v += __shfl_xor_sync(0xffffffffu, s2v, sft);
                }
                if (lane == 0) { red_s[2 * (tid >> 5)] = s1; red_s[2 * (tid >> 5) + 1] = s2v; }
                asm volatile("bar.sync 1, 96;");
                const double den = An_s[s] - ((red_s[0] + red_s[2]) + red_s[4]);     // celerite_solver.jl:92 on the chunk-local C
                const double z = yn_s[s] - ((red_s[1] + red_s[3]) + red_s[5]);       // celerite_solver.jl:141 on the chunk-local b
                const double rden = 1.0 / den;
                const double w = (v - cu) * rden;
                const double ph = Ps[s * SRW + j];
                const double ats = at * rden;
                w_s[j] = w; dw_s[j] = v - cu;
                atu_h[s * SRW + j] = at; atus_h[s * SRW + j] = ats;
                b_s[j] = ph * fma(w, z, bj);
                eta_s[j] = fma(-ats, z, eta_s[j]);
            }
            __syncthreads();

            // ---- phase C: rank-1 updates and decay of the two register tiles
            {
                const double* P = Ps + s * SRW;
                const double* atu = atu_h + s * SRW;
                double phc[3], wc[3], atc[3];
#pragma unroll
                for (int c = 0; c < 3; c++) { phc[c] = P[tx + 32 * c]; wc[c] = w_s[tx + 32 * c]; atc[c] = atu[tx + 32 * c]; }
#pragma unroll
                for (int e = 0; e < 6; e += 2) {
                    const int r0 = 32 * (e >> 1) + 2 * ty;
                    const double2 p = *reinterpret_cast<const double2*>(&P[r0]);
                    const double2 wv = *reinterpret_cast<const double2*>(&w_s[r0]);
                    const double2 dv = *reinterpret_cast<const double2*>(&dw_s[r0]);
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        C[e][c] = (p.x * phc[c]) * fma(dv.x, wc[c], C[e][c]);             // celerite_solver.jl:76,85
                        C[e + 1][c] = (p.y * phc[c]) * fma(dv.y, wc[c], C[e + 1][c]);
                        A[e][c] = p.x * fma(-wv.x, atc[c], A[e][c]);
                        A[e + 1][c] = p.y * fma(-wv.y, atc[c], A[e + 1][c]);
                    }
                }
            }
            // the next phase A reads registers and Us; phase B of the next step is behind its own barrier
        }
        flush_J(ns);       // reads the histories written in phase B (ordered by the barrier after it), writes only this thread's J tile
        __syncthreads();   // Us/Vs/Ps and the histories are rebuilt for the next SB steps
    }
    store_composite(args.elems + ((size_t)th * args.P + ch) * SELW);
}

// ------------------------------------------------------------------------------------------------ live-rank helpers
// Shared-memory matrices are Rr × LD row-major with LD = Rr | 1 (odd: column walks are conflict-free); global ones SRW-strided.
// 256 threads as a 16×16 grid, thread (ty, tx) owning the interleaved tile rows {ty + 16 i} × columns {tx + 16 j}, i, j < 6.
struct SwSmem {
    double* m[3];
    double* v[8];
    int Rr, LD;
};
__host__ __device__ inline int scanw_ld(int Rr) { return Rr | 1; }
__host__ __device__ inline size_t scanw_smem_bytes(int Rr) { return sizeof(double) * (3 * (size_t)Rr * scanw_ld(Rr) + 8 * SRW) + 64; }
__device__ __forceinline__ SwSmem scanw_smem(unsigned char* raw, int Rr) {
    SwSmem w;
    w.Rr = Rr; w.LD = scanw_ld(Rr);
    double* p = reinterpret_cast<double*>(raw);
    for (int k = 0; k < 3; k++) { w.m[k] = p; p += (size_t)Rr * w.LD; }
    for (int k = 0; k < 8; k++) { w.v[k] = p; p += SRW; }
    return w;
}
__device__ __forceinline__ void sw_load(const SwSmem& w, double* dst, const double* src) {       // global → shared (no barrier)
    for (int k = threadIdx.x; k < w.Rr * w.Rr; k += blockDim.x) { const int r = k / w.Rr, c = k - r * w.Rr; dst[r * w.LD + c] = src[r * SRW + c]; }
}
__device__ __forceinline__ void sw_store_full(const SwSmem& w, double* __restrict__ dst, const double* src) {   // shared → global, zero outside the live rank
    for (int k = threadIdx.x; k < SRW * SRW; k += blockDim.x) {
        const int r = k / SRW, c = k - r * SRW;
        dst[k] = (r < w.Rr && c < w.Rr) ? src[r * w.LD + c] : 0.0;
    }
}
// D = op(X)·op(Y) [+ addend (global, SRW-strided)] → dst_s (shared) and/or dst_g (global, SRW-strided; entries outside the live
// rank are not written).  dst_s must not alias X or Y; with symmetrise it is needed as scratch.  Ends with a barrier.
// 8 warps; warp w forms the row tiles w and w + 8 of D on the FP64 tensor pipe (fragments from shared memory, scan_dmma).
template <bool TX, bool TY>
__device__ __forceinline__ void sw_matmul(const SwSmem& w, double* dst_s, double* dst_g, const double* X, const double* Y,
                                          const double* addend, bool symmetrise) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, Rr = w.Rr, LD = w.LD;
    constexpr int NTW = SRW / 8;
    double acc[2][NTW][2];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int K = 0; K < NTW; K++) acc[i][K][0] = acc[i][K][1] = 0.0;
    const int ar0 = 8 * warp + g, ar1 = 8 * (warp + 8) + g;
    const bool r0ok = ar0 < Rr, r1ok = ar1 < Rr;
    for (int k0 = 0; k0 < Rr; k0 += 4) {
        const double a0 = r0ok ? (TX ? X[(k0 + t) * LD + ar0] : X[ar0 * LD + k0 + t]) : 0.0;
        const double a1 = r1ok ? (TX ? X[(k0 + t) * LD + ar1] : X[ar1 * LD + k0 + t]) : 0.0;
#pragma unroll
        for (int K = 0; K < NTW; K++) {
            if (8 * K < Rr) {
                const int col = 8 * K + g;
                const double b = col < Rr ? (TY ? Y[col * LD + k0 + t] : Y[(k0 + t) * LD + col]) : 0.0;
                scan_dmma(acc[0][K][0], acc[0][K][1], a0, b);
                if (8 * (warp + 8) < Rr) scan_dmma(acc[1][K][0], acc[1][K][1], a1, b);
            }
        }
    }
    if (symmetrise) {
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int K = 0; K < NTW; K++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int row = i ? ar1 : ar0, col = 8 * K + 2 * t + e;
                    if (row < Rr && col < Rr) dst_s[row * LD + col] = acc[i][K][e];
                }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int K = 0; K < NTW; K++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int row = i ? ar1 : ar0, col = 8 * K + 2 * t + e;
                    if (row < Rr && col < Rr) acc[i][K][e] = 0.5 * (acc[i][K][e] + dst_s[col * LD + row]);
                }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int K = 0; K < NTW; K++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int row = i ? ar1 : ar0, col = 8 * K + 2 * t + e;
                if (row < Rr && col < Rr) {
                    double v = acc[i][K][e];
                    if (addend) v += addend[row * SRW + col];
                    if (dst_s) dst_s[row * LD + col] = v;
                    if (dst_g) dst_g[row * SRW + col] = v;
                }
            }
    __syncthreads();
}
template <bool TX>
__device__ __forceinline__ double sw_matvec_row(const SwSmem& w, const double* X, const double* x, int row) {
    double acc = 0.0;
    if (row < w.Rr)
        for (int k = 0; k < w.Rr; k++) acc = fma(TX ? X[k * w.LD + row] : X[row * w.LD + k], x[k], acc);
    return acc;
}
// y_row = Σ_k Xg[row][k] x[k] with Xg in GLOBAL memory (SRW-strided; L2-resident), transposed when TX
template <bool TX>
__device__ __forceinline__ double sw_matvec_row_g(const SwSmem& w, const double* Xg, const double* x, int row) {
    double acc = 0.0;
    if (row < w.Rr)
        for (int k = 0; k < w.Rr; k++) acc = fma(TX ? Xg[k * SRW + row] : Xg[row * SRW + k], x[k], acc);
    return acc;
}
// X ← M⁻¹ X for up to two shared right-hand-side matrices and nvec shared vectors (vecs + v·SRW): Gauss–Jordan with partial
// pivoting on the augmented system, as sm_gj_solve of scan.cuh.  M is destroyed.
__device__ __forceinline__ void sw_gj_solve(const SwSmem& w, double* M, double* R1, double* R2, double* vecs, int nvec) {
    __shared__ int piv_s;
    __shared__ int used_s[SRW], prow_s[SRW];
    __shared__ double pinv_s[SRW];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, Rr = w.Rr, LD = w.LD;
    if (tid < SRW) used_s[tid] = 0;
    __syncthreads();
    for (int k = 0; k < Rr; k++) {
        if (tid < 32) {
            double best = -1.0; int bi = -1;
            for (int r = tid; r < Rr; r += 32) {
                const double v = used_s[r] ? -1.0 : fabs(M[r * LD + k]);
                if (v > best || (v == best && bi < 0)) { best = v; bi = r; }
            }
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, sft);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, sft);
                if (ob > best || (ob == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ob; bi = oi; }
            }
            if (tid == 0) {
                if (bi < 0)
                    for (int r = 0; r < Rr; r++) if (!used_s[r]) { bi = r; break; }
                piv_s = bi; used_s[bi] = 1; prow_s[k] = bi; pinv_s[k] = 1.0 / M[bi * LD + k];
            }
        }
        __syncthreads();
        const int p = piv_s;
        const double inv = pinv_s[k];
        double f[6];
#pragma unroll
        for (int q = 0; q < 6; q++) { const int r = ty + 16 * q; f[q] = (r < Rr && r != p) ? M[r * LD + k] * inv : 0.0; }
        // my 6×6 entries of each matrix, unrolled (the index arithmetic of the generic loops was most of the instruction count)
        bool rok[6];
#pragma unroll
        for (int q = 0; q < 6; q++) { const int r = ty + 16 * q; rok[q] = r < Rr && r != p; }
#pragma unroll
        for (int which = 0; which < 3; which++) {
            double* X = which == 0 ? M : (which == 1 ? R1 : R2);
            if (!X) continue;
            double* Xt = X + ty * LD + tx;
            const double* Xp = X + p * LD + tx;
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const int c = tx + 16 * j;
                if (c >= Rr || (which == 0 && c <= k)) continue;
                const double rk = Xp[16 * j];
#pragma unroll
                for (int q = 0; q < 6; q++)
                    if (rok[q]) { double* e = Xt + 16 * q * LD + 16 * j; *e = fma(-f[q], rk, *e); }
            }
        }
        if (tid < nvec * 32) {
            double* v = vecs + (size_t)(tid >> 5) * SRW;
            const double vk = v[p];
            for (int r = (tid & 31); r < Rr; r += 32)
                if (r != p) v[r] = fma(-(M[r * LD + k] * inv), vk, v[r]);
        }
        __syncthreads();
    }
    for (int which = 1; which < 3; which++) {
        double* X = which == 1 ? R1 : R2;
        if (!X) continue;
        for (int q = tid; q < Rr * Rr; q += blockDim.x) { const int k = q / Rr, c = q - k * Rr; M[k * LD + c] = X[prow_s[k] * LD + c] * pinv_s[k]; }
        __syncthreads();
        for (int q = tid; q < Rr * Rr; q += blockDim.x) { const int k = q / Rr, c = q - k * Rr; X[k * LD + c] = M[k * LD + c]; }
        __syncthreads();
    }
    for (int v = 0; v < nvec; v++) {
        double* x = vecs + (size_t)v * SRW;
        double val = 0.0;
        if (tid < Rr) val = x[prow_s[tid]] * pinv_s[tid];
        __syncthreads();
        if (tid < Rr) x[tid] = val;
    }
    __syncthreads();
}

// out = ei ⊗ ej (ei earlier), the algebra of scan_combine with three buffers (X, Y, Z) = w.m[0 … 2]; out's J slot is scratch
// until the last product.  out may alias neither input.
__device__ __forceinline__ void sw_combine(const SwSmem& w, const double* ei, const double* ej, double* out) {
    const int tid = threadIdx.x;
    constexpr int MM = SRW * SRW;
    const double *Ai = ei, *Ci = ei + MM, *Ji = ei + 2 * MM, *bi = ei + 3 * MM, *eti = bi + SRW;
    const double *Aj = ej, *Cj = ej + MM, *Jj = ej + 2 * MM, *bj = ej + 3 * MM, *etj = bj + SRW;
    double *Ao = out, *Co = out + MM, *Jo = out + 2 * MM, *bo = out + 3 * MM, *eto = bo + SRW;
    double *X = w.m[0], *Y = w.m[1], *Z = w.m[2];
    sw_load(w, X, Ci);
    sw_load(w, Y, Jj);
    if (tid < SRW) { w.v[0][tid] = tid < w.Rr ? bi[tid] : 0.0; w.v[1][tid] = tid < w.Rr ? etj[tid] : 0.0; }
    __syncthreads();
    sw_matmul<false, false>(w, Z, nullptr, X, Y, nullptr, false);                           // C_i J_j
    if (tid < w.Rr) {
        Z[tid * w.LD + tid] += 1.0;
        w.v[2][tid] = w.v[1][tid] - sw_matvec_row<false>(w, Y, w.v[0], tid);                // r = η_j − J_j b_i
        w.v[3][tid] = w.v[0][tid] + sw_matvec_row<false>(w, X, w.v[1], tid);                // b_i + C_i η_j
    }
    __syncthreads();
    if (tid < w.Rr) w.v[4][tid] = sw_matvec_row<false>(w, X, w.v[2], tid);                  // C_i r
    __syncthreads();
    sw_load(w, Y, Ai);                                                                      // J_j comes back from L2 later
    __syncthreads();
    sw_gj_solve(w, Z, X, Y, w.v[3], 2);          // X = M C_i;  Y = M 𝒜_i;  v3 = M (b_i + C_i η_j);  v4 = M C_i r
    sw_load(w, Z, Aj);
    __syncthreads();
    sw_matmul<false, false>(w, nullptr, Ao, Z, Y, nullptr, false);                          // 𝒜 = 𝒜_j (M 𝒜_i)
    if (tid < w.Rr) {
        bo[tid] = sw_matvec_row<false>(w, Z, w.v[3], tid) + bj[tid];                        // b
        w.v[5][tid] = w.v[2][tid] - sw_matvec_row_g<false>(w, Jj, w.v[4], tid);             // r − J_j M C_i r
    }
    for (int k = tid; k < w.Rr * w.Rr; k += blockDim.x) { const int r = k / w.Rr, c = k - r * w.Rr; Jo[r * SRW + c] = Y[r * w.LD + c]; }   // park M 𝒜_i
    __syncthreads();
    sw_matmul<false, false>(w, Y, nullptr, Z, X, nullptr, false);                           // 𝒜_j (M C_i)
    sw_matmul<false, true>(w, X, Co, Y, Z, Cj, true);                                       // C = (…) 𝒜_jᵀ + C_j
    sw_load(w, X, Jj);
    sw_load(w, Y, Jo);                                                                      // M 𝒜_i
    __syncthreads();
    sw_matmul<false, false>(w, Z, nullptr, X, Y, nullptr, false);                           // J_j (M 𝒜_i)
    sw_load(w, X, Ai);
    __syncthreads();
    if (tid < w.Rr) eto[tid] = sw_matvec_row<true>(w, X, w.v[5], tid) + eti[tid];           // η
    sw_matmul<true, false>(w, Y, Jo, X, Z, Ji, true);                                       // J = 𝒜_iᵀ (…) + J_i
}

// (S', g') = el applied to (S, g); in == nullptr: the zero state.  out may alias in.
__device__ __forceinline__ void sw_apply(const SwSmem& w, const double* el, const double* in, double* out) {
    const int tid = threadIdx.x;
    constexpr int MM = SRW * SRW;
    const double *Ae = el, *Ce = el + MM, *Je = el + 2 * MM, *be = el + 3 * MM, *ete = be + SRW;
    if (!in) {
        for (int k = tid; k < MM; k += blockDim.x) { const int r = k / SRW, c = k - r * SRW; out[k] = (r < w.Rr && c < w.Rr) ? Ce[k] : 0.0; }
        if (tid < SRW) out[MM + tid] = tid < w.Rr ? be[tid] : 0.0;
        __syncthreads();
        return;
    }
    double *X = w.m[0], *Y = w.m[1], *Z = w.m[2];
    sw_load(w, X, in);
    sw_load(w, Y, Je);
    if (tid < SRW) { w.v[0][tid] = tid < w.Rr ? in[MM + tid] : 0.0; w.v[1][tid] = tid < w.Rr ? ete[tid] : 0.0; }
    __syncthreads();
    sw_matmul<false, false>(w, Z, nullptr, X, Y, nullptr, false);                           // S J
    if (tid < w.Rr) {
        Z[tid * w.LD + tid] += 1.0;
        w.v[3][tid] = w.v[0][tid] + sw_matvec_row<false>(w, X, w.v[1], tid);                // g + S η
    }
    __syncthreads();
    sw_gj_solve(w, Z, X, nullptr, w.v[3], 1);    // X = (I + S J)⁻¹ S;  v3 = (I + S J)⁻¹ (g + S η)
    sw_load(w, Y, Ae);
    __syncthreads();
    sw_matmul<false, false>(w, Z, nullptr, Y, X, nullptr, false);                           // 𝒜 W
    if (tid < w.Rr) w.v[5][tid] = sw_matvec_row<false>(w, Y, w.v[3], tid) + be[tid];
    __syncthreads();
    sw_matmul<false, true>(w, X, out, Z, Y, Ce, true);                                      // S' = (𝒜 W) 𝒜ᵀ + C
    if (tid < w.Rr) out[MM + tid] = w.v[5][tid];
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ pass 2 kernels (as scan.cuh)
__global__ void __launch_bounds__(256, 1) scanw_ks_kernel(const double* in, double* out, int n, int seg, int d, int Rr) {
    extern __shared__ __align__(16) unsigned char raw[];
    const SwSmem w = scanw_smem(raw, Rr);
    const int th = blockIdx.y, i = blockIdx.x;
    const size_t base = (size_t)th * n;
    const double* src = in + (base + i) * SELW;
    double* dst = out + (base + i) * SELW;
    if (i % seg >= d) {
        sw_combine(w, in + (base + i - d) * SELW, src, dst);
    } else {
        for (int k = threadIdx.x; k < SELW; k += blockDim.x) dst[k] = src[k];
    }
}
__global__ void scanw_gather_kernel(const double* __restrict__ pref, double* __restrict__ tot, int P, int G2, int G1) {
    const int th = blockIdx.y, g = blockIdx.x;
    const int last = min(P, (g + 1) * G2) - 1;
    const double* src = pref + ((size_t)th * P + last) * SELW;
    double* dst = tot + ((size_t)th * G1 + g) * SELW;
    for (int k = threadIdx.x; k < SELW; k += blockDim.x) dst[k] = src[k];
}
__global__ void __launch_bounds__(256, 1) scanw_group_states_kernel(const double* tp, double* gstate, int G1, int Rr) {
    extern __shared__ __align__(16) unsigned char raw[];
    const SwSmem w = scanw_smem(raw, Rr);
    const int th = blockIdx.y, g = blockIdx.x;
    double* S = gstate + ((size_t)th * G1 + g) * SSTATEW;
    if (g == 0) {
        for (int k = threadIdx.x; k < SSTATEW; k += blockDim.x) S[k] = 0.0;
        return;
    }
    sw_apply(w, tp + ((size_t)th * G1 + g - 1) * SELW, nullptr, S);
}
__global__ void __launch_bounds__(256, 1) scanw_states_kernel(const double* pref, const double* gstate, double* cstate, int P,
                                                              int G2, int G1, int Rr) {
    extern __shared__ __align__(16) unsigned char raw[];
    const SwSmem w = scanw_smem(raw, Rr);
    const int th = blockIdx.y, ch = blockIdx.x;
    const int g = ch / G2;
    const double* Q = pref + (size_t)th * P * SELW;
    const double* Sg = gstate + ((size_t)th * G1 + g) * SSTATEW;
    double* out = cstate + ((size_t)th * P + ch) * SSTATEW;
    if (ch == g * G2) {
        for (int k = threadIdx.x; k < SSTATEW; k += blockDim.x) out[k] = Sg[k];
        return;
    }
    sw_apply(w, Q + (size_t)(ch - 1) * SELW, g == 0 ? nullptr : Sg, out);
}
__global__ void __launch_bounds__(256, 1) scanw_substates_kernel(const double* subel, const double* cstate, double* substate,
                                                                 int P, int SUB, int Rr) {
    extern __shared__ __align__(16) unsigned char raw[];
    const SwSmem w = scanw_smem(raw, Rr);
    const int th = blockIdx.y, ch = blockIdx.x / (SUB - 1), j = blockIdx.x % (SUB - 1);
    const size_t q = (size_t)th * P + ch;
    sw_apply(w, subel + (q * (SUB - 1) + j) * SELW, ch == 0 ? nullptr : cstate + q * SSTATEW, substate + (q * (SUB - 1) + j) * SSTATEW);
}

// ------------------------------------------------------------------------------------------------ Newton refinement (as scan.cuh)
// T = 𝒜 (I − W J) with W = (I + S J)⁻¹ S;  m = η − J ĝ,  ĝ = (I + S J)⁻¹ (g + S η).  grid = (P, B).
__global__ void __launch_bounds__(256, 1) scanw_newton_T_kernel(const double* elems, const double* cstate, double* tm, int P,
                                                                int Rr) {
    extern __shared__ __align__(16) unsigned char raw[];
    const SwSmem w = scanw_smem(raw, Rr);
    constexpr int MM = SRW * SRW;
    const int th = blockIdx.y, ch = blockIdx.x, tid = threadIdx.x;
    const size_t q = (size_t)th * P + ch;
    const double* el = elems + q * SELW;
    const double *Ae = el, *Je = el + 2 * MM, *ete = el + 3 * MM + SRW;
    double* To = tm + q * SNEWTW;
    double* mo = To + MM;
    if (ch == 0) {
        for (int k = tid; k < MM; k += blockDim.x) To[k] = Ae[k];
        if (tid < SRW) mo[tid] = ete[tid];
        return;
    }
    const double* in = cstate + q * SSTATEW;
    double *X = w.m[0], *Y = w.m[1], *Z = w.m[2];
    sw_load(w, X, in);
    sw_load(w, Y, Je);
    if (tid < SRW) { w.v[0][tid] = tid < w.Rr ? in[MM + tid] : 0.0; w.v[1][tid] = tid < w.Rr ? ete[tid] : 0.0; }
    __syncthreads();
    sw_matmul<false, false>(w, Z, nullptr, X, Y, nullptr, false);                           // S J
    if (tid < w.Rr) {
        Z[tid * w.LD + tid] += 1.0;
        w.v[3][tid] = w.v[0][tid] + sw_matvec_row<false>(w, X, w.v[1], tid);                // g + S η
    }
    __syncthreads();
    sw_gj_solve(w, Z, X, nullptr, w.v[3], 1);    // X = W;  v3 = ĝ
    if (tid < SRW) mo[tid] = tid < w.Rr ? w.v[1][tid] - sw_matvec_row<false>(w, Y, w.v[3], tid) : 0.0;
    sw_matmul<false, false>(w, Z, nullptr, X, Y, nullptr, false);                           // W J
    sw_load(w, X, Ae);
    __syncthreads();
    sw_matmul<false, false>(w, Y, nullptr, X, Z, nullptr, false);                           // 𝒜 (W J)
    for (int k = tid; k < MM; k += blockDim.x) {
        const int r = k / SRW, c = k - r * SRW;
        To[k] = (r < w.Rr && c < w.Rr) ? X[r * w.LD + c] - Y[r * w.LD + c] : 0.0;
    }
}
// The affine scan of scan.cuh (scan_newton_prep / _ks / _apply) on three shared buffers.
constexpr int SNELW = 2 * SRW * SRW + 2 * SRW;
__global__ void scanw_newton_prep_kernel(const double* __restrict__ tm, const double* __restrict__ exits,
                                         const double* __restrict__ cstate, double* __restrict__ nel, int P) {
    const int th = blockIdx.y, k = blockIdx.x;
    if (k == 0) return;
    constexpr int MM = SRW * SRW;
    const size_t q = (size_t)th * P + k;
    const double* T = tm + (q - 1) * SNEWTW;
    const double* E = exits + (q - 1) * SSTATEW;
    const double* S = cstate + q * SSTATEW;
    double* el = nel + q * SNELW;
    for (int i = threadIdx.x; i < MM; i += blockDim.x) { el[i] = T[i]; el[MM + i] = E[i] - S[i]; }
    for (int i = threadIdx.x; i < SRW; i += blockDim.x) {
        el[2 * MM + i] = T[MM + i];
        el[2 * MM + SRW + i] = E[MM + i] - S[MM + i];
    }
}
__global__ void __launch_bounds__(256, 1) scanw_newton_ks_kernel(const double* in, double* out, int P, int d, int Rr) {
    extern __shared__ __align__(16) unsigned char raw[];
    const SwSmem w = scanw_smem(raw, Rr);
    const int th = blockIdx.y, k = blockIdx.x, tid = threadIdx.x;
    if (k == 0) return;
    constexpr int MM = SRW * SRW;
    const size_t q = (size_t)th * P + k;
    const double* ej = in + q * SNELW;
    double* eo = out + q * SNELW;
    if (k - d < 1) {
        for (int i = tid; i < SNELW; i += blockDim.x) eo[i] = ej[i];
        return;
    }
    const double* ei = in + (q - d) * SNELW;
    double *X = w.m[0], *Y = w.m[1], *Z = w.m[2];
    sw_load(w, X, ej);                   // T_j
    sw_load(w, Y, ei);                   // T_i
    if (tid < SRW) { w.v[0][tid] = tid < w.Rr ? ej[2 * MM + tid] : 0.0; w.v[1][tid] = tid < w.Rr ? ei[2 * MM + SRW + tid] : 0.0; }   // m_j, r^g_i
    __syncthreads();
    if (tid < SRW) eo[2 * MM + tid] = tid < w.Rr ? ei[2 * MM + tid] + sw_matvec_row<true>(w, Y, w.v[0], tid) : 0.0;     // m = m_i + T_iᵀ m_j
    sw_matmul<false, false>(w, nullptr, eo, X, Y, nullptr, false);                               // T = T_j T_i
    sw_load(w, Y, ei + MM);              // r_i
    __syncthreads();
    if (tid < w.Rr) w.v[3][tid] = w.v[1][tid] + sw_matvec_row<false>(w, Y, w.v[0], tid);          // r^g_i + r_i m_j
    __syncthreads();
    if (tid < SRW) eo[2 * MM + SRW + tid] = tid < w.Rr ? sw_matvec_row<false>(w, X, w.v[3], tid) + ej[2 * MM + SRW + tid] : 0.0;   // r^g
    sw_matmul<false, false>(w, Z, nullptr, X, Y, nullptr, false);                                // T_j r_i
    sw_matmul<false, true>(w, Y, eo + MM, Z, X, ej + MM, true);                                  // r = (T_j r_i) T_jᵀ + r_j
}
__global__ void scanw_newton_apply_kernel(const double* __restrict__ nel, double* __restrict__ cstate, int P) {
    const int th = blockIdx.y, k = blockIdx.x;
    if (k == 0) return;
    constexpr int MM = SRW * SRW;
    const size_t q = (size_t)th * P + k;
    const double* el = nel + q * SNELW;
    double* S = cstate + q * SSTATEW;
    for (int i = threadIdx.x; i < MM; i += blockDim.x) S[i] += el[MM + i];
    for (int i = threadIdx.x; i < SRW; i += blockDim.x) S[MM + i] += el[2 * MM + SRW + i];
}

// ------------------------------------------------------------------------------------------------ pass 3
// The register-file CTA kernel of wide.cuh in chunked form: one CTA per work item — steps [n_begin, n_end) of one parameter
// vector from the injected state `init` (S | g with leading dimension SRW; nullptr: start of the series), the three self-check
// segments of the warp kernel (sums of the first n_head steps → chk[0..1], of the whole range → part, of the n_ext steps after
// n_end → chk[2..3]) and, when `exit` is set, the state entering step n_end (taken after the tile update of that step, which is
// executed as a look-ahead step even when n_ext = 0).
template <int TS>
__global__ void __launch_bounds__(WIDE_THREADS, TS <= 6 ? 2 : 1) celerite_wide_chunk_kernel(const BatchArgs args) {
    constexpr int LD = 16 * TS;
    __shared__ __align__(16) double tabU[WIDE_CH][LD], tabV[WIDE_CH][LD], tabP[WIDE_CH][LD];
    __shared__ __align__(16) double qphi[LD], wphi[LD], p_s[LD];
    __shared__ double red[2 * (WIDE_THREADS / 32)];

    const WorkItem wk = args.work[blockIdx.x];
    const int th = wk.theta_begin;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int Jt = args.Jt;
    const int64_t N = wk.N, n0 = wk.n_begin, n1 = wk.n_end;
    const double* ca = args.a + (size_t)th * Jt;
    const double* cb = args.b + (size_t)th * Jt;
    const double* cc = args.c + (size_t)th * Jt;
    const double* cd = args.d + (size_t)th * Jt;
    const size_t pi = (size_t)wk.par_begin;
    const double mu = args.mu ? args.mu[pi * args.pstride] : 0.0;
    const double nu = args.nu ? args.nu[pi * args.pstride] : 1.0;
    const double* init = wk.init;

    double suma = 0.0;
    for (int m = 0; m < Jt; m++) suma += ca[m];

    double Tm[TS][TS];
#pragma unroll
    for (int i = 0; i < TS; i++)
#pragma unroll
        for (int j = 0; j < TS; j++) {
            const int r = ty + 16 * i, c = tx + 16 * j;
            Tm[i][j] = (init && r < SRW && c < SRW) ? init[(size_t)r * SRW + c] : 0.0;
        }
    for (int k = tid; k < WIDE_CH * LD; k += WIDE_THREADS) { (&tabU[0][0])[k] = 0.0; (&tabV[0][0])[k] = 0.0; (&tabP[0][0])[k] = 0.0; }
    const bool owner = tid < LD;
    double g = (init && owner && tid < SRW) ? init[(size_t)SRW * SRW + tid] : 0.0;
    double q = 0.0, w = 0.0, zprev = 0.0;
    double chi2 = 0.0, logsum = 0.0;              // sums of the current segment (Σ log|D| on the last warp only: off the owners' path)
    const bool summer = warp == WIDE_THREADS / 32 - 1, writer = tid == WIDE_THREADS - 32;
    const bool b3 = (tx & 8) != 0, b2 = (tx & 4) != 0, b1 = (tx & 2) != 0;
    const int64_t n_head_end = n0 + wk.n_head;
    const bool want_exit = wk.exit != nullptr && n1 < N;
    const int64_t n_stop = n1 + max((int64_t)wk.n_ext, (int64_t)(want_exit ? 1 : 0));
    __syncthreads();

    for (int64_t nb = n0; nb < n_stop; nb += WIDE_CH) {
        const int ns = (int)((n_stop - nb) < WIDE_CH ? (n_stop - nb) : WIDE_CH);
        for (int idx = tid; idx < ns * Jt; idx += WIDE_THREADS) {
            const int s = idx / Jt, m = idx - s * Jt;
            const int64_t n = nb + s;
            const double tn = wk.t[n];
            double ph = (n >= 1) ? exp(-cc[m] * (tn - wk.t[n - 1])) : 0.0;
            if (init && n == n0) ph = 1.0;          // the injected state is already decayed to t_n0
            const int tr = args.term_row[m];
            if (tr < 0) {
                const int r0 = -tr - 1;
                tabU[s][r0] = ca[m]; tabV[s][r0] = 1.0; tabP[s][r0] = ph;
            } else {
                double si, co;
                sincos_large(cd[m] * tn, &si, &co);
                tabU[s][tr] = ca[m] * co + cb[m] * si;
                tabU[s][tr + 1] = ca[m] * si - cb[m] * co;
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
            // segment hand-overs (uniform): first n_head steps → chk[0..1]; range → part; look-ahead → chk[2..3]
            if (n == n_head_end && wk.n_head > 0 && writer) { wk.chk[0] = logsum; wk.chk[1] = chi2; }
            if (n == n1) {
                if (writer) { wk.part[0] = logsum; wk.part[1] = chi2; }
                logsum = 0.0; chi2 = 0.0;
            }
            if (owner) {
                const double ph = Pn[tid];
                qphi[tid] = q * ph;
                wphi[tid] = w * ph;
                g = ph * fma(w, zprev, g);
            }
            __syncthreads();
            double phr[TS], qr[TS], rs[8];
#pragma unroll
            for (int i = 0; i < TS; i++) { phr[i] = Pn[ty + 16 * i]; qr[i] = qphi[ty + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 8; i++) rs[i] = 0.0;
#pragma unroll
            for (int j = 0; j < TS; j++) {
                const double pc = Pn[tx + 16 * j], wc = wphi[tx + 16 * j], uc = Un[tx + 16 * j];
#pragma unroll
                for (int i = 0; i < TS; i++) {
                    const double t = fma(phr[i], pc * Tm[i][j], qr[i] * wc);
                    Tm[i][j] = t;
                    rs[i] = fma(t, uc, rs[i]);
                }
            }
            if (want_exit && n == n1) {        // the state entering step n1: tile, then g of the owners
                double* E = wk.exit;
#pragma unroll
                for (int i = 0; i < TS; i++)
#pragma unroll
                    for (int j = 0; j < TS; j++) {
                        const int r = ty + 16 * i, c = tx + 16 * j;
                        if (r < SRW && c < SRW) E[(size_t)r * SRW + c] = Tm[i][j];
                    }
                if (owner && tid < SRW) E[(size_t)SRW * SRW + tid] = g;
                if (wk.n_ext == 0) break;      // uniform: the look-ahead step was only needed for the state
            }
            {
                double e4[4], e2[2];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double recv = __shfl_xor_sync(FULL, b3 ? rs[k] : rs[k + 4], 8);
                    e4[k] = (b3 ? rs[k + 4] : rs[k]) + recv;
                }
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const double recv = __shfl_xor_sync(FULL, b2 ? e4[k] : e4[k + 2], 4);
                    e2[k] = (b2 ? e4[k + 2] : e4[k]) + recv;
                }
                const double recv = __shfl_xor_sync(FULL, b1 ? e2[0] : e2[1], 2);
                double tot = (b1 ? e2[1] : e2[0]) + recv;
                tot += __shfl_xor_sync(FULL, tot, 1);
                const int isel = 4 * (b3 ? 1 : 0) + 2 * (b2 ? 1 : 0) + (b1 ? 1 : 0);
                if ((tx & 1) == 0 && isel < TS) p_s[ty + 16 * isel] = tot;
            }
            __syncthreads();
            double p = 0.0, sred = 0.0, ured = 0.0;
            if (owner) {
                p = p_s[tid];
                const double u = Un[tid];
                sred = u * p;
                ured = u * g;
            }
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) {
                sred += __shfl_xor_sync(FULL, sred, sft);
                ured += __shfl_xor_sync(FULL, ured, sft);
            }
            if (lane == 0) { red[2 * warp] = sred; red[2 * warp + 1] = ured; }
            __syncthreads();
            double stot = 0.0, utot = 0.0;
#pragma unroll
            for (int k = 0; k < WIDE_THREADS / 32; k++) { stot += red[2 * k]; utot += red[2 * k + 1]; }
            const double D = fma(nu, wk.s2[n], suma) - stot;      // celerite_solver.jl:92
            const double z = (wk.y[n] - mu) - utot;               // celerite_solver.jl:141
            const double rD = fast_rcp(D);
            if (owner) { q = Vn[tid] - p; w = q * rD; }
            zprev = z;
            chi2 = fma(z * z, rD, chi2);
            if (summer) logsum += (n == 0) ? log(D) : log(fabs(D));   // celerite_solver.jl:126,140
        }
        __syncthreads();
    }
    if (writer) {
        if (n_stop <= n1) { wk.part[0] = logsum; wk.part[1] = chi2; }            // no look-ahead: the range sums are still open
        else if (wk.n_ext > 0) { wk.chk[2] = logsum; wk.chk[3] = chi2; }
        if (wk.n_ext == 0) { wk.chk[2] = 0.0; wk.chk[3] = 0.0; }
        if (wk.n_head == 0) { wk.chk[0] = 0.0; wk.chk[1] = 0.0; }
    }
}

}  // namespace pioran
