// scan_blocked.cuh — pass 1 of K3 (the fold of every chunk into its composite) on the FP64 tensor pipe.
//
// scan_fold_kernel / scanw_fold_kernel fold one time step at a time: three rank-1 updates and two matrix-vector products per
// step behind two CTA barriers — 2.2 µs per step and CTA, latency-bound, and more than half of a K3 call.  The composite of a
// chunk obeys the same 8-step blocked algebra as the sweep itself (blocked.cuh), in the scan's state convention (S enters a step
// already decayed to its time, the decay φ_{n+1} follows the update): with Ψ'_{a→b} = Π_{a<i≤b} φ_{n0+i},
//     Û[:,s] = Ψ'_{0→s}∘U_s,   V̂[:,s] = Ψ'_{s→8}∘V_s,   ψ8 = Ψ'_{0→8},   K_blk[s][s'] = Σ_j U_s Ψ'_{s'→s} V_{s'},  diagonal A_n
//     P0 = X·Û,  C8 = K_blk − Ûᵀ·P0 = L D Lᵀ,  E = L⁻¹,  Q̂ = (V̂ − ψ8∘P0)·Eᵀ,  Ŵ = Q̂ D⁻¹,   X ← (ψ8ψ8ᵀ)∘X + Q̂·Ŵᵀ          (C, and b as its data row)
//     Pa = Y·Û,  Ga = Pa·Eᵀ,                                                      Y ← Y∘(1 ψ8ᵀ) − Ga·Ŵᵀ             (Y = 𝒜ᵀ)
//                                                                                 Z ← Z − Gz·(Gz D⁻¹)ᵀ              (J, and η as its data row/column:
//                                                                                  Gz = Ga with the whitened innovations z = Q̂[data row] as data row)
// (derivation and numpy check against the rank-1 fold: tests/tools/proto/blocked_fold_math.py).  One CTA of NT warps per chunk,
// warp w owning row tile w of X, Y and Z as FULL rows in the DMMA accumulator layout (blocked_wide.cuh): 5 NT² tile products per
// block and warp-set, two CTA barriers per 8 steps.  The per-θ block table (Û, V̂, ψ8, K_blk with its diagonal, 9 KB per block at
// rank 60) is built by scan_block_table_kernel (one CTA per block: rows in parallel, K_blk reduced over the rows) and streamed by
// 1-D TMA; pass 3 can sweep from the same table.
#pragma once
#include "blocked_wide.cuh"
#include "scan.cuh"

namespace pioran {

// Block record of the scan's table (doubles): UT [NT][8 steps][8 rows] | VH [8·NT rows][8 steps] | PSI8 [8·NT] | KB [8][8] | y σ² mask pad [32]
__host__ __device__ constexpr int sblk_off_vh(int NT) { return 64 * NT; }
__host__ __device__ constexpr int sblk_off_psi(int NT) { return 128 * NT; }
__host__ __device__ constexpr int sblk_off_kb(int NT) { return 136 * NT; }
__host__ __device__ constexpr int sblk_off_sc(int NT) { return 136 * NT + 64; }
__host__ __device__ constexpr int sblk_doubles(int NT) { return 136 * NT + 96; }
__host__ __device__ constexpr int sblk_nt(int R) { return (R + 8) / 8; }       // rows 0 … R−1 and the data row RG = R

// grid = (blocks of the range, B); block = 128 threads: thread r < 8·NT fills physical row r (= logical row r; the data row at R).
// term_row as in the generic kernels.  The record is zero-filled before the launch.
__global__ void __launch_bounds__(128) scan_block_table_kernel(double* __restrict__ tables, int64_t table_stride,
                                                               const double* __restrict__ t, const double* __restrict__ y,
                                                               const double* __restrict__ s2, int64_t N,
                                                               const double* __restrict__ a, const double* __restrict__ b,
                                                               const double* __restrict__ c, const double* __restrict__ d, int Jt,
                                                               const int* __restrict__ row_term, const int* __restrict__ row_kind,
                                                               int R, int NT, const double* __restrict__ mu, const double* __restrict__ nu,
                                                               const int64_t blk_first) {
    __shared__ double kpart[4][36];
    const int64_t blk = blk_first + blockIdx.x;       // absolute block number; record blockIdx.x of the table
    const int th = blockIdx.y, r = threadIdx.x, lane = r & 31, warp = r >> 5;
    double* tab = tables + (size_t)th * table_stride + (size_t)blockIdx.x * sblk_doubles(NT);
    const int64_t n0 = blk * BLK;
    const int kind = (r < R) ? row_kind[r] : (r == R ? ROW_AUG : ROW_PAD);
    double ph[BLK], ut[BLK], vv[BLK];
    double ca = 0.0, cb = 0.0, cdec = 0.0, dfreq = 0.0;
    if (r < R) { const size_t k = (size_t)th * Jt + row_term[r]; ca = a[k]; cb = b[k]; cdec = c[k]; dfreq = d[k]; }
    const double muv = mu ? mu[th] : 0.0;
#pragma unroll
    for (int s = 0; s < BLK; s++) {
        const int64_t n = n0 + s;
        ph[s] = 1.0; ut[s] = 0.0; vv[s] = 0.0;
        if (kind == ROW_PAD) { ph[s] = 0.0; continue; }
        if (n >= N) continue;                                        // padded step: identity
        if (kind == ROW_AUG) { vv[s] = y[n] - muv; continue; }
        const double tn = t[n];
        ph[s] = (n + 1 < N) ? exp(-cdec * (t[n + 1] - tn)) : 0.0;    // φ_{n+1}: the decay that FOLLOWS step n (φ_N := 0)
        if (kind == ROW_REAL) { ut[s] = ca; vv[s] = 1.0; }
        else {
            double si, co;
            sincos_large(dfreq * tn, &si, &co);
            if (kind == ROW_COS) { ut[s] = fma(cb, si, ca * co); vv[s] = co; }      // celerite_solver.jl:60
            else                 { ut[s] = fma(-cb, co, ca * si); vv[s] = si; }     // celerite_solver.jl:59
        }
    }
    // Ψ'_{0→s} = Π_{i<s} ph[i],  Ψ'_{s→8} = Π_{i≥s} ph[i]
    double p0[BLK], pe[BLK];
    p0[0] = 1.0;
#pragma unroll
    for (int s = 1; s < BLK; s++) p0[s] = p0[s - 1] * ph[s - 1];
    pe[BLK - 1] = ph[BLK - 1];
#pragma unroll
    for (int s = BLK - 2; s >= 0; s--) pe[s] = pe[s + 1] * ph[s];
    const bool live = r < 8 * NT;
    if (live) {
        const int K = r >> 3, rr = r & 7;
#pragma unroll
        for (int s = 0; s < BLK; s++) {
            tab[K * 64 + s * 8 + rr] = p0[s] * ut[s];
            tab[sblk_off_vh(NT) + r * 8 + s] = pe[s] * vv[s];
        }
        tab[sblk_off_psi(NT) + r] = (kind == ROW_AUG) ? 1.0 : p0[BLK - 1] * ph[BLK - 1];
    }
    // K_blk: couplings of this row, summed over the rows (warp shuffles, then the four warps through shared memory)
    double kv[36];
    {
        int q = 0;
#pragma unroll
        for (int s = 1; s < BLK; s++) {
            double dec = 1.0;
#pragma unroll
            for (int sp = s - 1; sp >= 0; sp--) {
                dec *= ph[sp];
                kv[pair_slot(s, sp)] = (kind == ROW_COS || kind == ROW_SIN || kind == ROW_REAL) ? ut[s] * dec * vv[sp] : 0.0;
                q++;
            }
        }
        (void)q;
#pragma unroll
        for (int k = 28; k < 36; k++) kv[k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < 28; k++) {
        double v = kv[k];
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
        if (lane == 0) kpart[warp][k] = v;
    }
    __syncthreads();
    if (r < 64) {
        const int g = r >> 3, cc = r & 7;
        const int64_t n = n0 + g;
        double v;
        if (g == cc) {
            // diagonal A_n = Σa + ν σ²_n (celerite_solver.jl:92); padded steps are unit pivots
            double suma = 0.0;
            for (int m = 0; m < Jt; m++) suma += a[(size_t)th * Jt + m];
            v = (n < N) ? fma(nu ? nu[th] : 1.0, s2[n], suma) : 1.0;
        } else {
            const int k = pair_slot(max(g, cc), min(g, cc));
            v = (kpart[0][k] + kpart[1][k]) + (kpart[2][k] + kpart[3][k]);
        }
        tab[sblk_off_kb(NT) + r] = v;
    }
    if (r < 32) {
        const int s = r & 7, f = r >> 3;
        const int64_t n = n0 + s;
        double sc = 0.0;
        if (n < N) sc = (f == 0) ? y[n] : (f == 1) ? s2[n] : (f == 2) ? 1.0 : 0.0;
        tab[sblk_off_sc(NT) + r] = sc;
    }
}

constexpr int SFB_NSTAGE = 3;
template <int NT>
constexpr size_t sfb_smem_bytes() {
    return sizeof(double) * ((size_t)SFB_NSTAGE * sblk_doubles(NT) + (size_t)NT * 64 + 2 * (size_t)NT * 64) + SFB_NSTAGE * sizeof(uint64_t) + 16;
}

// Composite of the chunk so far → global (𝒜 | C | J | b | η with leading dimension LDS; entries outside the live rank are not written).
template <int NT>
__device__ __forceinline__ void sfb_store(double* __restrict__ E, const int LDS, const int R, const int warp, const int g, const int t,
                                          const double (&X)[NT][2], const double (&Y)[NT][2], const double (&Z)[NT][2]) {
    const size_t MM = (size_t)LDS * LDS;
    const int row = 8 * warp + g;
#pragma unroll
    for (int K = 0; K < NT; K++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int col = 8 * K + 2 * t + e;
            if (col < R) {
                if (row < R) {
                    E[(size_t)col * LDS + row] = Y[K][e];              // 𝒜 = Yᵀ
                    E[MM + (size_t)row * LDS + col] = X[K][e];         // C
                    E[2 * MM + (size_t)row * LDS + col] = Z[K][e];     // J
                } else if (row == R) {
                    E[3 * MM + col] = X[K][e];                         // b: the data row of X
                    E[3 * MM + LDS + col] = Z[K][e];                   // η: the data row of Z
                }
            }
        }
    }
}

// grid = (P, B); block = NT warps.  Chunk bounds are multiples of 8 (the table's block grid).
template <int NT>
__global__ void __launch_bounds__(NT * 32, 1) scan_fold_blocked_kernel(const ScanArgs args, const double* __restrict__ tables,
                                                                       const int64_t table_stride, const int R, const int LDS,
                                                                       const int SELr) {
    constexpr int BD = sblk_doubles(NT), O_VH = sblk_off_vh(NT), O_PSI = sblk_off_psi(NT), O_KB = sblk_off_kb(NT);
    constexpr uint32_t STAGE_BYTES = BD * sizeof(double);
    constexpr int W = NT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    double* cred = stages + SFB_NSTAGE * BD;             // [W][32][2]
    double* wpub = cred + W * 64;                        // [NT][32][2]  Ŵ
    double* gpub = wpub + NT * 64;                       // [NT][32][2]  Gz D⁻¹
    uint64_t* bars = reinterpret_cast<uint64_t*>(gpub + NT * 64);

    const int th = blockIdx.y, ch = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n0c = args.bounds[ch], n1c = args.bounds[ch + 1];
    const int64_t b0 = n0c / BLK, b1 = (n1c + BLK - 1) / BLK, nblocks = b1 - b0;
    const double* tbase = tables + (size_t)th * table_stride + (size_t)b0 * BD;

    if (threadIdx.x == 0) {
        for (int k = 0; k < SFB_NSTAGE; k++) mbar_init(&bars[k], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < SFB_NSTAGE && k < nblocks; k++) {
            mbar_arrive_expect_tx(&bars[k], STAGE_BYTES);
            tma_load_1d(stages + k * BD, tbase + (size_t)k * BD, STAGE_BYTES, &bars[k]);
        }
    }
    const BlkLane L = make_blk_lane(lane);
    const int g = L.g, t = L.t;
    const int I = warp, row = 8 * I + g;
    const bool isrg = (row == R);

    double X[NT][2], Y[NT][2], Z[NT][2];
#pragma unroll
    for (int K = 0; K < NT; K++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            X[K][e] = 0.0; Z[K][e] = 0.0;
            Y[K][e] = (row < R && row == 8 * K + 2 * t + e) ? 1.0 : 0.0;      // 𝒜 = I on the live rows
        }

    int next_sub = 1;
    int64_t next_bound = scan_sub_bound(n0c, n1c, 1, args.SUB);
    int sidx = 0;
    uint32_t parity = 0;
    for (int64_t bb = 0; bb < nblocks; bb++) {
        // running composite at an inner boundary (a multiple of 8 steps from the chunk start)
        while (next_sub < args.SUB && n0c + bb * BLK >= next_bound) {
            sfb_store<NT>(args.subel + (((size_t)th * args.P + ch) * (args.SUB - 1) + (next_sub - 1)) * SELr, LDS, R, warp, g, t, X, Y, Z);
            next_sub++;
            next_bound = scan_sub_bound(n0c, n1c, next_sub, args.SUB);
        }
        mbar_wait(&bars[sidx], parity);
        const double* tab = stages + sidx * BD;

        // ---- 1. P0 = X·Û, Pa = Y·Û on my row tile; my share of C8 = K_blk − Ûᵀ·P0
        double P0[2] = {0.0, 0.0}, Pa[2] = {0.0, 0.0};
#pragma unroll
        for (int K = 0; K < NT; K++) {
            const double2 u = *reinterpret_cast<const double2*>(tab + K * 64 + g * 8 + 2 * t);
            dmma(P0[0], P0[1], X[K][0], u.x);
            dmma(P0[0], P0[1], X[K][1], u.y);
            dmma(Pa[0], Pa[1], Y[K][0], u.x);
            dmma(Pa[0], Pa[1], Y[K][1], u.y);
        }
        {
            double ca0 = 0.0, ca1 = 0.0;
            const double2 u = *reinterpret_cast<const double2*>(tab + I * 64 + g * 8 + 2 * t);
            double p0, p1;
            tile_transpose(L, P0[0], P0[1], p0, p1);
            dmma(ca0, ca1, u.x, p0);
            dmma(ca0, ca1, u.y, p1);
            double2 kb = make_double2(0.0, 0.0);
            if (warp == 0) kb = *reinterpret_cast<const double2*>(tab + O_KB + g * 8 + 2 * t);
            *reinterpret_cast<double2*>(cred + (warp * 32 + lane) * 2) = make_double2(kb.x - ca0, kb.y - ca1);
        }
        __syncthreads();
        double cm0 = 0.0, cm1 = 0.0;
#pragma unroll
        for (int q = 0; q < W; q++) {
            const double2 c = *reinterpret_cast<const double2*>(cred + (q * 32 + lane) * 2);
            cm0 += c.x; cm1 += c.y;
        }

        // ---- 2. decays, Bm, the 8×8 LDLᵀ (redundant in every warp), Q̂, Ga
        const double psr = tab[O_PSI + row];
#pragma unroll
        for (int K = 0; K < NT; K++) {
            const double2 pc = *reinterpret_cast<const double2*>(tab + O_PSI + 8 * K + 2 * t);
            X[K][0] *= psr * pc.x; X[K][1] *= psr * pc.y;
            Y[K][0] *= pc.x;       Y[K][1] *= pc.y;
        }
        {
            const double2 vh = *reinterpret_cast<const double2*>(tab + O_VH + row * 8 + 2 * t);
            P0[0] = fma(-psr, P0[0], vh.x);
            P0[1] = fma(-psr, P0[1], vh.y);
        }
        double e0 = L.cdiag0 ? 1.0 : 0.0, e1 = L.cdiag1 ? 1.0 : 0.0;
        double rd0 = 0.0, rd1 = 0.0;
        const int rowbase = lane & ~3;
#pragma unroll
        for (int j = 0; j < BLK; j++) {
            const int tj = j >> 1;
            const double cme = (j & 1) ? cm1 : cm0;
            const double dj = __shfl_sync(FULL, cme, 4 * j + tj);
            const double cgj = __shfl_sync(FULL, cme, rowbase | tj);
            const double rdj = fast_rcp(dj);
            if (j & 1) rd1 = (t == tj) ? rdj : rd1;
            else       rd0 = (t == tj) ? rdj : rd0;
            const double l = cgj * rdj;
            const double cj0 = __shfl_sync(FULL, cm0, 4 * j + t), cj1 = __shfl_sync(FULL, cm1, 4 * j + t);
            const double ej0 = __shfl_sync(FULL, e0, 4 * j + t), ej1 = __shfl_sync(FULL, e1, 4 * j + t);
            const double lm = (g > j) ? l : 0.0;
            cm0 = fma(-lm, cj0, cm0); cm1 = fma(-lm, cj1, cm1);
            e0 = fma(-lm, ej0, e0);   e1 = fma(-lm, ej1, e1);
        }
        double Q[2] = {0.0, 0.0}, Ga[2] = {0.0, 0.0};
        dmma(Q[0], Q[1], P0[0], e0);
        dmma(Q[0], Q[1], P0[1], e1);
        dmma(Ga[0], Ga[1], Pa[0], e0);
        dmma(Ga[0], Ga[1], Pa[1], e1);
        // Gz: Ga with the whitened innovations of the data row (η rides in J as b rides in C)
        const double gz0 = isrg ? Q[0] : Ga[0], gz1 = isrg ? Q[1] : Ga[1];
        *reinterpret_cast<double2*>(wpub + (I * 32 + lane) * 2) = make_double2(Q[0] * rd0, Q[1] * rd1);
        *reinterpret_cast<double2*>(gpub + (I * 32 + lane) * 2) = make_double2(gz0 * rd0, gz1 * rd1);
        __syncthreads();
        if (threadIdx.x == 0 && bb + SFB_NSTAGE < nblocks) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&bars[sidx], STAGE_BYTES);
            tma_load_1d(stages + sidx * BD, tbase + (size_t)(bb + SFB_NSTAGE) * BD, STAGE_BYTES, &bars[sidx]);
        }
        // ---- 3. rank-8 updates of my row tile of X, Y, Z
        const double ng0 = -Ga[0], ng1 = -Ga[1], nz0 = -gz0, nz1 = -gz1;
#pragma unroll
        for (int K = 0; K < NT; K++) {
            const double2 wv = *reinterpret_cast<const double2*>(wpub + (K * 32 + lane) * 2);
            const double2 gv = *reinterpret_cast<const double2*>(gpub + (K * 32 + lane) * 2);
            dmma(X[K][0], X[K][1], Q[0], wv.x);
            dmma(X[K][0], X[K][1], Q[1], wv.y);
            dmma(Y[K][0], Y[K][1], ng0, wv.x);
            dmma(Y[K][0], Y[K][1], ng1, wv.y);
            dmma(Z[K][0], Z[K][1], nz0, gv.x);
            dmma(Z[K][0], Z[K][1], nz1, gv.y);
        }
        if (++sidx == SFB_NSTAGE) { sidx = 0; parity ^= 1; }
    }
    sfb_store<NT>(args.elems + ((size_t)th * args.P + ch) * SELr, LDS, R, warp, g, t, X, Y, Z);
}

}  // namespace pioran

namespace pioran {

// ------------------------------------------------------------------------------------------------ pass 3 on the tensor pipe
// Re-filter of one chunk from its injected state on the scan's block table: the blocked sweep of blocked_wide.cuh (four warps per
// chunk, warp w owning the row tiles I ≡ w mod 4 as full rows) in the scan's state convention, so the injected state IS X and the
// state after the last block IS the exit state of the Newton refinement — no pending factors.  Chunk bounds sit on the block grid;
// the self-check segments are whole blocks: the sums of the first block → chk[0..1], of the chunk → part, of the block after the
// chunk (swept from this chunk's own state, not stored) → chk[2..3].
constexpr int SSB_W = 4, SSB_NSTAGE = 3;
template <int NT>
constexpr size_t ssb_smem_bytes() {
    return sizeof(double) * ((size_t)SSB_NSTAGE * sblk_doubles(NT) + (size_t)SSB_W * 64 + (size_t)NT * 64) + SSB_NSTAGE * sizeof(uint64_t) + 16;
}
// grid = work items (one chunk of one parameter vector each); block = 4 warps.
template <int NT>
__global__ void __launch_bounds__(SSB_W * 32, 1) scan_sweep_blocked_kernel(const WorkItem* __restrict__ work, const double* __restrict__ tables,
                                                                           const int64_t table_stride, const int R, const int LDS) {
    constexpr int BD = sblk_doubles(NT), O_VH = sblk_off_vh(NT), O_PSI = sblk_off_psi(NT), O_KB = sblk_off_kb(NT), O_SC = sblk_off_sc(NT);
    constexpr int W = SSB_W, MR = (NT + W - 1) / W;
    constexpr uint32_t STAGE_BYTES = BD * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    double* cred = stages + SSB_NSTAGE * BD;
    double* wpub = cred + W * 64;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wpub + NT * 64);
    __shared__ double fin[2];

    const WorkItem wk = work[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t N = wk.N;
    const int64_t b0 = wk.n_begin / BLK, b1 = (wk.n_end + BLK - 1) / BLK;
    const bool look = wk.n_ext > 0 && wk.n_end < N;               // one more block for the self-check
    const int64_t nblocks = b1 - b0 + (look ? 1 : 0);
    const double* tbase = tables + (size_t)wk.theta_begin * table_stride + (size_t)b0 * BD;

    if (threadIdx.x == 0) {
        for (int k = 0; k < SSB_NSTAGE; k++) mbar_init(&bars[k], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < SSB_NSTAGE && k < nblocks; k++) {
            mbar_arrive_expect_tx(&bars[k], STAGE_BYTES);
            tma_load_1d(stages + k * BD, tbase + (size_t)k * BD, STAGE_BYTES, &bars[k]);
        }
    }
    const BlkLane L = make_blk_lane(lane);
    const int g = L.g, t = L.t;
    const size_t MM = (size_t)LDS * LDS;

    double x[MR][NT][2];
#pragma unroll
    for (int i = 0; i < MR; i++) {
        const int row = 8 * (warp + W * i) + g;
#pragma unroll
        for (int K = 0; K < NT; K++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int col = 8 * K + 2 * t + e;
                double v = 0.0;
                if (wk.init && col < R) {
                    if (row < R) v = wk.init[(size_t)row * LDS + col];
                    else if (row == R) v = wk.init[MM + col];
                }
                x[i][K][e] = v;
            }
    }
    double chi2 = 0.0, logsum = 0.0;            // sums of the current segment (every warp keeps the pivots; warp 0 reports Σ log|D|)
    constexpr int WOWN = (NT - 1) % W;          // the warp that owns the data row's tile (row tile NT − 1)

    int sidx = 0;
    uint32_t parity = 0;
    for (int64_t bb = 0; bb < nblocks; bb++) {
        // segment hand-overs (uniform): after the first block → chk[0..1]; after the chunk's last block → part and the exit state
        if (bb == 1 && wk.n_head > 0) {
            double ls = logsum, ch = chi2;
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) { ls += __shfl_xor_sync(FULL, ls, sft); ch += __shfl_xor_sync(FULL, ch, sft); }
            if (warp == 0 && lane == 0) wk.chk[0] = ls;
            if (warp == WOWN && lane == 0) wk.chk[1] = ch;
        }
        if (bb == b1 - b0) {
            double ls = logsum, ch = chi2;
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) { ls += __shfl_xor_sync(FULL, ls, sft); ch += __shfl_xor_sync(FULL, ch, sft); }
            if (warp == 0 && lane == 0) wk.part[0] = ls;
            if (warp == WOWN && lane == 0) wk.part[1] = ch;
            logsum = 0.0; chi2 = 0.0;
            if (wk.exit) {
#pragma unroll
                for (int i = 0; i < MR; i++) {
                    const int row = 8 * (warp + W * i) + g;
#pragma unroll
                    for (int K = 0; K < NT; K++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int col = 8 * K + 2 * t + e;
                            if (col < R) {
                                if (row < R) wk.exit[(size_t)row * LDS + col] = x[i][K][e];
                                else if (row == R) wk.exit[MM + col] = x[i][K][e];
                            }
                        }
                }
            }
        }
        mbar_wait(&bars[sidx], parity);
        const double* tab = stages + sidx * BD;

        double P0[MR][2];
#pragma unroll
        for (int i = 0; i < MR; i++) P0[i][0] = P0[i][1] = 0.0;
#pragma unroll
        for (int K = 0; K < NT; K++) {
            const double2 u = *reinterpret_cast<const double2*>(tab + K * 64 + g * 8 + 2 * t);
#pragma unroll
            for (int i = 0; i < MR; i++) {
                dmma(P0[i][0], P0[i][1], x[i][K][0], u.x);
                dmma(P0[i][0], P0[i][1], x[i][K][1], u.y);
            }
        }
        {
            double ca0 = 0.0, ca1 = 0.0;
#pragma unroll
            for (int i = 0; i < MR; i++) {
                const int I = warp + W * i;
                if (I < NT) {
                    const double2 u = *reinterpret_cast<const double2*>(tab + I * 64 + g * 8 + 2 * t);
                    double p0, p1;
                    tile_transpose(L, P0[i][0], P0[i][1], p0, p1);
                    dmma(ca0, ca1, u.x, p0);
                    dmma(ca0, ca1, u.y, p1);
                }
            }
            double2 kb = make_double2(0.0, 0.0);
            if (warp == 0) kb = *reinterpret_cast<const double2*>(tab + O_KB + g * 8 + 2 * t);
            *reinterpret_cast<double2*>(cred + (warp * 32 + lane) * 2) = make_double2(kb.x - ca0, kb.y - ca1);
        }
        __syncthreads();
        double cm0 = 0.0, cm1 = 0.0;
#pragma unroll
        for (int q = 0; q < W; q++) {
            const double2 c = *reinterpret_cast<const double2*>(cred + (q * 32 + lane) * 2);
            cm0 += c.x; cm1 += c.y;
        }
        double psr[MR];
#pragma unroll
        for (int i = 0; i < MR; i++) { const int I = warp + W * i; psr[i] = I < NT ? tab[O_PSI + 8 * I + g] : 0.0; }
#pragma unroll
        for (int K = 0; K < NT; K++) {
            const double2 pc = *reinterpret_cast<const double2*>(tab + O_PSI + 8 * K + 2 * t);
#pragma unroll
            for (int i = 0; i < MR; i++) {
                x[i][K][0] *= psr[i] * pc.x;
                x[i][K][1] *= psr[i] * pc.y;
            }
        }
#pragma unroll
        for (int i = 0; i < MR; i++) {
            const int I = warp + W * i;
            if (I < NT) {
                const double2 vh = *reinterpret_cast<const double2*>(tab + O_VH + (8 * I + g) * 8 + 2 * t);
                P0[i][0] = fma(-psr[i], P0[i][0], vh.x);
                P0[i][1] = fma(-psr[i], P0[i][1], vh.y);
            } else {
                P0[i][0] = 0.0; P0[i][1] = 0.0;
            }
        }
        double e0 = L.cdiag0 ? 1.0 : 0.0, e1 = L.cdiag1 ? 1.0 : 0.0;
        double rd0 = 0.0, rd1 = 0.0;
        const int rowbase = lane & ~3;
        double dmine = 1.0;                     // lane j < 8 keeps the pivot of step j
#pragma unroll
        for (int j = 0; j < BLK; j++) {
            const int tj = j >> 1;
            const double cme = (j & 1) ? cm1 : cm0;
            const double dj = __shfl_sync(FULL, cme, 4 * j + tj);          // pivot D_n (celerite_solver.jl:92)
            const double cgj = __shfl_sync(FULL, cme, rowbase | tj);
            const double rdj = fast_rcp(dj);
            if (j & 1) rd1 = (t == tj) ? rdj : rd1;
            else       rd0 = (t == tj) ? rdj : rd0;
            const double l = cgj * rdj;
            const double cj0 = __shfl_sync(FULL, cm0, 4 * j + t), cj1 = __shfl_sync(FULL, cm1, 4 * j + t);
            const double ej0 = __shfl_sync(FULL, e0, 4 * j + t), ej1 = __shfl_sync(FULL, e1, 4 * j + t);
            const double lm = (g > j) ? l : 0.0;
            cm0 = fma(-lm, cj0, cm0); cm1 = fma(-lm, cj1, cm1);
            e0 = fma(-lm, ej0, e0);   e1 = fma(-lm, ej1, e1);
            if (lane == j) dmine = dj;
        }
        // Σ log|D_n| (celerite_solver.jl:140; the first pivot of the series without abs, :126): 8 lanes of warp 0, one log each
        if (warp == 0 && lane < BLK) logsum += (b0 + bb == 0 && lane == 0) ? log(dmine) : log(fabs(dmine));

        double Q[MR][2];
#pragma unroll
        for (int i = 0; i < MR; i++) {
            const int I = warp + W * i;
            Q[i][0] = Q[i][1] = 0.0;
            dmma(Q[i][0], Q[i][1], P0[i][0], e0);
            dmma(Q[i][0], Q[i][1], P0[i][1], e1);
            if (I == NT - 1) {
                const bool isrg = (8 * I + g == R);
                const double zz = fma(Q[i][0] * rd0, Q[i][0], (Q[i][1] * rd1) * Q[i][1]);     // Σ z²/D (celerite_solver.jl:333)
                chi2 += isrg ? zz : 0.0;
            }
            if (I < NT) *reinterpret_cast<double2*>(wpub + (I * 32 + lane) * 2) = make_double2(Q[i][0] * rd0, Q[i][1] * rd1);
        }
        __syncthreads();
        if (threadIdx.x == 0 && bb + SSB_NSTAGE < nblocks) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&bars[sidx], STAGE_BYTES);
            tma_load_1d(stages + sidx * BD, tbase + (size_t)(bb + SSB_NSTAGE) * BD, STAGE_BYTES, &bars[sidx]);
        }
#pragma unroll
        for (int K = 0; K < NT; K++) {
            const double2 wv = *reinterpret_cast<const double2*>(wpub + (K * 32 + lane) * 2);
#pragma unroll
            for (int i = 0; i < MR; i++) {
                dmma(x[i][K][0], x[i][K][1], Q[i][0], wv.x);
                dmma(x[i][K][0], x[i][K][1], Q[i][1], wv.y);
            }
        }
        if (++sidx == SSB_NSTAGE) { sidx = 0; parity ^= 1; }
    }
    // the last segment: the chunk itself (no look-ahead block) or the look-ahead block
    {
        double ls = logsum, ch = chi2;
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) { ls += __shfl_xor_sync(FULL, ls, sft); ch += __shfl_xor_sync(FULL, ch, sft); }
        if (warp == 0 && lane == 0) fin[0] = ls;
        if (warp == WOWN && lane == 0) fin[1] = ch;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (look) { wk.chk[2] = fin[0]; wk.chk[3] = fin[1]; }
        else { wk.part[0] = fin[0]; wk.part[1] = fin[1]; wk.chk[2] = 0.0; wk.chk[3] = 0.0; }
        if (wk.n_head == 0) { wk.chk[0] = 0.0; wk.chk[1] = 0.0; }
    }
    (void)O_SC;
}

}  // namespace pioran
