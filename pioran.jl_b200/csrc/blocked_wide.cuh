// blocked_wide.cuh — K2tw: the blocked (tensor-pipe) celerite sweep of blocked.cuh for ranks 65 … 128, one CTA per evaluation.
//
// The warp kernel K2t keeps the symmetric state of one evaluation in the registers of ONE warp (ranks ≤ 63).  The reference's
// benchmark grid goes on to SHO J = 40, 50 and DRWCelerite J = 30, 40 (ranks 80 … 120, benchmark/benchmarks.jl:16-18), which round 1
// served with a rank-1 kernel at three CTA barriers per time step (wide.cuh).  Here the same blocked algebra as blocked.cuh —
//     P0 = X·Û,  C = K_blk − Ûᵀ·P0 = L D Lᵀ,  Q̂ = (amp∘V̂ − ψ8∘P0)·L⁻ᵀ,  Ŵ = Q̂ D⁻¹,  X ← (ψ8ψ8ᵀ)∘X + Q̂·Ŵᵀ      (8 steps per block)
// — runs on W warps that share one evaluation: warp w owns the ROW TILES I ≡ w (mod W) of the state as FULL rows (all column
// tiles, both triangles), in the DMMA accumulator layout.  With full rows P0[I] = Σ_K X[I][K]·Û[K] needs no transposed partner and
// no cross-warp sum; the price is the rank-8 update on both triangles (1.5× the DMMAs of the symmetric warp kernel).  Per block:
//   1. every warp: P0 of its rows, its share of K_blk (row pairs jp ≡ 2w mod 2W) minus its share of ÛᵀP0 → 2 doubles per lane
//      to shared memory;                                                                                    barrier
//   2. every warp sums the W shares (same order: identical C everywhere), runs the 8×8 LDLᵀ redundantly, forms Bm, Q̂ of its rows,
//      publishes Ŵ of its rows lane for lane (it is the B operand of the update as it stands);                 barrier
//   3. every warp updates its rows with the Ŵ of all column tiles.
// Two CTA barriers per 8 steps instead of three per step.  Same block table as K2t (blocked_table_kernel, any NT), staged by 1-D TMA.
#pragma once
#include "blocked.cuh"

namespace pioran {

constexpr int BLKW_NSTAGE = 2;

template <int NT, int NTR, int W>
constexpr size_t blkw_smem_bytes() {
    return sizeof(double) * ((size_t)BLKW_NSTAGE * blk_doubles(NT, NTR) + 8 * NTR + 8 * NT + (size_t)W * 64 + (size_t)NT * 64) +
           BLKW_NSTAGE * sizeof(uint64_t) + 16;
}

// grid = work items (one parameter vector each); block = W warps.
template <int NT, int NTR, int W>
__global__ void __launch_bounds__(W * 32, 1) celerite_blocked_wide_kernel(const BatchArgs args, const int R, const int amp_stride,
                                                                          const int RG) {
    constexpr int BD = blk_doubles(NT, NTR), RPT = 8 * NTR, MR = (NTR + W - 1) / W;
    constexpr int O_VH = blk_off_vh(NT, NTR), O_PSI = blk_off_psi(NT, NTR), O_H = blk_off_h(NT, NTR), O_SC = blk_off_sc(NT, NTR);
    constexpr uint32_t STAGE_BYTES = BD * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    double* amp_s = stages + BLKW_NSTAGE * BD;       // amplitudes by physical row (amp[RG] = 1)
    double* amp_l = amp_s + RPT;                     // amplitudes by logical row (K_blk)
    double* cred = amp_l + 8 * NT;                   // [W][32][2] shares of C
    double* wpub = cred + W * 64;                    // [NT][32][2] Ŵ of every column tile, lane for lane
    uint64_t* bars = reinterpret_cast<uint64_t*>(wpub + NT * 64);

    const WorkItem wk = args.work[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t N = wk.N;
    const int64_t nblocks = (N + BLK - 1) / BLK;
    const int th = wk.theta_begin;

    if (threadIdx.x == 0) {
        for (int k = 0; k < BLKW_NSTAGE; k++) mbar_init(&bars[k], 1);
        fence_mbar_init();
    }
    for (int k = threadIdx.x; k < RPT + 8 * NT; k += W * 32) amp_s[k] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < BLKW_NSTAGE && k < nblocks; k++) {
            mbar_arrive_expect_tx(&bars[k], STAGE_BYTES);
            tma_load_1d(stages + k * BD, wk.table + (size_t)k * BD, STAGE_BYTES, &bars[k]);
        }
    }
    for (int k = threadIdx.x; k < R; k += W * 32) {
        const double av = args.amp[(size_t)th * amp_stride + k];
        amp_s[blk_phys_row(k, R)] = av;
        amp_l[k] = av;
    }
    if (threadIdx.x == 0) amp_s[RG] = 1.0;
    __syncthreads();
    const BlkLane L = make_blk_lane(lane);
    const int g = L.g, t = L.t;
    const double suma = args.suma[th];
    const size_t pi = (size_t)wk.par_begin;
    const double mu = args.mu ? args.mu[pi * args.pstride] : 0.0;
    const double nu = args.nu ? args.nu[pi * args.pstride] : 1.0;
    const double* yb = args.y_batch ? args.y_batch + pi * args.ystride : nullptr;      // per-parameter-vector data (explicit coefficients)
    const double* sb = args.s2_batch ? args.s2_batch + pi * args.ystride : nullptr;

    double x[MR][NT][2];
#pragma unroll
    for (int i = 0; i < MR; i++)
#pragma unroll
        for (int K = 0; K < NT; K++) x[i][K][0] = x[i][K][1] = 0.0;
    double chi2 = 0.0, logacc = 0.0, dkeep = 1.0, dfirst = 1.0;

    int sidx = 0;
    uint32_t parity = 0;
    for (int64_t b = 0; b < nblocks; b++) {
        mbar_wait(&bars[sidx], parity);
        const double* tab = stages + sidx * BD;
        const int64_t n0 = b * BLK;

        // ---- 1a. P0 of my row tiles
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
        // ---- 1b. my share of C = K_blk − Ûᵀ·P0
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
            const double2* Hq = reinterpret_cast<const double2*>(tab + O_H) + lane;
            const double2* Aq = reinterpret_cast<const double2*>(amp_l);
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int jp = 2 * warp; jp < 4 * NT; jp += 2 * W) {
                const double2 h0 = Hq[jp * 32], h1 = Hq[(jp + 1) * 32];
                const double2 m0 = Aq[jp], m1 = Aq[jp + 1];
                a0 = fma(m0.x, h0.x, a0); a1 = fma(m0.y, h0.y, a1);
                a2 = fma(m1.x, h1.x, a2); a3 = fma(m1.y, h1.y, a3);
            }
            const double v = (a0 + a1) + (a2 + a3);
            const double o0 = __shfl_sync(FULL, v, L.csrc0);
            const double o1 = __shfl_sync(FULL, v, L.csrc1);
            // diagonal (warp 0): A_n = Σa + ν σ²_n (celerite_solver.jl:92); padded steps are unit pivots
            const double mk = tab[O_SC + 16 + g];
            const double s2v = sb ? (n0 + g < N ? sb[n0 + g] : 0.0) : tab[O_SC + 8 + g];
            const double dg = warp == 0 ? fma(fma(nu, s2v, suma), mk, 1.0 - mk) : 0.0;
            const double c0 = (L.cdiag0 ? dg : o0) - ca0, c1 = (L.cdiag1 ? dg : o1) - ca1;
            *reinterpret_cast<double2*>(cred + (warp * 32 + lane) * 2) = make_double2(c0, c1);
        }
        __syncthreads();
        double cm0 = 0.0, cm1 = 0.0;
#pragma unroll
        for (int q = 0; q < W; q++) {
            const double2 c = *reinterpret_cast<const double2*>(cred + (q * 32 + lane) * 2);
            cm0 += c.x; cm1 += c.y;
        }

        // ---- 2a. X ← (ψ8ψ8ᵀ)∘X on my rows;  Bm = amp∘V̂ − ψ8∘P0 (the data row carries (y − μ) − prediction)
        double psr[MR];
#pragma unroll
        for (int i = 0; i < MR; i++) { const int I = warp + W * i; psr[i] = I < NTR ? tab[O_PSI + 8 * I + g] : 0.0; }
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
            if (I < NTR) {
                const int row = 8 * I + g;
                double2 vh = *reinterpret_cast<const double2*>(tab + O_VH + row * 8 + 2 * t);
                const double am = amp_s[row];
                if (I == NTR - 1) {
                    const double2 mk2 = *reinterpret_cast<const double2*>(tab + O_SC + 16 + 2 * t);
                    if (yb && row == RG) {
                        const int64_t n = n0 + 2 * t;
                        vh.x = (n < N) ? yb[n] : 0.0;
                        vh.y = (n + 1 < N) ? yb[n + 1] : 0.0;
                    }
                    const double m_ = (row == RG) ? mu : 0.0;
                    vh.x = fma(-m_, mk2.x, vh.x);
                    vh.y = fma(-m_, mk2.y, vh.y);
                }
                P0[i][0] = fma(-psr[i], P0[i][0], am * vh.x);
                P0[i][1] = fma(-psr[i], P0[i][1], am * vh.y);
            } else {
                P0[i][0] = 0.0; P0[i][1] = 0.0;
            }
        }

        // ---- 2b. 8×8 LDLᵀ of C, redundantly in every warp (identical inputs, identical operations)
        double e0 = L.cdiag0 ? 1.0 : 0.0, e1 = L.cdiag1 ? 1.0 : 0.0;
        double rd0 = 0.0, rd1 = 0.0;
        const int rowbase = lane & ~3;
        const int ring = (int)(n0 & 31);
#pragma unroll
        for (int j = 0; j < BLK; j++) {
            const int tj = j >> 1;
            const double cme = (j & 1) ? cm1 : cm0;
            const double dj = __shfl_sync(FULL, cme, 4 * j + tj);          // pivot D_{n0+j} (celerite_solver.jl:92)
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
            if (j == 0 && n0 == 0) dfirst = dj;                            // celerite_solver.jl:126 (no abs on the first pivot)
            else if (lane == ring + j) dkeep = dj;
        }
        if (ring == 24) { logacc += log(fabs(dkeep)); dkeep = 1.0; }       // one log per 32 steps (kept by every warp; warp 0 reports)

        // ---- 2c. Q̂ = Bm·L⁻ᵀ of my rows; χ² on the data row; Ŵ published lane for lane
        double Q[MR][2];
#pragma unroll
        for (int i = 0; i < MR; i++) {
            const int I = warp + W * i;
            Q[i][0] = Q[i][1] = 0.0;
            dmma(Q[i][0], Q[i][1], P0[i][0], e0);
            dmma(Q[i][0], Q[i][1], P0[i][1], e1);
            if (I == NTR - 1) {
                const bool isrg = (8 * I + g == RG);
                const double zz = fma(Q[i][0] * rd0, Q[i][0], (Q[i][1] * rd1) * Q[i][1]);     // Σ z²/D (celerite_solver.jl:333)
                chi2 += isrg ? zz : 0.0;
            }
            if (I < NT) *reinterpret_cast<double2*>(wpub + (I * 32 + lane) * 2) = make_double2(Q[i][0] * rd0, Q[i][1] * rd1);
        }
        __syncthreads();
        // the table of this block is no longer read: refill its stage
        if (threadIdx.x == 0 && b + BLKW_NSTAGE < nblocks) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&bars[sidx], STAGE_BYTES);
            tma_load_1d(stages + sidx * BD, wk.table + (size_t)(b + BLKW_NSTAGE) * BD, STAGE_BYTES, &bars[sidx]);
        }
        // ---- 3. X ← X + Q̂·Ŵᵀ on my rows, all column tiles
#pragma unroll
        for (int K = 0; K < NT; K++) {
            const double2 wv = *reinterpret_cast<const double2*>(wpub + (K * 32 + lane) * 2);
#pragma unroll
            for (int i = 0; i < MR; i++) {
                dmma(x[i][K][0], x[i][K][1], Q[i][0], wv.x);
                dmma(x[i][K][0], x[i][K][1], Q[i][1], wv.y);
            }
        }
        if (++sidx == BLKW_NSTAGE) { sidx = 0; parity ^= 1; }
        // (the next block's writes to cred come after this block's second barrier, its writes to wpub after its own first one)
    }
    // Σ log|D_n| (warp 0) and the χ² term (the warp that owns the data row)
    constexpr int WOWN = (NTR - 1) % W;
    __shared__ double fin[2];
    if (warp == 0) {
        double la = logacc + log(fabs(dkeep));
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) la += __shfl_xor_sync(FULL, la, sft);
        const double df = __shfl_sync(FULL, dfirst, 0);
        if (lane == 0) fin[0] = log(df) + la;
    }
    if (warp == WOWN) {
        double ch = chi2;
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) ch += __shfl_xor_sync(FULL, ch, sft);
        if (lane == 0) fin[1] = ch;
    }
    __syncthreads();
    if (threadIdx.x == 0) args.out[wk.out_begin] = -fin[0] / 2 - (double)N * 1.8378770664093453 / 2 - fin[1] / 2;   // celerite_solver.jl:333
}

}  // namespace pioran
