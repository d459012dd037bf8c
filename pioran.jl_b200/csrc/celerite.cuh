// celerite.cuh — K2: batched celerite factorisation + solve + log-determinant (FP64, sm_100a).
//
// Replaces  logl(a,b,c,d,τ,y,σ2)  = init_semi_separable! + solve_prec!  of the reference
// (src/celerite_solver.jl:12-100, 115-158, 312-334) for a batch of parameter vectors.
//
// Algorithm (forward-only fused sweep).  With A_n = Σa + ν σ²_n and the amplitude-scaled state
// T = diag(amp) S diag(amp)  (S is the reference's S_n matrix, celerite_solver.jl:69-90):
//     T   ← (φ_n φ_nᵀ) ∘ (T + q_{n-1} w_{n-1}ᵀ)            q = D·w
//     p   = T Ũ_n ;  D_n = A_n − Ũ_nᵀ p                     (celerite_solver.jl:92)
//     q_n = amp∘V_n − p ;  w_n = q_n / D_n                   (W of the paper, celerite_solver.jl:95-98)
//     g   ← φ_n ∘ (g + w_{n-1} z_{n-1}) ;  z_n = (y_n − μ) − Ũ_nᵀ g      (celerite_solver.jl:132-142)
//     logL = −½ Σ log|D_n| − ½ Σ z_n²/D_n − (N/2) log 2π     (y'K⁻¹y = Σ z_n²/D_n, celerite_solver.jl:333)
// No U/W/φ matrices are materialised and there is no backward pass.
//
// Mapping.  One warp per (parameter vector, series).  The symmetric R_pad×R_pad state (R_pad = 8·BS) is cut
// into an 8×8 grid of BS×BS blocks.  Lane l = (i = l>>2, o = l&3) keeps in registers
//     o = 1,2,3 : the full off-diagonal block (I = i, K = (i+o) mod 8);
//     o = 0     : a composite block — strict lower triangle of the diagonal block (i,i) and, in the upper
//                 triangle, its half of the distance-4 block {i, i^4};
// the 8·BS true diagonal entries live with the row owners (lane (i,o) owns rows i·BS+o and i·BS+o+4).
// Every lane therefore runs the same BS×BS instruction stream (no divergence) with 4 FP64 issue slots per
// stored entry: 1 DMUL + 1 DFMA (rank-1 update and decay) + 2 DFMA (symmetric matrix-vector product).
// The per-entry decay product φ_j φ_k is never formed: the state is stored with one factor pending,
// alternately on the row side (odd steps) and the column side (even steps), and the pending factor is folded
// into the vectors (UH = φ∘Ũ, KAP = φ_n∘φ_{n−1} in the series table).
// Cross-lane traffic per step: BS+ceil(BS/2)+ceil(BS/4) 64-bit shuffles for the matvec reduce-scatter, one
// 2-value butterfly all-reduce for (ŨᵀTŨ, Ũᵀg), and 2 shared-memory vectors (q, w) per warp.
//
// θ-independent per-step vectors (Ũ, φ∘Ũ, φφ', φ, V, y, σ²) come from the series table (table.cuh), staged
// chunk-wise into shared memory by TMA bulk copies (cp.async.bulk + mbarrier) and shared by all warps of a
// CTA; in generic mode (explicit per-θ c,d) each warp builds its own chunk with sincos/exp.
#pragma once
#include "common.cuh"

#ifndef PIORAN_ABLATE
#define PIORAN_ABLATE 0   // timing experiments only (tools/ablate.sh): non-zero values remove one cost and break the results
#endif

namespace pioran {

// ------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D TMA bulk copy global → shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr unsigned FULL = 0xffffffffu;

// Per-lane persistent state of one (θ, series) evaluation.
template <int BS>
struct LaneState {
    double M[BS][BS];   // block of the (pending-scaled) state matrix
    double sjj[2];      // true diagonal entries of the owned rows
    double g[2];        // forward-substitution vector of the owned rows
    double amp[2];      // row amplitudes of the owned rows
    double chi2, logacc, dkeep, dfirst;
};

// Lane → block mapping (constant over the sweep); offsets are in the padded vector layout (common.cuh).
struct LaneMap {
    int rowI, colA, colB;  // padded offsets of block-row I and of the column blocks of the A / B parts
    int j0, j1;            // padded offsets of the two owned rows
    int src_lane;          // lane whose column sums belong to my block-row
    int o;
    bool valid1, hi;       // second owned row exists; lane ≥ 16
    bool dzero;            // this lane's copy of the distance-4 block diagonal is the masked one
};
template <int BS>
__device__ __forceinline__ LaneMap make_lane_map(int lane) {
    constexpr int BSP = bsp_of(BS);
    LaneMap lm;
    const int i = lane >> 2, o = lane & 3;
    lm.o = o;
    lm.rowI = i * BSP;
    lm.colA = ((i + o) & 7) * BSP;
    lm.colB = (o == 0) ? ((i ^ 4) * BSP) : lm.colA;
    lm.dzero = (o == 0 && i >= 4);
    lm.src_lane = (((i - (o ? o : 4)) & 7) << 2) | o;
    lm.j0 = lm.rowI + o;
    lm.valid1 = (o + 4) < BS;
    lm.j1 = lm.valid1 ? lm.j0 + 4 : lm.j0;
    lm.hi = lane >= 16;
    return lm;
}

// BS consecutive doubles from a 16-byte aligned shared-memory address: 128-bit loads, one 64-bit tail when BS is odd.
template <int BS>
__device__ __forceinline__ void load_slice(double (&dst)[BS], const double* __restrict__ p) {
#pragma unroll
    for (int r = 0; r + 1 < BS; r += 2) {
        const double2 v = *reinterpret_cast<const double2*>(p + r);
        dst[r] = v.x; dst[r + 1] = v.y;
    }
    if (BS & 1) dst[BS - 1] = p[BS - 1];
}

// 1/x to ~1 ulp: hardware seed (MUFU.RCP64H, ≥ 20 bits) + two Newton steps.  x = 0, ±Inf, NaN give non-finite
// results, which the likelihood treats as data (a θ whose covariance is not positive definite).
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// One time step.  ODD: the step index n is odd → the state leaves the step with the row factor pending.
//   T      : this step's table record in shared memory (padded vectors)
//   qs, ws : per-warp scratch vectors (q and w of the previous step, one of them pre-multiplied by φ_n)
//   PRE    : the state is kept PRE-DECAYED — the decay factor of the coming step (κ_c of an odd step, κ_r of an even
//            step, read from the next record Tnext) is applied at the END of a step, where it is independent of the
//            reductions still in flight, instead of inside the rank-1 update at the start of the next one.  Same FP64
//            count (DMUL κ·m + DFMA q·w+m instead of DMUL q·w + DFMA κ·m+qw), shorter critical path.
//   MODE   : STEP_LOGL — likelihood only;  STEP_STORE — also write D_n, the forward z_n and W_n (the factor the posterior
//            mean needs, celerite_solver.jl:381-388) to `aux`;  STEP_SIM — yn carries a standard-normal draw q_n and the
//            step emits y_n = U_nᵀ f_n + √D_n q_n (sim, celerite_solver.jl:536-546) instead of consuming data.
enum StepMode : int { STEP_LOGL = 0, STEP_STORE = 1, STEP_SIM = 2 };
struct StepAux {
    double* W;     // STEP_STORE: [N × 8·BS] (logical rows)
    double* D;     // STEP_STORE: [N]
    double* zf;    // STEP_STORE: [N] forward-substitution z (before the backward pass)
    double* ysim;  // STEP_SIM:   [N]
};
template <int BS, bool ODD, bool PRE = false, int MODE = STEP_LOGL>
__device__ __forceinline__ void celerite_step(LaneState<BS>& st, const double* __restrict__ T, double* __restrict__ qs,
                                              double* __restrict__ ws, const LaneMap& lm, const double yn,
                                              const double s2n, const double suma, const double mu, const double nu,
                                              const int64_t n, const int lane, const double* __restrict__ Tnext = nullptr,
                                              const StepAux* aux = nullptr) {
    constexpr int RP = rps_of(BS);
    const int o = lm.o;
    // ---- row-side operands of block-row I
    double qrow[BS], urow[BS], xrow[BS];
    load_slice<BS>(qrow, qs + lm.rowI);
    load_slice<BS>(urow, T + (ODD ? F_UH : F_UT) * RP + lm.rowI);    // weight of the column partial sums
    if (ODD || !PRE) load_slice<BS>(xrow, T + (ODD ? F_PHI : F_KAP) * RP + lm.rowI);  // ODD: post-scale of row sums; EVEN: decay of row r
    double prow[BS];                                                  // EVEN: φ of my rows = post-scale of column sums
    if (!ODD) load_slice<BS>(prow, T + F_PHI * RP + lm.rowI);
    double rowpart[BS], acc[BS];
#pragma unroll
    for (int r = 0; r < BS; r++) rowpart[r] = 0.0;

    // ---- block phase: rank-1 update + decay + symmetric matvec, two columns per pass (128-bit operand loads)
    const double* wsA = ws + lm.colA;
    const double* wsB = ws + lm.colB;
    const double* uAp = T + (ODD ? F_UT : F_UH) * RP + lm.colA;
    const double* uBp = T + (ODD ? F_UT : F_UH) * RP + lm.colB;
    const double* zAp = T + F_KAP * RP + lm.colA;   // ODD only: decay of column c
    const double* zBp = T + F_KAP * RP + lm.colB;
#pragma unroll
    for (int c0 = 0; c0 < BS; c0 += 2) {
        double wA2[2], wB2[2], uA2[2], uB2[2], zA2[2] = {0.0, 0.0}, zB2[2] = {0.0, 0.0};
        if (c0 + 1 < BS) {
#if PIORAN_ABLATE == 1
            const double2 a0 = *reinterpret_cast<const double2*>(wsA + c0), a1 = a0;
            const double2 a2 = *reinterpret_cast<const double2*>(uAp + c0), a3 = a2;
#else
            const double2 a0 = *reinterpret_cast<const double2*>(wsA + c0), a1 = *reinterpret_cast<const double2*>(wsB + c0);
            const double2 a2 = *reinterpret_cast<const double2*>(uAp + c0), a3 = *reinterpret_cast<const double2*>(uBp + c0);
#endif
            wA2[0] = a0.x; wA2[1] = a0.y; wB2[0] = a1.x; wB2[1] = a1.y;
            uA2[0] = a2.x; uA2[1] = a2.y; uB2[0] = a3.x; uB2[1] = a3.y;
            if (ODD && !PRE) {
                const double2 a4 = *reinterpret_cast<const double2*>(zAp + c0), a5 = *reinterpret_cast<const double2*>(zBp + c0);
                zA2[0] = a4.x; zA2[1] = a4.y; zB2[0] = a5.x; zB2[1] = a5.y;
            }
        } else {
            wA2[0] = wsA[c0]; wB2[0] = wsB[c0]; uA2[0] = uAp[c0]; uB2[0] = uBp[c0];
            if (ODD && !PRE) { zA2[0] = zAp[c0]; zB2[0] = zBp[c0]; }
            wA2[1] = wB2[1] = uA2[1] = uB2[1] = 0.0;
        }
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
            const int c = c0 + cc;
            if (c < BS) {
                const double wA = wA2[cc], wB = wB2[cc], uA = uA2[cc], uB = uB2[cc], zA = zA2[cc], zB = zB2[cc];
                double cA = 0.0, cB = 0.0;
#pragma unroll
                for (int r = 0; r < BS; r++) {
                    const bool useA = r > c;  // compile-time after unrolling
                    const double w = useA ? wA : wB;
                    const double u = useA ? uA : uB;
                    const double qr = (r == c) ? (lm.dzero ? 0.0 : qrow[r]) : qrow[r];
                    double m;
                    if (PRE)      m = fma(qr, w, st.M[r][c]);                   // the decay is already in M
                    else if (ODD) m = fma(useA ? zA : zB, st.M[r][c], qr * w);  // κ_c·M + q_r·(φ_c w_c)
                    else          m = fma(xrow[r], st.M[r][c], qr * w);         // κ_r·M + (φ_r q_r)·w_c
                    st.M[r][c] = m;
                    rowpart[r] = fma(m, u, rowpart[r]);
                    if (useA) cA = fma(m, urow[r], cA);
                    else      cB = fma(m, urow[r], cB);
                }
                // column sums go to the lane that owns block-row K; the diagonal-block part (o = 0) stays here.
                // EVEN steps: the column factor φ_c is still pending — the receiver applies it (prow) in `tot`.
#if PIORAN_ABLATE == 3
                acc[c] = cB + cA;
#else
                const double yv = __shfl_sync(FULL, cB + (o ? cA : 0.0), lm.src_lane);
                acc[c] = yv + (o ? 0.0 : cA);
#endif
            }
        }
    }

    // ---- ŨᵀTŨ from the block-level partial sums (each stored entry is off-diagonal and counts twice), so the
    //      all-reduce below does not wait for the reduce-scatter of p
    const double ut0 = T[F_UT * RP + lm.j0], ut1 = T[F_UT * RP + lm.j1];
    double sblk = 0.0, sblk2 = 0.0;   // two chains: the sum is on the critical path to D_n
#pragma unroll
    for (int r = 0; r < BS; r += 2) {
        sblk = fma(urow[r], rowpart[r], sblk);
        if (r + 1 < BS) sblk2 = fma(urow[r + 1], rowpart[r + 1], sblk2);
    }
    sblk += sblk2;
    double spart = fma(st.sjj[1] * ut1, ut1, fma(st.sjj[0] * ut0, ut0, sblk + sblk));
    double upart = fma(ut1, st.g[1], ut0 * st.g[0]);
    // all-reduce of (spart, upart) in 6 exchanges: halves swap roles first, so each half reduces one value
#if PIORAN_ABLATE != 2
    {
        const double recv = __shfl_xor_sync(FULL, lm.hi ? spart : upart, 16);
        double keep = (lm.hi ? upart : spart) + recv;
#pragma unroll
        for (int sft = 8; sft >= 1; sft >>= 1) keep += __shfl_xor_sync(FULL, keep, sft);
        const double other = __shfl_xor_sync(FULL, keep, 16);
        spart = lm.hi ? other : keep;
        upart = lm.hi ? keep : other;
    }
#endif

    // ---- PRE: decay of the coming step, independent of everything above that is still in flight
    if (PRE) {
        if (!ODD) {   // next step is odd: column factors κ_c(n+1)
            const double* nA = Tnext + F_KAP * RP + lm.colA;
            const double* nB = Tnext + F_KAP * RP + lm.colB;
            double zA[BS], zB[BS];
            load_slice<BS>(zA, nA);
#if PIORAN_ABLATE == 1
#pragma unroll
            for (int c = 0; c < BS; c++) zB[c] = zA[c];
#else
            load_slice<BS>(zB, nB);
#endif
#pragma unroll
            for (int c = 0; c < BS; c++)
#pragma unroll
                for (int r = 0; r < BS; r++) st.M[r][c] *= (r > c) ? zA[c] : zB[c];
        } else {      // next step is even: row factors κ_r(n+1)
            double xn[BS];
            load_slice<BS>(xn, Tnext + F_KAP * RP + lm.rowI);
#pragma unroll
            for (int r = 0; r < BS; r++)
#pragma unroll
                for (int c = 0; c < BS; c++) st.M[r][c] *= xn[r];
        }
    }

    // ---- matvec reduction: reduce-scatter of the BS row sums over the 4 lanes of the block-row
    double tot[BS];
#pragma unroll
    for (int c = 0; c < BS; c++) {
        if (ODD) tot[c] = fma(xrow[c], rowpart[c], acc[c]);
        else     tot[c] = fma(prow[c], acc[c], rowpart[c]);
    }
#if PIORAN_ABLATE == 4
    double f0 = 0.0, f1 = 0.0;
#pragma unroll
    for (int c = 0; c < BS; c++) { if (c & 1) f1 += tot[c]; else f0 += tot[c]; }
#else
    const bool bit0 = (o & 1) != 0, bit1 = (o & 2) != 0;
    double e[4];
#pragma unroll
    for (int m = 0; m < 4; m++) {
        if (2 * m < BS) {
            const double lo = tot[(2 * m < BS) ? 2 * m : 0];
            const double hi = (2 * m + 1 < BS) ? tot[(2 * m + 1 < BS) ? 2 * m + 1 : 0] : 0.0;
            const double recv = __shfl_xor_sync(FULL, bit0 ? lo : hi, 1);
            e[m] = (bit0 ? hi : lo) + recv;
        } else {
            e[m] = 0.0;
        }
    }
    double f0, f1 = 0.0;
    {
        const double recv = __shfl_xor_sync(FULL, bit1 ? e[0] : e[1], 2);
        f0 = (bit1 ? e[1] : e[0]) + recv;
    }
    if (BS > 4) {
        const double recv = __shfl_xor_sync(FULL, bit1 ? e[2] : e[3], 2);
        f1 = (bit1 ? e[3] : e[2]) + recv;
    }
#endif

    // ---- owner phase: rows j0 (and j1)
    const double v0 = T[F_V * RP + lm.j0], v1 = T[F_V * RP + lm.j1];
    const double pn0 = T[F_PHN * RP + lm.j0], pn1 = T[F_PHN * RP + lm.j1];
    const double p0 = fma(st.sjj[0], ut0, f0);
    const double p1 = fma(st.sjj[1], ut1, f1);
    const double An = fma(nu, s2n, suma);
    const double D = An - spart;              // celerite_solver.jl:92
    const double rD = fast_rcp(D);
    double z = (yn - mu) - upart;             // celerite_solver.jl:141
    if (MODE == STEP_SIM) {                   // celerite_solver.jl:539-545: the draw enters where the innovation would
        z = sqrt(D) * yn;
        if (lane == 0) aux->ysim[n] = upart + z;
    }
    st.chi2 = fma(z * z, rD, st.chi2);
    // log|D_n| (celerite_solver.jl:140; no abs on the first pivot, :126): lane n%32 keeps D_n, one log per 32 steps
    if (n == 0) st.dfirst = D;
    else if ((int)(n & 31) == lane) st.dkeep = D;
    if ((n & 31) == 31) { st.logacc += log(fabs(st.dkeep)); st.dkeep = 1.0; }

    const double q0 = fma(st.amp[0], v0, -p0), q1 = fma(st.amp[1], v1, -p1);
    const double w0 = q0 * rD, w1 = q1 * rD;
    if (MODE == STEP_STORE) {
        const int row0 = (lane >> 2) * BS + o;
        aux->W[n * (G * BS) + row0] = w0;
        if (lm.valid1) aux->W[n * (G * BS) + row0 + 4] = w1;
        if (lane == 0) { aux->D[n] = D; aux->zf[n] = z; }
    }
    st.g[0] = pn0 * fma(w0, z, st.g[0]);
    st.g[1] = pn1 * fma(w1, z, st.g[1]);
    st.sjj[0] = (pn0 * pn0) * fma(q0, w0, st.sjj[0]);   // celerite_solver.jl:85
    st.sjj[1] = (pn1 * pn1) * fma(q1, w1, st.sjj[1]);
    __syncwarp();
    // the NEXT step is odd iff this one is even: odd steps consume (q, φ∘w), even steps (φ∘q, w)
    if (!ODD) {
        qs[lm.j0] = q0; ws[lm.j0] = pn0 * w0;
        if (lm.valid1) { qs[lm.j1] = q1; ws[lm.j1] = pn1 * w1; }
    } else {
        qs[lm.j0] = pn0 * q0; ws[lm.j0] = w0;
        if (lm.valid1) { qs[lm.j1] = pn1 * q1; ws[lm.j1] = w1; }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------- paired step
// Two parameter vectors per warp (block sizes ≤ 5, where two lane states fit the register budget): the θ-independent
// operands of a step — Ũ, φ∘Ũ, κ, φ slices and the owner-phase scalars — are loaded from shared memory ONCE and used for
// both states, which removes about a quarter of the LSU wavefronts per evaluation (the co-limiter of the small-rank
// kernel, profiles/r01_k2_sho_*), and the two independent reduction chains interleave.  Same arithmetic per θ as
// celerite_step<BS, ODD, /*PRE=*/true>; the two (ŨᵀTŨ, Ũᵀg) pairs share one 4-value all-reduce (9 exchanges, not 12).
template <int BS, bool ODD>
__device__ __forceinline__ void celerite_step_pair(LaneState<BS> (&st)[2], const double* __restrict__ T,
                                                   double* __restrict__ sc0, double* __restrict__ sc1, const LaneMap& lm,
                                                   const double (&yn)[2], const double (&s2n)[2], const double (&suma)[2],
                                                   const double (&mu)[2], const double (&nu)[2], const int64_t n,
                                                   const int lane, const double* __restrict__ Tnext) {
    constexpr int RP = rps_of(BS);
    const int o = lm.o;
    double* const qsv[2] = {sc0, sc1};
    // ---- row-side operands of block-row I (shared by both states except q)
    double urow[BS], xrow[BS], prow[BS], qrow[2][BS];
    load_slice<BS>(urow, T + (ODD ? F_UH : F_UT) * RP + lm.rowI);
    if (ODD) load_slice<BS>(xrow, T + F_PHI * RP + lm.rowI);
    else     load_slice<BS>(prow, T + F_PHI * RP + lm.rowI);
    load_slice<BS>(qrow[0], sc0 + lm.rowI);
    load_slice<BS>(qrow[1], sc1 + lm.rowI);
    double rowpart[2][BS], acc[2][BS];
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
        for (int r = 0; r < BS; r++) rowpart[k][r] = 0.0;

    const double* uAp = T + (ODD ? F_UT : F_UH) * RP + lm.colA;
    const double* uBp = T + (ODD ? F_UT : F_UH) * RP + lm.colB;
#pragma unroll
    for (int c0 = 0; c0 < BS; c0 += 2) {
        double uA2[2], uB2[2], wA2[2][2], wB2[2][2];
        if (c0 + 1 < BS) {
            const double2 a2 = *reinterpret_cast<const double2*>(uAp + c0), a3 = *reinterpret_cast<const double2*>(uBp + c0);
            uA2[0] = a2.x; uA2[1] = a2.y; uB2[0] = a3.x; uB2[1] = a3.y;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const double* ws = qsv[k] + RP;
                const double2 a0 = *reinterpret_cast<const double2*>(ws + lm.colA + c0), a1 = *reinterpret_cast<const double2*>(ws + lm.colB + c0);
                wA2[k][0] = a0.x; wA2[k][1] = a0.y; wB2[k][0] = a1.x; wB2[k][1] = a1.y;
            }
        } else {
            uA2[0] = uAp[c0]; uB2[0] = uBp[c0]; uA2[1] = uB2[1] = 0.0;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const double* ws = qsv[k] + RP;
                wA2[k][0] = ws[lm.colA + c0]; wB2[k][0] = ws[lm.colB + c0]; wA2[k][1] = wB2[k][1] = 0.0;
            }
        }
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
            const int c = c0 + cc;
            if (c < BS) {
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    double cA = 0.0, cB = 0.0;
#pragma unroll
                    for (int r = 0; r < BS; r++) {
                        const bool useA = r > c;  // compile-time after unrolling
                        const double w = useA ? wA2[k][cc] : wB2[k][cc];
                        const double u = useA ? uA2[cc] : uB2[cc];
                        const double qr = (r == c) ? (lm.dzero ? 0.0 : qrow[k][r]) : qrow[k][r];
                        const double m = fma(qr, w, st[k].M[r][c]);      // the decay is already in M (pre-decayed state)
                        st[k].M[r][c] = m;
                        rowpart[k][r] = fma(m, u, rowpart[k][r]);
                        if (useA) cA = fma(m, urow[r], cA);
                        else      cB = fma(m, urow[r], cB);
                    }
                    const double yv = __shfl_sync(FULL, cB + (o ? cA : 0.0), lm.src_lane);
                    acc[k][c] = yv + (o ? 0.0 : cA);
                }
            }
        }
    }

    // ---- (ŨᵀTŨ, Ũᵀg) of both states: one 4-value all-reduce
    const double ut0 = T[F_UT * RP + lm.j0], ut1 = T[F_UT * RP + lm.j1];
    double sp[2], up[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        double sblk = 0.0, sblk2 = 0.0;
#pragma unroll
        for (int r = 0; r < BS; r += 2) {
            sblk = fma(urow[r], rowpart[k][r], sblk);
            if (r + 1 < BS) sblk2 = fma(urow[r + 1], rowpart[k][r + 1], sblk2);
        }
        sblk += sblk2;
        sp[k] = fma(st[k].sjj[1] * ut1, ut1, fma(st[k].sjj[0] * ut0, ut0, sblk + sblk));
        up[k] = fma(ut1, st[k].g[1], ut0 * st[k].g[0]);
    }
    {
        // halves: lanes < 16 reduce state 0's pair, lanes ≥ 16 state 1's; quarter bit (lane & 8) picks s or u
        const bool q8 = (lane & 8) != 0;
        const double r0 = __shfl_xor_sync(FULL, lm.hi ? sp[0] : sp[1], 16);
        const double r1 = __shfl_xor_sync(FULL, lm.hi ? up[0] : up[1], 16);
        const double ks = (lm.hi ? sp[1] : sp[0]) + r0, ku = (lm.hi ? up[1] : up[0]) + r1;
        const double r2 = __shfl_xor_sync(FULL, q8 ? ks : ku, 8);
        double keep = (q8 ? ku : ks) + r2;
#pragma unroll
        for (int sft = 4; sft >= 1; sft >>= 1) keep += __shfl_xor_sync(FULL, keep, sft);
        const double oth = __shfl_xor_sync(FULL, keep, 8);
        const double ms = q8 ? oth : keep, mu_ = q8 ? keep : oth;     // (s, u) of my half's state
        const double os = __shfl_xor_sync(FULL, ms, 16), ou = __shfl_xor_sync(FULL, mu_, 16);
        sp[0] = lm.hi ? os : ms; up[0] = lm.hi ? ou : mu_;
        sp[1] = lm.hi ? ms : os; up[1] = lm.hi ? mu_ : ou;
    }

    // ---- decay of the coming step (pre-decayed state), factors loaded once for both states
    if (!ODD) {   // next step is odd: column factors κ_c(n+1)
        double zA[BS], zB[BS];
        load_slice<BS>(zA, Tnext + F_KAP * RP + lm.colA);
        load_slice<BS>(zB, Tnext + F_KAP * RP + lm.colB);
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int c = 0; c < BS; c++)
#pragma unroll
                for (int r = 0; r < BS; r++) st[k].M[r][c] *= (r > c) ? zA[c] : zB[c];
    } else {      // next step is even: row factors κ_r(n+1)
        double xn[BS];
        load_slice<BS>(xn, Tnext + F_KAP * RP + lm.rowI);
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int r = 0; r < BS; r++)
#pragma unroll
                for (int c = 0; c < BS; c++) st[k].M[r][c] *= xn[r];
    }

    // ---- matvec reduction (per state) and owner phase; the table scalars of the owned rows are shared
    const double v0 = T[F_V * RP + lm.j0], v1 = T[F_V * RP + lm.j1];
    const double pn0 = T[F_PHN * RP + lm.j0], pn1 = T[F_PHN * RP + lm.j1];
    const bool bit0 = (o & 1) != 0, bit1 = (o & 2) != 0;
    double q0[2], q1[2], w0[2], w1[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        double tot[BS];
#pragma unroll
        for (int c = 0; c < BS; c++) {
            if (ODD) tot[c] = fma(xrow[c], rowpart[k][c], acc[k][c]);
            else     tot[c] = fma(prow[c], acc[k][c], rowpart[k][c]);
        }
        double e[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
            if (2 * m < BS) {
                const double lo = tot[(2 * m < BS) ? 2 * m : 0];
                const double hi = (2 * m + 1 < BS) ? tot[(2 * m + 1 < BS) ? 2 * m + 1 : 0] : 0.0;
                const double recv = __shfl_xor_sync(FULL, bit0 ? lo : hi, 1);
                e[m] = (bit0 ? hi : lo) + recv;
            } else {
                e[m] = 0.0;
            }
        }
        double f0, f1 = 0.0;
        {
            const double recv = __shfl_xor_sync(FULL, bit1 ? e[0] : e[1], 2);
            f0 = (bit1 ? e[1] : e[0]) + recv;
        }
        if (BS > 4) {
            const double recv = __shfl_xor_sync(FULL, bit1 ? e[2] : e[3], 2);
            f1 = (bit1 ? e[3] : e[2]) + recv;
        }
        LaneState<BS>& s_ = st[k];
        const double p0 = fma(s_.sjj[0], ut0, f0), p1 = fma(s_.sjj[1], ut1, f1);
        const double D = fma(nu[k], s2n[k], suma[k]) - sp[k];     // celerite_solver.jl:92
        const double rD = fast_rcp(D);
        const double z = (yn[k] - mu[k]) - up[k];                 // celerite_solver.jl:141
        s_.chi2 = fma(z * z, rD, s_.chi2);
        if (n == 0) s_.dfirst = D;
        else if ((int)(n & 31) == lane) s_.dkeep = D;
        if ((n & 31) == 31) { s_.logacc += log(fabs(s_.dkeep)); s_.dkeep = 1.0; }
        q0[k] = fma(s_.amp[0], v0, -p0); q1[k] = fma(s_.amp[1], v1, -p1);
        w0[k] = q0[k] * rD; w1[k] = q1[k] * rD;
        s_.g[0] = pn0 * fma(w0[k], z, s_.g[0]);
        s_.g[1] = pn1 * fma(w1[k], z, s_.g[1]);
        s_.sjj[0] = (pn0 * pn0) * fma(q0[k], w0[k], s_.sjj[0]);   // celerite_solver.jl:85
        s_.sjj[1] = (pn1 * pn1) * fma(q1[k], w1[k], s_.sjj[1]);
    }
    __syncwarp();
    // the NEXT step is odd iff this one is even: odd steps consume (q, φ∘w), even steps (φ∘q, w)
#pragma unroll
    for (int k = 0; k < 2; k++) {
        double* qs = qsv[k];
        double* ws = qsv[k] + RP;
        if (!ODD) {
            qs[lm.j0] = q0[k]; ws[lm.j0] = pn0 * w0[k];
            if (lm.valid1) { qs[lm.j1] = q1[k]; ws[lm.j1] = pn1 * w1[k]; }
        } else {
            qs[lm.j0] = pn0 * q0[k]; ws[lm.j0] = w0[k];
            if (lm.valid1) { qs[lm.j1] = pn1 * q1[k]; ws[lm.j1] = w1[k]; }
        }
    }
    __syncwarp();
}

// Per-θ inputs of the batched kernel.
struct BatchArgs {
    const WorkItem* work;
    // shared-table mode: row amplitudes [nθ × R_pad] and Σa [nθ]
    const double* amp;
    const double* suma;
    // generic mode: celerite coefficients [nθ × Jt]
    const double* a; const double* b; const double* c; const double* d;
    int Jt;
    const int* term_row;    // generic mode: first row of term m (≥0: complex, 2 rows; <0: real term at row −v−1)
    int R;                  // generic mode: number of live rows
    const double* mu;       // [nθ] (stride pstride) or nullptr
    const double* nu;       // [nθ] (stride pstride) or nullptr
    int pstride;
    const double* y_batch;  // [nθ × ystride] or nullptr
    const double* s2_batch; // [nθ × ystride] or nullptr
    int64_t ystride;
    double* out;            // logL
    // generic kernel, STEP_STORE: per-θ factor [nθ × N × 8·BS], [nθ × N], [nθ × N];  STEP_SIM: draws come in through y_batch,
    // the realisation goes to ysim [nθ × N]
    double* W_out; double* D_out; double* zf_out; double* ysim_out;
};

template <int BS>
__device__ __forceinline__ void lane_init(LaneState<BS>& st) {
#pragma unroll
    for (int r = 0; r < BS; r++)
#pragma unroll
        for (int c = 0; c < BS; c++) st.M[r][c] = 0.0;
    st.sjj[0] = st.sjj[1] = 0.0;
    st.g[0] = st.g[1] = 0.0;
    st.chi2 = 0.0; st.logacc = 0.0; st.dkeep = 1.0; st.dfirst = 1.0;
}

// Σ log|D_n| over the steps this warp swept (log D_1 without abs, celerite_solver.jl:126).
template <int BS>
__device__ __forceinline__ double lane_logdet(LaneState<BS>& st) {
    // flush the log|D| ring: lanes without a pending pivot hold 1.0 (log 1 = 0)
    double la = st.logacc + log(fabs(st.dkeep));
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) la += __shfl_xor_sync(FULL, la, sft);
    return log(st.dfirst) + la;
}
template <int BS>
__device__ __forceinline__ double lane_finish(LaneState<BS>& st, int64_t N, int lane) {
    const double logdet = lane_logdet(st);
    // celerite_solver.jl:333
    return -logdet / 2 - (double)N * 1.8378770664093453 / 2 - st.chi2 / 2;
}

// ------------------------------------------------------------------------------------------- shared-table kernel
// grid = number of work items; block = NW warps; each warp one θ of the item's series.
// dynamic smem: 2 stages × CHUNK_STEPS × SD doubles | NW × 2·RPS scratch | 2 mbarriers | 2 stage counters
// Every warp runs the full sweep (warps beyond the item's count redo its last θ and skip the store): the hot loop
// then has no lane-dependent control flow and the shuffles compile to plain SHFL (no WARPSYNC.COLLECTIVE).
template <int BS, int NW>
__global__ void __launch_bounds__(NW * 32, 1) celerite_shared_kernel(const BatchArgs args) {
    constexpr int RP = G * BS, RPS = rps_of(BS), SD = table_step_doubles(RPS), STAGE = CHUNK_STEPS * SD;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    double* scratch = stages + 2 * STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + NW * 2 * RPS);
    int* done = reinterpret_cast<int*>(bars + 2);     // per stage: number of warps that have finished reading it

    const WorkItem wk = args.work[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t N = wk.N;
    const int64_t nchunks = (N + CHUNK_STEPS - 1) / CHUNK_STEPS;
    constexpr uint32_t STAGE_BYTES = STAGE * sizeof(double);

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        done[0] = done[1] = 0;
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 2 && k < nchunks; k++) {
            mbar_arrive_expect_tx(&bars[k], STAGE_BYTES);
            tma_load_1d(stages + k * STAGE, wk.table + (size_t)k * STAGE, STAGE_BYTES, &bars[k]);
        }
    }

    const bool active = warp < wk.count;
    const int slot = active ? warp : wk.count - 1;
    const int th = wk.theta_begin + slot;
    const LaneMap lm = make_lane_map<BS>(lane);
    const int i = lane >> 2, o = lane & 3;

    LaneState<BS> st;
    lane_init(st);
    st.amp[0] = args.amp[(size_t)th * RP + i * BS + o];
    st.amp[1] = lm.valid1 ? args.amp[(size_t)th * RP + i * BS + o + 4] : 0.0;
    const double suma = args.suma[th];
    const size_t pi = (size_t)wk.par_begin + slot;
    const double mu = args.mu ? args.mu[pi * args.pstride] : 0.0;
    const double nu = args.nu ? args.nu[pi * args.pstride] : 1.0;
    const double* yb = args.y_batch ? args.y_batch + pi * args.ystride : nullptr;
    const double* sb = args.s2_batch ? args.s2_batch + pi * args.ystride : nullptr;

    double* qs = scratch + warp * 2 * RPS;
    double* ws = qs + RPS;
    for (int k = lane; k < 2 * RPS; k += 32) qs[k] = 0.0;
    __syncwarp();

    mbar_wait(&bars[0], 0);
    for (int64_t k = 0; k < nchunks; k++) {
        const int sidx = (int)(k & 1);
        const double* stage = stages + sidx * STAGE;
        const double* nstage = stages + (sidx ^ 1) * STAGE;     // chunk k+1 (read by the last step's look-ahead)
        const bool has_next = k + 1 < nchunks;
        const int64_t nbeg = k * CHUNK_STEPS;
        const int nsteps = (int)((N - nbeg) < CHUNK_STEPS ? (N - nbeg) : CHUNK_STEPS);
        for (int s = 0; s < nsteps; s += 2) {
            const double* T0 = stage + s * SD;
            const double* T1 = T0 + SD;
            const int64_t n = nbeg + s;
            // chunk k+1 was requested a whole chunk ago; its arrival is checked right before the first read
            if (has_next && s + 2 >= nsteps) mbar_wait(&bars[sidx ^ 1], (uint32_t)(((k + 1) >> 1) & 1));
            double yn = yb ? yb[n] : T0[6 * RPS + 0];
            double s2n = sb ? sb[n] : T0[6 * RPS + 1];
            // look-ahead record: the next step's, or (end of the series) this one — its factors are then never used
            const double* Ta = (s + 1 < nsteps) ? T1 : (has_next ? nstage : T0);
            celerite_step<BS, false, true>(st, T0, qs, ws, lm, yn, s2n, suma, mu, nu, n, lane, Ta);
            if (s + 1 < nsteps) {
                yn = yb ? yb[n + 1] : T1[6 * RPS + 0];
                s2n = sb ? sb[n + 1] : T1[6 * RPS + 1];
                const double* Tb = (s + 2 < nsteps) ? T1 + SD : (has_next ? nstage : T1);
                celerite_step<BS, true, true>(st, T1, qs, ws, lm, yn, s2n, suma, mu, nu, n + 1, lane, Tb);
            }
        }
        // Stage hand-back without a CTA barrier: every warp counts itself out of the stage and the LAST one refills it with
        // chunk k+2, so no warp waits for its siblings unless it is a whole chunk ahead of them.
        __syncwarp();
        if (lane == 0 && k + 2 < nchunks) {
            __threadfence_block();
            if (atomicAdd(&done[sidx], 1) == NW - 1) {
                done[sidx] = 0;
                fence_proxy_async();
                mbar_arrive_expect_tx(&bars[sidx], STAGE_BYTES);
                tma_load_1d(stages + sidx * STAGE, wk.table + (size_t)(k + 2) * STAGE, STAGE_BYTES, &bars[sidx]);
            }
        }
    }
    const double res = lane_finish(st, N, lane);
    if (active && lane == 0) args.out[wk.out_begin + warp] = res;
}

// Paired variant of the shared-table kernel: each warp sweeps TWO parameter vectors of the item's series (block sizes ≤ 5).
// grid = number of work items (≤ 2·NW parameter vectors each); same TMA staging and stage hand-back as above.
template <int BS, int NW>
__global__ void __launch_bounds__(NW * 32, 1) celerite_shared_pair_kernel(const BatchArgs args) {
    constexpr int RP = G * BS, RPS = rps_of(BS), SD = table_step_doubles(RPS), STAGE = CHUNK_STEPS * SD;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    double* scratch = stages + 2 * STAGE;                      // per warp: (q, w) of both states, 4·RPS doubles
    uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + NW * 4 * RPS);
    int* done = reinterpret_cast<int*>(bars + 2);

    const WorkItem wk = args.work[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t N = wk.N;
    const int64_t nchunks = (N + CHUNK_STEPS - 1) / CHUNK_STEPS;
    constexpr uint32_t STAGE_BYTES = STAGE * sizeof(double);

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        done[0] = done[1] = 0;
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 2 && k < nchunks; k++) {
            mbar_arrive_expect_tx(&bars[k], STAGE_BYTES);
            tma_load_1d(stages + k * STAGE, wk.table + (size_t)k * STAGE, STAGE_BYTES, &bars[k]);
        }
    }

    const LaneMap lm = make_lane_map<BS>(lane);
    const int i = lane >> 2, o = lane & 3;
    LaneState<BS> st[2];
    bool active[2];
    double suma[2], mu[2], nu[2];
    const double* yb[2];
    const double* sb[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int want = 2 * warp + k;
        active[k] = want < wk.count;
        const int slot = active[k] ? want : wk.count - 1;     // surplus slots redo the item's last θ and skip the store
        const int th = wk.theta_begin + slot;
        lane_init(st[k]);
        st[k].amp[0] = args.amp[(size_t)th * RP + i * BS + o];
        st[k].amp[1] = lm.valid1 ? args.amp[(size_t)th * RP + i * BS + o + 4] : 0.0;
        suma[k] = args.suma[th];
        const size_t pi = (size_t)wk.par_begin + slot;
        mu[k] = args.mu ? args.mu[pi * args.pstride] : 0.0;
        nu[k] = args.nu ? args.nu[pi * args.pstride] : 1.0;
        yb[k] = args.y_batch ? args.y_batch + pi * args.ystride : nullptr;
        sb[k] = args.s2_batch ? args.s2_batch + pi * args.ystride : nullptr;
    }
    double* sc0 = scratch + warp * 4 * RPS;
    double* sc1 = sc0 + 2 * RPS;
    for (int k = lane; k < 4 * RPS; k += 32) sc0[k] = 0.0;
    __syncwarp();

    mbar_wait(&bars[0], 0);
    for (int64_t k = 0; k < nchunks; k++) {
        const int sidx = (int)(k & 1);
        const double* stage = stages + sidx * STAGE;
        const double* nstage = stages + (sidx ^ 1) * STAGE;
        const bool has_next = k + 1 < nchunks;
        const int64_t nbeg = k * CHUNK_STEPS;
        const int nsteps = (int)((N - nbeg) < CHUNK_STEPS ? (N - nbeg) : CHUNK_STEPS);
        for (int s = 0; s < nsteps; s += 2) {
            const double* T0 = stage + s * SD;
            const double* T1 = T0 + SD;
            const int64_t n = nbeg + s;
            if (has_next && s + 2 >= nsteps) mbar_wait(&bars[sidx ^ 1], (uint32_t)(((k + 1) >> 1) & 1));
            double yn[2], s2n[2];
#pragma unroll
            for (int q = 0; q < 2; q++) { yn[q] = yb[q] ? yb[q][n] : T0[6 * RPS + 0]; s2n[q] = sb[q] ? sb[q][n] : T0[6 * RPS + 1]; }
            const double* Ta = (s + 1 < nsteps) ? T1 : (has_next ? nstage : T0);
            celerite_step_pair<BS, false>(st, T0, sc0, sc1, lm, yn, s2n, suma, mu, nu, n, lane, Ta);
            if (s + 1 < nsteps) {
#pragma unroll
                for (int q = 0; q < 2; q++) { yn[q] = yb[q] ? yb[q][n + 1] : T1[6 * RPS + 0]; s2n[q] = sb[q] ? sb[q][n + 1] : T1[6 * RPS + 1]; }
                const double* Tb = (s + 2 < nsteps) ? T1 + SD : (has_next ? nstage : T1);
                celerite_step_pair<BS, true>(st, T1, sc0, sc1, lm, yn, s2n, suma, mu, nu, n + 1, lane, Tb);
            }
        }
        __syncwarp();
        if (lane == 0 && k + 2 < nchunks) {
            __threadfence_block();
            if (atomicAdd(&done[sidx], 1) == NW - 1) {
                done[sidx] = 0;
                fence_proxy_async();
                mbar_arrive_expect_tx(&bars[sidx], STAGE_BYTES);
                tma_load_1d(stages + sidx * STAGE, wk.table + (size_t)(k + 2) * STAGE, STAGE_BYTES, &bars[sidx]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const double res = lane_finish(st[k], N, lane);
        if (active[k] && lane == 0) args.out[wk.out_begin + 2 * warp + k] = res;
    }
}

// ------------------------------------------------------------------------------------------- generic kernel
// Explicit per-θ (a,b,c,d): every term gets a cos-row and a sin-row (R = 2·Jt, like the reference), the warp
// computes its own table chunk (GCH steps) with sincos/exp — celerite_solver.jl:51-64 — then runs the same steps.
constexpr int GCH = 4;

// CHUNKED = true is the K3 pass-3 variant: one work item per warp, a step range and an injected initial state.
template <int BS, int NW, bool CHUNKED = false, int MODE = STEP_LOGL>
__global__ void __launch_bounds__(NW * 32, 1) celerite_generic_kernel(const BatchArgs args) {
    constexpr int RPS = rps_of(BS), SD = table_step_doubles(RPS);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // per warp: GCH·SD table | 2·RPS scratch | (GCH+2)·Jt φ side array (rounded up to an even count: 16-byte alignment)
    const int Jt = args.Jt;
    const int per_warp = GCH * SD + 2 * RPS + (((GCH + 2) * Jt + 1) & ~1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* base = reinterpret_cast<double*>(smem_raw) + (size_t)warp * per_warp;
    double* tab = base;
    double* qs = base + GCH * SD;
    double* ws = qs + RPS;
    double* phs = ws + RPS;

    const WorkItem wk = args.work[CHUNKED ? blockIdx.x * NW + warp : blockIdx.x];
    const bool active = CHUNKED ? true : warp < wk.count;
    const int slot = CHUNKED ? 0 : (active ? warp : wk.count - 1);
    const int th = wk.theta_begin + slot;
    const int64_t N = wk.N;
    const int64_t n0 = CHUNKED ? wk.n_begin : 0, n1 = CHUNKED ? wk.n_end : N;
    const double* const init = CHUNKED ? wk.init : nullptr;
    const double* ca = args.a + (size_t)th * Jt;
    const double* cb = args.b + (size_t)th * Jt;
    const double* cc = args.c + (size_t)th * Jt;
    const double* cd = args.d + (size_t)th * Jt;

    const LaneMap lm = make_lane_map<BS>(lane);
    const int i = lane >> 2, o = lane & 3;
    const int R = args.R;
    const int* term_row = args.term_row;

    LaneState<BS> st;
    lane_init(st);
    st.amp[0] = (i * BS + o < R) ? 1.0 : 0.0;
    st.amp[1] = (lm.valid1 && i * BS + o + 4 < R) ? 1.0 : 0.0;
    double suma = 0.0;
    for (int m = 0; m < Jt; m++) suma += ca[m];   // celerite_solver.jl:21
    const size_t pi = (size_t)wk.par_begin + slot;
    const double mu = args.mu ? args.mu[pi * args.pstride] : 0.0;
    const double nu = args.nu ? args.nu[pi * args.pstride] : 1.0;
    const double* yb = args.y_batch ? args.y_batch + pi * args.ystride : wk.y;
    const double* sb = args.s2_batch ? args.s2_batch + pi * args.ystride : wk.s2;

    StepAux aux{nullptr, nullptr, nullptr, nullptr};
    if (MODE == STEP_STORE) {
        aux.W = args.W_out + pi * (size_t)N * (G * BS); aux.D = args.D_out + pi * (size_t)N; aux.zf = args.zf_out + pi * (size_t)N;
    }
    if (MODE == STEP_SIM) aux.ysim = args.ysim_out + pi * (size_t)N;
    for (int k = lane; k < 2 * RPS; k += 32) qs[k] = 0.0;
    // zero the table once: padded slots and rows ≥ R never change
    for (int k = lane; k < GCH * SD; k += 32) tab[k] = 0.0;
    __syncwarp();
    if (CHUNKED && init) {
        // K3 pass 3: inject the state entering step n0 (already decayed to t_n0).  The first local step is an EVEN
        // step with φ_n0 = φ_n0−1 := 1 below, so the blocks hold the true values with no factor pending.
        constexpr int RL = SCAN_LD;
        const double* S0 = init;
        const int rI = i * BS, cA = ((i + o) & 7) * BS, cB = (o == 0) ? ((i ^ 4) * BS) : cA;
#pragma unroll
        for (int r = 0; r < BS; r++)
#pragma unroll
            for (int c = 0; c < BS; c++) {
                double v = S0[(size_t)(rI + r) * RL + ((r > c) ? cA : cB) + c];
                if (r == c && lm.dzero) v = 0.0;     // the masked copy of the distance-4 diagonal
                st.M[r][c] = v;
            }
        st.sjj[0] = S0[(size_t)(rI + o) * RL + rI + o];
        st.g[0] = S0[(size_t)RL * RL + rI + o];
        if (lm.valid1) {
            st.sjj[1] = S0[(size_t)(rI + o + 4) * RL + rI + o + 4];
            st.g[1] = S0[(size_t)RL * RL + rI + o + 4];
        }
    }

    // table records of the steps [nbeg, nbeg + nsteps) in this warp's shared-memory chunk (celerite_solver.jl:51-64)
    auto build_rows = [&](const int64_t nbeg, const int nsteps) {
        // pass 1: φ for steps nbeg-1 … nbeg+GCH  (φ_0 = 0, φ_N = 0; chunk start of K3: φ := 1 up to step n0)
        for (int idx = lane; idx < (GCH + 2) * Jt; idx += 32) {
            const int s = idx / Jt, m = idx - s * Jt;
            const int64_t n = nbeg - 1 + s;
            double ph = 0.0;
            if (n >= 1 && n < N) ph = exp(-cc[m] * (wk.t[n] - wk.t[n - 1]));
            if (CHUNKED && init && n <= n0) ph = 1.0;
            phs[idx] = ph;
        }
        __syncwarp();
        // pass 2: rows
        for (int idx = lane; idx < nsteps * Jt; idx += 32) {
            const int s = idx / Jt, m = idx - s * Jt;
            const int64_t n = nbeg + s;
            const double php = phs[s * Jt + m], ph = phs[(s + 1) * Jt + m], phn = phs[(s + 2) * Jt + m];
            double* Tn = tab + s * SD;
            const int tr = term_row[m];
            if (tr < 0) {  // real term (b = d = 0): U = a, V = 1; the sin-row is identically zero and is dropped
                const int r0 = pad_index(-tr - 1, BS);
                Tn[F_UT * RPS + r0] = ca[m];       Tn[F_UH * RPS + r0] = ph * ca[m];
                Tn[F_KAP * RPS + r0] = ph * php;   Tn[F_PHI * RPS + r0] = ph;
                Tn[F_V * RPS + r0] = 1.0;          Tn[F_PHN * RPS + r0] = phn;
            } else {
                double si, co;
                sincos_large(cd[m] * wk.t[n], &si, &co);
                const double u0 = ca[m] * co + cb[m] * si;   // celerite_solver.jl:60
                const double u1 = ca[m] * si - cb[m] * co;   // celerite_solver.jl:59
                const int r0 = pad_index(tr, BS), r1 = pad_index(tr + 1, BS);
                Tn[F_UT * RPS + r0] = u0;        Tn[F_UT * RPS + r1] = u1;
                Tn[F_UH * RPS + r0] = ph * u0;   Tn[F_UH * RPS + r1] = ph * u1;
                Tn[F_KAP * RPS + r0] = ph * php; Tn[F_KAP * RPS + r1] = ph * php;
                Tn[F_PHI * RPS + r0] = ph;       Tn[F_PHI * RPS + r1] = ph;
                Tn[F_V * RPS + r0] = co;         Tn[F_V * RPS + r1] = si;
                Tn[F_PHN * RPS + r0] = phn;      Tn[F_PHN * RPS + r1] = phn;
            }
        }
        __syncwarp();
    };
    // CHUNKED: three segments — [nb, nb + n_head) | … n1) | [n1, n1 + n_ext) — with the sums taken and reset in between (every
    // segment but the last has an even length, so the even/odd alternation of the steps runs through)
    // (seg = −1: the run-up of a refinement pass, [n0, n0 + n_warm), sums discarded; the sub-chunk proper starts at nb)
    const int64_t nb = CHUNKED ? n0 + wk.n_warm : n0;
    for (int seg = (CHUNKED && wk.n_warm > 0) ? -1 : 0; seg < (CHUNKED ? 3 : 1); seg++) {
    const int64_t sbeg = !CHUNKED ? n0 : seg < 0 ? n0 : seg == 0 ? nb : seg == 1 ? nb + wk.n_head : n1;
    const int64_t send = !CHUNKED ? n1 : seg < 0 ? nb : seg == 0 ? nb + wk.n_head : seg == 1 ? n1 : n1 + wk.n_ext;
    for (int64_t nbeg = sbeg; nbeg < send; nbeg += GCH) {
        const int nsteps = (int)((send - nbeg) < GCH ? (send - nbeg) : GCH);
        build_rows(nbeg, nsteps);
        for (int s = 0; s < nsteps; s += 2) {
            const double* T0 = tab + s * SD;
            const int64_t n = nbeg + s;
            celerite_step<BS, false, false, MODE>(st, T0, qs, ws, lm, yb[n], sb[n], suma, mu, nu, n, lane, nullptr, &aux);
            if (s + 1 < nsteps)
                celerite_step<BS, true, false, MODE>(st, T0 + SD, qs, ws, lm, yb[n + 1], sb[n + 1], suma, mu, nu, n + 1, lane,
                                                     nullptr, &aux);
        }
        __syncwarp();
    }
    if (CHUNKED && seg == 1 && wk.exit && n1 < N) {
        // K3 Newton refinement: the state entering step n1.  The last step was a (local) odd one: M holds the state with its
        // row factor pending and (φ∘q, w) of that step sit in the scratch vectors — what the even step n1 would do to its
        // blocks, then the pending column factor φ_c(n1); the diagonal and g are kept fully decayed already.
        build_rows(n1, 1);
        constexpr int RL = SCAN_LD;
        double* E = wk.exit;
        const int rI = i * BS, cA = ((i + o) & 7) * BS, cB = (o == 0) ? ((i ^ 4) * BS) : cA;
#pragma unroll
        for (int r = 0; r < BS; r++)
#pragma unroll
            for (int c = 0; c < BS; c++) {
                if (r == c && lm.dzero) continue;
                const bool useA = r > c;
                const int colp = useA ? lm.colA : lm.colB, col = useA ? cA : cB;
                const double m = fma(tab[F_KAP * RPS + lm.rowI + r], st.M[r][c], qs[lm.rowI + r] * ws[colp + c]);
                const double v = tab[F_PHI * RPS + colp + c] * m;
                E[(size_t)(rI + r) * RL + col + c] = v;
                E[(size_t)(col + c) * RL + rI + r] = v;
            }
        E[(size_t)(rI + o) * RL + rI + o] = st.sjj[0];
        E[(size_t)RL * RL + rI + o] = st.g[0];
        if (lm.valid1) {
            E[(size_t)(rI + o + 4) * RL + rI + o + 4] = st.sjj[1];
            E[(size_t)RL * RL + rI + o + 4] = st.g[1];
        }
        __syncwarp();
    }
    if (CHUNKED) {
        const double logdet = lane_logdet(st);
        if (lane == 0 && seg >= 0) {
            if (seg == 0) { wk.chk[0] = logdet; wk.chk[1] = st.chi2; }
            else if (seg == 1) { wk.part[0] = wk.chk[0] + logdet; wk.part[1] = wk.chk[1] + st.chi2; }
            else { wk.chk[2] = logdet; wk.chk[3] = st.chi2; }
        }
        st.chi2 = 0.0; st.logacc = 0.0; st.dkeep = 1.0; st.dfirst = 1.0;
    }
    }
    if (!CHUNKED) {
        const double res = lane_finish(st, N, lane);
        if (active && lane == 0) args.out[wk.out_begin + warp] = res;
    }
}

}  // namespace pioran
