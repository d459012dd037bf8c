// grad_pipe.cuh — K5p: gradient kernel for block sizes ≥ 6 (ranks 41 … 64), one CTA per parameter vector.
//
// celerite_grad_kernel (grad.cuh) gives every (θ, direction) warp its own copy of the value recursion next to the tangent; at
// block sizes ≥ 6 the pair (2·BS² doubles per lane) does not fit the register file and the tangent block had to live in shared
// memory.  Here the value recursion runs ONCE per θ, in warp 0, and one warp per direction carries the tangent block alone
// (BS² doubles per lane — it fits).  The tangent of step n needs, from the value side, only what the value warp leaves in its
// per-step scratch anyway — q_{n−1}, w_{n−1} (for the product rule of the rank-1 update) and q_n, w_n, 1/D_n, z_n (for the
// owner phase) — so that scratch becomes a 3-slot ring in shared memory and the CTA runs as a two-stage pipeline: at tick τ the
// value warp sweeps step τ while the tangent warps sweep step τ−1, one __syncthreads per tick, no other handshake.
// Same step arithmetic as celerite_step<BS, ODD, /*PRE=*/false> (value) and celerite_step_dual (tangent part).
#pragma once
#include "grad.cuh"

namespace pioran {

template <int BS> __host__ __device__ constexpr int pipe_slot_doubles() { return 2 * rps_of(BS) + 2; }   // q | w | 1/D, z
constexpr int PIPE_CS = 8;      // steps per TMA stage (two to three CTAs share an SM)

// One lane state for both roles (a warp is either the value warp or a tangent warp for the whole sweep; two separate structs
// would make every thread carry two BS×BS blocks).  Value warp: gmu/chimu = the μ derivative, logacc/dkeep/dfirst = the log|D|
// ring.  Tangent warp: dlog = Σ D'_n/D_n; gmu, chimu, the ring unused.
template <int BS>
struct PipeState {
    double M[BS][BS];
    double sjj[2], g[2], amp[2], gmu[2];
    double chi2, chimu, logacc, dkeep, dfirst, dlog;
};
template <int BS> using PipeValueState = PipeState<BS>;
template <int BS> using PipeTangentState = PipeState<BS>;

// Value step n.  in: ring slot of step n−1 (q at [0, RPS), w at [RPS, 2·RPS), in the pending-φ form of celerite_step);
// out: slot of step n (same form, plus 1/D_n and z_n behind the vectors).
template <int BS, bool ODD>
__device__ __forceinline__ void pipe_value_step(PipeValueState<BS>& st, const double* __restrict__ T,
                                                const double* __restrict__ in, double* __restrict__ out, const LaneMap& lm,
                                                const double yn, const double s2n, const double suma, const double mu,
                                                const double nu, const int64_t n, const int lane) {
    constexpr int RP = rps_of(BS);
    const int o = lm.o;
    const double* qs = in;
    const double* ws = in + RP;
    double qrow[BS], urow[BS], xrow[BS], prow[BS];
    load_slice<BS>(qrow, qs + lm.rowI);
    load_slice<BS>(urow, T + (ODD ? F_UH : F_UT) * RP + lm.rowI);
    load_slice<BS>(xrow, T + (ODD ? F_PHI : F_KAP) * RP + lm.rowI);
    if (!ODD) load_slice<BS>(prow, T + F_PHI * RP + lm.rowI);
    double rowpart[BS], acc[BS];
#pragma unroll
    for (int r = 0; r < BS; r++) rowpart[r] = 0.0;
    const double* uAp = T + (ODD ? F_UT : F_UH) * RP + lm.colA;
    const double* uBp = T + (ODD ? F_UT : F_UH) * RP + lm.colB;
    const double* zAp = T + F_KAP * RP + lm.colA;
    const double* zBp = T + F_KAP * RP + lm.colB;
#pragma unroll
    for (int c = 0; c < BS; c++) {
        const double wA = ws[lm.colA + c], wB = ws[lm.colB + c], uA = uAp[c], uB = uBp[c];
        const double zA = ODD ? zAp[c] : 0.0, zB = ODD ? zBp[c] : 0.0;
        double cA = 0.0, cB = 0.0;
#pragma unroll
        for (int r = 0; r < BS; r++) {
            const bool useA = r > c;
            const double qr = (r == c) ? (lm.dzero ? 0.0 : qrow[r]) : qrow[r];
            const double m = fma(ODD ? (useA ? zA : zB) : xrow[r], st.M[r][c], qr * (useA ? wA : wB));
            st.M[r][c] = m;
            rowpart[r] = fma(m, useA ? uA : uB, rowpart[r]);
            if (useA) cA = fma(m, urow[r], cA);
            else      cB = fma(m, urow[r], cB);
        }
        const double yv = __shfl_sync(FULL, cB + (o ? cA : 0.0), lm.src_lane);
        acc[c] = yv + (o ? 0.0 : cA);
    }
    const double ut0 = T[F_UT * RP + lm.j0], ut1 = T[F_UT * RP + lm.j1];
    double sblk = 0.0;
#pragma unroll
    for (int r = 0; r < BS; r++) sblk = fma(urow[r], rowpart[r], sblk);
    double spart = fma(st.sjj[1] * ut1, ut1, fma(st.sjj[0] * ut0, ut0, sblk + sblk));
    double upart = fma(ut1, st.g[1], ut0 * st.g[0]);
    double umu = fma(ut1, st.gmu[1], ut0 * st.gmu[0]);
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) {
        spart += __shfl_xor_sync(FULL, spart, sft);
        upart += __shfl_xor_sync(FULL, upart, sft);
        umu += __shfl_xor_sync(FULL, umu, sft);
    }
    double tot[BS];
#pragma unroll
    for (int c = 0; c < BS; c++) {
        if (ODD) tot[c] = fma(xrow[c], rowpart[c], acc[c]);
        else     tot[c] = fma(prow[c], acc[c], rowpart[c]);
    }
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
    const double v0 = T[F_V * RP + lm.j0], v1 = T[F_V * RP + lm.j1];
    const double pn0 = T[F_PHN * RP + lm.j0], pn1 = T[F_PHN * RP + lm.j1];
    const double p0 = fma(st.sjj[0], ut0, f0), p1 = fma(st.sjj[1], ut1, f1);
    const double D = fma(nu, s2n, suma) - spart;      // celerite_solver.jl:92
    const double rD = fast_rcp(D);
    const double z = (yn - mu) - upart;               // celerite_solver.jl:141
    st.chi2 = fma(z * z, rD, st.chi2);
    const double dzmu = -1.0 - umu;                   // ∂z_n/∂μ
    st.chimu = fma(2.0 * z * dzmu, rD, st.chimu);
    if (n == 0) st.dfirst = D;
    else if ((int)(n & 31) == lane) st.dkeep = D;
    if ((n & 31) == 31) { st.logacc += log(fabs(st.dkeep)); st.dkeep = 1.0; }
    const double q0 = fma(st.amp[0], v0, -p0), q1 = fma(st.amp[1], v1, -p1);
    const double w0 = q0 * rD, w1 = q1 * rD;
    st.g[0] = pn0 * fma(w0, z, st.g[0]);
    st.g[1] = pn1 * fma(w1, z, st.g[1]);
    st.gmu[0] = pn0 * fma(w0, dzmu, st.gmu[0]);
    st.gmu[1] = pn1 * fma(w1, dzmu, st.gmu[1]);
    st.sjj[0] = (pn0 * pn0) * fma(q0, w0, st.sjj[0]);   // celerite_solver.jl:85
    st.sjj[1] = (pn1 * pn1) * fma(q1, w1, st.sjj[1]);
    // the NEXT step is odd iff this one is even: odd steps consume (q, φ∘w), even steps (φ∘q, w)
    double* qo = out;
    double* wo = out + RP;
    if (!ODD) {
        qo[lm.j0] = q0; wo[lm.j0] = pn0 * w0;
        if (lm.valid1) { qo[lm.j1] = q1; wo[lm.j1] = pn1 * w1; }
    } else {
        qo[lm.j0] = pn0 * q0; wo[lm.j0] = w0;
        if (lm.valid1) { qo[lm.j1] = pn1 * q1; wo[lm.j1] = w1; }
    }
    if (lane == 0) { out[2 * RP] = rD; out[2 * RP + 1] = z; }
}

// Tangent of step n along one direction.  prev / cur: value-side ring slots of steps n−1 / n; sd: this warp's own scratch
// (tangents of q and w in the same pending-φ form; q at [0, RPS), w at [RPS, 2·RPS)).
template <int BS, bool ODD>
__device__ __forceinline__ void pipe_tangent_step(PipeTangentState<BS>& st, const double* __restrict__ T,
                                                  const double* __restrict__ prev, const double* __restrict__ cur,
                                                  double* __restrict__ sd, const LaneMap& lm, const double s2n,
                                                  const double dsuma, const double dnu, const int lane) {
    constexpr int RP = rps_of(BS);
    const int o = lm.o;
    double qv[BS], qd[BS], urow[BS], xrow[BS], prow[BS];
    load_slice<BS>(qv, prev + lm.rowI);
    load_slice<BS>(qd, sd + lm.rowI);
    load_slice<BS>(urow, T + (ODD ? F_UH : F_UT) * RP + lm.rowI);
    load_slice<BS>(xrow, T + (ODD ? F_PHI : F_KAP) * RP + lm.rowI);
    if (!ODD) load_slice<BS>(prow, T + F_PHI * RP + lm.rowI);
    double rowpart[BS], acc[BS];
#pragma unroll
    for (int r = 0; r < BS; r++) rowpart[r] = 0.0;
    const double* wvp = prev + RP;
    const double* wdp = sd + RP;
    const double* uAp = T + (ODD ? F_UT : F_UH) * RP + lm.colA;
    const double* uBp = T + (ODD ? F_UT : F_UH) * RP + lm.colB;
    const double* zAp = T + F_KAP * RP + lm.colA;
    const double* zBp = T + F_KAP * RP + lm.colB;
#pragma unroll
    for (int c = 0; c < BS; c++) {
        const double wAv = wvp[lm.colA + c], wBv = wvp[lm.colB + c], wAd = wdp[lm.colA + c], wBd = wdp[lm.colB + c];
        const double uA = uAp[c], uB = uBp[c];
        const double zA = ODD ? zAp[c] : 0.0, zB = ODD ? zBp[c] : 0.0;
        double cA = 0.0, cB = 0.0;
#pragma unroll
        for (int r = 0; r < BS; r++) {
            const bool useA = r > c;
            const bool masked = (r == c) && lm.dzero;
            const double qrv = masked ? 0.0 : qv[r], qrd = masked ? 0.0 : qd[r];
            // d(q_r w_c) = q'_r w_c + q_r w'_c, then the decay of the entry
            const double prod = fma(qrd, useA ? wAv : wBv, qrv * (useA ? wAd : wBd));
            const double m = fma(ODD ? (useA ? zA : zB) : xrow[r], st.M[r][c], prod);
            st.M[r][c] = m;
            rowpart[r] = fma(m, useA ? uA : uB, rowpart[r]);
            if (useA) cA = fma(m, urow[r], cA);
            else      cB = fma(m, urow[r], cB);
        }
        const double yv = __shfl_sync(FULL, cB + (o ? cA : 0.0), lm.src_lane);
        acc[c] = yv + (o ? 0.0 : cA);
    }
    const double ut0 = T[F_UT * RP + lm.j0], ut1 = T[F_UT * RP + lm.j1];
    double sblk = 0.0;
#pragma unroll
    for (int r = 0; r < BS; r++) sblk = fma(urow[r], rowpart[r], sblk);
    double spart = fma(st.sjj[1] * ut1, ut1, fma(st.sjj[0] * ut0, ut0, sblk + sblk));
    double upart = fma(ut1, st.g[1], ut0 * st.g[0]);
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) {
        spart += __shfl_xor_sync(FULL, spart, sft);
        upart += __shfl_xor_sync(FULL, upart, sft);
    }
    double tot[BS];
#pragma unroll
    for (int c = 0; c < BS; c++) {
        if (ODD) tot[c] = fma(xrow[c], rowpart[c], acc[c]);
        else     tot[c] = fma(prow[c], acc[c], rowpart[c]);
    }
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
    // owner phase: tangents of rows j0 (and j1); value-side quantities of step n come from `cur`
    const double v0 = T[F_V * RP + lm.j0], v1 = T[F_V * RP + lm.j1];
    const double pn0 = T[F_PHN * RP + lm.j0], pn1 = T[F_PHN * RP + lm.j1];
    const double rDv = cur[2 * RP], zv = cur[2 * RP + 1];
    // cur holds (q, φ∘w) after an even step and (φ∘q, w) after an odd one: the raw one gives w' through
    // w' = q'/D − q D'/D² (even) or w' = (q' − w D')/D (odd); the products below need only the φ-scaled forms
    // (a lane without a second row has j1 == j0: its second-row quantities must stay zero, as they do on the value side)
    const double cq0 = cur[lm.j0], cw0 = cur[RP + lm.j0];
    const double cq1 = lm.valid1 ? cur[lm.j1] : 0.0, cw1 = lm.valid1 ? cur[RP + lm.j1] : 0.0;
    const double dp0 = fma(st.sjj[0], ut0, f0), dp1 = fma(st.sjj[1], ut1, f1);
    const double dD = fma(dnu, s2n, dsuma) - spart;
    const double drD = -dD * rDv * rDv;
    const double dz = -upart;
    st.chi2 = fma(zv * zv, drD, fma(2.0 * zv * dz, rDv, st.chi2));
    st.dlog = fma(dD, rDv, st.dlog);
    const double dq0 = fma(st.amp[0], v0, -dp0), dq1 = fma(st.amp[1], v1, -dp1);
    double dw0, dw1, pq0, pw0, pq1, pw1;        // w'; φ_{n+1}∘q, φ_{n+1}∘w (value side)
    if (!ODD) {
        dw0 = fma(dq0, rDv, cq0 * drD); dw1 = fma(dq1, rDv, cq1 * drD);
        pq0 = pn0 * cq0; pw0 = cw0; pq1 = pn1 * cq1; pw1 = cw1;
    } else {
        dw0 = fma(-cw0, dD, dq0) * rDv; dw1 = fma(-cw1, dD, dq1) * rDv;
        pq0 = cq0; pw0 = pn0 * cw0; pq1 = cq1; pw1 = pn1 * cw1;
    }
    st.g[0] = fma(pw0, dz, pn0 * fma(dw0, zv, st.g[0]));
    st.g[1] = fma(pw1, dz, pn1 * fma(dw1, zv, st.g[1]));
    st.sjj[0] = pn0 * fma(dq0, pw0, fma(pq0, dw0, pn0 * st.sjj[0]));
    st.sjj[1] = pn1 * fma(dq1, pw1, fma(pq1, dw1, pn1 * st.sjj[1]));
    __syncwarp();
    if (!ODD) {
        sd[lm.j0] = dq0; sd[RP + lm.j0] = pn0 * dw0;
        if (lm.valid1) { sd[lm.j1] = dq1; sd[RP + lm.j1] = pn1 * dw1; }
    } else {
        sd[lm.j0] = pn0 * dq0; sd[RP + lm.j0] = dw0;
        if (lm.valid1) { sd[lm.j1] = pn1 * dq1; sd[RP + lm.j1] = dw1; }
    }
    __syncwarp();
}

// grid = number of parameter vectors (one work item each); block = (1 + ND + 1) warps: warp 0 the value recursion, warp 1 + k the
// tangent along direction k (k < ND: PSD parameter k; k = ND: ν).  ∂/∂μ rides in the value warp, ∂/∂norm follows from ∂/∂ν
// (grad.cuh).  Shared memory: 2 TMA stages of PIPE_CS table records | 3 ring slots | (ND + 1) tangent scratch vectors pairs |
// 2 mbarriers | 2 result slots.
// NWARPS = 1 + ND + 1 (5 for three PSD parameters, 7 for five).  One CTA per SM: a BS = 8 block plus the step's operands needs
// the full 255 registers (capping them for two CTAs per SM spills 6 KB per thread and runs 1.75× slower, measured).
template <int BS, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1) celerite_grad_pipe_kernel(const GradArgs args) {
    constexpr int RP = G * BS, RPS = rps_of(BS), SD = table_step_doubles(RPS), STAGE = PIPE_CS * SD, SLOT = pipe_slot_doubles<BS>();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NT = args.ND + 1;
    double* stages = reinterpret_cast<double*>(smem_raw);
    double* ring = stages + 2 * STAGE;
    double* tsc = ring + 3 * SLOT;                         // [NT][2·RPS]
    double* res = tsc + (size_t)NT * 2 * RPS;              // chi2 of the value warp, for the norm derivative
    uint64_t* bars = reinterpret_cast<uint64_t*>(res + 2);

    const WorkItem wk = args.work[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t N = wk.N;
    const int64_t nchunks = (N + PIPE_CS - 1) / PIPE_CS;
    constexpr uint32_t STAGE_BYTES = STAGE * sizeof(double);
    const int th = wk.theta_begin;
    const int P = args.ND + 3;

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    for (int k = threadIdx.x; k < 3 * SLOT + NT * 2 * RPS + 2; k += blockDim.x) ring[k] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 2 && k < nchunks; k++) {
            mbar_arrive_expect_tx(&bars[k], STAGE_BYTES);
            tma_load_1d(stages + k * STAGE, wk.table + (size_t)k * STAGE, STAGE_BYTES, &bars[k]);
        }
    }
    const LaneMap lm = make_lane_map<BS>(lane);
    const int i = lane >> 2, o = lane & 3;
    const double* trow = args.theta + (size_t)th * args.pstride;
    const double nu = trow[args.ND + 1], mu = trow[args.ND + 2];
    const double* av = args.amp + (size_t)th * RP;

    PipeState<BS> st;
    double dsuma = 0.0, dnu = 0.0;
    const double suma = args.suma[th];
    double* sd = tsc + (size_t)(warp > 0 ? warp - 1 : 0) * 2 * RPS;
#pragma unroll
    for (int r = 0; r < BS; r++)
#pragma unroll
        for (int c = 0; c < BS; c++) st.M[r][c] = 0.0;
    st.sjj[0] = st.sjj[1] = st.g[0] = st.g[1] = st.gmu[0] = st.gmu[1] = 0.0;
    st.chi2 = st.chimu = st.logacc = st.dlog = 0.0; st.dkeep = 1.0; st.dfirst = 1.0;
    if (warp == 0) {
        st.amp[0] = av[i * BS + o];
        st.amp[1] = lm.valid1 ? av[i * BS + o + 4] : 0.0;
    } else {
        const int k = min(warp - 1, NT - 1);
        const bool amp_dir = k < args.ND;
        const double* ad = args.damp + ((size_t)th * args.ND + (amp_dir ? k : 0)) * RP;
        st.amp[0] = amp_dir ? ad[i * BS + o] : 0.0;
        st.amp[1] = (amp_dir && lm.valid1) ? ad[i * BS + o + 4] : 0.0;
        dsuma = amp_dir ? args.dsuma[(size_t)th * args.ND + k] : 0.0;
        dnu = amp_dir ? 0.0 : 1.0;
    }

    // tick τ: the value warp sweeps step τ (τ < N), the tangent warps step τ − 1 (τ ≥ 1); slot of step n = ring[(n mod 3)],
    // the slot "before step 0" is ring[2] (zeros)
    for (int64_t tau = 0; tau <= N; tau++) {
        if (warp == 0) {
            if (tau < N) {
                const int64_t kc = tau / PIPE_CS;
                if (tau % PIPE_CS == 0) mbar_wait(&bars[kc & 1], (uint32_t)((kc >> 1) & 1));
                const double* T = stages + (kc & 1) * STAGE + (tau % PIPE_CS) * SD;
                const double* in = ring + ((tau + 2) % 3) * SLOT;
                double* out = ring + (tau % 3) * SLOT;
                if (tau & 1) pipe_value_step<BS, true>(st, T, in, out, lm, T[6 * RPS], T[6 * RPS + 1], suma, mu, nu, tau, lane);
                else         pipe_value_step<BS, false>(st, T, in, out, lm, T[6 * RPS], T[6 * RPS + 1], suma, mu, nu, tau, lane);
            }
        } else if (warp <= NT) {
            if (tau >= 1) {
                const int64_t n = tau - 1, kc = n / PIPE_CS;
                if (n % PIPE_CS == 0) mbar_wait(&bars[kc & 1], (uint32_t)((kc >> 1) & 1));
                const double* T = stages + (kc & 1) * STAGE + (n % PIPE_CS) * SD;
                const double* prev = ring + ((n + 2) % 3) * SLOT;
                const double* cur = ring + (n % 3) * SLOT;
                if (n & 1) pipe_tangent_step<BS, true>(st, T, prev, cur, sd, lm, T[6 * RPS + 1], dsuma, dnu, lane);
                else       pipe_tangent_step<BS, false>(st, T, prev, cur, sd, lm, T[6 * RPS + 1], dsuma, dnu, lane);
            }
        }
        __syncthreads();
        // the tangents have left chunk kc = τ/CS − 1 when τ is a positive multiple of CS: its stage takes chunk kc + 2
        if (threadIdx.x == 0 && tau >= 1 && tau % PIPE_CS == 0) {
            const int64_t kc = tau / PIPE_CS - 1;
            if (kc + 2 < nchunks) {
                fence_proxy_async();
                mbar_arrive_expect_tx(&bars[kc & 1], STAGE_BYTES);
                tma_load_1d(stages + (kc & 1) * STAGE, wk.table + (size_t)(kc + 2) * STAGE, STAGE_BYTES, &bars[kc & 1]);
            }
        }
    }
    if (warp == 0) {
        double la = st.logacc + log(fabs(st.dkeep));
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) la += __shfl_xor_sync(FULL, la, sft);
        const double logdet = log(st.dfirst) + la;
        if (lane == 0) {
            if (args.logl) args.logl[th] = -logdet / 2 - (double)N * 1.8378770664093453 / 2 - st.chi2 / 2;
            args.grad[(size_t)th * P + args.ND + 2] = -st.chimu / 2;
            res[0] = st.chi2;
        }
    }
    __syncthreads();
    if (warp >= 1 && warp <= NT && lane == 0) {
        const int k = warp - 1;
        const double gk = -st.dlog / 2 - st.chi2 / 2;
        if (k < args.ND) {
            args.grad[(size_t)th * P + k] = gk;
        } else {      // the ν warp also reports ∂/∂norm (Euler's identity, grad.cuh)
            args.grad[(size_t)th * P + args.ND + 1] = gk;
            args.grad[(size_t)th * P + args.ND] = (0.5 * res[0] - 0.5 * (double)N - nu * gk) / trow[args.ND];
        }
    }
}

}  // namespace pioran
