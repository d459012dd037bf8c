// wide.cuh — K2w: celerite log-likelihood for ranks above 64 (FP64, sm_100a): state in shared memory (ranks up to 160, first
// kernel below) or in the register file of one CTA (ranks up to 128, second kernel — the one the dispatcher prefers).
//
// Same recursion as celerite.cuh (reference src/celerite_solver.jl:12-158, 312-334, forward-only fused sweep), for the
// part of the reference's own benchmark grid that the register-resident kernel cannot hold: SHO J = 40, 50 (rank 80, 100),
// DRWCelerite J = 30 … 50 (rank 90 … 150), explicit coefficient sets of 64 terms (rank 128) — benchmark/benchmarks.jl:16-18.
// One CTA per (parameter vector, series); the R×R state lives in SHARED memory (full square, both triangles, so that the
// mat-vec needs no cross-thread reduction along a column):
//     T   ← (φ_n φ_nᵀ) ∘ (T + q_{n-1} w_{n-1}ᵀ)      q = V − p,  w = q / D
//     p   = Tᵀ U_n  (= T U_n);   D_n = A_n − U_nᵀ p;   z_n = (y_n − μ) − U_nᵀ g,   g ← φ_n ∘ (g + w_{n-1} z_{n-1})
// Thread (cp, h) owns the column pair (2cp, 2cp+1) on the row range h: 128-bit shared loads/stores of the state, the
// row-side operands (φ_r, q_r φ_r, U_r) are warp-wide broadcasts.  The kernel is bound by shared-memory bandwidth
// (16 B moved per entry and step against 4 FP64 issues), about 4× the FP64 time — the price of a state that does not fit
// the register file; it is the wide-rank completion of the path, not its headline.
#pragma once
#include "celerite.cuh"

namespace pioran {

constexpr int WIDE_THREADS = 256;
constexpr int WIDE_CH = 4;          // steps whose U, V, φ vectors are prepared at once
constexpr int WIDE_MAX_RANK = 160;

struct WideGeom {
    int R;        // live rows
    int CP;       // column pairs (multiple of 16); leading dimension of the state = 2·CP
    int PARTS;    // row ranges per column pair; CP · PARTS ≤ 256
    int RPP;      // rows per range
};
__host__ __device__ inline WideGeom wide_geom(int R) {
    WideGeom g;
    g.R = R;
    g.CP = (((R + 1) / 2 + 15) / 16) * 16;
    g.PARTS = WIDE_THREADS / g.CP;
    g.RPP = (R + g.PARTS - 1) / g.PARTS;
    return g;
}
// doubles of dynamic shared memory: state | U, V, φ tables (WIDE_CH steps) | bro (4 per row) | wφ | p partials | 16 reduction slots
__host__ __device__ inline size_t wide_smem_doubles(const WideGeom& g) {
    const int LD = 2 * g.CP;
    return (size_t)g.R * LD + 3 * (size_t)WIDE_CH * LD + 4 * (size_t)LD + LD + (size_t)g.PARTS * LD + 32;
}

// grid = number of work items (one parameter vector each, WorkItem.count == 1); block = 256.
__global__ void __launch_bounds__(WIDE_THREADS, 1) celerite_wide_kernel(const BatchArgs args) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const WideGeom gm = wide_geom(args.R);
    const int R = gm.R, LD = 2 * gm.CP;
    double* T = reinterpret_cast<double*>(smem_raw);
    double* tabU = T + (size_t)R * LD;
    double* tabV = tabU + WIDE_CH * LD;
    double* tabP = tabV + WIDE_CH * LD;
    double* bro = tabP + WIDE_CH * LD;          // [row][4]: φ_r, q_r φ_r, U_r, unused
    double* wphi = bro + 4 * LD;                // w_c φ_c of the coming step
    double* ppart = wphi + LD;                  // [PARTS][LD]
    double* red = ppart + gm.PARTS * LD;        // [8 warps][2] + 2 result slots

    const WorkItem wk = args.work[blockIdx.x];
    const int th = wk.theta_begin;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Jt = args.Jt;
    const int64_t N = wk.N;
    const double* ca = args.a + (size_t)th * Jt;
    const double* cb = args.b + (size_t)th * Jt;
    const double* cc = args.c + (size_t)th * Jt;
    const double* cd = args.d + (size_t)th * Jt;
    const size_t pi = (size_t)wk.par_begin;
    const double mu = args.mu ? args.mu[pi * args.pstride] : 0.0;
    const double nu = args.nu ? args.nu[pi * args.pstride] : 1.0;
    const double* yb = args.y_batch ? args.y_batch + pi * args.ystride : wk.y;
    const double* sb = args.s2_batch ? args.s2_batch + pi * args.ystride : wk.s2;

    double suma = 0.0;
    for (int m = 0; m < Jt; m++) suma += ca[m];   // celerite_solver.jl:21

    const int cp = tid % gm.CP, h = tid / gm.CP;
    const bool worker = h < gm.PARTS;
    const int c0 = 2 * cp;                        // my column pair
    const int r_lo = h * gm.RPP, r_hi = min(R, r_lo + gm.RPP);
    const bool owner = (h == 0);                  // owner of columns c0, c0 + 1 in the vector phases

    for (size_t k = tid; k < wide_smem_doubles(gm); k += WIDE_THREADS) T[k] = 0.0;
    double g0 = 0.0, g1 = 0.0;                    // forward-substitution vector of my columns (owner threads)
    double w0 = 0.0, w1 = 0.0, q0 = 0.0, q1 = 0.0, zprev = 0.0;
    double chi2 = 0.0, logdet = 0.0;
    __syncthreads();

    for (int64_t nb = 0; nb < N; nb += WIDE_CH) {
        const int ns = (int)((N - nb) < WIDE_CH ? (N - nb) : WIDE_CH);
        // ---- U, V, φ of the next ns steps (celerite_solver.jl:51-64), one thread per (step, term)
        for (int idx = tid; idx < ns * Jt; idx += WIDE_THREADS) {
            const int s = idx / Jt, m = idx - s * Jt;
            const int64_t n = nb + s;
            const double tn = wk.t[n];
            const double ph = (n >= 1) ? exp(-cc[m] * (tn - wk.t[n - 1])) : 0.0;
            const int tr = args.term_row[m];
            if (tr < 0) {     // real term: one row, U = a, V = 1
                const int r0 = -tr - 1;
                tabU[s * LD + r0] = ca[m]; tabV[s * LD + r0] = 1.0; tabP[s * LD + r0] = ph;
            } else {
                double si, co;
                sincos_large(cd[m] * tn, &si, &co);
                tabU[s * LD + tr] = ca[m] * co + cb[m] * si;      // celerite_solver.jl:60
                tabU[s * LD + tr + 1] = ca[m] * si - cb[m] * co;  // celerite_solver.jl:59
                tabV[s * LD + tr] = co; tabV[s * LD + tr + 1] = si;
                tabP[s * LD + tr] = ph; tabP[s * LD + tr + 1] = ph;
            }
        }
        __syncthreads();
        for (int s = 0; s < ns; s++) {
            const int64_t n = nb + s;
            const double* Un = tabU + s * LD;
            const double* Vn = tabV + s * LD;
            const double* Pn = tabP + s * LD;
            // ---- phase 0: row-side broadcast operands and w φ of this step (owners), from the previous step's q, w
            if (owner) {
                const double p0 = Pn[c0], p1 = Pn[c0 + 1];
                *reinterpret_cast<double2*>(bro + 4 * c0) = make_double2(p0, q0 * p0);
                bro[4 * c0 + 2] = Un[c0];
                *reinterpret_cast<double2*>(bro + 4 * (c0 + 1)) = make_double2(p1, q1 * p1);
                bro[4 * (c0 + 1) + 2] = Un[c0 + 1];
                *reinterpret_cast<double2*>(wphi + c0) = make_double2(w0 * p0, w1 * p1);
                g0 = p0 * fma(w0, zprev, g0);         // celerite_solver.jl:137
                g1 = p1 * fma(w1, zprev, g1);
            }
            __syncthreads();
            // ---- phase 1: state update and column sums over my row range
            double pa0 = 0.0, pa1 = 0.0;
            if (worker) {
                const double2 wp = *reinterpret_cast<const double2*>(wphi + c0);
                const double2 pc = *reinterpret_cast<const double2*>(Pn + c0);
                double* Tc = T + (size_t)r_lo * LD + c0;
#pragma unroll 4
                for (int r = r_lo; r < r_hi; r++, Tc += LD) {
                    const double2 b01 = *reinterpret_cast<const double2*>(bro + 4 * r);   // φ_r, q_r φ_r
                    const double ur = bro[4 * r + 2];
                    double2 t = *reinterpret_cast<double2*>(Tc);
                    t.x = fma(pc.x, b01.x * t.x, b01.y * wp.x);     // φ_c (φ_r T) + (q_r φ_r)(w_c φ_c)   celerite_solver.jl:76
                    t.y = fma(pc.y, b01.x * t.y, b01.y * wp.y);
                    *reinterpret_cast<double2*>(Tc) = t;
                    pa0 = fma(t.x, ur, pa0);
                    pa1 = fma(t.y, ur, pa1);
                }
                *reinterpret_cast<double2*>(ppart + h * LD + c0) = make_double2(pa0, pa1);
            }
            __syncthreads();
            // ---- phase 2: p, and the two inner products UᵀTU, Uᵀg
            double p0 = 0.0, p1 = 0.0, sred = 0.0, ured = 0.0;
            if (owner) {
                for (int k = 0; k < gm.PARTS; k++) {
                    const double2 v = *reinterpret_cast<const double2*>(ppart + k * LD + c0);
                    p0 += v.x; p1 += v.y;
                }
                const double u0 = Un[c0], u1 = Un[c0 + 1];
                sred = fma(u1, p1, u0 * p0);
                ured = fma(u1, g1, u0 * g0);
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
            // ---- phase 3: pivot, innovation, next q and w
            const double D = fma(nu, sb[n], suma) - stot;     // celerite_solver.jl:92
            const double z = (yb[n] - mu) - utot;             // celerite_solver.jl:141
            const double rD = 1.0 / D;
            if (owner) {
                q0 = Vn[c0] - p0; q1 = Vn[c0 + 1] - p1;       // celerite_solver.jl:95-98 (W = q / D)
                w0 = q0 * rD; w1 = q1 * rD;
            }
            zprev = z;
            chi2 = fma(z * z, rD, chi2);
            logdet += (n == 0) ? log(D) : log(fabs(D));       // celerite_solver.jl:126,140
            // the barrier at the top of the next step (after phase 0) orders red/ppart reuse
        }
        __syncthreads();   // the tables are rebuilt for the next WIDE_CH steps
    }
    if (tid == 0) args.out[wk.out_begin] = -logdet / 2 - (double)N * 1.8378770664093453 / 2 - chi2 / 2;   // celerite_solver.jl:333
}

// ------------------------------------------------------------------------------------------------ ranks 65 … 128 in registers
// The register file of ONE CTA holds a full 128×128 FP64 matrix (256 threads × 64 doubles): for ranks up to 128 the state
// leaves shared memory altogether.  Thread (ty, tx) of a 16×16 grid owns the interleaved TS×TS tile rows {ty + 16·i} × columns
// {tx + 16·j} (TS = ⌈R/16⌉ = 5 … 8) of the FULL square (both triangles: row sums need no transposed partner).  Per step:
//   phase 0  owners (one thread per row) publish q φ, w φ of the step and advance g;
//   phase 1  every thread updates its tile, T ← φ_r (φ_c T + q_r (w_c φ_c)), and accumulates the row sums Σ_c T_rc U_c; a
//            reduce-scatter over the 16 lanes of a half-warp (8 shuffles) leaves one row total per lane pair → p;
//   phase 2  owners form U_r p_r and U_r g_r, block reduction → D_n, z_n;   phase 3  owners form the next q, w.
// Column-side vectors are read as 16 consecutive doubles per half-warp (conflict-free), row-side ones are broadcasts.  FP64-bound
// at 4 issues per entry of the full square (twice the symmetric kernel's count), 3–4× faster than the shared-memory state.
// MODE (celerite.cuh: StepMode): STEP_STORE also writes D_n, the forward z_n and W_n ([N × 16·TS] per parameter vector) for the
// posterior mean; STEP_SIM takes standard-normal draws through y_batch and emits the realisation (celerite_solver.jl:536-546).
template <int TS, int MODE = STEP_LOGL>
__global__ void __launch_bounds__(WIDE_THREADS, TS <= 6 ? 2 : 1) celerite_wide_reg_kernel(const BatchArgs args) {
    constexpr int LD = 16 * TS;
    __shared__ __align__(16) double tabU[WIDE_CH][LD], tabV[WIDE_CH][LD], tabP[WIDE_CH][LD];
    __shared__ __align__(16) double qphi[LD], wphi[LD], p_s[LD];
    __shared__ double red[2 * (WIDE_THREADS / 32)];

    const WorkItem wk = args.work[blockIdx.x];
    const int th = wk.theta_begin;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int Jt = args.Jt;
    const int64_t N = wk.N;
    const double* ca = args.a + (size_t)th * Jt;
    const double* cb = args.b + (size_t)th * Jt;
    const double* cc = args.c + (size_t)th * Jt;
    const double* cd = args.d + (size_t)th * Jt;
    const size_t pi = (size_t)wk.par_begin;
    const double mu = args.mu ? args.mu[pi * args.pstride] : 0.0;
    const double nu = args.nu ? args.nu[pi * args.pstride] : 1.0;
    const double* yb = args.y_batch ? args.y_batch + pi * args.ystride : wk.y;
    const double* sb = args.s2_batch ? args.s2_batch + pi * args.ystride : wk.s2;

    double suma = 0.0;
    for (int m = 0; m < Jt; m++) suma += ca[m];   // celerite_solver.jl:21

    double Tm[TS][TS];
#pragma unroll
    for (int i = 0; i < TS; i++)
#pragma unroll
        for (int j = 0; j < TS; j++) Tm[i][j] = 0.0;
    for (int k = tid; k < WIDE_CH * LD; k += WIDE_THREADS) { (&tabU[0][0])[k] = 0.0; (&tabV[0][0])[k] = 0.0; (&tabP[0][0])[k] = 0.0; }
    const bool owner = tid < LD;                  // owner of row `tid` in the vector phases
    double* const Wst = MODE == STEP_STORE ? args.W_out + pi * (size_t)N * LD : nullptr;
    double* const Dst = MODE == STEP_STORE ? args.D_out + pi * (size_t)N : nullptr;
    double* const zst = MODE == STEP_STORE ? args.zf_out + pi * (size_t)N : nullptr;
    double* const ysim = MODE == STEP_SIM ? args.ysim_out + pi * (size_t)N : nullptr;
    double g = 0.0, q = 0.0, w = 0.0, zprev = 0.0, chi2 = 0.0;
    double logacc = 0.0, dkeep = 1.0, dfirst = 1.0;   // Σ log|D_n|: the last warp keeps D_n in lane n % 32, one log per 32 steps
    const bool b3 = (tx & 8) != 0, b2 = (tx & 4) != 0, b1 = (tx & 2) != 0;
    __syncthreads();

    for (int64_t nb = 0; nb < N; nb += WIDE_CH) {
        const int ns = (int)((N - nb) < WIDE_CH ? (N - nb) : WIDE_CH);
        // ---- U, V, φ of the next ns steps (celerite_solver.jl:51-64), one thread per (step, term)
        for (int idx = tid; idx < ns * Jt; idx += WIDE_THREADS) {
            const int s = idx / Jt, m = idx - s * Jt;
            const int64_t n = nb + s;
            const double tn = wk.t[n];
            const double ph = (n >= 1) ? exp(-cc[m] * (tn - wk.t[n - 1])) : 0.0;
            const int tr = args.term_row[m];
            if (tr < 0) {     // real term: one row, U = a, V = 1
                const int r0 = -tr - 1;
                tabU[s][r0] = ca[m]; tabV[s][r0] = 1.0; tabP[s][r0] = ph;
            } else {
                double si, co;
                sincos_large(cd[m] * tn, &si, &co);
                tabU[s][tr] = ca[m] * co + cb[m] * si;      // celerite_solver.jl:60
                tabU[s][tr + 1] = ca[m] * si - cb[m] * co;  // celerite_solver.jl:59
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
            // ---- phase 0
            if (owner) {
                const double ph = Pn[tid];
                qphi[tid] = q * ph;
                wphi[tid] = w * ph;
                g = ph * fma(w, zprev, g);            // celerite_solver.jl:137
            }
            __syncthreads();
            // ---- phase 1: tile update and row sums
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
                    // φ_r (φ_c T) + (q_r φ_r)(w_c φ_c)   celerite_solver.jl:76
                    const double t = fma(phr[i], pc * Tm[i][j], qr[i] * wc);
                    Tm[i][j] = t;
                    rs[i] = fma(t, uc, rs[i]);
                }
            }
            {   // reduce-scatter of the (padded) 8 row sums over the 16 lanes that share ty
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
            // ---- phase 2: the two inner products UᵀTU, Uᵀg
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
            // ---- phase 3: pivot, innovation, next q and w
            const double D = fma(nu, sb[n], suma) - stot;     // celerite_solver.jl:92
            double z = (yb[n] - mu) - utot;                   // celerite_solver.jl:141
            if (MODE == STEP_SIM) {                           // celerite_solver.jl:539-545: the draw enters where the innovation would
                z = sqrt(D) * yb[n];
                if (tid == 0) ysim[n] = utot + z;
            }
            const double rD = fast_rcp(D);
            if (owner) { q = Vn[tid] - p; w = q * rD; }       // celerite_solver.jl:95-98 (W = q / D)
            if (MODE == STEP_STORE) {
                if (owner) Wst[n * LD + tid] = w;
                if (tid == 0) { Dst[n] = D; zst[n] = z; }
            }
            zprev = z;
            chi2 = fma(z * z, rD, chi2);
            if (warp == WIDE_THREADS / 32 - 1) {              // off the owners' critical path (celerite_solver.jl:126,140)
                if (n == 0) dfirst = D;
                else if ((int)(n & 31) == lane) dkeep = D;
                if ((n & 31) == 31) { logacc += log(fabs(dkeep)); dkeep = 1.0; }
            }
        }
        __syncthreads();   // the tables are rebuilt for the next WIDE_CH steps
    }
    if (warp == WIDE_THREADS / 32 - 1) {
        double la = logacc + log(fabs(dkeep));
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) la += __shfl_xor_sync(FULL, la, sft);
        const double logdet = log(dfirst) + la;               // no abs on the first pivot (celerite_solver.jl:126)
        if (lane == 0) args.out[wk.out_begin] = -logdet / 2 - (double)N * 1.8378770664093453 / 2 - chi2 / 2;   // celerite_solver.jl:333
    }
}

}  // namespace pioran
