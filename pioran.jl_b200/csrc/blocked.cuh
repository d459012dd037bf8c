// blocked.cuh — K2t: batched celerite factorisation + solve + log-determinant on the FP64 tensor pipe (sm_100a).
//
// Same job as celerite.cuh (logl = init_semi_separable! + solve_prec!, src/celerite_solver.jl:12-100, 115-158, 312-334)
// for the shared-table (approx) path, but the recursion is BLOCKED over 8 time steps so that its two O(R²) operations
// become small matrix products that run on DMMA (mma.sync.m8n8k4.f64; tcgen05 has no FP64 kind):
//
// With X = T_n + q_n w_nᵀ (the amplitude-scaled state of celerite.cuh after step n0, before its decay) and for the
// steps s = 1…8 of a block, Ψ_{a→b} = Π_{a<i≤b} φ_{n0+i} (all factors ≤ 1: nothing can overflow),
//     Û[:,s]  = Ψ_{0→s}∘Ũ_s          V̂[:,s] = Ψ_{s→8}∘V_s          ψ8 = Ψ_{0→8}            (θ-independent table)
//     P0      = X·Û                                                   (R×R · R×8  → DMMA)
//     C       = K_blk − Ûᵀ·P0                                         (8×8 conditional covariance of the block;  DMMA)
//               K_blk[s][s'] = Σ_j amp_j·H[s][s'][j],  H = Ũ_s Ψ_{s'→s} V_{s'} (table);  diagonal Σa + ν σ²_s
//     C       = L D Lᵀ                                                (D_s are the pivots D_n of celerite_solver.jl:92)
//     Q̂       = (amp∘V̂ − ψ8∘P0)·L⁻ᵀ ,  Ŵ = Q̂ D⁻¹                     (R×8 · 8×8 → DMMA;  q̂_s = Ψ_{s→8}∘q_s, celerite_solver.jl:95-98)
//     X       ← (ψ8ψ8ᵀ)∘X + Q̂·Ŵᵀ                                     (rank-8 update → DMMA;  celerite_solver.jl:69-90)
// The forward substitution (celerite_solver.jl:132-142) rides along as one more ROW of the state: row RG holds g⁺ᵀ,
// its "amp∘V̂" entry is (y_s − μ), its decay factor 1; then P0[RG][s] is the prediction Û_sᵀg⁺, Q̂[RG][s] the innovation
// z_s, and yᵀK⁻¹y = Σ z_s²/D_s = Σ Q̂[RG][s]·Ŵ[RG][s].
//
// Mapping.  One warp per (parameter vector, series).  X is symmetric: the lower-triangular 8×8 tiles (I ≥ K) live in
// registers in the DMMA accumulator layout (lane (g = lane>>2, t = lane&3) holds rows 8I+g, columns 8K+2t, 8K+2t+1), which
// is directly the A operand of P0 = X·Û when the contraction index of chunk c is read as column 8K+2t+c (the B operand —
// the table — is stored in that order).  The upper tiles are produced on the fly by a 2-exchange register transpose.
// Every product of the block chains without leaving registers: P0 (accumulator layout) → Bm → A operand of Q̂ = Bm·L⁻ᵀ →
// Q̂ is the A operand and Ŵ the B operand of the update.  Only C needs P0ᵀ (one more transpose per row tile).
// The 8×8 LDLᵀ is distributed over the warp (row g on lanes (g,·)); the inverse factor is carried along by the same
// eliminations, so the sequential part of a block is 8 pivots of ~100 cycles, overlapped by the other warps' products.
//
// FP64 work per block of 8 steps at NT = 8 row tiles: 232 DMMA (= 1 856 DFMA-equivalents) + ≈ 360 scalar issues, against
// 2 688 issues for 8 steps of the scalar kernel — and no per-entry operand traffic: 1 operand load per 8 DMMAs.
#pragma once
#include "celerite.cuh"

namespace pioran {

constexpr int BLK = 8;            // steps per block
constexpr int BLK_NSTAGE = 3;     // TMA stages (one block each)
constexpr int ROW_AUG = 4;        // RowKind of the augmented (data) row
constexpr int ROW_AUG2 = 5;       // second data row, right-hand side −1 per step: its innovations are ∂z/∂μ (gradient kernel)

// Rows of the blocked state: R celerite rows + the augmented row RG = R.  NT = tiles that carry celerite rows (columns of
// the state), NTR = row tiles (NT, or NT+1 when R is a multiple of 8 and the augmented row needs a tile of its own).
__host__ __device__ constexpr int blk_nt(int R) { return (R + 7) / 8; }
__host__ __device__ constexpr int blk_ntr(int R) { return (R + 8) / 8; }
// Block record (doubles):  UT [NT][8 steps][8 rows] | VH [8·NTR rows][8 steps] | PSI8 [8·NTR] | H2 [4·NT row pairs][32 slots][2] |
// y[8] σ²[8] mask[8] pad[8].  Rows of UT/VH/PSI8 are PHYSICAL rows (blk_phys_row); H2 is indexed by the LOGICAL celerite row.
__host__ __device__ constexpr int blk_off_vh(int NT, int NTR) { return 64 * NT; }
__host__ __device__ constexpr int blk_off_psi(int NT, int NTR) { return 64 * NT + 64 * NTR; }
__host__ __device__ constexpr int blk_off_h(int NT, int NTR) { return 64 * NT + 72 * NTR; }
__host__ __device__ constexpr int blk_off_sc(int NT, int NTR) { return 64 * NT + 72 * NTR + 256 * NT; }
__host__ __device__ constexpr int blk_doubles(int NT, int NTR) { return 64 * NT + 72 * NTR + 256 * NT + 32; }
// HALF mode: the last column tile carries at most 4 celerite rows.  They sit at its even positions and the augmented row at
// position 1, so the odd contraction chunk of that tile multiplies zeros only and its products are skipped.
__host__ __device__ constexpr bool blk_half(int R) { return (R % 8) >= 1 && (R % 8) <= 4; }
__host__ __device__ constexpr int blk_phys_row(int r, int R) {          // logical row (R = the augmented row) → physical row
    const int base = 8 * ((R + 7) / 8 - 1);
    if (!blk_half(R) || r < base) return r;
    return r == R ? base + 1 : base + 2 * (r - base);
}
__host__ __device__ constexpr int pair_slot(int s, int sp) { return s * (s - 1) / 2 + sp; }   // s > sp

// Physical row of the second data row (blocked_grad.cuh), or −1 when the last tile has no free row left (R ≡ 7 mod 8).
__host__ __device__ constexpr int blk_phys_row2(int R) {
    const int base = 8 * ((R + 7) / 8 - 1);
    if (R % 8 == 0) return R + 1;                 // the data rows have a tile of their own
    if (blk_half(R)) return base + 3;
    return (R % 8 == 7) ? -1 : R + 1;
}
// Row layout of a block table.  The likelihood needs one data row (RG), the gradient two (RG, RM); when the last column tile
// has no room for them (R ≡ 0 mod 8; for the gradient also R ≡ 7) both move to a row tile of their own.
struct BlkLayout { int NT, NTR, RG, RM; bool half; };
__host__ __device__ constexpr BlkLayout blk_layout(int R, bool grad) {
    const int NT = (R + 7) / 8;
    if (R % 8 == 0 || (grad && R % 8 == 7)) return BlkLayout{NT, NT + 1, 8 * NT, 8 * NT + 1, false};
    return BlkLayout{NT, NT, blk_phys_row(R, R), blk_phys_row2(R), blk_half(R)};
}

// ------------------------------------------------------------------------------------------- K0b: block table
// Record entries of one (block, physical row).  kind: RowKind / ROW_AUG / ROW_AUG2; (cdec, dfreq): decay rate and angular frequency
// of the row's term; (pa, pb): Ũ = pa·cos + pb·sin on the cos-row, pa·sin − pb·cos on the sin-row, pa on a real row — the
// reference's U (celerite_solver.jl:59-60) divided by the row's amplitude; lr: logical row index (K_blk table).  The table is
// zero-filled before the launch.
__device__ __forceinline__ void blocked_table_fill(double* __restrict__ tab, const double* __restrict__ t,
                                                   const double* __restrict__ y, const double* __restrict__ s2, const int64_t N,
                                                   const int64_t n0, const int r, const int kind, const double cdec,
                                                   const double dfreq, const double pa, const double pb, const int lr,
                                                   const int NT, const int NTR) {
    const int RPT = 8 * NTR, RP = 8 * NT;
    double ph[BLK], ut[BLK], vv[BLK];
#pragma unroll
    for (int s = 0; s < BLK; s++) {
        const int64_t n = n0 + s;
        ph[s] = 1.0; ut[s] = 0.0; vv[s] = 0.0;
        if (kind == ROW_PAD) { ph[s] = 0.0; continue; }
        if (n >= N) continue;                         // padded step: no decay, no coupling
        if (kind == ROW_AUG) { vv[s] = y[n]; continue; }
        if (kind == ROW_AUG2) { vv[s] = -1.0; continue; }
        const double tn = t[n];
        ph[s] = (n >= 1) ? exp(-cdec * (tn - t[n - 1])) : 0.0;      // celerite_solver.jl:54 (φ_0 := 0: nothing precedes step 0)
        if (kind == ROW_REAL) { ut[s] = pa; vv[s] = 1.0; }
        else {
            double si, co;
            sincos_large(dfreq * tn, &si, &co);                       // celerite_solver.jl:52-53 (absolute time)
            if (kind == ROW_COS) { ut[s] = fma(pb, si, pa * co); vv[s] = co; }     // (a·co + b·si)/amp
            else                 { ut[s] = fma(-pb, co, pa * si); vv[s] = si; }    // (a·si − b·co)/amp
        }
    }
    // Ψ_{0→s} and Ψ_{s→8}
    double p0[BLK], pe[BLK];
    p0[0] = ph[0];
#pragma unroll
    for (int s = 1; s < BLK; s++) p0[s] = p0[s - 1] * ph[s];
    pe[BLK - 1] = 1.0;
#pragma unroll
    for (int s = BLK - 2; s >= 0; s--) pe[s] = pe[s + 1] * ph[s + 1];
    if (r < RP) {
        const int K = r >> 3, rr = r & 7;
#pragma unroll
        for (int s = 0; s < BLK; s++) tab[K * 64 + s * 8 + rr] = p0[s] * ut[s];
    }
    // H2[logical row pair][slot][2]: coupling of steps (s, s') through this row; rows without a logical index stay zero (memset)
    if (kind == ROW_COS || kind == ROW_SIN || kind == ROW_REAL) {
        double* H = tab + blk_off_h(NT, NTR) + (lr >> 1) * 64 + (lr & 1);
#pragma unroll
        for (int s = 1; s < BLK; s++) {
            double dec = 1.0;
#pragma unroll
            for (int sp = s - 1; sp >= 0; sp--) {
                dec *= ph[sp + 1];
                H[pair_slot(s, sp) * 2] = ut[s] * dec * vv[sp];
            }
        }
    }
#pragma unroll
    for (int s = 0; s < BLK; s++) tab[blk_off_vh(NT, NTR) + r * 8 + s] = pe[s] * vv[s];
    tab[blk_off_psi(NT, NTR) + r] = (kind == ROW_AUG || kind == ROW_AUG2) ? 1.0 : p0[BLK - 1];
    for (int q = r; q < 32; q += RPT) {
        const int s = q & 7, f = q >> 3;
        const int64_t n = n0 + s;
        double sc = 0.0;
        if (n < N) sc = (f == 0) ? y[n] : (f == 1) ? s2[n] : (f == 2) ? 1.0 : 0.0;
        tab[blk_off_sc(NT, NTR) + q] = sc;
    }
}

// Shared table of the approx path: one thread per (block, physical row).  rows[] describes the 8·NTR physical rows (ROW_PAD where
// nothing lives, ROW_AUG / ROW_AUG2 at the data rows); RowDesc::term holds the logical row index of a celerite row.
__global__ void blocked_table_kernel(double* __restrict__ table, const double* __restrict__ t, const double* __restrict__ y,
                                     const double* __restrict__ s2, int64_t N, int64_t nblocks,
                                     const RowDesc* __restrict__ rows, int NT, int NTR) {
    const int RPT = 8 * NTR;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nblocks * RPT) return;
    const int64_t b = gid / RPT;
    const int r = (int)(gid - b * RPT);
    const RowDesc rd = rows[r];
    blocked_table_fill(table + b * (int64_t)blk_doubles(NT, NTR), t, y, s2, N, b * BLK, r, rd.kind, rd.c, rd.d, 1.0, rd.ratio, rd.term,
                       NT, NTR);
}

// Row amplitude of an explicit-coefficient term: a, or b when a = 0 (a pure sine term), or 1 for an empty term.
__host__ __device__ inline double blk_term_scale(double a, double b) { return a != 0.0 ? a : (b != 0.0 ? b : 1.0); }

// Per-θ tables for explicit coefficients (CARMA, PSD features, the `:celerite_gpu` drop-in): the decay rates and frequencies
// depend on θ, so every parameter vector gets its own table.  One thread per (θ, block, physical row).  prow[r] = {kind, term,
// logical row} of physical row r (the pattern of real / complex terms is the batch's, make_term_rows).
struct BlkRowMap { int kind, term, lrow; };
__global__ void blocked_table_theta_kernel(double* __restrict__ tables, int64_t table_stride, int B, const double* __restrict__ t,
                                           const double* __restrict__ y, const double* __restrict__ s2, int64_t N, int64_t nblocks,
                                           const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
                                           const double* __restrict__ d, int Jt, const BlkRowMap* __restrict__ prow, int NT, int NTR) {
    const int RPT = 8 * NTR;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)B * nblocks * RPT) return;
    const int64_t i = gid / (nblocks * RPT), rem = gid - i * nblocks * RPT;
    const int64_t blk = rem / RPT;
    const int r = (int)(rem - blk * RPT);
    const BlkRowMap rm = prow[r];
    double cdec = 0.0, dfreq = 0.0, pa = 1.0, pb = 0.0;
    if (rm.term >= 0) {
        const size_t k = (size_t)i * Jt + rm.term;
        const double sc = blk_term_scale(a[k], b[k]);
        cdec = c[k]; dfreq = d[k]; pa = a[k] / sc; pb = b[k] / sc;
    }
    blocked_table_fill(tables + i * table_stride + blk * (int64_t)blk_doubles(NT, NTR), t, y, s2, N, blk * BLK, r, rm.kind, cdec, dfreq,
                       pa, pb, rm.lrow, NT, NTR);
}
// Row amplitudes (logical order, stride RPA) and Σa (celerite_solver.jl:21) of the same batch.  One thread per θ.
__global__ void blocked_amp_theta_kernel(double* __restrict__ amp, double* __restrict__ suma, int B, int RPA, int R,
                                         const double* __restrict__ a, const double* __restrict__ b, int Jt,
                                         const int* __restrict__ lrow_term) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double sa = 0.0;
    for (int m = 0; m < Jt; m++) sa += a[(size_t)i * Jt + m];
    suma[i] = sa;
    for (int k = 0; k < RPA; k++) {
        double v = 0.0;
        if (k < R) { const size_t q = (size_t)i * Jt + lrow_term[k]; v = blk_term_scale(a[q], b[q]); }
        amp[(size_t)i * RPA + k] = v;
    }
}

// ------------------------------------------------------------------------------------------- device helpers
// D(8×8) += A(8×4)·B(4×8), FP64 tensor pipe.  Fragments: A lane(g,t) = A[g][t];  B lane(g,t) = B[t][g];  C/D lane(g,t) = [g][2t], [g][2t+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct BlkLane {
    int g, t;
    bool odd;          // g & 1
    int src1, src2;    // transpose: source lanes of the two exchanges
    int csrc0, csrc1;  // lane whose K_blk slot is C[g][2t], C[g][2t+1]
    bool cdiag0, cdiag1;
};
__device__ __forceinline__ BlkLane make_blk_lane(int lane) {
    BlkLane L;
    L.g = lane >> 2; L.t = lane & 3;
    L.odd = (L.g & 1) != 0;
    L.src1 = 4 * (2 * L.t + (L.g & 1)) + (L.g >> 1);
    L.src2 = 4 * (2 * L.t + 1 - (L.g & 1)) + (L.g >> 1);
    const int c0 = 2 * L.t, c1 = 2 * L.t + 1;
    const int h0 = max(L.g, c0), l0 = min(L.g, c0), h1 = max(L.g, c1), l1 = min(L.g, c1);
    L.cdiag0 = (L.g == c0); L.cdiag1 = (L.g == c1);
    L.csrc0 = L.cdiag0 ? 0 : pair_slot(h0, l0);
    L.csrc1 = L.cdiag1 ? 0 : pair_slot(h1, l1);
    return L;
}
// Transpose of an 8×8 tile held in the accumulator layout: 2 exchanges (each lane's two values come from two lanes).
__device__ __forceinline__ void tile_transpose(const BlkLane& L, const double e0, const double e1, double& f0, double& f1) {
    const double r1 = __shfl_sync(FULL, L.odd ? e1 : e0, L.src1);
    const double r2 = __shfl_sync(FULL, L.odd ? e0 : e1, L.src2);
    f0 = L.odd ? r2 : r1;
    f1 = L.odd ? r1 : r2;
}

// Per-lane persistent state of one evaluation.
template <int NT, int NTR>
struct BlkState {
    double x[NTR][NT][2];   // tiles I ≥ K only (the others are never touched and take no registers)
    double chi2, logacc, dkeep, dfirst;
};

#ifndef PIORAN_BLK_PREFETCH
#define PIORAN_BLK_PREFETCH 0      // 1: K_blk of block b+1 is formed during block b (measured slower: 41.0 vs 38.5 ms, profiles/r02_k2t_variants.txt)
#endif
#ifndef PIORAN_BLK_EARLY_DECAY
#define PIORAN_BLK_EARLY_DECAY 1   // (ψ8ψ8ᵀ)∘X is applied right after P0, ahead of the pivot chain, not inside the update
#endif

// K_blk of one block in the accumulator layout (lane (g,t): C[g][2t], C[g][2t+1]).  Lane l < 28 sums pair slot l over all
// celerite rows (4 accumulation chains, no cross-lane reduction); two exchanges then hand every lane its two entries.  The
// diagonal is A_n = Σa + ν σ²_n.  amp_l: per-warp amplitudes by LOGICAL row (8·NT doubles, zero-padded).
template <int NT, int NTR, bool TAN = false>
__device__ __forceinline__ void blk_kblk(const double* __restrict__ tab, const double* __restrict__ amp_l, const BlkLane& L,
                                         const int lane, const double suma, const double nu, const int64_t n0,
                                         const int64_t N, const double* __restrict__ sb, double& cm0, double& cm1,
                                         const bool diag_given = false, const double diag_value = 0.0) {
    constexpr int O_H = blk_off_h(NT, NTR), O_SC = blk_off_sc(NT, NTR);
    const int g = L.g;
    const double2* Hq = reinterpret_cast<const double2*>(tab + O_H) + lane;
    const double2* Aq = reinterpret_cast<const double2*>(amp_l);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int jp = 0; jp < 4 * NT; jp += 2) {
        const double2 h0 = Hq[jp * 32], h1 = Hq[(jp + 1) * 32];
        const double2 m0 = Aq[jp], m1 = Aq[jp + 1];
        a0 = fma(m0.x, h0.x, a0); a1 = fma(m0.y, h0.y, a1);
        a2 = fma(m1.x, h1.x, a2); a3 = fma(m1.y, h1.y, a3);
    }
    const double v = (a0 + a1) + (a2 + a3);
    const double o0 = __shfl_sync(FULL, v, L.csrc0);
    const double o1 = __shfl_sync(FULL, v, L.csrc1);
    // diagonal: A_n = Σa + ν σ²_n (celerite_solver.jl:92); padded steps are unit pivots
    const int64_t n = n0 + g;
    const double mk = tab[O_SC + 16 + g];
    const double s2v = sb ? (n < N ? sb[n] : 0.0) : tab[O_SC + 8 + g];
    // TAN: (suma, nu) are the tangents of (Σa, ν); the unit pivot of a padded step has no tangent
    const double dg = diag_given ? diag_value * mk : TAN ? fma(nu, s2v, suma) * mk : fma(fma(nu, s2v, suma), mk, 1.0 - mk);
    cm0 = L.cdiag0 ? dg : o0;
    cm1 = L.cdiag1 ? dg : o1;
}

// One block of 8 steps.  tab: this block's record in shared memory; amp_s: per-warp amplitudes (8·NTR doubles, amp[RG] = 1);
// (cm0, cm1): K_blk of this block on entry, of the NEXT block (record tabn, first step n0 + 8) on exit.
// PUB (gradient kernel): the block's L⁻¹, 1/D, Bm and Q̂ are also written to `pub` (item i of lane l at pub[32·i + l]: the
// tangent warps read them back lane for lane) and Σ 2 z ż/D of the second data row RM is added to *chimu.
template <int NT, int NTR, bool HALF, bool PUB = false>
__device__ __forceinline__ void blocked_step(BlkState<NT, NTR>& st, const double* __restrict__ tab,
                                             const double* __restrict__ tabn, const double* __restrict__ amp_s,
                                             const double* __restrict__ amp_l, const BlkLane& L, const int lane,
                                             const double suma, const double mu, const double nu,
                                             const int64_t n0, const int64_t N, const double* __restrict__ yb,
                                             const double* __restrict__ sb, const int RG, double& cm0_io, double& cm1_io,
                                             double* __restrict__ pub = nullptr, const int RM = -1, double* chimu = nullptr) {
    constexpr int O_VH = blk_off_vh(NT, NTR), O_PSI = blk_off_psi(NT, NTR), O_SC = blk_off_sc(NT, NTR);
    const int g = L.g, t = L.t;
    double cm0 = cm0_io, cm1 = cm1_io;
#if !PIORAN_BLK_PREFETCH
    blk_kblk<NT, NTR>(tab, amp_l, L, lane, suma, nu, n0, N, sb, cm0, cm1);
#endif

    // ---- P0 = X·Û (row tiles I, contraction over column tiles K)
    double P0[NTR][2];
#pragma unroll
    for (int I = 0; I < NTR; I++) P0[I][0] = P0[I][1] = 0.0;
#pragma unroll
    for (int K = 0; K < NT; K++) {
        const double2 u = *reinterpret_cast<const double2*>(tab + K * 64 + g * 8 + 2 * t);
#pragma unroll
        for (int I = 0; I < NTR; I++) {
            double a0, a1;
            if (I >= K) { a0 = st.x[I][K][0]; a1 = st.x[I][K][1]; }
            else tile_transpose(L, st.x[K][I][0], st.x[K][I][1], a0, a1);
            dmma(P0[I][0], P0[I][1], a0, u.x);
            if (!(HALF && K == NT - 1)) dmma(P0[I][0], P0[I][1], a1, u.y);   // HALF: the odd chunk of the last tile is all zeros
        }
    }

    // ---- C = K_blk − Ûᵀ·P0 (needs P0ᵀ as the B operand), two accumulation chains
    {
        double ca0 = 0.0, ca1 = 0.0, cb0 = 0.0, cb1 = 0.0;
#pragma unroll
        for (int J = 0; J < NT; J++) {
            const double2 u = *reinterpret_cast<const double2*>(tab + J * 64 + g * 8 + 2 * t);
            double p0, p1;
            tile_transpose(L, P0[J][0], P0[J][1], p0, p1);
            const bool both = !(HALF && J == NT - 1);
            if (J & 1) { dmma(cb0, cb1, u.x, p0); if (both) dmma(cb0, cb1, u.y, p1); }
            else       { dmma(ca0, ca1, u.x, p0); if (both) dmma(ca0, ca1, u.y, p1); }
        }
        cm0 -= ca0 + cb0;
        cm1 -= ca1 + cb1;
    }

    double psr[NTR];
#pragma unroll
    for (int I = 0; I < NTR; I++) psr[I] = tab[O_PSI + 8 * I + g];
#if PIORAN_BLK_EARLY_DECAY
    // ---- X ← (ψ8ψ8ᵀ)∘X now: the un-decayed state is no longer needed, and these products fill the pivot chain's latency
#pragma unroll
    for (int K = 0; K < NT; K++) {
        const double2 pc = *reinterpret_cast<const double2*>(tab + O_PSI + 8 * K + 2 * t);
#pragma unroll
        for (int I = K; I < NTR; I++) {
            st.x[I][K][0] *= psr[I] * pc.x;
            st.x[I][K][1] *= psr[I] * pc.y;
        }
    }
#endif
#if PIORAN_BLK_PREFETCH
    double cn0, cn1;
    blk_kblk<NT, NTR>(tabn, amp_l, L, lane, suma, nu, n0 + BLK, N, sb, cn0, cn1);
    cm0_io = cn0; cm1_io = cn1;
#endif

    // ---- Bm = amp∘V̂ − ψ8∘P0 (in place); the augmented row carries (y − μ) − prediction
#pragma unroll
    for (int I = 0; I < NTR; I++) {
        const int row = 8 * I + g;
        double2 vh = *reinterpret_cast<const double2*>(tab + O_VH + row * 8 + 2 * t);
        const double am = amp_s[row];
        if (I == NTR - 1) {
            const bool isrg = (row == RG);
            const double2 mk = *reinterpret_cast<const double2*>(tab + O_SC + 16 + 2 * t);
            if (yb && isrg) {
                const int64_t n = n0 + 2 * t;
                vh.x = (n < N) ? yb[n] : 0.0;
                vh.y = (n + 1 < N) ? yb[n + 1] : 0.0;
            }
            const double m_ = isrg ? mu : 0.0;
            vh.x = fma(-m_, mk.x, vh.x);
            vh.y = fma(-m_, mk.y, vh.y);
        }
        P0[I][0] = fma(-psr[I], P0[I][0], am * vh.x);
        P0[I][1] = fma(-psr[I], P0[I][1], am * vh.y);
    }

    // ---- 8×8 LDLᵀ of C, distributed: lane (g,t) holds C[g][2t], C[g][2t+1]; E becomes L⁻¹ by the same eliminations
    double e0 = L.cdiag0 ? 1.0 : 0.0, e1 = L.cdiag1 ? 1.0 : 0.0;
    double rd0 = 0.0, rd1 = 0.0;      // 1/D of this lane's two steps (2t, 2t+1)
    const int rowbase = lane & ~3;
    const int ring = (int)(n0 & 31);
#pragma unroll
    for (int j = 0; j < BLK; j++) {
        const int tj = j >> 1;
        const double cme = (j & 1) ? cm1 : cm0;
        const double dj = __shfl_sync(FULL, cme, 4 * j + tj);          // pivot D_{n0+j} (celerite_solver.jl:92)
        const double cgj = __shfl_sync(FULL, cme, rowbase | tj);       // C[g][j]
        const double rdj = fast_rcp(dj);
        if (j & 1) rd1 = (t == tj) ? rdj : rd1;
        else       rd0 = (t == tj) ? rdj : rd0;
        const double l = cgj * rdj;
        const double cj0 = __shfl_sync(FULL, cm0, 4 * j + t), cj1 = __shfl_sync(FULL, cm1, 4 * j + t);
        const double ej0 = __shfl_sync(FULL, e0, 4 * j + t), ej1 = __shfl_sync(FULL, e1, 4 * j + t);
        const double lm = (g > j) ? l : 0.0;
        cm0 = fma(-lm, cj0, cm0); cm1 = fma(-lm, cj1, cm1);
        e0 = fma(-lm, ej0, e0);   e1 = fma(-lm, ej1, e1);
        // log|D_n| (celerite_solver.jl:140; no abs on the first pivot, :126): lane n%32 keeps D_n, one log per 32 steps
        if (j == 0 && n0 == 0) st.dfirst = dj;
        else if (lane == ring + j) st.dkeep = dj;
    }
    if (ring == 24) { st.logacc += log(fabs(st.dkeep)); st.dkeep = 1.0; }

    // ---- Q̂ = Bm·L⁻ᵀ
    double Q[NTR][2];
#pragma unroll
    for (int I = 0; I < NTR; I++) {
        Q[I][0] = Q[I][1] = 0.0;
        dmma(Q[I][0], Q[I][1], P0[I][0], e0);
        dmma(Q[I][0], Q[I][1], P0[I][1], e1);
    }

    if (PUB) {
        double* p = pub + lane;
        p[0] = e0; p[32] = e1; p[64] = rd0; p[96] = rd1;
#pragma unroll
        for (int I = 0; I < NTR; I++) {
            p[(4 + 2 * I) * 32] = P0[I][0]; p[(5 + 2 * I) * 32] = P0[I][1];                      // Bm
            p[(4 + 2 * NTR + 2 * I) * 32] = Q[I][0]; p[(5 + 2 * NTR + 2 * I) * 32] = Q[I][1];    // Q̂
        }
        // ∂χ²/∂μ = Σ_s 2 z_s ż_s / D_s with ż the innovations of the second data row (same tile, row RM)
        const int srcm = ((RM & 7) << 2) | t;
        const double zm0 = __shfl_sync(FULL, Q[NTR - 1][0], srcm), zm1 = __shfl_sync(FULL, Q[NTR - 1][1], srcm);
        const bool isrg2 = (8 * (NTR - 1) + g == RG);
        const double cmv = 2.0 * fma(Q[NTR - 1][0] * rd0, zm0, (Q[NTR - 1][1] * rd1) * zm1);
        *chimu += isrg2 ? cmv : 0.0;
    }

    // ---- yᵀK⁻¹y += Σ_s z_s²/D_s  (celerite_solver.jl:333) on the lanes that hold the augmented row
    {
        const bool isrg = (8 * (NTR - 1) + g == RG);
        const double zz = fma(Q[NTR - 1][0] * rd0, Q[NTR - 1][0], (Q[NTR - 1][1] * rd1) * Q[NTR - 1][1]);
        st.chi2 += isrg ? zz : 0.0;
    }

    // ---- X ← (ψ8ψ8ᵀ)∘X + Q̂·Ŵᵀ
#pragma unroll
    for (int K = 0; K < NT; K++) {
        const double w0 = Q[K][0] * rd0, w1 = Q[K][1] * rd1;
#if !PIORAN_BLK_EARLY_DECAY
        const double2 pc = *reinterpret_cast<const double2*>(tab + O_PSI + 8 * K + 2 * t);
#endif
#pragma unroll
        for (int I = K; I < NTR; I++) {
#if PIORAN_BLK_EARLY_DECAY
            double x0 = st.x[I][K][0], x1 = st.x[I][K][1];
#else
            double x0 = st.x[I][K][0] * (psr[I] * pc.x), x1 = st.x[I][K][1] * (psr[I] * pc.y);
#endif
            dmma(x0, x1, Q[I][0], w0);
            dmma(x0, x1, Q[I][1], w1);
            st.x[I][K][0] = x0; st.x[I][K][1] = x1;
        }
    }
}

// ------------------------------------------------------------------------------------------- kernel
// grid = work items; block = NW warps, one parameter vector each; dynamic smem: BLK_NSTAGE block records | NW × 8·NTR amplitudes |
// BLK_NSTAGE mbarriers | BLK_NSTAGE stage counters.  Same stage hand-back as celerite_shared_kernel (last warp out refills).
template <int NT, int NTR, bool HALF, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) celerite_blocked_kernel(const BatchArgs args, const int R, const int amp_stride, const int RG) {
    constexpr int BD = blk_doubles(NT, NTR), RPT = 8 * NTR, APW = RPT + 8 * NT;   // per warp: amplitudes by physical and by logical row
    constexpr uint32_t STAGE_BYTES = BD * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    double* amps = stages + BLK_NSTAGE * BD;
    uint64_t* bars = reinterpret_cast<uint64_t*>(amps + NW * APW);
    int* done = reinterpret_cast<int*>(bars + BLK_NSTAGE);

    const WorkItem wk = args.work[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t N = wk.N;
    const int64_t nblocks = (N + BLK - 1) / BLK;

    if (threadIdx.x == 0) {
        for (int k = 0; k < BLK_NSTAGE; k++) { mbar_init(&bars[k], 1); done[k] = 0; }
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < BLK_NSTAGE && k < nblocks; k++) {
            mbar_arrive_expect_tx(&bars[k], STAGE_BYTES);
            tma_load_1d(stages + k * BD, wk.table + (size_t)k * BD, STAGE_BYTES, &bars[k]);
        }
    }

    // Surplus warps of a partly filled item leave: the stage hand-back counts the item's own warps, and a tail item of
    // wk.count <= 4 then runs one warp per sub-partition at the single-warp rate instead of sharing the pipe with redundant sweeps.
    const int nact = wk.count < NW ? wk.count : NW;
    if (warp >= nact) return;
    const int slot = warp;
    const int th = wk.theta_begin + slot;
    const BlkLane L = make_blk_lane(lane);

    double* amp_s = amps + warp * APW;
    double* amp_l = amp_s + RPT;
    for (int k = lane; k < APW; k += 32) amp_s[k] = 0.0;
    __syncwarp();
    for (int k = lane; k < R; k += 32) {
        const double av = args.amp[(size_t)th * amp_stride + k];
        amp_s[blk_phys_row(k, R)] = av;
        amp_l[k] = av;
    }
    if (lane == 0) amp_s[RG] = 1.0;
    __syncwarp();
    const double suma = args.suma[th];
    const size_t pi = (size_t)wk.par_begin + slot;
    const double mu = args.mu ? args.mu[pi * args.pstride] : 0.0;
    const double nu = args.nu ? args.nu[pi * args.pstride] : 1.0;
    const double* yb = args.y_batch ? args.y_batch + pi * args.ystride : nullptr;
    const double* sb = args.s2_batch ? args.s2_batch + pi * args.ystride : nullptr;

    BlkState<NT, NTR> st;
#pragma unroll
    for (int I = 0; I < NTR; I++)
#pragma unroll
        for (int K = 0; K < NT; K++) st.x[I][K][0] = st.x[I][K][1] = 0.0;
    st.chi2 = 0.0; st.logacc = 0.0; st.dkeep = 1.0; st.dfirst = 1.0;

    int sidx = 0;
    uint32_t parity = 0;
    double cm0 = 0.0, cm1 = 0.0;
    mbar_wait(&bars[0], 0);
#if PIORAN_BLK_PREFETCH
    blk_kblk<NT, NTR>(stages, amp_l, L, lane, suma, nu, 0, N, sb, cm0, cm1);
#endif
    for (int64_t b = 0; b < nblocks; b++) {
        // the next block's record (its K_blk is formed during this block); the last block points at itself
        int nidx = sidx + 1;
        uint32_t npar = parity;
        if (nidx == BLK_NSTAGE) { nidx = 0; npar ^= 1; }
        if (b + 1 < nblocks) mbar_wait(&bars[nidx], npar);
        else nidx = sidx;
        blocked_step<NT, NTR, HALF>(st, stages + sidx * BD, stages + nidx * BD, amp_s, amp_l, L, lane, suma, mu, nu, b * BLK, N,
                                    yb, sb, RG, cm0, cm1);
        __syncwarp();
        if (lane == 0 && b + BLK_NSTAGE < nblocks) {
            __threadfence_block();
            if (atomicAdd(&done[sidx], 1) == nact - 1) {
                done[sidx] = 0;
                fence_proxy_async();
                mbar_arrive_expect_tx(&bars[sidx], STAGE_BYTES);
                tma_load_1d(stages + sidx * BD, wk.table + (size_t)(b + BLK_NSTAGE) * BD, STAGE_BYTES, &bars[sidx]);
            }
        }
        if (++sidx == BLK_NSTAGE) { sidx = 0; parity ^= 1; }
    }
    // Σ log|D_n| (first pivot without abs) and the χ² term
    double la = st.logacc + log(fabs(st.dkeep));
    double ch = st.chi2;
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) {
        la += __shfl_xor_sync(FULL, la, sft);
        ch += __shfl_xor_sync(FULL, ch, sft);
    }
    const double dfirst = __shfl_sync(FULL, st.dfirst, 0);
    const double logdet = log(dfirst) + la;
    // celerite_solver.jl:333
    const double res = -logdet / 2 - (double)N * 1.8378770664093453 / 2 - ch / 2;
    if (lane == 0) args.out[wk.out_begin + warp] = res;
}

}  // namespace pioran
