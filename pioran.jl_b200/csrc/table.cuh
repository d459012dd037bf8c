// table.cuh — K0: θ-independent per-series table for the shared-table celerite kernel.
//
// On the approx() path the decay rates c_j and frequencies d_j depend only on the spectral grid
// (src/psd.jl:250,266-267), never on θ, so every transcendental of the celerite sweep
// (cos/sin/exp of src/celerite_solver.jl:52-54) is evaluated ONCE per (series, grid) here and shared by the
// whole parameter batch.  Record layout: common.cuh (table_step_doubles / TableField).
#pragma once
#include "common.cuh"

namespace pioran {

// grid.x covers N_pad·R_pad (one thread per (step, row)); table must hold N_pad·SD doubles, N_pad a multiple
// of CHUNK_STEPS; steps n ≥ N are zero-filled.
__global__ void table_build_kernel(double* __restrict__ table, const double* __restrict__ t,
                                   const double* __restrict__ y, const double* __restrict__ s2, int64_t N,
                                   int64_t N_pad, const RowDesc* __restrict__ rows, int BS) {
    const int RP = rps_of(BS), BSP = bsp_of(BS);   // one thread per padded slot; slots ≥ BS of a block are zero
    const int SD = table_step_doubles(RP);
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= N_pad * RP) return;
    const int64_t n = gid / RP;
    const int j = (int)(gid - n * RP);
    double* Tn = table + n * SD;
    const int jb = j / BSP, jr = j - jb * BSP;
    RowDesc rd{0, 0, 0, ROW_PAD, 0};
    if (jr < BS) rd = rows[jb * BS + jr];
    double ut = 0, v = 0, ph = 0, php = 0, phn = 0;
    if (n < N && rd.kind != ROW_PAD) {
        const double tn = t[n];
        if (n >= 1) ph = exp(-rd.c * (tn - t[n - 1]));           // celerite_solver.jl:54
        if (n >= 2) php = exp(-rd.c * (t[n - 1] - t[n - 2]));
        if (n + 1 < N) phn = exp(-rd.c * (t[n + 1] - tn));
        if (rd.kind == ROW_REAL) {
            ut = 1.0; v = 1.0;                                   // d = 0: cos = 1, the sin-row vanishes
        } else {
            double si, co;
            sincos_large(rd.d * tn, &si, &co);                         // celerite_solver.jl:52-53 (absolute time)
            if (rd.kind == ROW_COS) { ut = fma(rd.ratio, si, co); v = co; }    // (a·co + b·si)/a
            else                    { ut = fma(-rd.ratio, co, si); v = si; }   // (a·si − b·co)/a
        }
    }
    Tn[F_UT * RP + j] = ut;
    Tn[F_UH * RP + j] = ph * ut;
    Tn[F_KAP * RP + j] = ph * php;
    Tn[F_PHI * RP + j] = ph;
    Tn[F_V * RP + j] = v;
    Tn[F_PHN * RP + j] = phn;
    if (j < 8) {
        double sc = 0.0;
        if (n < N) sc = (j == 0) ? y[n] : (j == 1) ? s2[n] : (j == 2) ? t[n] : 0.0;
        Tn[6 * RP + j] = sc;
    }
}

}  // namespace pioran
