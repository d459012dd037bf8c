// common.cuh — shared definitions of the B200 celerite backend (device + host).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pioran {

// ---------------------------------------------------------------------------------------------------------
// Row layout.  A sum of celerite terms is turned into R "rows" (the rank of the semiseparable part):
//   complex term (b≠0 or d≠0) → 2 rows (cos-row, sin-row);  real term (b=d=0) → 1 row.
// The reference always uses 2 rows per term (src/celerite_solver.jl:20, R = 2J) and carries identically-zero
// rows for real terms; dropping them changes no arithmetic result.
// Rows are grouped into G = 8 blocks of BS rows: R_pad = 8·BS ≥ R, padded rows are all-zero.
// ---------------------------------------------------------------------------------------------------------
constexpr int G = 8;

enum RowKind : int { ROW_COS = 0, ROW_SIN = 1, ROW_REAL = 2, ROW_PAD = 3 };

// θ-independent description of one row (shared-table mode): which celerite term, which component.
struct RowDesc {
    double c;      // decay rate of the term
    double d;      // angular frequency of the term
    double ratio;  // b/a of the term (SHO: 1, DRWCelerite celerite part: √3, real terms: 0)
    int kind;      // RowKind
    int term;      // index of the amplitude this row scales with
};

// Shared-memory layout of a row vector: block g (BS rows) starts at g·BSP doubles with BSP even and
// BSP/2 odd, so that (i) every block starts 16-byte aligned (128-bit shared loads) and (ii) the 8 block starts fall
// into 8 different 16-byte bank groups — a warp reading "its" block (8 distinct blocks, 4 lanes each) is
// conflict-free.  (BS = 8 unpadded: stride 64 B → 4-way bank conflict on every load, measured in
// profiles/r01_k2_bs8_before_padding.txt.)
__host__ __device__ constexpr int bsp_of(int bs) { return bs <= 6 ? 6 : 10; }
__host__ __device__ constexpr int rps_of(int bs) { return G * bsp_of(bs); }
__host__ __device__ constexpr int pad_index(int j, int bs) { return (j / bs) * bsp_of(bs) + (j % bs); }

// Per-step record of the series table (doubles).  Six row vectors of RPS = 8·BSP entries (padded layout) then 8 scalars.
//   UT  : Ũ_n            (U_n = amp ∘ Ũ_n;  src/celerite_solver.jl:59-60 with a factored out)
//   UH  : φ_n ∘ Ũ_n
//   KAP : φ_n ∘ φ_{n-1}
//   PHI : φ_n = exp(-c (t_n - t_{n-1}))   (φ_0 = 0;  src/celerite_solver.jl:54-57)
//   V   : cos/sin(d t_n) (or 1 for real rows;  src/celerite_solver.jl:62-63)
//   PHN : φ_{n+1}        (0 at the last step)
//   scalars: [0] y_n  [1] σ²_n  [2] t_n
__host__ __device__ constexpr int table_step_doubles(int rps) { return 6 * rps + 8; }
enum TableField : int { F_UT = 0, F_UH = 1, F_KAP = 2, F_PHI = 3, F_V = 4, F_PHN = 5 };

constexpr int CHUNK_STEPS = 16;  // steps per TMA stage
constexpr int SCAN_LD = 64;      // leading dimension of the K3 state matrices (rank ≤ 64)

// One unit of work of the batched kernels: `count` parameter vectors of one series (or, for the chunked re-filter of
// the long-series path K3, one parameter vector on the step range [n_begin, n_end) starting from a given state).
struct WorkItem {
    const double* table;   // series table (shared mode) or nullptr (generic mode)
    const double* t;       // series arrays (device)
    const double* y;
    const double* s2;
    int64_t N;
    int theta_begin;       // first row of the per-(series,θ) arrays (amp, Σa) handled by this CTA
    int par_begin;         // first row of the per-θ scalar arrays (μ, ν, y_batch) handled by this CTA
    int count;             // number of θ (≤ warps per CTA)
    int out_begin;         // first index of logl_out
    // K3 pass 3 (generic kernel, one work item per warp): sweep steps [n_begin, n_end) only, starting from the state
    // `init` = S (SCAN_LD×SCAN_LD row-major over the logical rows) followed by g (SCAN_LD) entering step n_begin, and write the
    // partial sums (Σ log|D_n|, Σ z_n²/D_n) to `part` instead of a log-likelihood.  init == nullptr: whole series.
    int64_t n_begin, n_end;
    const double* init;
    double* part;
    // K3 self-check: the sums of the first n_head steps go to chk[0..1] as well, and the sweep continues over the n_ext
    // steps after n_end into chk[2..3] (not part of `part`).  The next work item sweeps those same steps from the state the
    // scan handed it: the two pairs agree to rounding exactly when that state is the one this sweep arrives at.
    int n_head, n_ext;
    double* chk;
    // K3 refinement: the sweep starts n_warm steps before the sub-chunk (n_begin and `init` refer to that earlier point) and
    // the sums of those steps are discarded — the filter forgets the error of the injected state while it runs up.
    int64_t n_warm;
    // K3 Newton refinement: when non-null, the state entering step n_end (S | g, same layout as `init`; entries outside the
    // 8·BS logical rows are not written) is stored here before the look-ahead steps — the exit state of the sweep, i.e. the
    // exact recursion applied to `init`.  Needs an even number of steps in [n_begin, n_end) and n_end < N.
    double* exit;
};

// sin and cos of a large FP64 argument.  The celerite rows take cos/sin(d_j·t_n) at ABSOLUTE times
// (src/celerite_solver.jl:52-53): arguments reach 1e5…1e9 rad, where CUDA's sincos() leaves its fast path (|x| > 105 615)
// for a Payne–Hanek reduction that costs an order of magnitude more.  Here: k = rint(x·2/π), three-constant Cody–Waite
// reduction with FMAs (π/2 split into 53+53+53 bits: absolute error of the reduced angle ≤ ~4e-16 for |x| < 2^31), then
// sincos() of |r| ≤ π/4 (fast path) and the quadrant fix-up.  Absolute accuracy of the results ≈ 5e-16 — what the
// covariance entries need; beyond 2^31 rad the library routine is used.
__device__ __forceinline__ void sincos_large(double x, double* s, double* c) {
    if (!(fabs(x) < 2147483648.0)) { sincos(x, s, c); return; }
    const double k = rint(x * 0.63661977236758138);                 // 2/π
    double r = fma(-k, 1.5707963267948966, x);                      // π/2 = P1 + P2 + P3
    r = fma(-k, 6.123233995736766e-17, r);
    r = fma(-k, -1.4973849048591698e-33, r);
    double sr, cr;
    sincos(r, &sr, &cr);
    const int q = (int)(long long)k & 3;
    const double ss = (q & 1) ? cr : sr, cc = (q & 1) ? sr : cr;
    *s = (q & 2) ? -ss : ss;
    *c = ((q + 1) & 2) ? -cc : cc;
}

}  // namespace pioran
