"""Seeded synthetic inputs of the BASELINE.json configurations.  numpy only — used by bench.py and the
GPU tests; no reference code, no oracle.

Series: irregular sampling with gaps 0.05 + Exp(1), σ ~ U(0.01, 0.05), y = draw of the celerite GP that
approx(SingleBendingPowerLaw(0.82, 0.01, 3.3), …, SHO basis) describes (values of benchmark/benchmarks.jl:36-37) plus
white noise, drawn through the semiseparable factorisation of the covariance (numpy, O(N·R²)).

Parameter vectors: the prior transform of examples/ultranest/single_pl.jl:96-104 on a seeded unit cube.
"""
import numpy as np


def sho_coefficients(theta, f_min, f_max, J=20, variance=1.0, S_low=20.0, S_high=20.0):
    """numpy statement of approx(SBPL, …; basis_function="SHO") used only to draw synthetic data."""
    a1, f1, a2 = theta
    f0, fM = f_min / S_low, f_max * S_high
    fj = f0 * (fM / f0) ** (np.arange(J) / (J - 1))
    psd = lambda f: (f / f1) ** (-a1) / (1 + (f / f1) ** (a2 - a1))
    Bm = 1.0 / (1.0 + (fj[:, None] / fj[None, :]) ** 4)
    amp = np.linalg.solve(Bm, psd(fj) / psd(fj[0]))

    def prim(x):
        c = fj
        return c * amp / (4 * np.sqrt(2)) * (
            np.log((x * x + np.sqrt(2) * c * x + c * c) / (x * x - np.sqrt(2) * c * x + c * c))
            + 2 * np.arctan2(np.sqrt(2) * c * x, c * c - x * x))

    amp = amp * variance / np.sum(prim(f_max) - prim(f_min))
    a = amp * fj * np.pi / np.sqrt(2)
    c = np.sqrt(2) * np.pi * fj
    return a, a.copy(), c, c.copy()


def draw_celerite(t, a, b, c, d, s2, rng):
    """One realisation y ~ N(0, K + diag σ²) with K_ij = Σ exp(−cτ)(a cos dτ + b sin dτ), τ = |t_i − t_j|, in O(N·R²):
    K + diag σ² = L D Lᵀ by the semiseparable (celerite) recursion, y = L·(√D ∘ g) with g ~ N(0, I) — the same
    construction the reference's simulate uses (src/celerite_solver.jl:497-549), written in the forward form
    z_n = y_n − U_nᵀ f_n  ⇔  y_n = z_n + U_nᵀ f_n."""
    N, J = len(t), len(a)
    S = np.zeros((2 * J, 2 * J))
    f = np.zeros(2 * J)
    W = np.zeros(2 * J)
    y = np.empty(N)
    suma = a.sum()
    Dp = zp = 0.0
    g = rng.normal(size=N)
    for n in range(N):
        co, si = np.cos(d * t[n]), np.sin(d * t[n])
        U = np.concatenate([a * co + b * si, a * si - b * co])
        V = np.concatenate([co, si])
        if n == 0:
            D = suma + s2[0]
            W = V / D
        else:
            ph = np.tile(np.exp(-c * (t[n] - t[n - 1])), 2)
            S = np.outer(ph, ph) * (S + Dp * np.outer(W, W))
            f = ph * (f + W * zp)
            SU = S @ U
            D = suma + s2[n] - U @ SU
            W = (V - SU) / D
        if not D > 0:
            raise ValueError("kernel is not positive definite at the simulation parameters")
        z = np.sqrt(D) * g[n]
        y[n] = z + U @ f
        Dp, zp = D, z
    return y


def make_series(N, seed, theta0=(0.82, 0.01, 3.3), variance=1.0, J=20):
    """→ t, y, σ², f_min, f_max  (f_min, f_max as in examples/ultranest/single_pl.jl:48)."""
    rng = np.random.default_rng(seed)
    t = np.cumsum(0.05 + rng.exponential(1.0, N))
    t -= t[0]
    sig = rng.uniform(0.01, 0.05, N)
    f_min, f_max = 1.0 / (t[-1] - t[0]), 1.0 / np.min(np.diff(t)) / 2.0
    a, b, c, d = sho_coefficients(theta0, f_min, f_max, J, variance)
    y = draw_celerite(t, a, b, c, d, sig ** 2, rng)
    return t, y, sig ** 2, f_min, f_max


def prior_theta(B, f_min, f_max, ybar, ysd, seed, alpha2_max=4.0):
    """[B × 6] rows (α₁, f₁, α₂, variance, ν, μ): α₁~U(0,1.5), f₁~LogU(4f₀, f_M/4), α₂~U(α₁, α₂max),
    variance~LogNormal(−3, √2), ν~Gamma(2, 0.5), μ~N(ȳ, 5·sd)."""
    from scipy import special
    rng = np.random.default_rng(seed)
    u = rng.uniform(size=(B, 6))
    f0, fM = f_min / 20.0, f_max * 20.0
    a1 = 1.5 * u[:, 0]
    f1 = np.exp(np.log(4 * f0) + u[:, 1] * (np.log(fM / 4) - np.log(4 * f0)))
    a2 = a1 + u[:, 2] * (alpha2_max - a1)
    var = np.exp(-3.0 + np.sqrt(2.0) * special.ndtri(u[:, 3]))
    nu = 0.5 * special.gammaincinv(2.0, u[:, 4])
    mu = ybar + 5 * ysd * special.ndtri(u[:, 5])
    return np.ascontiguousarray(np.column_stack([a1, f1, a2, var, nu, mu]))


def flops_per_step(R):
    """Algorithmic FP64 flops of one time step of the forward-only celerite sweep at rank R."""
    return 4 * R * R + 13 * R + 40


def rank_of(basis, J):
    return 2 * J if basis == "SHO" else 3 * J


def make_series_fast(N, seed, taus=(3.0, 60.0, 2000.0), amps=(0.4, 0.7, 1.0)):
    """Long irregular series for the N ~ 1e6 configuration (C4) without the O(N·R²) GP draw: y = sum of three
    Ornstein–Uhlenbeck processes (red-noise-like, exact irregular-step recurrence) + white noise.
    → t, y, σ², f_min, f_max"""
    rng = np.random.default_rng(seed)
    dt = 0.05 + rng.exponential(1.0, N)
    t = np.cumsum(dt)
    t -= t[0]
    sig = rng.uniform(0.01, 0.05, N)
    y = sig * rng.normal(size=N)
    for tau, amp in zip(taus, amps):
        e = np.exp(-dt / tau)
        drive = amp * np.sqrt(1.0 - e * e) * rng.normal(size=N)
        x = np.empty(N)
        acc = amp * rng.normal()
        el, dl = e.tolist(), drive.tolist()
        out = [0.0] * N
        for n in range(N):
            acc = el[n] * acc + dl[n]
            out[n] = acc
        y += np.asarray(out)
    f_min, f_max = 1.0 / (t[-1] - t[0]), 1.0 / np.min(np.diff(t)) / 2.0
    return t, y, sig ** 2, f_min, f_max
