import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def rel_err(x, ref):
    """|Δ| / max(1, |ref|): relative error that stays meaningful near logL ≈ 0."""
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.abs(x - ref) / np.maximum(1.0, np.abs(ref))


class GoldenRun:
    """One shipped ultranest run of the reference: the series it used and the (θ, logL) pairs it produced."""

    def __init__(self, stem, model, n_psd_par, log_transform):
        ts = np.loadtxt(os.path.join(GOLDEN, f"{stem}_subset_time_series.txt"))
        self.t, self.y_raw, self.yerr = (np.ascontiguousarray(c) for c in ts.T)
        chains = np.load(os.path.join(GOLDEN, "chains.npz"))
        self.chain = chains[stem]
        self.columns = [str(c) for c in chains[stem + "_columns"]]
        self.model, self.n_psd_par = model, n_psd_par
        # examples/ultranest/single_pl.jl:48: f_min, f_max = 1/(t[end]-t[1]), 1/minimum(diff(t))/2
        self.f_min = 1.0 / (self.t[-1] - self.t[0])
        self.f_max = 1.0 / np.min(np.diff(self.t)) / 2.0
        if log_transform:  # single_pl.jl:70-73: σ² = ν σ²/y², yn = log y
            self.y = np.log(self.y_raw)
            self.s2 = self.yerr ** 2 / self.y_raw ** 2
        else:              # single_pl_periodicity.jl: σ² = ν σ²
            self.y = self.y_raw.copy()
            self.s2 = self.yerr ** 2
        self.logl = self.chain[:, 1]
        self.theta = self.chain[:, 2:2 + n_psd_par + 3]  # psd parameters…, variance, ν, μ


@pytest.fixture(scope="session")
def golden_single():
    return GoldenRun("simu_single", "SingleBendingPowerLaw", 3, True)


@pytest.fixture(scope="session")
def golden_double():
    return GoldenRun("simu_double", "DoubleBendingPowerLaw", 5, True)


@pytest.fixture(scope="session")
def golden_periodic():
    g = GoldenRun("simu_periodic_rednoise", "SingleBendingPowerLaw", 3, False)
    return g


def periodic_mean(t, row):
    """examples/ultranest/single_pl_periodicity.jl: A·sin(2πt/T₀ + ϕ) + μ ; columns … μ A ϕ T₀."""
    mu, A, phi, T0 = row[7], row[8], row[9], row[10]
    return A * np.sin(2 * np.pi * t / T0 + phi) + mu


def synthetic_series(N, seed, theta0=(0.82, 0.01, 3.3), variance=1.0, basis="SHO", J=20):
    """SURVEY §8d generator: gaps 0.05+Exp(1), σ~U(0.01,0.05); y = GP draw at θ₀ (state-space sampling of the
    celerite model, equivalent in distribution to Pioran.sim, src/celerite_solver.jl:515-549) + N(0,σ²)."""
    from oracle import oracle as orc
    rng = np.random.default_rng(seed)
    t = np.cumsum(0.05 + rng.exponential(1.0, N))
    t -= t[0]
    sig = rng.uniform(0.01, 0.05, N)
    f_min, f_max = 1.0 / (t[-1] - t[0]), 1.0 / np.min(np.diff(t)) / 2.0
    a, b, c, d = orc.approx("SBPL", theta0, f_min, f_max, J, variance, basis=basis)
    y = np.zeros(N)
    for aj, bj, cj, dj in zip(a, b, c, d):
        # complex term: x = (x1,x2) with stationary covariance [[a,-b],[-b,a]]... sample through the exact
        # discretised OU-rotation; for simplicity draw each term as a stationary complex AR(1) with the right ACVF
        if aj <= 0:
            continue
        rr = min(abs(bj) / aj, 1.0) if aj > 0 else 0.0
        s = np.sign(bj) if bj != 0 else 1.0
        # P∞ = [[a, -b],[-b, a]] has eigenvalues a∓b ≥ 0 when |b| ≤ a
        P = np.array([[aj, -s * rr * aj], [-s * rr * aj, aj]])
        w, V = np.linalg.eigh(P)
        Lc = V @ np.diag(np.sqrt(np.clip(w, 0, None)))
        x = Lc @ rng.normal(size=2)
        y[0] += x[0]
        for n in range(1, N):
            dt = t[n] - t[n - 1]
            e = np.exp(-cj * dt)
            co, si = np.cos(dj * dt), np.sin(dj * dt)
            F = e * np.array([[co, -si], [si, co]])
            Q = P - F @ P @ F.T
            Q = 0.5 * (Q + Q.T)
            wq, Vq = np.linalg.eigh(Q)
            x = F @ x + Vq @ (np.sqrt(np.clip(wq, 0, None)) * rng.normal(size=2))
            y[n] += x[0]
    y += sig * rng.normal(size=N)
    return t, y, sig ** 2, f_min, f_max


def prior_theta(B, f_min, f_max, ybar, ysd, seed, alpha2_max=4.0):
    """Prior transform of examples/ultranest/single_pl.jl:96-104 on a seeded unit cube."""
    rng = np.random.default_rng(seed)
    u = rng.uniform(size=(B, 6))
    f0, fM = f_min / 20.0, f_max * 20.0
    a1 = 1.5 * u[:, 0]
    f1 = np.exp(np.log(4 * f0) + u[:, 1] * (np.log(fM / 4) - np.log(4 * f0)))
    a2 = a1 + u[:, 2] * (alpha2_max - a1)
    from scipy import stats
    var = stats.lognorm(s=np.sqrt(2.0), scale=np.exp(-3.0)).ppf(u[:, 3])
    nu = stats.gamma(a=2, scale=0.5).ppf(u[:, 4])
    mu = stats.norm(loc=ybar, scale=5 * ysd).ppf(u[:, 5])
    return np.column_stack([a1, f1, a2, var, nu, mu])
