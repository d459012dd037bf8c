"""Pins the CPU oracle (oracle/pioran_oracle.c) against every golden vector, known-answer test and fixture the
reference holds for the likelihood path.  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, periodic_mean, rel_err
from oracle import oracle as orc

# test/test_psd.jl:38 — 20 SHO amplitudes for SingleBendingPowerLaw(0.3, 0.02, 2.93), f0=0.02, fM=152
GOLD_AMPLITUDES = np.array([
    1.3749158408973243, 0.26031747510091013, 0.06961116778917277, 0.013679642568525807, 0.0037949128465199307,
    0.0008858780578830132, 0.00023278915565955668, 5.714159750636342e-5, 1.463191298808472e-5, 3.6532013241322788e-6,
    9.262211884550235e-7, 2.3267166983266322e-7, 5.877072005450016e-8, 1.4801031386988674e-8, 3.728877337268077e-9,
    9.44575715327315e-10, 2.3313738171903584e-10, 6.377629826311069e-11, 1.119218106083312e-11, 6.962520986945091e-12])

# test/test_psd.jl:101-105 and 120-124 parameter sets
A1 = [0.2, 0.03, 0.1, 0.46, 0.1, 0.21, 0.74, 0.1, 0.03, 0.92]
F1 = [1.3e-2, 1.32e-1, 5.53e-2, 3.3, 0.342, 3.2e1, 1.3, 4.0e1, 1.0e-2, 0.5]
A2_SHO = [3.2, 3.1, 2.3, 2.57, 3.6, 2.3, 2.1, 2.79, 3.3, 3.8]
A2_DRW = [4.2, 3.1, 4.3, 5.57, 4.6, 2.3, 5.1, 2.79, 4.3, 5.8]
VARS = [1.32, 35.3, 242.2, 46.6, 0.3, 0.244, 9.64, 0.75, 0.193, 0.21]
MUS = [1.2, 0.3, 0.1, 0.46, 0.1, 0.21, 0.74, 0.1, 0.03, 0.92]


def test_psd_closed_forms():
    """test/test_psd.jl:3-13"""
    f = 10.0 ** np.linspace(-3, 2, 1000)
    got = orc.psd_eval("SBPL", [0.3, 0.02, 2.93], f)
    want = (f / 0.02) ** (-0.3) / (1 + (f / 0.02) ** (2.93 - 0.3))
    assert np.allclose(got, want, rtol=2e-16 * 8, atol=0)
    f = 10.0 ** np.linspace(-3, 3, 1000)
    got = orc.psd_eval("DBPL", [0.3, 0.02, 1.4, 10.2, 2.93], f)
    want = (f / 0.02) ** (-0.3) / (1 + (f / 0.02) ** (1.4 - 0.3)) / (1 + (f / 10.2) ** (2.93 - 1.4))
    assert np.allclose(got, want, rtol=2e-16 * 8, atol=0)


def test_spectral_grid():
    """test/test_psd.jl:24-29"""
    f0, fM, J = 0.02, 1.52e2, 20
    fj, B = orc.build_approx(J, f0, fM)
    assert len(fj) == 20
    assert np.allclose(f0 * ((fM / f0) ** (1 / (J - 1))) ** np.arange(J), fj, rtol=1e-13)
    assert np.allclose(np.diag(B), 0.5)


def test_golden_amplitudes():
    """test/test_psd.jl:32-39"""
    amp, _ = orc.get_approx_coefficients("SBPL", [0.3, 0.02, 2.93], 0.02, 1.52e2, 20)
    assert np.all(np.isfinite(amp))
    # the reference's `≈` is rtol √eps; the restatement is far inside it
    assert np.max(np.abs(amp / GOLD_AMPLITUDES - 1)) < 1e-12


@pytest.mark.parametrize("basis,a2set,J", [("SHO", A2_SHO, 25), ("DRWCelerite", A2_DRW, 25)])
def test_approx_variance(basis, a2set, J):
    """test/test_psd.jl:100-153: approx(...; is_integrated_power=false)(0,0) ≈ va  ⇒ Σa = va."""
    for i in range(10):
        a, b, c, d = orc.approx("SBPL", [A1[i], F1[i], a2set[i]], 2.0e-3, 3.52e2, J, VARS[i],
                                is_integrated_power=False, basis=basis)
        assert len(a) == (J if basis == "SHO" else 2 * J)
        assert abs(a.sum() / VARS[i] - 1) < 1e-12


@pytest.mark.parametrize("basis,a2set,J", [("SHO", A2_SHO, 25), ("DRWCelerite", A2_DRW, 30)])
def test_approx_integral(basis, a2set, J):
    """test/test_psd.jl:155-203: ∫ of the recovered basis functions over [f_min,f_max] ≈ va, rtol 1e-8."""
    f_min, f_max = 1.0e-3, 3.52e2
    for i in range(10):
        a, b, c, d = orc.approx("SBPL", [A1[i], F1[i], a2set[i]], f_min, f_max, J, VARS[i], basis=basis)
        if basis == "SHO":
            sp = c / (np.sqrt(2) * np.pi)
            amp = a / (sp * np.pi / np.sqrt(2))
        else:
            sp = c[:J] / np.pi
            amp = a[:J] / (sp * np.pi / 3)
        assert abs(orc.integrate_basis(amp, sp, f_min, f_max, basis) / VARS[i] - 1) < 1e-8


def test_celerite_equals_dense_n6():
    """test/test_scalablegp.jl:109-132 — 10 literal cases, logpdf ≈ −log_likelihood_direct (rtol √eps)."""
    t = np.array([0.0, 3.0, 3.2, 3.4, 45.5, 101.2])
    y = np.array([1.3, 2.2, 4.21, 2.5, 3.3, 5.2])
    yerr = np.array([0.1, 0.2, 0.1, 0.1, 0.2, 0.1])
    for i in range(10):
        a, b, c, d = orc.approx("SBPL", [A1[i], F1[i], A2_SHO[i]], 1.0e-4, 1.0e1, 30, VARS[i])
        ll = orc.celerite_logl(a, b, c, d, t, y - MUS[i], yerr ** 2)
        nll, info = orc.direct_nll(a, b, c, d, t, y - MUS[i], yerr ** 2)
        assert info == 0 and np.isfinite(ll)
        assert abs(ll + nll) <= 1.5e-8 * abs(nll)


def test_celerite_equals_dense_simu_log():
    """test/test_likelihood.jl:7-61 on test/data/simu_log.txt (N=490), SHO and DRWCelerite, J=20."""
    t, y, yerr = np.loadtxt(os.path.join(GOLDEN, "simu_log.txt")).T
    f0 = 1 / (t[-1] - t[0]) / 100
    fM = 1 / np.min(np.diff(t)) / 2 * 20
    variance = np.var(y, ddof=1)
    for basis in ("SHO", "DRWCelerite"):
        a, b, c, d = orc.approx("SBPL", [0.82, 0.01, 3.3], f0, fM, 20, variance, basis=basis)
        ll = orc.celerite_logl(a, b, c, d, t, y, yerr ** 2)
        nll, info = orc.direct_nll(a, b, c, d, t, y, yerr ** 2)
        assert info == 0
        assert abs(ll + nll) <= 1.5e-8 * abs(nll)


def _check_chain(g, got, bound=1e-9):
    r = rel_err(got, g.logl)
    assert np.all(np.isfinite(got))
    assert r.max() <= bound, f"max rel {r.max():.3e} at row {r.argmax()}"
    assert np.median(r) < 1e-13


def test_chain_simu_single(golden_single):
    """≈6.5 k (θ, logL) pairs produced by the reference itself (examples/ultranest/single_pl.jl)."""
    g = golden_single
    got = orc.approx_logl_batch("SBPL", g.theta, g.f_min, g.f_max, 20, g.t, g.y, g.s2, nthreads=0)
    _check_chain(g, got)


def test_chain_simu_double(golden_double):
    """examples/ultranest/double_pl.jl (DoubleBendingPowerLaw)."""
    g = golden_double
    got = orc.approx_logl_batch("DBPL", g.theta, g.f_min, g.f_max, 20, g.t, g.y, g.s2, nthreads=0)
    _check_chain(g, got)


def test_chain_simu_periodic(golden_periodic):
    """examples/ultranest/single_pl_periodicity.jl (sinusoidal CustomMean); every 8th row keeps it fast."""
    g = golden_periodic
    rows = np.arange(0, len(g.chain), 8)
    got = np.empty(len(rows))
    for k, i in enumerate(rows):
        th = g.theta[i]
        a, b, c, d = orc.approx("SBPL", th[:3], g.f_min, g.f_max, 20, th[3])
        got[k] = orc.celerite_logl(a, b, c, d, g.t, g.y - periodic_mean(g.t, g.chain[i]), th[4] * g.s2)
    r = rel_err(got, g.logl[rows])
    assert r.max() <= 1e-9 and np.median(r) < 1e-13


def test_maximum_likelihood_point():
    """examples/ultranest/inference/simu_single/info/results.json maximum_likelihood."""
    from conftest import GoldenRun
    g = GoldenRun("simu_single", "SingleBendingPowerLaw", 3, True)
    with open(os.path.join(GOLDEN, "simu_single_maximum_likelihood.json")) as fh:
        ml = json.load(fh)
    got = orc.approx_logl_batch("SBPL", np.array([ml["point"]]), g.f_min, g.f_max, 20, g.t, g.y, g.s2)[0]
    assert abs(got - ml["logl"]) <= 1e-9 * abs(ml["logl"])


def test_long_double_twin_agrees():
    """The 80-bit twin used for conditioning triage evaluates the same likelihood."""
    from conftest import GoldenRun
    g = GoldenRun("simu_single", "SingleBendingPowerLaw", 3, True)
    th = g.theta[3000]
    a, b, c, d = orc.approx("SBPL", th[:3], g.f_min, g.f_max, 20, th[3])
    v = orc.celerite_logl(a, b, c, d, g.t, g.y - th[5], th[4] * g.s2)
    vl = float(orc.celerite_logl(a, b, c, d, g.t, g.y - th[5], th[4] * g.s2, long_double=True))
    assert abs(v - vl) <= 1e-9 * abs(v)


def test_negative_pivot_semantics():
    """log|D_n| keeps negative pivots finite (src/celerite_solver.jl:140), but the first pivot has no abs (:126)."""
    t = np.array([0.0, 1.0, 2.5, 3.0])
    y = np.array([0.1, -0.2, 0.3, 0.0])
    s2 = np.full(4, 1e-2)
    # a < 0 for one term: D_1 = Σa + σ² < 0 → log(negative) = NaN in the reference
    v = orc.celerite_logl([-1.0], [0.0], [0.5], [0.0], t, y, s2)
    assert np.isnan(v)
    # Σa > 0 but indefinite: later pivots may be negative, result stays finite
    v = orc.celerite_logl([2.0, -1.5], [0.0, 0.0], [0.1, 5.0], [0.0, 0.0], t, y, s2)
    assert np.isfinite(v)


# ---------------------------------------------------------------------------------------------- gradient oracle
@pytest.mark.parametrize("basis", ["SHO", "DRWCelerite"])
def test_gradient_oracle_value_and_finite_differences(golden_single, basis):
    """pioran_oracle_grad.c (forward mode, what ForwardDiff does in test/test_likelihood.jl:55): the value part is the
    pinned oracle bit for bit, the derivative part agrees with central differences of it, and is finite (:60)."""
    g = golden_single
    theta = g.theta[np.linspace(0, len(g.theta) - 1, 6).astype(int)].copy()
    if basis == "DRWCelerite":
        theta[:, 2] += 1.0
    val, grad = orc.approx_logl_grad_batch("SBPL", theta, g.f_min, g.f_max, 20, g.t, g.y, g.s2, basis=basis, nthreads=0)
    plain = orc.approx_logl_batch("SBPL", theta, g.f_min, g.f_max, 20, g.t, g.y, g.s2, basis=basis, nthreads=0)
    assert np.array_equal(val, plain)
    assert np.all(np.isfinite(grad))
    for k in range(theta.shape[1]):
        h = 1e-6 * np.maximum(1.0, np.abs(theta[:, k]))
        tp, tm = theta.copy(), theta.copy()
        tp[:, k] += h
        tm[:, k] -= h
        fd = (orc.approx_logl_batch("SBPL", tp, g.f_min, g.f_max, 20, g.t, g.y, g.s2, basis=basis, nthreads=0)
              - orc.approx_logl_batch("SBPL", tm, g.f_min, g.f_max, 20, g.t, g.y, g.s2, basis=basis, nthreads=0)) / (2 * h)
        assert np.allclose(grad[:, k], fd, rtol=2e-5, atol=1e-6 * np.abs(grad[:, k]).max()), (k, grad[:, k], fd)
    # the 80-bit twin (pioran_oracle_grad_ld.c: the same source compiled with long double — conditioning triage for gradient
    # comparisons) agrees with the FP64 oracle far below the parity tolerance on these well-conditioned rows
    val_ld, grad_ld = orc.approx_logl_grad_batch("SBPL", theta, g.f_min, g.f_max, 20, g.t, g.y, g.s2, basis=basis, nthreads=0,
                                                  long_double=True)
    assert np.max(np.abs(val_ld - val) / np.abs(val)) <= 1e-10
    assert np.max(np.abs(grad_ld - grad) / np.maximum(np.abs(grad), np.abs(grad).max(axis=0))) <= 1e-8


def test_logshift_gradient_oracle_against_central_differences():
    """The log-shift branch of the gradient oracle (pioran_oracle_grad.c: y and σ² carry the tangent of c) has no reference
    literal; its value part is the pinned log-likelihood oracle on the transformed data and its derivative part matches
    central differences of that same oracle."""
    rng = np.random.default_rng(3)
    N = 120
    t = np.cumsum(0.1 + rng.exponential(1, N)); y = np.exp(rng.normal(0, 0.3, N)) + 0.5; s2 = (0.02 * y) ** 2
    fm, fx = 1 / (t[-1] - t[0]), 1 / np.min(np.diff(t)) / 2
    th = np.array([[0.6, 0.05, 3.0, 0.1, 1.3, 0.1, 0.2]])
    for basis in ("SHO", "DRWCelerite"):
        l, g = orc.approx_logl_logshift_grad_batch("SBPL", th, fm, fx, 12, t, y, s2, basis=basis)

        def f(x):
            c = x[6]
            return orc.approx_logl_batch("SBPL", x[None, :6], fm, fx, 12, t, np.log(y - c), s2 / (y - c) ** 2, basis=basis)[0]
        assert l[0] == f(th[0])
        for k in range(7):
            h = 1e-6 * max(1.0, abs(th[0, k]))
            xp, xm = th[0].copy(), th[0].copy()
            xp[k] += h; xm[k] -= h
            fd = (f(xp) - f(xm)) / (2 * h)
            assert abs(g[0, k] - fd) <= 2e-6 * max(1.0, abs(fd)), (basis, k, g[0, k], fd)
