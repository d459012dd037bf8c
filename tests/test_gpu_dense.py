"""GPU tests of K4 — the batched dense-Cholesky drop-in for log_likelihood_direct (src/direct_solver.jl:6-21) — through
the C ABI: against the oracle's dense restatement, against the reference's own celerite ≡ −dense identities
(test/test_likelihood.jl:57-61, test/test_scalablegp.jl:109-132), and BASELINE config 5 (N = 2 000) against K2."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, prior_theta, rel_err, synthetic_series
from oracle import oracle as orc
from test_oracle import A1, A2_SHO, F1, MUS, VARS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import pioran_b200
    return pioran_b200


@pytest.fixture(scope="module")
def ctx(pb):
    return pb.get_context(0)


def test_dense_n6_literals_match_reference_identity(pb, ctx):
    """test/test_scalablegp.jl:109-132: logpdf(fx, y) ≈ −log_likelihood_direct(…), 10 literal parameter sets, N = 6."""
    t = np.array([0.0, 3.0, 3.2, 3.4, 45.5, 101.2])
    y = np.array([1.3, 2.2, 4.21, 2.5, 3.3, 5.2])
    yerr = np.array([0.1, 0.2, 0.1, 0.1, 0.2, 0.1])
    for i in range(10):
        P = pb.SingleBendingPowerLaw(A1[i], F1[i], A2_SHO[i])
        R = pb.approx(P, 1.0e-4, 1.0e1, 30, VARS[i], ctx=ctx)
        ll = pb.logpdf(pb.ScalableGP(MUS[i], R)(t, yerr ** 2), y, ctx=ctx)
        nll = pb.log_likelihood_direct(R, t, y - MUS[i], yerr ** 2, ctx=ctx)
        want, info = orc.direct_nll(R.a, R.b, R.c, R.d, t, y - MUS[i], yerr ** 2)
        assert info == 0
        assert rel_err(nll, want) <= 1e-10
        assert abs(ll + nll) <= 1.5e-8 * abs(nll)          # the reference's own bar: isapprox, rtol = √eps
        # ScalableGP(…, :direct) routes logpdf through the dense solver
        assert rel_err(pb.logpdf(pb.ScalableGP(MUS[i], R, "direct")(t, yerr ** 2), y, ctx=ctx), -want) <= 1e-10


@pytest.mark.parametrize("basis", ["SHO", "DRWCelerite"])
def test_dense_simu_log_celerite_identity(pb, ctx, basis):
    """test/test_likelihood.jl:7-61 on test/data/simu_log.txt (N = 490): celerite ≡ −dense, both on the GPU."""
    t, y, yerr = np.loadtxt(os.path.join(GOLDEN, "simu_log.txt")).T
    f0 = 1 / (t[-1] - t[0]) / 100
    fM = 1 / np.min(np.diff(t)) / 2 * 20
    variance = np.var(y, ddof=1)
    R = pb.approx(pb.SingleBendingPowerLaw(0.82, 0.01, 3.3), f0, fM, 20, variance, basis_function=basis, ctx=ctx)
    ll = pb.log_likelihood(R, t, y, yerr ** 2, ctx=ctx)
    nll = pb.log_likelihood_direct(R, t, y, yerr ** 2, ctx=ctx)
    want, info = orc.direct_nll(R.a, R.b, R.c, R.d, t, y, yerr ** 2)
    assert info == 0 and rel_err(nll, want) <= 1e-10
    assert abs(ll + nll) <= 1.5e-8 * abs(nll)


@pytest.mark.parametrize("N", [1, 2, 63, 64, 65, 127, 128, 200, 255, 256, 330, 580])
def test_dense_block_edges_vs_oracle(pb, ctx, N):
    """Sizes around the 64-wide panel (the augmented row lands first/last in a block) and around the groups of four panels
    of the trailing update (4, 5, 6, 10 blocks), batched with μ and ν."""
    rng = np.random.default_rng(N)
    t = np.cumsum(0.1 + rng.exponential(1.0, N))
    y = rng.normal(0, 1, N)
    s2 = rng.uniform(0.01, 0.1, N)
    B, Jt = 5, 3
    # positive-definite by construction: a real term, a pure damped cosine, an SHO-type term (a = b, c = d)
    a = rng.uniform(0.5, 2.0, (B, Jt)); b = np.zeros((B, Jt))
    c = rng.uniform(0.05, 1.0, (B, Jt)); d = rng.uniform(0, 2.0, (B, Jt))
    d[:, 0] = 0
    b[:, 2] = a[:, 2]; d[:, 2] = c[:, 2]
    mu = rng.normal(0, 0.3, B); nu = rng.uniform(0.5, 2, B)
    ser = ctx.upload_series(t, y, s2)
    got, info = ctx.direct_logl(ser, a, b, c, d, mu=mu, nu=nu)
    ser.free()
    for i in range(B):
        want, oi = orc.direct_nll(a[i], b[i], c[i], d[i], t, y - mu[i], nu[i] * s2)
        assert oi == 0 and info[i] == 0
        assert rel_err(got[i], want) <= 1e-10, (N, i)


def test_dense_many_terms_beyond_the_factor_tables(pb, ctx):
    """Jt = 120 terms: the separable fill's shared-memory tables stop at 97 terms; beyond that every entry comes from the
    reference's formula (src/Celerite.jl:42-44).  Same oracle, same bar."""
    rng = np.random.default_rng(120)
    N, B, Jt = 150, 3, 120
    t = np.cumsum(0.1 + rng.exponential(1.0, N))
    y = rng.normal(0, 1, N)
    s2 = rng.uniform(0.01, 0.1, N)
    a = rng.uniform(0.01, 0.05, (B, Jt)); b = np.zeros((B, Jt))
    c = rng.uniform(0.05, 1.0, (B, Jt)); d = rng.uniform(0, 2.0, (B, Jt))
    d[:, ::2] = 0                                   # real terms and pure damped cosines: positive-definite by construction
    ser = ctx.upload_series(t, y, s2)
    got, info = ctx.direct_logl(ser, a, b, c, d)
    ser.free()
    for i in range(B):
        want, oi = orc.direct_nll(a[i], b[i], c[i], d[i], t, y, s2)
        assert oi == 0 and info[i] == 0
        assert rel_err(got[i], want) <= 1e-10, i


def test_dense_not_positive_definite_reports_like_posdef_exception(pb, ctx):
    """A negative amplitude that makes K indefinite: the reference throws PosDefException (direct_solver.jl:14);
    the C ABI returns NaN + the order of the failing leading minor, the Python mirror raises."""
    t = np.linspace(0, 10, 40)
    y = np.sin(t)
    s2 = np.full(40, 1e-4)
    ser = ctx.upload_series(t, y, s2)
    a = np.array([[1.0], [-1.0]]); z = np.zeros((2, 1)); c = np.array([[0.3], [0.3]])
    got, info = ctx.direct_logl(ser, a, z, c, z)
    ser.free()
    assert np.isfinite(got[0]) and info[0] == 0
    assert np.isnan(got[1]) and info[1] == 1
    _, oi = orc.direct_nll(a[1], z[1], c[1], z[1], t, y, s2)
    assert oi != 0
    with pytest.raises(np.linalg.LinAlgError):
        pb.log_likelihood_direct(pb.Exp(-2.0, 0.3), t, y, s2, ctx=ctx)


def test_config_c5_dense_n2000_vs_celerite_tolerance_report(pb, ctx):
    """BASELINE configs[4]: batched N = 2 000 full-covariance logpdf (K4) vs the celerite kernel (K2), B = 64 θ from the
    C2 prior, SHO J = 20.  The reference's bar for this identity is rtol √eps (test/test_likelihood.jl:58)."""
    t, y, s2, f_min, f_max = synthetic_series(2000, seed=5)
    th = prior_theta(64, f_min, f_max, y.mean(), y.std(), seed=7)
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20)
    ser = ctx.upload_series(t, y, s2)
    a, b, c, d = ctx.approx_coeffs(spec, th[:, :4])
    ll = ctx.approx_logl(ser, spec, th)[0]
    nll, info = ctx.direct_logl(ser, a, b, c, d, mu=th[:, 5], nu=th[:, 4])
    ok = info == 0
    assert ok.mean() > 0.9
    r = rel_err(-nll[ok], ll[ok])
    print(f"\nC5 tolerance report: {ok.sum()} PD of 64; |Δ|/max(1,|logL|): max {r.max():.3e}, median {np.median(r):.3e}; "
          f"K4 device time {ctx.last_kernel_ms():.1f} ms")
    assert r.max() <= 1.5e-8
    # spot-check two rows against the CPU dense restatement
    for i in np.flatnonzero(ok)[:2]:
        want, _ = orc.direct_nll(a[i], b[i], c[i], d[i], t, y - th[i, 5], th[i, 4] * s2)
        assert rel_err(nll[i], want) <= 1e-9
