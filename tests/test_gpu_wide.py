"""GPU parity of the wide-rank kernel (K2w, ranks 65 … 160: the upper part of the reference's own benchmark grid,
benchmark/benchmarks.jl:16-18) against the oracle, through the C ABI."""
import numpy as np
import pytest

from conftest import assert_parity, prior_theta, rel_err, synthetic_series
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def pb():
    import pioran_b200
    return pioran_b200


@pytest.fixture(scope="module")
def ctx(pb):
    return pb.get_context(0)


@pytest.mark.parametrize("basis,J", [("SHO", 40), ("SHO", 50), ("DRWCelerite", 30), ("DRWCelerite", 40), ("DRWCelerite", 50),
                                     ("SHO", 33), ("DRWCelerite", 22), ("SHO", 36), ("SHO", 64), ("DRWCelerite", 35), ("SHO", 60),
                                     ("DRWCelerite", 42), ("SHO", 44), ("DRWCelerite", 27), ("SHO", 56)])
def test_wide_fused_vs_oracle(pb, ctx, golden_single, basis, J):
    """approx + logpdf at ranks 66 … 150 on the reference's own series (up to 128: the blocked one-CTA-per-evaluation kernel
    of csrc/blocked_wide.cuh — every column-tile count 9 … 16, with and without a row tile of its own for the data row)."""
    g = golden_single
    rows = np.linspace(0, len(g.theta) - 1, 24).astype(int)
    theta = g.theta[rows].copy()
    if basis == "DRWCelerite":
        theta[:, 2] += 1.0
    like = pb.BatchedLikelihood(g.t, g.y, g.s2, "SingleBendingPowerLaw", J, basis, f_min=g.f_min, f_max=g.f_max, ctx=ctx)
    got = like(theta)
    like.close()
    want = orc.approx_logl_batch("SBPL", theta, g.f_min, g.f_max, J, g.t, g.y, g.s2, basis=basis, nthreads=0)

    def ld(row):
        a, b, c, d = orc.approx("SBPL", row[:3], g.f_min, g.f_max, J, row[3], basis=basis)
        return orc.celerite_logl(a, b, c, d, g.t, g.y - row[5], row[4] * g.s2, long_double=True)
    assert_parity(got, want, theta, ld, tol=TOL)


@pytest.mark.parametrize("Jt", [33, 48, 64, 80])
def test_wide_generic_benchmark_grid(ctx, Jt):
    """celerite_likelihood group of benchmark/benchmarks.jl:76-91: a = 5·U(0,1), b, c, d ~ U(0,1), explicit coefficients."""
    rng = np.random.default_rng(1234 + Jt)
    B = 6
    a = 5 * rng.uniform(size=(B, Jt))
    b, c, d = (rng.uniform(size=(B, Jt)) for _ in range(3))
    b = np.minimum(b, a * c / np.maximum(d, 1e-12) * 0.9)     # keep every term a valid covariance (|b d| < a c)
    t, y, s2, _, _ = synthetic_series(700, 77)
    ser = ctx.upload_series(t, y, s2)
    got = ctx.celerite_logl(ser, a, b, c, d)
    ser.free()
    want = orc.celerite_logl_batch(a, b, c, d, t, y, s2, nthreads=0)
    assert rel_err(got, want).max() <= TOL, rel_err(got, want).max()


def test_wide_mixed_real_terms_and_scalars(ctx):
    """Real terms (b = d = 0) take one row; μ, ν and per-θ data vectors go through the same arguments as the narrow kernel."""
    rng = np.random.default_rng(5)
    B, Jt = 4, 50
    a = rng.uniform(0.1, 2.0, size=(B, Jt))
    c = rng.uniform(0.01, 1.0, size=(B, Jt))
    b = rng.uniform(0.0, 0.5, size=(B, Jt)) * a
    d = rng.uniform(0.05, 1.0, size=(B, Jt))
    b[:, 30:] = 0.0
    d[:, 30:] = 0.0                                       # 20 real terms: rank 2·30 + 20 = 80
    c[:, :30] = np.maximum(c[:, :30], d[:, :30] * 0.6)    # |b d| < a c
    t, y, s2, _, _ = synthetic_series(500, 78)
    mu, nu = rng.normal(size=B), rng.uniform(0.5, 2.0, size=B)
    yb = y[None, :] + 0.01 * rng.normal(size=(B, len(y)))
    ser = ctx.upload_series(t, y, s2)
    got = ctx.celerite_logl(ser, a, b, c, d, mu=mu, nu=nu, y_batch=yb)
    ser.free()
    want = np.array([orc.celerite_logl(a[i], b[i], c[i], d[i], t, yb[i] - mu[i], nu[i] * s2) for i in range(B)])
    assert rel_err(got, want).max() <= TOL, rel_err(got, want).max()


def test_rank_limit_is_reported(pb, ctx):
    t, y, s2, f_min, f_max = synthetic_series(64, 3)
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", 60, "DRWCelerite", f_min=f_min, f_max=f_max, ctx=ctx)
    with pytest.raises(Exception, match="exceeds this build's limit"):
        like(prior_theta(2, f_min, f_max, y.mean(), y.std(), 1))
    like.close()
