"""GPU parity of the gradient path (K5) against the oracle's forward-mode restatement, through the C ABI."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import pioran_b200
    return pioran_b200


@pytest.fixture(scope="module")
def ctx(pb):
    return pb.get_context(0)

# |Δ| ≤ RTOL · max(|g|_∞ over the batch column, |g|) per entry: the gradient is a sum over N steps of terms of both
# signs, so entries that nearly cancel are compared on the scale of their column.
RTOL = 1e-8


def _check(got, want, rtol=RTOL):
    scale = np.maximum(np.abs(want), np.abs(want).max(axis=0, keepdims=True))
    err = np.abs(got - want) / scale
    assert np.all(np.isfinite(got)), "non-finite gradient"   # test/test_likelihood.jl:60
    assert err.max() <= rtol, f"gradient parity {err.max():.3e}"
    return err.max()


@pytest.mark.parametrize("basis", ["SHO", "DRWCelerite"])
@pytest.mark.parametrize("J", [20, 12, 16])
def test_grad_vs_oracle(pb, ctx, golden_single, basis, J):
    g = golden_single
    t, y, s2, f_min, f_max = g.t, g.y, g.s2, g.f_min, g.f_max
    rows = np.linspace(0, len(g.theta) - 1, 48).astype(int)
    theta = g.theta[rows].copy()
    if basis == "DRWCelerite":
        theta[:, 2] += 1.0
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", J, basis, f_min=f_min, f_max=f_max, ctx=ctx)
    val, grad = like.value_and_gradient(theta)
    oval, ograd = orc.approx_logl_grad_batch("SBPL", theta, f_min, f_max, J, t, y, s2, basis=basis, nthreads=0)
    verr = np.abs(val - oval) / np.maximum(1.0, np.abs(oval))
    assert verr.max() <= 1e-9, f"value parity {verr.max():.3e}"
    plain = like(theta)
    assert np.abs(val - plain).max() <= 1e-9 * np.maximum(1.0, np.abs(plain)).max()
    _check(grad, ograd)
    like.close()


def test_grad_double_bending(pb, ctx, golden_double):
    g = golden_double
    t, y, s2, f_min, f_max = g.t, g.y, g.s2, g.f_min, g.f_max
    rows = np.linspace(0, len(g.theta) - 1, 24).astype(int)
    theta = g.theta[rows].copy()
    like = pb.BatchedLikelihood(t, y, s2, "DoubleBendingPowerLaw", 20, "SHO", f_min=f_min, f_max=f_max, ctx=ctx)
    val, grad = like.value_and_gradient(theta)
    oval, ograd = orc.approx_logl_grad_batch("DBPL", theta, f_min, f_max, 20, t, y, s2, basis="SHO", nthreads=0)
    assert np.abs(val - oval).max() <= 1e-9 * np.maximum(1.0, np.abs(oval)).max()
    _check(grad, ograd)
    like.close()


def test_grad_matches_central_differences(pb, ctx, golden_single):
    """Independent of the oracle's dual arithmetic: central differences of the GPU log-likelihood itself."""
    g = golden_single
    t, y, s2, f_min, f_max = g.t, g.y, g.s2, g.f_min, g.f_max
    theta = g.theta[-8:].copy()
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", 20, "SHO", f_min=f_min, f_max=f_max, ctx=ctx)
    grad = like.gradient(theta)
    for k in range(theta.shape[1]):
        h = 1e-6 * np.maximum(1.0, np.abs(theta[:, k]))
        tp, tm = theta.copy(), theta.copy()
        tp[:, k] += h
        tm[:, k] -= h
        fd = (like(tp) - like(tm)) / (2 * h)
        # rounding noise of a difference of two log-likelihoods: a few ulp of |log L| ≈ 1e3, divided by 2h
        noise = 2e-15 * np.abs(like(theta)).max() / h.min()
        assert np.allclose(grad[:, k], fd, rtol=2e-5, atol=1e-6 * np.abs(grad[:, k]).max() + noise), k
    like.close()


def test_grad_batch_sizes(pb, ctx, golden_single):
    """Ragged item boundaries: B·P not a multiple of the warps per CTA, and a single parameter vector."""
    g = golden_single
    t, y, s2, f_min, f_max = g.t, g.y, g.s2, g.f_min, g.f_max
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", 20, "SHO", f_min=f_min, f_max=f_max, ctx=ctx)
    ref_val, ref_grad = like.value_and_gradient(g.theta[-301:])
    for B in (1, 3, 301):
        v, gr = like.value_and_gradient(g.theta[-B:])
        assert np.array_equal(v, ref_val[-B:]) and np.array_equal(gr, ref_grad[-B:])
    like.close()


@pytest.mark.parametrize("basis", ["SHO", "DRWCelerite"])
@pytest.mark.parametrize("N", [1, 2, 7, 8, 9, 17, 33])
def test_grad_short_series(pb, ctx, golden_single, basis, N):
    """Edge lengths around the TMA stage sizes (8 and 16 steps) and the two-stage pipeline's start-up (N = 1, 2)."""
    g = golden_single
    t, y, s2 = g.t[:N].copy(), g.y[:N].copy(), g.s2[:N].copy()
    theta = g.theta[[10, 2000, 6000]].copy()
    if basis == "DRWCelerite":
        theta[:, 2] += 1.0
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", 20, basis, f_min=g.f_min, f_max=g.f_max, ctx=ctx)
    val, grad = like.value_and_gradient(theta)
    like.close()
    oval, ograd = orc.approx_logl_grad_batch("SBPL", theta, g.f_min, g.f_max, 20, t, y, s2, basis=basis, nthreads=0)
    assert np.abs(val - oval).max() <= 1e-9 * np.maximum(1.0, np.abs(oval)).max()
    _check(grad, ograd)


@pytest.mark.parametrize("basis,J", [("SHO", 20), ("DRWCelerite", 20), ("SHO", 7), ("DRWCelerite", 13)])
def test_logshift_gradient_vs_oracle(pb, ctx, golden_single, basis, J):
    """Gradient of the log-normal likelihood (docs/src/ultranest.md:197-217 under ForwardDiff, test/test_likelihood.jl:55):
    pioran_approx_logl_logshift_grad against the oracle's dual-number sweep with y and σ² carrying the tangent of c, on the
    reference's simu_single flux (untransformed) and prior-like shifts below min(y)."""
    g = golden_single
    t, y, s2, f_min, f_max = g.t, g.y_raw, g.yerr ** 2, g.f_min, g.f_max
    rows = np.linspace(0, len(g.theta) - 1, 40).astype(int)
    rng = np.random.default_rng(17)
    theta = np.column_stack([g.theta[rows], rng.uniform(-0.5, 0.9, len(rows)) * y.min()])
    if basis == "DRWCelerite":
        theta[:, 2] += 1.0
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", J, basis, f_min=f_min, f_max=f_max, ctx=ctx, log_shift=True)
    val, grad = like.value_and_gradient(theta)
    oval, ograd = orc.approx_logl_logshift_grad_batch("SBPL", theta, f_min, f_max, J, t, y, s2, basis=basis, nthreads=0)
    assert grad.shape == (len(rows), 7)
    verr = np.abs(val - oval) / np.maximum(1.0, np.abs(oval))
    assert verr.max() <= 1e-9, f"value parity {verr.max():.3e}"
    assert np.abs(val - like(theta)).max() <= 1e-9 * np.maximum(1.0, np.abs(val)).max()      # same value as the likelihood entry
    _check(grad, ograd)
    # c = 0 reduces to the plain model on log(y): the six common columns agree with the plain gradient entry
    th0 = theta[:8].copy(); th0[:, 6] = 0.0
    _, g0 = like.value_and_gradient(th0)
    plain = pb.BatchedLikelihood(t, np.log(y), s2 / y ** 2, "SingleBendingPowerLaw", J, basis, f_min=f_min, f_max=f_max, ctx=ctx)
    _, gp = plain.value_and_gradient(th0[:, :6])
    _check(g0[:, :6], gp, 1e-9)
    like.close(); plain.close()


@pytest.mark.parametrize("basis,J", [("SHO", 40), ("DRWCelerite", 30), ("DRWCelerite", 32), ("DRWCelerite", 24)])
def test_grad_wide_ranks_vs_oracle(pb, ctx, golden_single, basis, J):
    """Ranks 65 … 96 — the reference's benchmark grid above the warp kernels' limit (benchmark/benchmarks.jl:16-18: SHO J = 40,
    DRWCelerite J = 30) — through the register-file CTA kernel on (value, tangent) pairs (csrc/wide_grad.cuh)."""
    g = golden_single
    t, y, s2, f_min, f_max = g.t, g.y, g.s2, g.f_min, g.f_max
    rows = np.linspace(0, len(g.theta) - 1, 12).astype(int)
    theta = g.theta[rows].copy()
    if basis == "DRWCelerite":
        theta[:, 2] += 1.0
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", J, basis, f_min=f_min, f_max=f_max, ctx=ctx)
    val, grad = like.value_and_gradient(theta)
    oval, ograd = orc.approx_logl_grad_batch("SBPL", theta, f_min, f_max, J, t, y, s2, basis=basis, nthreads=0)
    verr = np.abs(val - oval) / np.maximum(1.0, np.abs(oval))
    assert verr.max() <= 1e-9, f"value parity {verr.max():.3e}"
    plain = like(theta)
    assert np.abs(val - plain).max() <= 1e-9 * np.maximum(1.0, np.abs(plain)).max()
    _check(grad, ograd)
    like.close()


def test_grad_above_rank_96_is_refused(pb, ctx, golden_single):
    g = golden_single
    like = pb.BatchedLikelihood(g.t, g.y, g.s2, "SingleBendingPowerLaw", 50, "SHO", f_min=g.f_min, f_max=g.f_max, ctx=ctx)
    with pytest.raises(pb.PioranError):
        like.gradient(g.theta[:2].copy())
    like.close()
