"""Widening rows SURVEY §8f #2/#3 — posterior mean (`pred`) and GP draws (`sim`).

CPU part: the oracle restatements of src/celerite_solver.jl:376-483 and :515-549 are pinned the way the reference pins
them itself — pred ≈ predict_direct on the four prediction grids of test/test_prediction.jl:44-58 and on the single
Celerite / Exp kernels of test/test_predict_celerite.jl — plus sim ≡ (dense Cholesky factor)·q.
GPU part (-m gpu): the CUDA path through the C ABI against the oracle on the same inputs, batched."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as orc

A = np.loadtxt(os.path.join(GOLDEN, "simu.txt"))
T, Y, YERR = (np.ascontiguousarray(c) for c in A.T)
S2 = YERR ** 2
# test/test_prediction.jl:11-18
F0 = 1 / (T[-1] - T[0]) / 100
FM = 1 / np.min(np.diff(T)) / 2 * 20
VAR = np.var(Y, ddof=1)


def grids():
    rng = np.random.default_rng(7)
    return {
        "data times": T,                                                              # test_prediction.jl:51
        "fine grid": np.linspace(T.min(), T.max(), 1000),                             # :44, :53
        "beyond the data": np.linspace(T.min() - 30, T.max() + 30, 1000),             # :45, :55
        "random": np.sort(rng.random(1000)) * (T[-1] - T[0]) * 2 + (T[0] - T[-1] / 2),  # :46, :57
    }


def coefs(basis="SHO", theta=(0.82, 0.01, 3.3), J=20):
    return orc.approx("SBPL", list(theta), F0, FM, J, VAR, basis=basis)


@pytest.mark.parametrize("basis", ["SHO", "DRWCelerite"])
def test_oracle_pred_matches_predict_direct(basis):
    a, b, c, d = coefs(basis)
    for name, tau in grids().items():
        p1 = orc.celerite_predict(a, b, c, d, tau, T, Y, S2)
        p2 = orc.direct_predict(a, b, c, d, tau, T, Y, S2)
        err = np.max(np.abs(p1 - p2)) / np.max(np.abs(p2))
        assert err < 1e-10, f"{basis}, {name}: pred vs predict_direct {err:.2e}"   # reference: ≈ (rtol √eps)


def test_oracle_pred_single_kernels():
    """test/test_predict_celerite.jl: Exp(1.0, 2.4) and Celerite(3.2, 0.2, 3.0, 0.2) on data/simu.txt."""
    tp = np.linspace(T.min(), T.max(), 1000)
    for a, b, c, d in (([1.0], [0.0], [2.4], [0.0]), ([3.2], [0.2], [3.0], [0.2])):
        p1 = orc.celerite_predict(a, b, c, d, tp, T, Y, S2)
        p2 = orc.direct_predict(a, b, c, d, tp, T, Y, S2)
        assert np.max(np.abs(p1 - p2)) / np.max(np.abs(p2)) < 1e-10


def test_oracle_pred_repeated_and_edge_points():
    """Several prediction points inside one data gap, on data times, before the first and after the last point."""
    a, b, c, d = coefs()
    tau = np.sort(np.concatenate([[T[0] - 5, T[0] - 1e-3, T[0]], np.linspace(T[3], T[4], 7), T[10:13],
                                  [0.5 * (T[20] + T[21])] * 3, [T[-1], T[-1] + 1e-3, T[-1] + 40]]))
    p1 = orc.celerite_predict(a, b, c, d, tau, T, Y, S2)
    p2 = orc.direct_predict(a, b, c, d, tau, T, Y, S2)
    assert np.max(np.abs(p1 - p2)) / np.max(np.abs(p2)) < 1e-10


@pytest.mark.parametrize("basis", ["SHO", "DRWCelerite"])
def test_oracle_sim_is_cholesky_times_q(basis):
    a, b, c, d = coefs(basis)
    q = np.random.default_rng(3).standard_normal(len(T))
    ys = orc.celerite_simulate(a, b, c, d, T, S2, q)
    tau = np.abs(T[:, None] - T[None, :])
    K = sum(np.exp(-cc * tau) * (aa * np.cos(dd * tau) + bb * np.sin(dd * tau)) for aa, bb, cc, dd in zip(a, b, c, d))
    L = np.linalg.cholesky(K + np.diag(S2))
    assert np.max(np.abs(L @ q - ys)) / np.max(np.abs(ys)) < 1e-10


# ------------------------------------------------------------------------------------------------------------ GPU
def _theta_batch(B, seed):
    rng = np.random.default_rng(seed)
    a1 = rng.uniform(0.0, 1.2, B)
    return np.stack([a1, 10 ** rng.uniform(-2.5, -1.0, B), a1 + rng.uniform(0.5, 2.5, B)], axis=1), rng


@pytest.mark.gpu
@pytest.mark.parametrize("basis", ["SHO", "DRWCelerite"])
def test_gpu_predict_matches_oracle(basis):
    import pioran_b200 as pb
    ctx = pb.get_context(0)
    B = 12
    th, rng = _theta_batch(B, 11)
    mu, nu = rng.normal(Y.mean(), 0.3, B), rng.uniform(0.5, 2.0, B)
    co = [orc.approx("SBPL", list(th[i]), F0, FM, 20, VAR * rng.uniform(0.5, 2.0), basis=basis) for i in range(B)]
    a, b, c, d = (np.stack([x[k] for x in co]) for k in range(4))
    ser = ctx.upload_series(T, Y, S2)
    try:
        for name, tau in grids().items():
            got = ctx.celerite_predict(ser, a, b, c, d, tau, mu=mu, nu=nu)
            for i in range(B):
                want = orc.celerite_predict(a[i], b[i], c[i], d[i], tau, T, Y - mu[i], nu[i] * S2) + mu[i]
                err = np.max(np.abs(got[i] - want)) / max(1.0, np.max(np.abs(want)))
                assert err < 1e-9, f"{basis}, {name}, theta {i}: {err:.2e}"
    finally:
        ser.free()


@pytest.mark.gpu
def test_gpu_predict_edge_points_and_api():
    import pioran_b200 as pb
    a, b, c, d = coefs()
    tau = np.sort(np.concatenate([[T[0] - 5, T[0]], np.linspace(T[3], T[4], 7), T[10:13], [0.5 * (T[20] + T[21])] * 3,
                                  [T[-1], T[-1] + 40]]))
    cov = pb.SumOfCelerite(a, b, c, d)
    got = pb.predict(cov, tau, T, Y, S2)
    want = orc.direct_predict(a, b, c, d, tau, T, Y, S2)
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 1e-9
    # mean(posterior(f(t, σ²), y), τ) with a constant mean (src/scalable_GP.jl:61-67)
    f = pb.ScalableGP(0.7, cov)
    fp = pb.posterior(f(T, S2), Y)
    got = pb.mean(fp, tau)
    want = orc.celerite_predict(a, b, c, d, tau, T, Y - 0.7, S2) + 0.7
    assert np.max(np.abs(got - want)) < 1e-9 * max(1.0, np.max(np.abs(want)))
    assert np.max(np.abs(pb.mean(fp) - (orc.celerite_predict(a, b, c, d, T, T, Y - 0.7, S2) + 0.7))) < 1e-8
    with pytest.raises(pb.PioranError):
        pb.predict(cov, tau[::-1].copy(), T, Y, S2)        # descending τ is rejected, not silently mis-swept


@pytest.mark.gpu
@pytest.mark.parametrize("basis", ["SHO", "DRWCelerite"])
def test_gpu_simulate_matches_oracle(basis):
    import pioran_b200 as pb
    ctx = pb.get_context(0)
    B = 9
    th, rng = _theta_batch(B, 5)
    nu = rng.uniform(0.5, 2.0, B)
    co = [orc.approx("SBPL", list(th[i]), F0, FM, 20, VAR, basis=basis) for i in range(B)]
    a, b, c, d = (np.stack([x[k] for x in co]) for k in range(4))
    q = rng.standard_normal((B, len(T)))
    ser = ctx.upload_series(T, np.zeros_like(T), S2)
    try:
        got = ctx.celerite_simulate(ser, a, b, c, d, q, nu=nu)
    finally:
        ser.free()
    for i in range(B):
        want = orc.celerite_simulate(a[i], b[i], c[i], d[i], T, nu[i] * S2, q[i])
        ok = np.isfinite(want)
        assert np.array_equal(ok, np.isfinite(got[i]))
        err = np.max(np.abs(got[i][ok] - want[ok])) / max(1.0, np.max(np.abs(want[ok])))
        assert err < 1e-9, f"{basis}, theta {i}: {err:.2e}"
    # rand(f(t, σ²)) adds the mean; rand(f(t, σ²), t') draws without noise on other times (src/scalable_GP.jl:133-142)
    cov = pb.SumOfCelerite(a[0], b[0], c[0], d[0])
    fx = pb.ScalableGP(1.5, cov)(T, S2)
    assert np.allclose(pb.rand(fx, q[0]), orc.celerite_simulate(a[0], b[0], c[0], d[0], T, S2, q[0]) + 1.5, rtol=0, atol=1e-9)
    t2 = np.linspace(T[0], T[-1], 300)
    assert np.allclose(pb.rand(fx, q[0][:300], t2), orc.celerite_simulate(a[0], b[0], c[0], d[0], t2, np.zeros(300), q[0][:300]) + 1.5,
                       rtol=0, atol=1e-8)


# ranks 65 … 128 (the register-file CTA kernel of csrc/wide.cuh in store / draw mode): the upper part of the reference's own
# benchmark grid — SHO J = 40 (rank 80), DRWCelerite J = 30 (rank 90), DRWCelerite J = 40 (rank 120, 80 terms)
WIDE_CASES = [("SHO", 40), ("DRWCelerite", 30), ("DRWCelerite", 40)]


@pytest.mark.gpu
@pytest.mark.parametrize("basis,J", WIDE_CASES)
def test_gpu_predict_wide_ranks(basis, J):
    import pioran_b200 as pb
    ctx = pb.get_context(0)
    B = 5
    th, rng = _theta_batch(B, 17)
    mu, nu = rng.normal(Y.mean(), 0.3, B), rng.uniform(0.5, 2.0, B)
    co = [orc.approx("SBPL", list(th[i]), F0, FM, J, VAR * rng.uniform(0.5, 2.0), basis=basis) for i in range(B)]
    a, b, c, d = (np.stack([x[k] for x in co]) for k in range(4))
    ser = ctx.upload_series(T, Y, S2)
    try:
        for name, tau in grids().items():
            got = ctx.celerite_predict(ser, a, b, c, d, tau, mu=mu, nu=nu)
            for i in range(B):
                want = orc.celerite_predict(a[i], b[i], c[i], d[i], tau, T, Y - mu[i], nu[i] * S2) + mu[i]
                err = np.max(np.abs(got[i] - want)) / max(1.0, np.max(np.abs(want)))
                assert err < 1e-9, f"{basis} J={J}, {name}, theta {i}: {err:.2e}"
    finally:
        ser.free()


@pytest.mark.gpu
@pytest.mark.parametrize("basis,J", WIDE_CASES)
def test_gpu_simulate_wide_ranks(basis, J):
    import pioran_b200 as pb
    ctx = pb.get_context(0)
    B = 5
    th, rng = _theta_batch(B, 23)
    nu = rng.uniform(0.5, 2.0, B)
    co = [orc.approx("SBPL", list(th[i]), F0, FM, J, VAR, basis=basis) for i in range(B)]
    a, b, c, d = (np.stack([x[k] for x in co]) for k in range(4))
    q = rng.standard_normal((B, len(T)))
    ser = ctx.upload_series(T, np.zeros_like(T), S2)
    try:
        got = ctx.celerite_simulate(ser, a, b, c, d, q, nu=nu)
    finally:
        ser.free()
    for i in range(B):
        want = orc.celerite_simulate(a[i], b[i], c[i], d[i], T, nu[i] * S2, q[i])
        ok = np.isfinite(want)
        assert np.array_equal(ok, np.isfinite(got[i]))
        err = np.max(np.abs(got[i][ok] - want[ok])) / max(1.0, np.max(np.abs(want[ok])))
        assert err < 1e-9, f"{basis} J={J}, theta {i}: {err:.2e}"
