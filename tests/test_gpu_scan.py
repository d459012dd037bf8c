"""GPU tests of K3 — the parallel-in-time (chunked associative scan) evaluation of the celerite log-likelihood for long
single series — through the C ABI: same value as logl(a,b,c,d,τ,y,σ2) (src/celerite_solver.jl:312-334) computed by the
sequential kernel K2 and by the CPU oracle."""
import time

import numpy as np
import pytest

from conftest import rel_err, synthetic_series
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def pb():
    import pioran_b200
    return pioran_b200


@pytest.fixture(scope="module")
def ctx(pb):
    c = pb.get_context(0)
    yield c
    c.set_scan_chunks(0)


# ranks above 64 (scan_wide.cuh): SHO J = 40 (80), DRWCelerite J = 30 (90), J = 32 (96), J = 24 (72: the 5-row tile shape)
@pytest.mark.parametrize("basis,J", [("SHO", 20), ("DRWCelerite", 20), ("SHO", 30), ("SHO", 8), ("DRWCelerite", 30), ("SHO", 40),
                                     ("DRWCelerite", 32), ("DRWCelerite", 24)])
def test_scan_equals_sequential_and_oracle(pb, ctx, basis, J):
    t, y, s2, f_min, f_max = synthetic_series(3000, seed=11)
    a, b, c, d = orc.approx("SBPL", [0.82, 0.01, 3.3], f_min, f_max, J, 1.0, basis=basis)
    want = orc.celerite_logl(a, b, c, d, t, y, s2)
    ser = ctx.upload_series(t, y, s2)
    ctx.set_auto_scan(False)                     # one evaluation of 3 000 steps would be routed to the scan path
    seq = ctx.celerite_logl(ser, a, b, c, d)[0]
    ctx.set_auto_scan(True)
    assert rel_err(seq, want) <= TOL
    for chunks in (1, 2, 3, 7, 16, 37):
        ctx.set_scan_chunks(chunks)
        got = ctx.celerite_logl_scan(ser, a, b, c, d)[0]
        assert rel_err(got, want) <= TOL, (chunks, got, want)
        assert rel_err(got, seq) <= 1e-11, (chunks, got, seq)
    ser.free()


def test_scan_batch_with_mean_and_variance_scale(pb, ctx):
    t, y, s2, f_min, f_max = synthetic_series(2500, seed=12)
    rng = np.random.default_rng(0)
    B = 3
    th = np.array([[0.5, 0.02, 2.8], [0.82, 0.01, 3.3], [1.1, 0.2, 3.9]])
    co = [orc.approx("SBPL", th[i], f_min, f_max, 20, [0.5, 1.0, 2.0][i]) for i in range(B)]
    a, b, c, d = (np.stack([x[k] for x in co]) for k in range(4))
    mu = rng.normal(0, 0.2, B); nu = rng.uniform(0.5, 2.0, B)
    ser = ctx.upload_series(t, y, s2)
    ctx.set_scan_chunks(12)
    got = ctx.celerite_logl_scan(ser, a, b, c, d, mu=mu, nu=nu)
    want = orc.celerite_logl_batch(a, b, c, d, t, y, s2, mu=mu, nu=nu, nthreads=0)
    assert rel_err(got, want).max() <= TOL
    ser.free()


def test_scan_short_series_falls_back_to_one_chunk(pb, ctx):
    t, y, s2, f_min, f_max = synthetic_series(50, seed=13)
    a, b, c, d = orc.approx("SBPL", [0.82, 0.01, 3.3], f_min, f_max, 20, 1.0)
    ser = ctx.upload_series(t, y, s2)
    ctx.set_scan_chunks(0)
    assert rel_err(ctx.celerite_logl_scan(ser, a, b, c, d)[0], orc.celerite_logl(a, b, c, d, t, y, s2)) <= TOL
    ser.free()


def test_config_c4_long_series_n1e6_j30(pb, ctx):
    """BASELINE configs[3]: N = 1e6, J = 30 (SHO, rank 60), parallel associative scan vs the sequential sweep."""
    import workloads as wl
    N = 1_000_000
    t, y, s2, f_min, f_max = wl.make_series_fast(N, seed=4)
    a, b, c, d = orc.approx("SBPL", [0.82, 0.01, 3.3], f_min, f_max, 30, float(np.var(y)))
    ser = ctx.upload_series(t, y, s2)
    ctx.set_scan_chunks(0)
    got = ctx.celerite_logl_scan(ser, a, b, c, d)[0]          # warm-up (buffers)
    t0 = time.perf_counter()
    got = ctx.celerite_logl_scan(ser, a, b, c, d)[0]
    wall_scan, dev_scan = time.perf_counter() - t0, ctx.last_kernel_ms()
    ctx.set_auto_scan(False)                                   # the plain entry would route this call to the scan path
    t0 = time.perf_counter()
    seq = ctx.celerite_logl(ser, a, b, c, d)[0]
    wall_seq = time.perf_counter() - t0
    ctx.set_auto_scan(True)
    t0 = time.perf_counter()
    want = orc.celerite_logl(a, b, c, d, t, y, s2)
    wall_cpu = time.perf_counter() - t0
    print(f"\nC4 N=1e6 R=60: scan {dev_scan:.2f} ms device ({wall_scan * 1e3:.1f} ms wall), sequential GPU sweep "
          f"{wall_seq * 1e3:.0f} ms, CPU restatement {wall_cpu:.1f} s; logL {got:.6f}; "
          f"rel(scan, seq) {rel_err(got, seq):.2e}, rel(scan, cpu) {rel_err(got, want):.2e}")
    assert np.isfinite(got)
    assert rel_err(got, seq) <= TOL
    assert rel_err(got, want) <= TOL
    ser.free()


def test_config_c4_long_series_n1e6_drwcelerite_j30(pb, ctx):
    """C4 with the DRWCelerite basis at J = 30 — rank 90, the reference's own benchmark grid (benchmark/benchmarks.jl:16-18) —
    through the wide-rank kernels of the scan path, against the sequential wide-rank sweep and the CPU restatement."""
    import workloads as wl
    N = 1_000_000
    t, y, s2, f_min, f_max = wl.make_series_fast(N, seed=4)
    a, b, c, d = orc.approx("SBPL", [0.82, 0.01, 3.3], f_min, f_max, 30, float(np.var(y)), basis="DRWCelerite")
    ser = ctx.upload_series(t, y, s2)
    ctx.set_scan_chunks(0)
    got = ctx.celerite_logl_scan(ser, a, b, c, d)[0]          # warm-up (buffers)
    got = ctx.celerite_logl_scan(ser, a, b, c, d)[0]
    dev_scan = ctx.last_kernel_ms()
    sc = ctx.last_scan_check()
    ctx.set_auto_scan(False)
    t0 = time.perf_counter()
    seq = ctx.celerite_logl(ser, a, b, c, d)[0]
    wall_seq = time.perf_counter() - t0
    ctx.set_auto_scan(True)
    t0 = time.perf_counter()
    want = orc.celerite_logl(a, b, c, d, t, y, s2)
    wall_cpu = time.perf_counter() - t0
    print(f"\nC4 N=1e6 DRWCelerite J=30 (R=90): scan {dev_scan:.2f} ms device, sequential GPU sweep {wall_seq * 1e3:.0f} ms, "
          f"CPU restatement {wall_cpu:.1f} s; rel(scan, seq) {rel_err(got, seq):.2e}, rel(scan, cpu) {rel_err(got, want):.2e}; {sc}")
    assert np.isfinite(got) and sc.fallback == 0
    assert rel_err(got, seq) <= TOL
    assert rel_err(got, want) <= TOL
    ser.free()


def test_config_c4_long_series_n1e6_drwcelerite_prior_draws(pb, ctx):
    """C4 at full size over 32 PRIOR-DRAWN parameter vectors with the DRWCelerite basis (rank 60, slopes up to 6: the
    ill-conditioned part of the prior included), one call each as a sampler on a single long series would make them: every
    value within 1e-9 of the sequential sweep (or at the rounding floor of its covariance, triaged with the 80-bit twin),
    the Newton refinement instead of the sequential fallback, mean device time reported."""
    import workloads as wl
    N, B, J = 1_000_000, 32, 20
    t, y, s2, f_min, f_max = wl.make_series_fast(N, seed=4)
    th = wl.prior_theta(B, f_min, f_max, y.mean(), y.std(), 12, 6.0)
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, J, basis_function="DRWCelerite")
    a, b, c, d = ctx.approx_coeffs(spec, th[:, :4])
    ser = ctx.upload_series(t, y, s2)
    ctx.set_scan_chunks(0)
    ctx.set_auto_scan(False)
    t0 = time.perf_counter()
    seq = ctx.celerite_logl(ser, a, b, c, d, mu=th[:, 5], nu=th[:, 4])       # 32 sequential sweeps side by side
    wall_seq = time.perf_counter() - t0
    ctx.set_auto_scan(True)
    ctx.celerite_logl_scan(ser, a[:1], b[:1], c[:1], d[:1], mu=th[:1, 5], nu=th[:1, 4])     # warm-up (buffers)
    got, ms, passes, nfb, nrf = np.empty(B), [], [], 0, 0
    for i in range(B):
        got[i] = ctx.celerite_logl_scan(ser, a[i:i + 1], b[i:i + 1], c[i:i + 1], d[i:i + 1], mu=th[i:i + 1, 5], nu=th[i:i + 1, 4])[0]
        ms.append(ctx.last_kernel_ms())
        sc = ctx.last_scan_check()
        nfb += sc.fallback; nrf += sc.refined
        passes.append(len(ctx.last_scan_history(0)[0]))
    ser.free()
    ok = np.isfinite(seq)
    assert np.array_equal(np.isfinite(got), ok)
    err = np.abs(got[ok] - seq[ok]) / np.maximum(1.0, np.abs(seq[ok]))
    ms = np.array(ms)
    print(f"\nC4 N=1e6 DRWCelerite J={J} over {B} prior draws ({int(ok.sum())} finite): device ms mean {ms.mean():.2f}, median {np.median(ms):.2f}, "
          f"max {ms.max():.2f} (one sequential sweep: {wall_seq * 1e3:.0f} ms for all {B} side by side); passes per call {np.bincount(passes).tolist()}; "
          f"{nrf} accepted after Newton refinement, {nfb} to the sequential sweep; max |scan - seq| {err.max():.2e}")
    for i in np.flatnonzero(ok)[err > TOL]:
        ld = float(orc.celerite_logl(a[i], b[i], c[i], d[i], t, y - th[i, 5], th[i, 4] * s2, long_double=True))
        floor = abs(seq[i] - ld) / max(1.0, abs(ld))
        assert floor > 1e-10 and abs(got[i] - ld) / max(1.0, abs(ld)) <= 30 * floor + TOL, (i, err.max(), floor)
    assert nfb <= 1, nfb          # the sequential sweep (2.6 s here) is the exception, not the ladder
    assert ms.mean() < 40.0


@pytest.mark.parametrize("world,chunks,N", [(2, 0, 6000), (3, 5, 6001), (8, 0, 5995), (4, 1, 6000)])
def test_scan_time_axis_split_across_ranks(pb, ctx, world, chunks, N):
    """the time axis split over `world` ranks — emulated on one GPU with one context per rank, the
    collectives replaced by plain lists — must reproduce the single-GPU scan and the oracle."""
    from pioran_b200.parallel import scan_logl_sharded
    t, y, s2, f_min, f_max = synthetic_series(N, seed=21)
    a, b, c, d = orc.approx("SBPL", [0.82, 0.01, 3.3], f_min, f_max, 30, 1.0, basis="SHO")
    mu, nu = 0.13, 1.7
    want = orc.celerite_logl_batch(a[None], b[None], c[None], d[None], t, y, s2, mu=np.array([mu]), nu=np.array([nu]))[0]
    ctxs = [pb.Context(0) for _ in range(world)]
    sers = [cx.upload_series(t, y, s2) for cx in ctxs]
    for cx in ctxs:
        cx.set_scan_chunks(chunks)
    from pioran_b200.parallel import scan_bounds
    off = scan_bounds(len(t), world)
    comps = [ctxs[r].scan_range_begin(sers[r], a, b, c, d, off[r], off[r + 1], mu=mu, nu=nu, max_prev=world) for r in range(world)]
    sums = [ctxs[r].scan_range_end(np.stack(comps[:r]) if r else None) for r in range(world)]
    tot = np.sum(sums, axis=0)
    got = -0.5 * tot[0] - 0.5 * tot[1] - 0.5 * len(t) * np.log(2 * np.pi)
    assert rel_err(got, want) <= TOL, (got, want)
    # self-check rows: inner estimates + the hand-overs between the ranks (8 steps swept by both neighbours)
    from pioran_b200.parallel import scan_check_total
    rows = np.stack([ctxs[r].scan_range_check() for r in range(world)])
    assert np.all(rows[1:, 5] == 8) and np.all(rows[:-1, 6] == 8) and rows[0, 5] == 0 and rows[-1, 6] == 0
    assert 0.0 <= scan_check_total(rows) / abs(want) <= 1e-11
    # the orchestration function with list-backed collectives gives the same number on every rank
    for r in range(world):
        val = scan_logl_sharded(lambda lo, hi: ctxs[r].scan_range_begin(sers[r], a, b, c, d, lo, hi, mu=mu, nu=nu, max_prev=world),
                                ctxs[r].scan_range_end, len(t), rank=r, world=world,
                                all_gather=lambda x: np.stack(comps), all_reduce_sum=lambda s_: s_ - sums[r] + tot)
        assert rel_err(val, want) <= TOL
    with pytest.raises(pb.PioranError):
        ctxs[0].scan_range_end(None)            # no range in progress
    for s_, cx in zip(sers, ctxs):
        s_.free()
        cx.close()


def test_auto_dispatch_to_scan(pb, ctx):
    """A single evaluation of a long series through the PLAIN entries (generic and fused) takes the parallel-in-time path
    (include/pioran_b200.h: pioran_ctx_set_auto_scan) and returns the sequential kernel's value."""
    t, y, s2, f_min, f_max = synthetic_series(12000, seed=21)
    theta = np.array([[0.82, 0.01, 3.3, 1.0, 1.3, 0.2]])
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20)
    a, b, c, d = ctx.approx_coeffs(spec, theta[:, :4])
    ser = ctx.upload_series(t, y, s2)
    n0 = ctx.launch_count
    auto_gen = ctx.celerite_logl(ser, a, b, c, d, mu=theta[:, 5], nu=theta[:, 4])[0]
    launches_auto = ctx.launch_count - n0
    auto_fused = ctx.approx_logl(ser, spec, theta)[0, 0]
    ctx.set_auto_scan(False)
    n0 = ctx.launch_count
    seq_gen = ctx.celerite_logl(ser, a, b, c, d, mu=theta[:, 5], nu=theta[:, 4])[0]
    launches_seq = ctx.launch_count - n0
    seq_fused = ctx.approx_logl(ser, spec, theta)[0, 0]
    ctx.set_auto_scan(True)
    ser.free()
    # the scan path is many kernels (fold, Kogge–Stone levels, applies, re-sweep); the sequential sweep is one kernel, or three when
    # the explicit-coefficient call builds its per-θ block table first (table, amplitudes, tensor-pipe sweep)
    assert launches_auto > launches_seq and launches_seq in (1, 3)
    want = orc.celerite_logl(a[0], b[0], c[0], d[0], t, y - theta[0, 5], theta[0, 4] * s2)
    for v in (auto_gen, auto_fused, seq_gen, seq_fused):
        assert rel_err(v, want) <= TOL, (v, want)


def _steep_prior_draws(B, f_min, f_max, seed, alpha2_max):
    """(α₁, f₁, α₂, variance) rows of the reference's example prior (examples/ultranest/single_pl.jl:96-118) with slopes up to
    alpha2_max: the steep ones make the celerite covariance ill-conditioned."""
    rng = np.random.default_rng(seed)
    a1 = rng.uniform(0.0, 1.5, B)
    f1 = np.exp(rng.uniform(np.log(f_min / 5), np.log(f_max * 5), B))
    a2 = a1 + rng.uniform(size=B) * (alpha2_max - a1)
    return np.stack([a1, f1, a2, np.ones(B)], axis=1)


@pytest.mark.parametrize("basis,J", [("DRWCelerite", 5), ("DRWCelerite", 2), ("DRWCelerite", 20), ("DRWCelerite", 30)])
def test_scan_self_check_keeps_sequential_accuracy(pb, ctx, basis, J):
    """The composites of the scan lose accuracy on ill-conditioned covariances (steep slopes: up to 1e-6, far more on the
    J = 2 grid).  Every call checks itself at the chunk boundaries and re-evaluates the offending parameter vectors with
    the sequential sweep (include/pioran_b200.h: pioran_ctx_set_scan_tolerance), so the caller sees the sequential value."""
    t, y, s2, f_min, f_max = synthetic_series(4200, seed=33)
    psd = _steep_prior_draws(64, f_min, f_max, seed=J, alpha2_max=6.0)
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, J, basis_function=basis)
    a, b, c, d = ctx.approx_coeffs(spec, psd)
    ser = ctx.upload_series(t, y, s2)
    ctx.set_auto_scan(False)
    seq = ctx.celerite_logl(ser, a, b, c, d)
    ctx.set_auto_scan(True)
    ctx.set_scan_chunks(0)
    got, nfb, nrf = [], 0, 0
    for i in range(0, 64, 4):
        got.append(ctx.celerite_logl_scan(ser, a[i:i + 4], b[i:i + 4], c[i:i + 4], d[i:i + 4]))   # the scan path with its ladder (4 200 steps: below the auto-routing threshold at some ranks)
        sc = ctx.last_scan_check()
        nfb += sc.fallback; nrf += sc.refined
        # what is returned from the scan path passed its check, or went through the Newton refinement / the sequential sweep
        assert not (sc.estimate > 1e-10) or sc.fallback > 0 or sc.refined > 0
        assert not (sc.estimate > 1e-4) or sc.fallback > 0       # a stalled refinement is accepted up to 1000 × the floor cap on its (pessimistic) estimate
    got = np.concatenate(got)
    ctx.set_scan_tolerance(0.0)
    raw = np.concatenate([ctx.celerite_logl_scan(ser, a[i:i + 4], b[i:i + 4], c[i:i + 4], d[i:i + 4]) for i in range(0, 64, 4)])
    assert ctx.last_scan_check().fallback == 0 and ctx.last_scan_check().refined == 0
    ctx.set_scan_tolerance(1e-10)
    ser.free()
    ok = np.isfinite(seq)
    err = np.abs(got[ok] - seq[ok]) / np.maximum(1.0, np.abs(seq[ok]))
    err_raw = np.abs(raw[ok] - seq[ok]) / np.maximum(1.0, np.abs(seq[ok]))
    print(f"\n{basis} J={J}: checked max {np.nanmax(err):.1e} ({nrf} of 64 accepted after a Newton refinement, {nfb} re-evaluated sequentially), raw scan max {np.nanmax(err_raw):.1e}")
    # beyond 1e-9 only where the FP64 sequential sweep is itself off an 80-bit evaluation (rounding floor of the covariance: a
    # converged Newton iteration is accepted there, and two FP64 routes differ by a multiple of what the 80-bit twin shows for
    # one of them); never beyond the floor cap.  The J = 2 grid is barely positive definite: its floor rows scatter more.
    assert np.all(err <= 1e-6), err.max()
    for i in np.flatnonzero(ok)[err > TOL]:
        ld = float(orc.celerite_logl(a[i], b[i], c[i], d[i], t, y, s2, long_double=True))
        floor = abs(seq[i] - ld) / max(1.0, abs(ld))
        assert floor > (1e-10 if J > 2 else 1e-11) and abs(got[i] - ld) / max(1.0, abs(ld)) <= (30 if J > 2 else 300) * floor + TOL, (i, err.max(), floor)
    if np.nanmax(err_raw) > TOL:
        assert nfb + nrf > 0
    assert np.array_equal(np.isfinite(got), ok)


def test_scan_self_check_leaves_well_conditioned_calls_alone(pb, ctx):
    t, y, s2, f_min, f_max = synthetic_series(12000, seed=21)
    theta = np.array([[0.82, 0.01, 3.3, 1.0, 1.3, 0.2], [0.3, 0.05, 2.5, 0.7, 1.0, -0.1]])
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20)
    ser = ctx.upload_series(t, y, s2)
    ctx.approx_logl(ser, spec, theta)
    sc = ctx.last_scan_check()
    ser.free()
    assert sc.fallback == 0 and sc.refined == 0 and 0.0 <= sc.estimate <= 1e-12, sc


def test_scan_time_axis_split_self_check_and_fallback(pb, ctx):
    """Ill-conditioned covariance with the time axis split over 3 ranks: the gathered self-check rows (inner estimates and
    hand-overs) flag the deviation, and scan_logl_sharded then returns the sequential sweep's value on every rank."""
    from pioran_b200.parallel import scan_logl_sharded, scan_check_total, scan_bounds
    world = 3
    t, y, s2, f_min, f_max = synthetic_series(6000, seed=33)
    psd = _steep_prior_draws(48, f_min, f_max, seed=5, alpha2_max=6.0)
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 5, basis_function="DRWCelerite")
    a, b, c, d = ctx.approx_coeffs(spec, psd)
    ser = ctx.upload_series(t, y, s2)
    ctx.set_auto_scan(False)
    seq = ctx.celerite_logl(ser, a, b, c, d)
    ctx.set_auto_scan(True)
    ser.free()
    ctxs = [pb.Context(0) for _ in range(world)]
    sers = [cx.upload_series(t, y, s2) for cx in ctxs]
    off = scan_bounds(len(t), world)
    flagged = 0
    for i in np.flatnonzero(np.isfinite(seq)):
        ai, bi, ci, di = a[i], b[i], c[i], d[i]
        comps = [ctxs[r].scan_range_begin(sers[r], ai, bi, ci, di, off[r], off[r + 1], max_prev=world) for r in range(world)]
        sums = [ctxs[r].scan_range_end(np.stack(comps[:r]) if r else None) for r in range(world)]
        rows = np.stack([ctxs[r].scan_range_check() for r in range(world)])
        tot = np.sum(sums, axis=0)
        raw = -0.5 * tot[0] - 0.5 * tot[1] - 0.5 * len(t) * np.log(2 * np.pi)
        rel = scan_check_total(rows) / max(1.0, abs(raw))
        dev = rel_err(raw, seq[i])
        assert dev <= 1e-9 or not (rel <= 1e-10), (i, dev, rel)       # no large deviation goes unnoticed
        if not (rel <= 1e-10):
            flagged += 1
            if flagged <= 2:            # the orchestration function, rank by rank, with list-backed collectives
                for r in range(world):
                    info = {}
                    gathers = iter([np.stack(comps), rows])
                    val = scan_logl_sharded(lambda lo, hi: ctxs[r].scan_range_begin(sers[r], ai, bi, ci, di, lo, hi, max_prev=world),
                                            ctxs[r].scan_range_end, len(t), rank=r, world=world,
                                            all_gather=lambda x: next(gathers), all_reduce_sum=lambda s_: s_ - sums[r] + tot,
                                            range_check=ctxs[r].scan_range_check, sequential=lambda: seq[i], info=info)
                    assert info["fallback"] and val == seq[i]
    print(f"\ntime axis over {world} ranks, DRWCelerite J=5 steep draws: {flagged} of {int(np.isfinite(seq).sum())} flagged")
    assert flagged >= 1
    for s_, cx in zip(sers, ctxs):
        s_.free()
        cx.close()
