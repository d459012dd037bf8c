"""CPU-side tests: the C-ABI library loads and exports every symbol the header declares, host-side argument
checking mirrors the reference's error behaviour, and the N>1 sharding plumbing works (gloo, world_size 2).
No GPU compute is called here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def _build():
    import pioran_b200 as pb
    pb.build.build()
    return pb


def test_header_symbols_exported():
    pb = _build()
    hdr = open(os.path.join(ROOT, "include", "pioran_b200.h")).read()
    declared = set(re.findall(r"\b(pioran_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"pioran_ctx", "pioran_approx_spec"}
    assert len(declared) >= 15
    lib = ctypes.CDLL(pb._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/pioran_b200.h but not exported"
    # and the Python binding covers the same set
    assert declared == set(pb._lib.SYMBOLS), declared ^ set(pb._lib.SYMBOLS)


def test_library_has_sm100a_tma_code():
    """The cubin inside the .so is sm_100a and its celerite kernel stages the table with TMA (UBLKCP in SASS)."""
    pb = _build()
    out = subprocess.run(["cuobjdump", "-lelf", pb._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN6pioran22celerite_shared_kernelILi5ELi12EEEvNS_9BatchArgsE",
                           pb._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "DFMA" in sass and "SYNCS" in sass


def test_no_gpu_fails_loudly():
    """No CPU fallback: without a usable device ctx creation returns PIORAN_ECUDA with a message."""
    pb = _build()
    lib = pb._lib.load()
    assert lib.pioran_version() >= 100
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(pb.PioranError) as ei:
        pb.Context(0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "pioran.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"
    # measurement helpers under tools/ stay clear of it too (the checker scripts that need it live under tests/tools/)
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith((".py", ".sh")):
            txt = open(os.path.join(ROOT, "tools", f), errors="replace").read()
            assert not re.search(r"(from|import)\s+oracle", txt), f"tools/{f} imports the oracle"


def test_spec_validation_mirrors_reference():
    pb = _build()
    with pytest.raises(ValueError, match="not implemented"):   # src/psd.jl:285
        pb.make_spec("SingleBendingPowerLaw", 1e-3, 1.0, 20, basis_function="Matern")
    sp = pb.make_spec("SingleBendingPowerLaw", 1e-3, 1.0)
    assert (sp.n_components, sp.S_low, sp.S_high, sp.is_integrated_power, sp.basis) == (20, 20.0, 20.0, 1, 0)
    with pytest.raises(TypeError):
        pb.approx("not a psd", 1e-3, 1.0)
    with pytest.raises(ValueError, match="not recognised"):    # src/celerite_solver.jl:268
        pb.log_likelihood(pb.Celerite(1, 0, 1, 0), [0.0, 1.0], [0.0, 0.0], [1.0, 1.0], solver="bogus")


def test_acvf_types():
    """(a,b,c,d) conventions of the kernel types (test/test_covariancefunctions.jl, test/test_acvf.jl)."""
    pb = _build()
    e = pb.Exp(2.0, 0.5)
    assert pb.celerite_coefs(e) == (np.array([1.0]), np.array([0.0]), np.array([0.5]), np.array([0.0]))  # a = A/2 (src/Exp.jl:30)
    s = pb.SHO(1.5, 2.0)
    a, b, c, d = pb.celerite_coefs(s)
    assert a[0] == b[0] == 1.5 and c[0] == d[0] == 2.0 / np.sqrt(2)
    tot = pb.SumOfCelerite([1.0], [0.0], [0.3], [0.0]) + pb.Celerite(2.0, 1.0, 0.1, 3.0)
    assert len(tot.a) == 2 and abs(tot(0.0, 0.0) - 3.0) < 1e-15
    τ = 0.7
    want = np.exp(-0.3 * τ) * 1.0 + np.exp(-0.1 * τ) * (2.0 * np.cos(3 * τ) + np.sin(3 * τ))
    assert abs(tot(1.0, 1.7) - want) < 1e-15


def test_shard_bounds():
    from pioran_b200.parallel import shard_bounds, shard_series
    off = shard_bounds(4096, 8)
    assert off[0] == 0 and off[-1] == 4096 and np.all(np.diff(off) == 512)
    off = shard_bounds(10, 4)
    assert list(np.diff(off)) == [3, 3, 2, 2]
    lens = np.array([3000, 1000, 2000, 2500, 1500, 1200, 2800, 1900])
    parts = shard_series(lens, 2)
    assert sorted(np.concatenate(parts)) == list(range(8))
    loads = [lens[p].sum() for p in parts]
    assert abs(loads[0] - loads[1]) <= 600


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import pioran_b200 as pb
from pioran_b200.parallel import ShardedEvaluator
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
calls = []
def fake_eval(theta):            # stands in for the per-GPU kernel call: any pure function of the rows
    calls.append(theta.shape[0])
    return (theta ** 2).sum(dim=1) - theta[:, 0]
B = 11                            # odd: shards of 6 and 5
g = torch.Generator().manual_seed(0)
theta = torch.randn(B, 6, generator=g, dtype=torch.float64)
out = ShardedEvaluator(fake_eval)(theta)
want = (theta ** 2).sum(dim=1) - theta[:, 0]
assert out.shape == (B,) and torch.equal(out, want), (out, want)
assert calls == [6 if dist.get_rank() == 0 else 5]
dist.destroy_process_group()
print("OK", flush=True)
"""


def test_sharded_allgather_gloo_world2(tmp_path):
    """N>1 path on CPU: two gloo ranks evaluate disjoint slices and all-gather the full logL vector."""
    _build()
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "OK" in o, o


_SCAN_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from pioran_b200.parallel import scan_logl_sharded, torch_collectives, scan_bounds
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), 2
# stand-in for the device scan with the same data flow: the "composite" of a range is the sum of its x, the state entering a
# range is the sum of the earlier composites, and the two returned sums depend on that incoming state
rng = np.random.default_rng(5)
N = 1001
x, wgt = rng.normal(size=N), rng.uniform(0.5, 1.5, N)
seen = dict()
def range_begin(lo, hi):
    seen["range"] = (lo, hi)
    return np.array([x[lo:hi].sum(), float(hi - lo)])
def range_end(prev):
    lo, hi = seen["range"]
    carry = 0.0 if prev is None or len(prev) == 0 else float(np.asarray(prev)[:, 0].sum())
    run = carry + np.cumsum(x[lo:hi])
    return np.array([np.sum(run * wgt[lo:hi]), np.sum(run ** 2)])
ag, ar = torch_collectives()
got = scan_logl_sharded(range_begin, range_end, N, rank=rank, world=world, all_gather=ag, all_reduce_sum=ar)
run = np.cumsum(x)
want = -0.5 * np.sum(run * wgt) - 0.5 * np.sum(run ** 2) - 0.5 * N * np.log(2 * np.pi)
off = scan_bounds(N, world)
assert off[1] % 2 == 0
assert seen["range"] == (int(off[rank]), int(off[rank + 1]))
assert abs(got - want) <= 1e-9 * abs(want), (got, want)
# self-check rows (Context.scan_range_check): consistent hand-over -> the scan's value; inconsistent -> every rank falls back
for delta, expect_fallback in ((0.0, False), (1e-3, True)):
    def range_check():
        if rank == 0:
            return np.array([1e-9, 0.0, 0.0, 3.5, 7.25, 0.0, 8.0, 10.0])
        return np.array([2e-9, 3.5 + delta, 7.25, 0.0, 0.0, 8.0, 0.0, 10.0])
    info = dict()
    got = scan_logl_sharded(range_begin, range_end, N, rank=rank, world=world, all_gather=ag, all_reduce_sum=ar,
                            range_check=range_check, sequential=lambda: 123.0, info=info)
    assert info["fallback"] == expect_fallback, info
    assert got == 123.0 if expect_fallback else abs(got - want) <= 1e-9 * abs(want)
    assert abs(info["estimate"] * max(1.0, abs(want)) - (3e-9 + 0.5 * 10.0 * delta)) <= 1e-12
dist.destroy_process_group()
print("OK", flush=True)
"""


def test_scan_check_total_rows():
    """parallel.scan_check_total: inner estimates + hand-over differences; a hand-over that was not swept twice, or a
    non-finite sum, gives NaN (which reads as 'not verified')."""
    from pioran_b200.parallel import scan_check_total
    rows = np.array([[1e-9, 0, 0, 2.0, 5.0, 0, 8, 4.0], [2e-9, 2.5, 5.0, 1.0, 1.0, 8, 8, 4.0], [0.0, 1.0, 1.5, 0, 0, 8, 0, 4.0]])
    assert abs(scan_check_total(rows) - (3e-9 + 0.5 * 4.0 * 0.5 + 0.5 * 4.0 * 0.5)) < 1e-15
    assert scan_check_total(rows[:1]) == 1e-9
    bad = rows.copy(); bad[1, 5] = 6
    assert np.isnan(scan_check_total(bad))
    bad = rows.copy(); bad[2, 1] = np.nan
    assert np.isnan(scan_check_total(bad))


def test_scan_time_axis_sharding_gloo_world2(tmp_path):
    """Host logic of the K3 time-axis split: ranges, all-gather of the composites, earlier composites handed to
    each rank in time order, 2-value all-reduce — two gloo ranks around a stand-in for the device calls."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "scan_worker.py"
    script.write_text(_SCAN_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "OK" in o, o


def test_prior_transform_matches_reference_prior():
    """sampler.PriorTransform restates prior_transform of examples/ultranest/single_pl.jl:96-104 (Distributions.jl quantiles):
    checked against the closed forms / scipy distributions column by column, scalar and batched calls agreeing."""
    from scipy import stats
    import pioran_b200 as pb
    f_min, f_max, xbar, va = 1e-3, 2.0, 0.3, 0.04
    tr = pb.sampler.single_bending_power_law_prior(f_min, f_max, xbar, va)
    u = np.random.default_rng(3).uniform(size=(257, 6))
    th = tr(u)
    f0, fM = f_min / 20.0, f_max * 20.0
    assert np.allclose(th[:, 0], 1.5 * u[:, 0], rtol=1e-15)
    assert np.allclose(th[:, 1], stats.loguniform(4 * f0, fM / 4).ppf(u[:, 1]), rtol=1e-12)
    assert np.allclose(th[:, 2], th[:, 0] + u[:, 2] * (4.0 - th[:, 0]), rtol=1e-15)       # α₂ ~ U(α₁, 4)
    assert np.allclose(th[:, 3], stats.lognorm(s=np.sqrt(2.0), scale=np.exp(-3.0)).ppf(u[:, 3]), rtol=1e-10)
    assert np.allclose(th[:, 4], stats.gamma(a=2, scale=0.5).ppf(u[:, 4]), rtol=1e-10)
    assert np.allclose(th[:, 5], stats.norm(xbar, 5 * np.sqrt(va)).ppf(u[:, 5]), rtol=1e-10, atol=1e-12)
    assert np.array_equal(tr(u[7]), th[7])
    assert np.all(th[:, 2] >= th[:, 0]) and np.all(th[:, 3] > 0) and np.all(th[:, 4] > 0)


def test_log_normal_prior_transform():
    """sampler.log_normal_prior restates the seven-parameter prior of docs/src/ultranest.md:220-229: the six columns of the
    single-bending-power-law prior (α₁ ≤ 1.25 there) plus c ~ LogUniform(1e-6, 0.99·min y)."""
    from scipy import stats
    import pioran_b200 as pb
    rng = np.random.default_rng(8)
    y = np.exp(rng.normal(0.3, 0.4, 200)) + 0.2
    f_min, f_max = 1e-3, 2.0
    tr = pb.sampler.log_normal_prior(f_min, f_max, y)
    u = rng.uniform(size=(129, 7))
    th = tr(u)
    assert np.allclose(th[:, 0], 1.25 * u[:, 0], rtol=1e-15)
    assert np.allclose(th[:, 2], th[:, 0] + u[:, 2] * (4.0 - th[:, 0]), rtol=1e-15)
    assert np.allclose(th[:, 5], stats.norm(np.log(y).mean(), 5 * np.log(y).std(ddof=1)).ppf(u[:, 5]), rtol=1e-10)
    assert np.allclose(th[:, 6], stats.loguniform(1e-6, 0.99 * y.min()).ppf(u[:, 6]), rtol=1e-12)
    assert np.all(th[:, 6] < y.min())
    assert np.allclose(tr(u[3]), th[3])


_CARMA32 = dict(rα=[-0.042163209825323775 + 1.1115603157767922j, -0.042163209825323775 - 1.1115603157767922j,
                    -0.7599101571312047 + 0.0j], β=[3.9413022090550216, 11.38193903188344, 1.0])


def test_carma_front_end_golden_vectors():
    """Host-side CARMA → celerite conversion against the reference's own literals: test/test_carma.jl:3-17
    (quad2roots), :19-33 (roots2coeffs), :53-70 (celerite_coefs), :96-113 (celerite ACVF ≡ CARMA ACVF), and the constructor's
    argument checks (src/CARMA.jl:28-37)."""
    import pioran_b200 as pb
    r = pb.quad2roots([0.025443151049354032, 0.04252858046335997, 2.5980088198563633])
    assert np.allclose(r, [-0.021264290231679986 + 0.1580853598860341j, -0.021264290231679986 - 0.1580853598860341j,
                           -2.5980088198563633 + 0.0j], rtol=1e-14)
    α = pb.roots2coeffs([-0.012721575524677016 + 0.20583182936448363j, -0.012721575524677016 - 0.20583182936448363j,
                         -2.5980088198563633 + 0.0j])
    assert np.allclose(α, [0.11048962713978024, 0.10863011129451944, 2.6234519709057174, 1], rtol=1e-13)
    cov = pb.CARMA(3, 2, _CARMA32["rα"], _CARMA32["β"], 1.3)
    a, b, c, d = pb.celerite_coefs(cov)
    assert np.allclose(a, [1.332733901854476, -0.03273390185447589], rtol=1e-12)
    assert np.allclose(b, [-0.026820976815752837, 0.0], rtol=1e-12)
    assert np.allclose(c, [0.042163209825323775, 0.7599101571312047], rtol=1e-15)
    assert np.allclose(d, [-1.1115603157767922, 0.0], rtol=1e-15)
    rep = pb.celerite_repr(cov)
    assert isinstance(rep, pb.SumOfCelerite) and np.array_equal(rep.a, a)
    t = np.linspace(0, 150, 1000)
    assert np.allclose([rep(x, 0.0) for x in t], cov.covariance(t), rtol=1e-10, atol=1e-14)
    assert abs(np.sum(a) - 1.3) < 1e-14                                   # integrated power: Σa = norm
    a2, _, _, _ = pb.celerite_coefs(pb.CARMA(3, 2, _CARMA32["rα"], _CARMA32["β"], 1.0, False))
    assert np.allclose(a2 / a2[0], a / a[0], rtol=1e-13)
    # even order: every term is a conjugate pair
    a4, b4, c4, d4 = pb.carma_celerite_coefs(4, pb.quad2roots([0.5, 0.2, 3.0, 0.4]), [1.0, 0.3])
    assert a4.shape == (2,) and np.all(d4 != 0.0) and abs(a4.sum() - 1.0) < 1e-14
    for bad in ((0, 0, [], [1.0]), (2, 3, [1j, -1j], [1, 1, 1, 1]), (3, 2, [1j, -1j], [1, 1, 1]), (3, 2, _CARMA32["rα"], [1, 1])):
        with pytest.raises(ValueError):
            pb.CARMA(*bad)


def test_batched_carma_coefficients_host_logic():
    """BatchedCARMALikelihood.coefficients (no device needed): quad → roots → bounds check → celerite coefficients, with the
    MA coefficients β = roots2coeffs(quad2roots(qb)) (docs/src/carma.md:26-42); the literal of test/test_carma.jl:53-70 comes back."""
    import pioran_b200 as pb
    like = pb.BatchedCARMALikelihood.__new__(pb.BatchedCARMALikelihood)
    like.p, like.q, like.f_min, like.f_max, like.log_shift, like.n_par = 3, 2, 1e-3, 50.0, False, 8
    rα = np.array(_CARMA32["rα"])
    qa = [abs(rα[0]) ** 2, -2 * rα[0].real, -rα[2].real]            # x² + qa[1] x + qa[0], x + qa[2]
    rβ = np.roots([1.0, 11.38193903188344, 3.9413022090550216])
    qb = [np.prod(rβ).real, -np.sum(rβ).real]
    th = np.array([qa + qb + [1.3, 1.0, 0.0], qa[:2] + [100.0] + qb + [1.3, 1.0, 0.0]])
    ok, a, b, c, d = like.coefficients(th)
    assert ok.tolist() == [True, False]
    assert np.allclose(a[0], [1.332733901854476, -0.03273390185447589], rtol=1e-10)
    assert np.allclose(b[0], [-0.026820976815752837, 0.0], rtol=1e-9)
    assert np.allclose(c[0], [0.042163209825323775, 0.7599101571312047], rtol=1e-12)
    assert np.allclose(d[0], [-1.1115603157767922, 0.0], rtol=1e-12)
    with pytest.raises(ValueError):
        like.coefficients(th[:, :5])
