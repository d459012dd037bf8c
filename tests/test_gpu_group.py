"""Device groups (pioran_ctx_create_multi, include/pioran_b200.h): one process, one context over several GPUs — what a sampler
that is a single process with one callback needs (reference: examples/ultranest/single_pl.jl:113-117).  A group of one device
exercises the whole forwarding layer on the single-GPU box; the two-device case runs where two GPUs are visible."""
import numpy as np
import pytest

from conftest import prior_theta, rel_err, synthetic_series

pytestmark = pytest.mark.gpu


def _ndev():
    import ctypes
    try:
        rt = ctypes.CDLL("libcudart.so")
    except OSError:
        import torch
        return torch.cuda.device_count()
    n = ctypes.c_int(0)
    rt.cudaGetDeviceCount(ctypes.byref(n))
    return n.value


@pytest.mark.parametrize("devices", [[0], [0, 1]])
def test_group_context_matches_single_device(devices):
    import pioran_b200 as pb
    if len(devices) > _ndev():
        pytest.skip("needs %d GPUs" % len(devices))
    one = pb.get_context(0)
    grp = pb.Context(devices)
    assert grp.device_count == len(devices)
    try:
        t, y, s2, f_min, f_max = synthetic_series(700, 3)
        th = prior_theta(1001, f_min, f_max, y.mean(), y.std(), 11, alpha2_max=3.5)
        for basis in ("SHO", "DRWCelerite"):
            spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20, basis_function=basis)
            s1, sg = one.upload_series(t, y, s2), grp.upload_series(t, y, s2)
            want = one.approx_logl(s1, spec, th)[0]
            got = grp.approx_logl(sg, spec, th)[0]
            assert np.array_equal(got, want), basis           # same kernels, same inputs per row
            # gradients and explicit coefficients split the same way
            lw, gw = one.approx_logl_grad(s1, spec, th[:37])
            lg, gg = grp.approx_logl_grad(sg, spec, th[:37])
            assert np.array_equal(lg, lw) and np.array_equal(gg, gw)
            a, b, c, d = grp.approx_coeffs(spec, th[:53, :4])
            assert np.array_equal(grp.celerite_logl(sg, a, b, c, d, mu=th[:53, 5], nu=th[:53, 4]),
                                  one.celerite_logl(s1, a, b, c, d, mu=th[:53, 5], nu=th[:53, 4]))
            # the dense cross-check splits its parameter vectors the same way
            a9, b9, c9, d9 = grp.approx_coeffs(spec, th[:9, :4])
            nw, iw = one.direct_logl(s1, a9, b9, c9, d9, mu=th[:9, 5], nu=th[:9, 4])
            ng, ig = grp.direct_logl(sg, a9, b9, c9, d9, mu=th[:9, 5], nu=th[:9, 4])
            assert np.array_equal(ng, nw) and np.array_equal(ig, iw)
            s1.free(); sg.free()
        # several series in one call
        sers1, sersg, specs = [], [], []
        for k in range(3):
            tk, yk, sk, fm, fx = synthetic_series(200 + 37 * k, 20 + k)
            sers1.append(one.upload_series(tk, yk, sk)); sersg.append(grp.upload_series(tk, yk, sk))
            specs.append(pb.make_spec("SingleBendingPowerLaw", fm, fx, 12))
        want = one.approx_logl(sers1, specs, th[:101])
        got = grp.approx_logl(sersg, specs, th[:101])
        assert got.shape == (3, 101) and np.array_equal(got, want)
        # device-pointer entries and stream injection need a single-device context
        with pytest.raises(pb.PioranError) as e:
            grp.set_stream(0x1234)
        assert e.value.code == -5
        with pytest.raises(pb.PioranError):
            grp.approx_logl(pb.backend.Series(grp, 99, 10), specs[0], th[:4])      # unknown series id
    finally:
        grp.close()


def test_resident_time_vector_drop_in():
    """log_likelihood(cov, τ, y, σ²; solver = :celerite_gpu) keeps τ on the device between calls (julia/b200_solver.jl:
    b200_resident_series; Python mirror api._resident_series): the second call uploads only y and σ²."""
    import pioran_b200 as pb
    from oracle import oracle as orc
    t, y, s2, f_min, f_max = synthetic_series(300, 5)
    cov = pb.approx(pb.SingleBendingPowerLaw(0.6, 0.02, 3.1), f_min, f_max, 20, 1.3, basis_function="DRWCelerite")
    a, b, c, d = pb.celerite_coefs(cov)
    pb.api.release_resident_series()
    for k, (mu, nu) in enumerate(((0.0, 1.0), (0.3, 1.7), (-0.2, 0.6))):
        got = pb.log_likelihood(cov, t, y - mu, nu * s2, solver="celerite_gpu")
        want = orc.celerite_logl(a, b, c, d, t, y - mu, nu * s2)
        assert rel_err(got, want) <= 1e-9
        assert len(pb.api._resident) == 1
    assert rel_err(pb.log_likelihood(cov, t, y, s2, solver=":celerite_matrix"), orc.celerite_logl(a, b, c, d, t, y, s2)) <= 1e-9
    pb.api.release_resident_series()
    assert len(pb.api._resident) == 0
