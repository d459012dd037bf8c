"""ctypes binding of oracle/libpioran_oracle.so — the CPU restatement of Pioran.jl's likelihood path.

TEST INFRASTRUCTURE ONLY (see the header of pioran_oracle.c).  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libpioran_oracle.so")

PSD_MODELS = {"SingleBendingPowerLaw": 0, "DoubleBendingPowerLaw": 1, "SBPL": 0, "DBPL": 1, 0: 0, 1: 1}
BASES = {"SHO": 0, "DRWCelerite": 1, 0: 0, 1: 1}
N_PSD_PAR = {0: 3, 1: 5}

_dp = C.POINTER(C.c_double)


def build():
    """Compile the oracle with its Makefile (gcc only)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _load():
    srcs = [os.path.join(_HERE, f) for f in ("pioran_oracle.c", "pioran_oracle_grad.c", "pioran_oracle_grad_ld.c")]
    if not os.path.exists(_LIB) or any(os.path.getmtime(_LIB) < os.path.getmtime(src) for src in srcs):
        build()
    lib = C.CDLL(_LIB)
    lib.orc_psd_eval.restype = C.c_double
    lib.orc_psd_eval.argtypes = [C.c_int, _dp, C.c_double]
    lib.orc_build_approx.restype = None
    lib.orc_build_approx.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, _dp, _dp]
    lib.orc_get_approx_coefficients.restype = C.c_int
    lib.orc_get_approx_coefficients.argtypes = [C.c_int, _dp, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp]
    lib.orc_integrate_basis.restype = C.c_double
    lib.orc_integrate_basis.argtypes = [C.c_int, _dp, _dp, C.c_double, C.c_double, C.c_int]
    lib.orc_approx.restype = C.c_int
    lib.orc_approx.argtypes = [C.c_int, _dp, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double,
                               C.c_int, C.c_int, _dp, _dp, _dp, _dp]
    lib.orc_approx_features.restype = C.c_int
    lib.orc_approx_features.argtypes = [C.c_int, _dp, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double,
                                        C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
    lib.orc_integral_celerite.restype = C.c_double
    lib.orc_integral_celerite.argtypes = [C.c_double] * 5
    for name, rt in (("orc_celerite_logl", C.c_double), ("orc_celerite_logl_ld", C.c_longdouble)):
        fn = getattr(lib, name)
        fn.restype = rt
        fn.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_int64, _dp, _dp, _dp]
    lib.orc_direct_nll.restype = C.c_double
    lib.orc_direct_nll.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_int64, _dp, _dp, _dp, C.POINTER(C.c_int)]
    lib.orc_approx_logl_batch.restype = None
    lib.orc_approx_logl_batch.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_double, C.c_double, C.c_int, C.c_double,
                                          C.c_double, C.c_int, C.c_int, C.c_int64, _dp, _dp, _dp, _dp, C.c_int]
    for name in ("orc_approx_logl_grad_batch", "orc_approx_logl_grad_batch_ld", "orc_approx_logl_logshift_grad_batch",
                 "orc_approx_logl_logshift_grad_batch_ld"):
        fn = getattr(lib, name)
        fn.restype = None
        fn.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int,
                       C.c_int, C.c_int64, _dp, _dp, _dp, _dp, _dp, C.c_int]
    lib.orc_celerite_logl_batch.restype = None
    lib.orc_celerite_logl_batch.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int64, _dp, _dp, _dp,
                                            _dp, C.c_int]
    for name in ("orc_celerite_predict", "orc_direct_predict"):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_int64, _dp, C.c_int64, _dp, _dp, _dp, _dp]
    lib.orc_celerite_simulate.restype = C.c_int
    lib.orc_celerite_simulate.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_int64, _dp, _dp, _dp, _dp]
    lib.orc_max_threads.restype = C.c_int
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _arr(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _p(x):
    return x.ctypes.data_as(_dp) if x is not None else None


def psd_eval(model, params, f):
    m = PSD_MODELS[model]
    par = _arr(params)
    f = np.atleast_1d(_arr(f))
    return np.array([lib().orc_psd_eval(m, _p(par), float(x)) for x in f])


def build_approx(J, f0, fM, basis="SHO"):
    fj = np.empty(J)
    B = np.empty((J, J), order="F")
    lib().orc_build_approx(J, f0, fM, BASES[basis], _p(fj), B.ctypes.data_as(_dp))
    return fj, B


def get_approx_coefficients(model, params, f0, fM, J=20, basis="SHO"):
    amp = np.empty(J)
    fj = np.empty(J)
    par = _arr(params)
    rc = lib().orc_get_approx_coefficients(PSD_MODELS[model], _p(par), f0, fM, J, BASES[basis], _p(amp), _p(fj))
    if rc:
        raise RuntimeError(f"oracle LU failed rc={rc}")
    return amp, fj


def integrate_basis(a, c, x1, x2, basis="SHO"):
    a, c = _arr(a), _arr(c)
    return lib().orc_integrate_basis(len(a), _p(a), _p(c), x1, x2, BASES[basis])


def approx(model, params, f_min, f_max, J=20, norm=1.0, S_low=20.0, S_high=20.0, is_integrated_power=True, basis="SHO"):
    """Restatement of Pioran.approx (src/psd.jl:214-289) → (a, b, c, d) celerite coefficient vectors."""
    par = _arr(params)
    a, b, c, d = (np.empty(2 * J) for _ in range(4))
    Jt = lib().orc_approx(PSD_MODELS[model], _p(par), f_min, f_max, J, norm, S_low, S_high, int(is_integrated_power),
                          BASES[basis], _p(a), _p(b), _p(c), _p(d))
    if Jt < 0:
        raise RuntimeError(f"oracle approx failed rc={Jt}")
    return a[:Jt].copy(), b[:Jt].copy(), c[:Jt].copy(), d[:Jt].copy()


def approx_features(model, params, features, f_min, f_max, J=20, norm=1.0, S_low=20.0, S_high=20.0, is_integrated_power=True,
                    basis="SHO"):
    """approx of continuum + QPO features [(S0, f0, Q), …] (src/psd.jl:214-289 with :15-44, :229-243) → (a, b, c, d)."""
    par = _arr(params)
    feat = _arr(np.asarray(features, dtype=np.float64).ravel())
    nf = len(feat) // 3
    a, b, c, d = (np.empty(2 * J + nf) for _ in range(4))
    Jt = lib().orc_approx_features(PSD_MODELS[model], _p(par), f_min, f_max, J, norm, S_low, S_high, int(is_integrated_power),
                                   BASES[basis], nf, _p(feat), _p(a), _p(b), _p(c), _p(d))
    if Jt < 0:
        raise RuntimeError(f"oracle approx_features failed rc={Jt}")
    return a[:Jt].copy(), b[:Jt].copy(), c[:Jt].copy(), d[:Jt].copy()


def integral_celerite(a, b, c, d, x):
    return lib().orc_integral_celerite(a, b, c, d, x)


def celerite_logl(a, b, c, d, t, y, s2, long_double=False):
    """Restatement of Pioran.logl (src/celerite_solver.jl:312-334)."""
    a, b, c, d, t, y, s2 = map(_arr, (a, b, c, d, t, y, s2))
    fn = lib().orc_celerite_logl_ld if long_double else lib().orc_celerite_logl
    return fn(len(a), _p(a), _p(b), _p(c), _p(d), len(t), _p(t), _p(y), _p(s2))


def direct_nll(a, b, c, d, t, y, s2):
    """Restatement of Pioran.log_likelihood_direct (src/direct_solver.jl:6-21): returns (+NLL, info)."""
    a, b, c, d, t, y, s2 = map(_arr, (a, b, c, d, t, y, s2))
    info = C.c_int(0)
    v = lib().orc_direct_nll(len(a), _p(a), _p(b), _p(c), _p(d), len(t), _p(t), _p(y), _p(s2), C.byref(info))
    return v, info.value


def approx_logl_batch(model, theta, f_min, f_max, J, t, y, s2, basis="SHO", S_low=20.0, S_high=20.0,
                      is_integrated_power=True, nthreads=1):
    """theta rows = [psd params…, norm, ν, μ]; returns logL[B] (approx + logpdf per row)."""
    m = PSD_MODELS[model]
    theta = np.atleast_2d(_arr(theta))
    npar = N_PSD_PAR[m]
    assert theta.shape[1] == npar + 3
    t, y, s2 = map(_arr, (t, y, s2))
    out = np.empty(theta.shape[0])
    lib().orc_approx_logl_batch(m, npar, theta.shape[0], _p(theta), f_min, f_max, J, S_low, S_high,
                                int(is_integrated_power), BASES[basis], len(t), _p(t), _p(y), _p(s2), _p(out), nthreads)
    return out


def approx_logl_grad_batch(model, theta, f_min, f_max, J, t, y, s2, basis="SHO", S_low=20.0, S_high=20.0,
                           is_integrated_power=True, nthreads=1, long_double=False):
    """Forward-mode gradient (pioran_oracle_grad.c): returns (logL[B], ∂logL/∂θ [B × (npar+3)]).  long_double: the same
    code in 80-bit arithmetic (pioran_oracle_grad_ld.c), results rounded to FP64 — conditioning triage only."""
    m = PSD_MODELS[model]
    theta = np.atleast_2d(_arr(theta))
    npar = N_PSD_PAR[m]
    assert theta.shape[1] == npar + 3
    t, y, s2 = map(_arr, (t, y, s2))
    out = np.empty(theta.shape[0])
    grad = np.empty(theta.shape)
    fn = lib().orc_approx_logl_grad_batch_ld if long_double else lib().orc_approx_logl_grad_batch
    fn(m, npar, theta.shape[0], _p(theta), f_min, f_max, J, S_low, S_high, int(is_integrated_power), BASES[basis], len(t),
       _p(t), _p(y), _p(s2), _p(out), _p(grad), nthreads)
    return out, grad


def approx_logl_logshift_grad_batch(model, theta, f_min, f_max, J, t, y, s2, basis="SHO", S_low=20.0, S_high=20.0,
                                    is_integrated_power=True, nthreads=1, long_double=False):
    """Log-normal model: theta rows = [psd params…, norm, ν, μ, c]; y, s2 the UNtransformed flux and its variance.
    Returns (logL[B], ∂logL/∂θ [B × (npar+4)])."""
    m = PSD_MODELS[model]
    theta = np.atleast_2d(_arr(theta))
    npar = N_PSD_PAR[m]
    assert theta.shape[1] == npar + 4
    t, y, s2 = map(_arr, (t, y, s2))
    out = np.empty(theta.shape[0])
    grad = np.empty(theta.shape)
    fn = lib().orc_approx_logl_logshift_grad_batch_ld if long_double else lib().orc_approx_logl_logshift_grad_batch
    fn(m, npar, theta.shape[0], _p(theta), f_min, f_max, J, S_low, S_high, int(is_integrated_power), BASES[basis], len(t),
       _p(t), _p(y), _p(s2), _p(out), _p(grad), nthreads)
    return out, grad


def celerite_logl_batch(a, b, c, d, t, y, s2, mu=None, nu=None, nthreads=1):
    a, b, c, d = (np.atleast_2d(_arr(x)) for x in (a, b, c, d))
    t, y, s2 = map(_arr, (t, y, s2))
    B, Jt = a.shape
    mu = _arr(mu) if mu is not None else None
    nu = _arr(nu) if nu is not None else None
    out = np.empty(B)
    lib().orc_celerite_logl_batch(B, Jt, _p(a), _p(b), _p(c), _p(d), _p(mu), _p(nu), len(t), _p(t), _p(y), _p(s2), _p(out),
                                  nthreads)
    return out


def celerite_predict(a, b, c, d, tau, t, y, s2):
    """pred(a, b, c, d, τ, t, y, σ²) — src/celerite_solver.jl:376-483 (posterior mean at ascending τ)."""
    a, b, c, d, tau, t, y, s2 = map(_arr, (a, b, c, d, tau, t, y, s2))
    out = np.empty(len(tau))
    rc = lib().orc_celerite_predict(len(a), _p(a), _p(b), _p(c), _p(d), len(tau), _p(tau), len(t), _p(t), _p(y), _p(s2), _p(out))
    if rc:
        raise MemoryError("orc_celerite_predict")
    return out


def direct_predict(a, b, c, d, tau, t, y, s2):
    """predict_direct (mean) — src/direct_solver.jl:74-119."""
    a, b, c, d, tau, t, y, s2 = map(_arr, (a, b, c, d, tau, t, y, s2))
    out = np.empty(len(tau))
    rc = lib().orc_direct_predict(len(a), _p(a), _p(b), _p(c), _p(d), len(tau), _p(tau), len(t), _p(t), _p(y), _p(s2), _p(out))
    if rc:
        raise ArithmeticError("orc_direct_predict: matrix not positive definite" if rc == 1 else "allocation failed")
    return out


def celerite_simulate(a, b, c, d, t, s2, q):
    """sim(rng, a, b, c, d, τ, σ²) with the normal draws q supplied — src/celerite_solver.jl:515-549."""
    a, b, c, d, t, s2, q = map(_arr, (a, b, c, d, t, s2, q))
    out = np.empty(len(t))
    rc = lib().orc_celerite_simulate(len(a), _p(a), _p(b), _p(c), _p(d), len(t), _p(t), _p(s2), _p(q), _p(out))
    if rc:
        raise MemoryError("orc_celerite_simulate")
    return out


def max_threads():
    return lib().orc_max_threads()
