"""The C ABI as a foreign-language binding sees it: tests/c_abi/abi_check.c is compiled with gcc -std=c11 against
include/pioran_b200.h (no CUDA headers, no C++), dlopens the library and (a) reports the layout of pioran_approx_spec and the
error codes, which must agree with the ctypes binding (and with struct B200ApproxSpec of julia/b200_solver.jl), (b) on the GPU,
evaluates one parameter vector of the reference's shipped simu_single run through pioran_approx_logl and reproduces its logL."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "c_abi", "abi_check.c")


@pytest.fixture(scope="module")
def abi_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("abi") / "abi_check")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-ldl", "-o", exe])
    return exe


def _lib_path():
    import pioran_b200
    pioran_b200.build.build()
    return pioran_b200._lib.LIB_PATH


def test_struct_layout_and_error_codes_match_the_bindings(abi_check):
    out = json.loads(subprocess.check_output([abi_check, "layout", _lib_path()], text=True))
    from pioran_b200._lib import ApproxSpec
    assert out["sizeof_spec"] == C.sizeof(ApproxSpec) == 48
    for f in ("psd_model", "n_components", "basis", "is_integrated_power", "f_min", "f_max", "S_low", "S_high"):
        assert out["off_" + f] == getattr(ApproxSpec, f).offset, f
    assert [out[k] for k in ("PIORAN_OK", "PIORAN_EINVAL", "PIORAN_ECUDA", "PIORAN_ENOMEM", "PIORAN_ESINGULAR", "PIORAN_EUNSUPPORTED")] \
        == [0, -1, -2, -3, -4, -5]
    from pioran_b200._lib import PriorSpec
    assert out["sizeof_prior"] == C.sizeof(PriorSpec) == 24
    for f in ("kind", "ref_col", "p0", "p1"):
        assert out["off_prior_" + f] == getattr(PriorSpec, f).offset, f
    assert out["version"] >= 200
    # the Julia struct of the shim declares the same fields in the same order
    jl = open(os.path.join(ROOT, "julia", "b200_solver.jl")).read()
    body = jl[jl.index("struct B200ApproxSpec"):]
    body = body[:body.index("end")]
    order = [body.index(f) for f in ("psd_model", "n_components", "basis", "is_integrated_power", "f_min", "f_max", "S_low", "S_high")]
    assert order == sorted(order)


@pytest.mark.gpu
def test_c_program_reproduces_a_shipped_julia_loglikelihood(abi_check, golden_single, tmp_path):
    g = golden_single
    series = tmp_path / "series.txt"
    np.savetxt(series, np.column_stack([g.t, g.y, g.s2]), fmt="%.17g")
    for row in (0, 1234, 4000):
        th = g.theta[row]
        out = subprocess.check_output([abi_check, "call", _lib_path(), str(series), "20", "0"] + [f"{v:.17g}" for v in th], text=True)
        got, want = float(out.strip()), g.logl[row]
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (row, got, want)
