"""PSD features (QPO) through approx(): the reference's feature branch (src/psd.jl:15-44 convert_feature /
get_covariance_from_psd, :221-243, :254-259, :277-282, :330-334 integral_celerite, :380-388) — oracle restatement, host mirror
and the K1 kernel's feature branch; term counts as in test/test_psd.jl:206-285."""
import numpy as np
import pytest

from conftest import rel_err, synthetic_series
from oracle import oracle as orc

# test/test_psd.jl:207-211, 249-253
F_MIN, F_MAX, J = 2.0e-3, 3.52e2, 25
CONT_SHO, CONT_DRW = [0.2, 1.3e-2, 3.2], [0.2, 1.3e-2, 4.2]
VA = 1.32


def test_oracle_feature_term_counts_and_normalisation():
    # test/test_psd.jl:214-224, 235-245, 255-265, 275-285: length(Rapprox.cov) == J + nfeat (SHO), 2J + nfeat (DRWCelerite)
    for basis, cont, feats, want in (("SHO", CONT_SHO, [2.0, 1.0e-2, 14.2], J + 1),
                                     ("SHO", CONT_SHO, [2.0, 1.0e-2, 14.2, 4.0, 1.0e-1, 4.2], J + 2),
                                     ("DRWCelerite", CONT_DRW, [1.4, 1.0e-2, 10.2], 2 * J + 1),
                                     ("DRWCelerite", CONT_DRW, [1.4, 1.0e-2, 10.2, 2.4, 5.0e-2, 12.2], 2 * J + 2)):
        a, b, c, d = orc.approx_features("SBPL", cont, feats, F_MIN, F_MAX, J, VA, is_integrated_power=False, basis=basis)
        assert len(a) == want
        # without integrated power the features do not enter the norm (src/psd.jl:389-395): the continuum terms are those of the plain approx
        a0, b0, c0, d0 = orc.approx("SBPL", cont, F_MIN, F_MAX, J, VA, is_integrated_power=False, basis=basis)
        nf = len(feats) // 3
        assert np.array_equal(a[:-nf], a0) and np.array_equal(c[:-nf], c0)
        # feature terms: c = ω₀/(2Q), d = c·sqrt(4Q² − 1), b = a/sqrt(4Q² − 1)  (src/psd.jl:17-24)
        for k in range(nf):
            S0, f0, Q = feats[3 * k:3 * k + 3]
            dl = np.sqrt(4 * Q * Q - 1)
            assert c[-nf + k] == 2 * np.pi * f0 / Q / 2 and np.isclose(d[-nf + k], c[-nf + k] * dl, rtol=1e-15)
            assert np.isclose(b[-nf + k], a[-nf + k] / dl, rtol=1e-14)
    # integrated power: the celerite-PSD integrals of ALL terms over [f_min, f_max] add up to 2·norm (two-sided convention:
    # integrate_basis_function + Σ integrate_psd_feature = norm, and every emitted term carries the factor 2 of src/psd.jl:254-259)
    a, b, c, d = orc.approx_features("SBPL", CONT_SHO, [2.0, 1.0e-2, 14.2, 4.0, 1.0e-1, 4.2], F_MIN, F_MAX, J, VA, basis="SHO")
    tot = sum(orc.integral_celerite(a[k], b[k], c[k], d[k], F_MAX) - orc.integral_celerite(a[k], b[k], c[k], d[k], F_MIN)
              for k in range(len(a)))
    assert abs(tot - 2 * VA) < 1e-12 * VA


def test_host_mirror_of_the_feature_helpers():
    import pioran_b200 as pb
    m = pb.SingleBendingPowerLaw(*CONT_SHO) + pb.QPO(2.0, 1.0e-2, 14.2) + pb.QPO(4.0, 1.0e-1, 4.2)
    cont, feats = pb.separate_psd(m)
    assert isinstance(cont, pb.SingleBendingPowerLaw) and len(feats) == 2
    cov = pb.get_covariance_from_psd(feats)
    assert cov.shape == (4, 2)
    S0, f0, Q = 2.0, 1.0e-2, 14.2
    assert np.allclose(cov[:, 0], [S0 * 2 * np.pi * f0 * Q / 4, S0 * 2 * np.pi * f0 * Q / 4 / np.sqrt(4 * Q * Q - 1), 2 * np.pi * f0 / Q / 2,
                                   2 * np.pi * f0 / Q / 2 * np.sqrt(4 * Q * Q - 1)], rtol=1e-15)
    assert pb.separate_psd(pb.SingleBendingPowerLaw(*CONT_SHO)) [1] is None
    with pytest.raises(ValueError):
        pb.convert_feature(pb.SingleBendingPowerLaw(*CONT_SHO))          # "Feature … not implemented" (src/psd.jl:26)
    with pytest.raises(ValueError):
        pb.separate_psd(pb.SingleBendingPowerLaw(*CONT_SHO) + pb.DoubleBendingPowerLaw(0.1, 1e-2, 2.0, 1.0, 3.0))


@pytest.mark.gpu
@pytest.mark.parametrize("basis,cont", [("SHO", CONT_SHO), ("DRWCelerite", CONT_DRW)])
def test_k1_feature_branch_and_fused_logl_vs_oracle(basis, cont):
    import pioran_b200 as pb
    ctx = pb.get_context(0)
    rng = np.random.default_rng(7)
    for integrated in (True, False):
        spec = pb.make_spec("SingleBendingPowerLaw", F_MIN, F_MAX, J, is_integrated_power=integrated, basis_function=basis)
        for nf in (1, 2, 3):
            B = 23
            feats = np.column_stack([v for _ in range(nf) for v in (rng.uniform(0.5, 5, B), np.exp(rng.uniform(np.log(5e-3), np.log(5.0), B)),
                                                                    rng.uniform(0.8, 20, B))])
            cth = np.column_stack([rng.uniform(0.0, 1.0, B), np.exp(rng.uniform(np.log(5e-3), np.log(10.0), B)),
                                   rng.uniform(cont[2] - 1.0, cont[2], B), np.exp(rng.normal(0, 1, B))])
            a, b, c, d = ctx.approx_coeffs_features(spec, nf, np.column_stack([cth, feats]))
            Jt = (J if basis == "SHO" else 2 * J) + nf
            assert a.shape == (B, Jt)
            for i in range(B):
                oa, ob, oc, od = orc.approx_features("SBPL", cth[i, :3], feats[i], F_MIN, F_MAX, J, cth[i, 3], is_integrated_power=integrated,
                                                     basis=basis)
                sc = np.abs(oa).max()
                assert np.abs(a[i] - oa).max() <= 1e-11 * sc and np.abs(b[i] - ob).max() <= 1e-11 * max(sc, np.abs(ob).max())
                assert np.allclose(c[i], oc, rtol=1e-14) and np.allclose(d[i], od, rtol=1e-14)
    # the reference's own shapes through the mirror of the Julia interface (test/test_psd.jl:217-223, 258-264)
    PS = pb.SingleBendingPowerLaw(*cont) + pb.QPO(2.0, 1.0e-2, 14.2) + pb.QPO(4.0, 1.0e-1, 4.2)
    R = pb.approx(PS, F_MIN, F_MAX, J, VA, is_integrated_power=False, basis_function=basis)
    assert isinstance(R, pb.SumOfCelerite) and len(R.a) == (J if basis == "SHO" else 2 * J) + 2
    # fused likelihood of a batch with one QPO against the oracle (coefficients + reference-order sweep)
    t, y, s2, f_min, f_max = synthetic_series(400, 9)
    spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 20, basis_function=basis)
    like = pb.BatchedLikelihood(t, y, s2, "SingleBendingPowerLaw", 20, basis, f_min=f_min, f_max=f_max, ctx=ctx, n_features=1)
    B = 40
    th = np.column_stack([rng.uniform(0.0, 1.0, B), np.exp(rng.uniform(np.log(f_min), np.log(f_max), B)), rng.uniform(2.0, 3.5, B),
                          np.exp(rng.normal(-1, 0.5, B)), rng.gamma(2, 0.5, B), rng.normal(0, 0.3, B),
                          rng.uniform(0.5, 3, B), np.exp(rng.uniform(np.log(2 * f_min), np.log(f_max / 2), B)), rng.uniform(1.0, 15, B)])
    got = like(th)
    for i in range(B):
        a, b, c, d = orc.approx_features("SBPL", th[i, :3], th[i, 6:9], f_min, f_max, 20, th[i, 3], basis=basis)
        want = orc.celerite_logl(a, b, c, d, t, y - th[i, 5], th[i, 4] * s2)
        ld = float(orc.celerite_logl(a, b, c, d, t, y - th[i, 5], th[i, 4] * s2, long_double=True))
        assert rel_err(got[i], want) <= 1e-9 or rel_err(got[i], ld) <= 4 * rel_err(want, ld), (i, got[i], want)
    like.close()
