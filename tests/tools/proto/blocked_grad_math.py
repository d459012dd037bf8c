"""Math-level prototype of the forward-mode tangent of the blocked celerite sweep (csrc/blocked_grad.cuh), checked against
central differences of the blocked sweep itself.  Directions: row amplitudes (with their Σa) and ν; μ through a second data row."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from blocked_math import rows_from_coeffs, ROOT
sys.path.insert(0, ROOT)
from oracle import oracle as orc

BK = 8

def sweep(amp, suma, rows, t, y, s2, mu, nu, damp=None, dsuma=0.0, dnu=0.0):
    """returns logL and (if damp is given) its derivative along (damp, dsuma, dnu), plus dlogL/dmu"""
    R = len(rows)
    cr = np.array([r[1] for r in rows]); dr = np.array([r[2] for r in rows]); ratio = np.array([r[3] for r in rows]); kind = np.array([r[4] for r in rows])
    N = len(t)
    arg = np.outer(t, dr); co, si = np.cos(arg), np.sin(arg)
    Ut = np.where(kind == 0, co + ratio * si, np.where(kind == 1, si - ratio * co, 1.0))
    V = np.where(kind == 0, co, np.where(kind == 1, si, 1.0))
    phi = np.zeros((N, R)); phi[1:] = np.exp(-np.outer(np.diff(t), cr))
    tang = damp is not None
    X = np.zeros((R, R)); g = np.zeros(R); gm = np.zeros(R)       # gm: second data row (d/dmu)
    Xd = np.zeros((R, R)); gd = np.zeros(R)
    logdet = chi2 = dlog = dchi = chimu = 0.0
    for n1 in range(0, N, BK):
        idx = np.arange(n1, min(n1 + BK, N)); nb = len(idx)
        Psi0 = np.cumprod(phi[idx], axis=0); Uh = (Psi0 * Ut[idx]).T
        PsiE = np.ones((nb, R))
        for s in range(nb - 2, -1, -1): PsiE[s] = PsiE[s + 1] * phi[idx[s + 1]]
        psi8 = Psi0[-1]; Vh = (PsiE * V[idx]).T
        H = np.zeros((nb, nb, R))
        for s in range(nb):
            dec = np.ones(R)
            for sp in range(s - 1, -1, -1):
                dec = dec * phi[idx[sp + 1]]; H[s, sp] = H[sp, s] = Ut[idx[s]] * dec * V[idx[sp]]
        def kblk(am, sa, nn):
            K = H @ am
            K[np.arange(nb), np.arange(nb)] = sa + nn * s2[idx]
            return K
        P0 = X @ Uh
        C = kblk(amp, suma, nu) - Uh.T @ P0
        L = np.eye(nb); D = np.zeros(nb)
        for j in range(nb):
            D[j] = C[j, j] - np.sum(L[j, :j] ** 2 * D[:j])
            for i in range(j + 1, nb):
                L[i, j] = (C[i, j] - np.sum(L[i, :j] * L[j, :j] * D[:j])) / D[j]
        E = np.linalg.inv(L); rd = 1.0 / D
        Bm = amp[:, None] * Vh - psi8[:, None] * P0
        r = y[idx] - mu - Uh.T @ g           # data row
        rm = -np.ones(nb) - Uh.T @ gm        # second data row: d(y - mu)/dmu = -1
        Q = Bm @ E.T; z = E @ r; zm = E @ rm
        W = Q * rd[None, :]
        chi2 += np.sum(z * z * rd); logdet += np.sum(np.log(D)); chimu += np.sum(2 * z * zm * rd)
        if tang:
            P0d = Xd @ Uh
            Cd = kblk(damp, dsuma, dnu) - Uh.T @ P0d
            M = E @ Cd @ E.T
            Dd = np.diag(M).copy()
            G = np.tril(M, -1) * rd[None, :]
            Ed = -G @ E
            Bmd = damp[:, None] * Vh - psi8[:, None] * P0d
            Qd = Bmd @ E.T + Bm @ Ed.T
            rdot = -Uh.T @ gd
            zd = E @ rdot + Ed @ r
            Wd = Qd * rd[None, :] - Q * (Dd * rd * rd)[None, :]
            dlog += np.sum(Dd * rd); dchi += np.sum(2 * z * zd * rd - z * z * Dd * rd * rd)
            Xd = np.outer(psi8, psi8) * Xd + Qd @ W.T + Q @ Wd.T
            gd = psi8 * gd + Wd @ z + W @ zd
        X = np.outer(psi8, psi8) * X + Q @ W.T
        g = psi8 * g + W @ z
        gm = psi8 * gm + W @ zm
    logl = -0.5 * logdet - 0.5 * chi2 - 0.5 * N * np.log(2 * np.pi)
    return logl, (-0.5 * dlog - 0.5 * dchi), -0.5 * chimu

if __name__ == "__main__":
    ts = np.loadtxt(os.path.join(ROOT, "tests", "golden", "simu_single_subset_time_series.txt"))
    t, y_raw, yerr = (np.ascontiguousarray(c) for c in ts.T)
    t, y_raw, yerr = t[:150], y_raw[:150], yerr[:150]
    y, s2 = np.log(y_raw), yerr ** 2 / y_raw ** 2
    f_min, f_max = 1.0 / (t[-1] - t[0]), 1.0 / np.min(np.diff(t)) / 2.0
    rng = np.random.default_rng(1)
    for basis in ("SHO", "DRWCelerite"):
        a, b, c, d = orc.approx("SBPL", [0.7, 0.02, 3.0], f_min, f_max, 10, 0.05, basis=basis)
        rows = rows_from_coeffs(a, b, c, d)
        amp = np.array([r[0] for r in rows]); suma = np.sum(a)
        # a random amplitude direction consistent over the rows of a term
        da = a * rng.normal(0, 1, len(a))
        drows = rows_from_coeffs(da, b, c, d) if False else None
        damp = []
        for aj, bj, dj, dd in zip(da, b, d, d):
            damp += [aj] if (bj == 0 and dj == 0) else [aj, aj]
        damp = np.array(damp); dsuma = np.sum(da)
        mu, nu = 0.3, 1.4
        for (dam, dsa, dnu, name) in ((damp, dsuma, 0.0, "amp direction"), (np.zeros_like(amp), 0.0, 1.0, "nu")):
            l0, dl, dmu = sweep(amp, suma, rows, t, y, s2, mu, nu, dam, dsa, dnu)
            h = 1e-6
            lp, _, _ = sweep(amp + h * dam, suma + h * dsa, rows, t, y, s2, mu, nu + h * dnu)
            lm_, _, _ = sweep(amp - h * dam, suma - h * dsa, rows, t, y, s2, mu, nu - h * dnu)
            fd = (lp - lm_) / (2 * h)
            print(basis, name, "analytic", dl, "central diff", fd, "rel", abs(dl - fd) / max(1, abs(fd)))
        lp, _, _ = sweep(amp, suma, rows, t, y, s2, mu + 1e-6, nu); lm_, _, _ = sweep(amp, suma, rows, t, y, s2, mu - 1e-6, nu)
        print(basis, "mu analytic", dmu, "central diff", (lp - lm_) / 2e-6)
