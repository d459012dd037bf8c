"""Math-level prototype of the 8-step blocked (tensor-core) celerite sweep, checked against the CPU oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc

BK = 8

def rows_from_coeffs(a, b, c, d):
    """rows: (amp, c, d, ratio, kind) ; kind 0 cos, 1 sin, 2 real"""
    rows = []
    for aj, bj, cj, dj in zip(a, b, c, d):
        if bj == 0 and dj == 0:
            rows.append((aj, cj, 0.0, 0.0, 2))
        else:
            rows.append((aj, cj, dj, bj / aj, 0))
            rows.append((aj, cj, dj, bj / aj, 1))
    return rows

def blocked_logl(a, b, c, d, t, y, s2, mu=0.0, nu=1.0):
    rows = rows_from_coeffs(a, b, c, d)
    R = len(rows)
    amp = np.array([r[0] for r in rows]); cr = np.array([r[1] for r in rows]); dr = np.array([r[2] for r in rows])
    ratio = np.array([r[3] for r in rows]); kind = np.array([r[4] for r in rows])
    N = len(t)
    suma = np.sum(a)
    # per step Ũ, V, φ
    arg = np.outer(t, dr)
    co, si = np.cos(arg), np.sin(arg)
    Ut = np.where(kind == 0, co + ratio * si, np.where(kind == 1, si - ratio * co, 1.0))
    V = np.where(kind == 0, co, np.where(kind == 1, si, 1.0))
    phi = np.zeros((N, R))
    phi[1:] = np.exp(-np.outer(np.diff(t), cr))
    X = np.zeros((R, R)); gp = np.zeros(R)
    logdet = 0.0; chi2 = 0.0
    for n1 in range(0, N, BK):
        idx = np.arange(n1, min(n1 + BK, N)); nb = len(idx)
        # cumulative decays
        Psi0 = np.cumprod(phi[idx], axis=0)                  # Psi_{0->s}
        Uh = (Psi0 * Ut[idx]).T                              # R x nb
        # Psi_{s->end}
        PsiE = np.ones((nb, R))
        for s in range(nb - 2, -1, -1):
            PsiE[s] = PsiE[s + 1] * phi[idx[s + 1]]
        psi8 = Psi0[-1]
        Vh = (PsiE * V[idx]).T                               # R x nb
        P0 = X @ Uh
        C2 = Uh.T @ P0
        Kb = np.zeros((nb, nb))
        for s in range(nb):
            Kb[s, s] = suma + nu * s2[idx[s]]
            dec = np.ones(R)
            for sp in range(s - 1, -1, -1):
                dec = dec * phi[idx[sp + 1]]
                Kb[s, sp] = Kb[sp, s] = np.sum(amp * Ut[idx[s]] * dec * V[idx[sp]])
        C = Kb - C2
        # LDL^T
        L = np.eye(nb); D = np.zeros(nb)
        for j in range(nb):
            D[j] = C[j, j] - np.sum(L[j, :j] ** 2 * D[:j])
            for i in range(j + 1, nb):
                L[i, j] = (C[i, j] - np.sum(L[i, :j] * L[j, :j] * D[:j])) / D[j]
        Linv = np.linalg.inv(L)
        r = y[idx] - mu - Uh.T @ gp
        z = Linv @ r
        chi2 += np.sum(z * z / D); logdet += np.sum(np.log(D))
        Bm = amp[:, None] * Vh - psi8[:, None] * P0
        Qh = Bm @ Linv.T
        Wh = Qh / D[None, :]
        X = np.outer(psi8, psi8) * X + Qh @ Wh.T
        gp = psi8 * gp + Wh @ z
    return -0.5 * logdet - 0.5 * chi2 - 0.5 * N * np.log(2 * np.pi)

if __name__ == "__main__":
    ts = np.loadtxt(os.path.join(ROOT, "tests", "golden", "simu_single_subset_time_series.txt"))
    chain = np.load(os.path.join(ROOT, "tests", "golden", "chains.npz"))["simu_single"]
    t, y_raw, yerr = (np.ascontiguousarray(c) for c in ts.T)
    y, s2 = np.log(y_raw), yerr ** 2 / y_raw ** 2
    f_min, f_max = 1.0 / (t[-1] - t[0]), 1.0 / np.min(np.diff(t)) / 2.0
    rng = np.random.default_rng(0)
    for basis in ("SHO", "DRWCelerite"):
        worst = 0
        for k in rng.choice(len(chain), 12, replace=False):
            th = chain[k, 2:8].copy()
            if basis == "DRWCelerite": th[2] += 1.0
            a, b, c, d = orc.approx("SBPL", th[:3], f_min, f_max, 20, th[3], basis=basis)
            ref = orc.celerite_logl(a, b, c, d, t, y - th[5], th[4] * s2)
            ld = float(orc.celerite_logl(a, b, c, d, t, y - th[5], th[4] * s2, long_double=True))
            got = blocked_logl(a, b, c, d, t, y, s2, mu=th[5], nu=th[4])
            e = abs(got - ref) / max(1, abs(ref)); e2 = abs(ref - ld) / max(1, abs(ld)); e3 = abs(got - ld) / max(1, abs(ld))
            worst = max(worst, e)
            print(basis, k, ref, f"blocked-vs-oracle {e:.2e}  oracle-vs-ld {e2:.2e} blocked-vs-ld {e3:.2e}")
        print(basis, "worst", worst)
