"""numpy check of the BLOCKED fold of the parallel-in-time path: the chunk composite (A, b, C, eta, J) of csrc/scan.cuh built 8
steps at a time with the matrix products of the blocked sweep (csrc/blocked.cuh), in the state convention of the scan (S enters a
step already decayed to its time: per-step decay phi_{n+1} AFTER the update):
    U^[:,s] = Psi'_{0->s} o U_s,  V^[:,s] = Psi'_{s->8} o V_s,  psi8' = Psi'_{0->8},   Psi'_{a->b} = prod_{a<i<=b} phi_{n0+i}
    P0 = X U^,  C8 = K_blk - U^T P0 = L D L^T,  E = L^-1,  Q^ = (V^ - psi8' o P0) E^T,  W^ = Q^ D^-1,  X <- psi8' psi8'^T o X + Q^ W^T
    Pa = A^T U^,  Ga = Pa E^T,   A^T <- A^T o (1 psi8'^T) - Ga W^T,   J <- J - Ga D^-1 Ga^T,   eta <- eta - Ga D^-1 z
with the data row carried as one more row of X (b) and z = Q^[data row].  Compared with the rank-1 fold of scan_newton_math.py."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools"); sys.path.insert(0, "tests/tools/proto")
import workloads as wl
from oracle import oracle as orc
from scan_newton_math import step_vectors, fold

BK = 8


def blocked_fold(U, V, P, A_n, y, n0, n1):
    """U, V: [N, R] rows; P[n] = phi_{n+1} (decay after step n); A_n = sum a + nu s2_n."""
    R = U.shape[1]
    X = np.zeros((R, R)); b = np.zeros(R); At = np.eye(R); J = np.zeros((R, R)); eta = np.zeros(R)
    for m0 in range(n0, n1, BK):
        idx = np.arange(m0, min(m0 + BK, n1)); nb = len(idx)
        ph = P[idx]                                        # ph[s] = decay AFTER step s of the block
        Psi0 = np.ones((nb, R))                            # Psi'_{0->s}
        for s in range(1, nb): Psi0[s] = Psi0[s - 1] * ph[s - 1]
        PsiE = np.ones((nb, R))                            # Psi'_{s->nb}
        PsiE[nb - 1] = ph[nb - 1]
        for s in range(nb - 2, -1, -1): PsiE[s] = PsiE[s + 1] * ph[s]
        psi8 = Psi0[nb - 1] * ph[nb - 1]
        Uh = (Psi0 * U[idx]).T; Vh = (PsiE * V[idx]).T
        Kb = np.zeros((nb, nb))
        for s in range(nb):
            Kb[s, s] = A_n[idx[s]]
            dec = np.ones(R)
            for sp in range(s - 1, -1, -1):
                dec = dec * ph[sp]
                Kb[s, sp] = Kb[sp, s] = np.sum(U[idx[s]] * dec * V[idx[sp]])
        P0 = X @ Uh
        C8 = Kb - Uh.T @ P0
        L = np.eye(nb); D = np.zeros(nb)
        for j in range(nb):
            D[j] = C8[j, j] - np.sum(L[j, :j] ** 2 * D[:j])
            for i in range(j + 1, nb):
                L[i, j] = (C8[i, j] - np.sum(L[i, :j] * L[j, :j] * D[:j])) / D[j]
        E = np.linalg.inv(L)
        Qh = (Vh - psi8[:, None] * P0) @ E.T
        Wh = Qh / D[None, :]
        z = E @ (y[idx] - Uh.T @ b)
        Pa = At @ Uh
        Ga = Pa @ E.T
        J = J - (Ga / D[None, :]) @ Ga.T
        eta = eta - (Ga / D[None, :]) @ z
        At = At * psi8[None, :] - Ga @ Wh.T
        X = np.outer(psi8, psi8) * X + Qh @ Wh.T
        b = psi8 * b + Wh @ z
    return At.T, b, X, eta, J


if __name__ == "__main__":
    N = 700
    t, y, s2, f_min, f_max = wl.make_series_fast(N, seed=5)
    th = wl.prior_theta(8, f_min, f_max, y.mean(), y.std(), 9, 6.0)
    for basis, Jn in (("DRWCelerite", 5), ("SHO", 6)):
        for i in range(4):
            a, b, c, d = orc.approx("SingleBendingPowerLaw", th[i, :3], f_min, f_max, Jn, th[i, 3], basis=basis)
            U, V, P = step_vectors(a, b, c, d, t)
            A_n = a.sum() + th[i, 4] * s2
            yy = y - th[i, 5]
            for (n0, n1) in ((0, 160), (160, 403), (403, N)):
                ref = fold(U, V, P, A_n, yy, n0, n1)
                got = blocked_fold(U, V, P, A_n, yy, n0, n1)
                errs = [np.abs(g - r).max() / max(np.abs(r).max(), 1e-300) for g, r in zip(got, ref)]
                print(basis, Jn, i, (n0, n1), "rel err (A, b, C, eta, J):", " ".join(f"{e:.1e}" for e in errs))
