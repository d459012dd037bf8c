"""numpy derivation of the Newton (multiple-shooting) refinement of the parallel-in-time path (K3, csrc/scan.cuh).

The chunk composites of the scan lose digits in their combine on ill-conditioned covariances (steep DRWCelerite slopes):
the states the scan hands to the re-filter sweeps are ~1e-8 off.  One Newton step on the boundary states fixes that with
the EXACT recursion as the residual:  r_{k+1} = F_k(S~_k) - S~_{k+1}  (F_k = the sequential sweep over chunk k), and the
correction obeys the linear recurrence  dS_{k+1} = T_k dS_k T_k^T + r_{k+1},  T_k = A_k (I + S~_k J_k)^-1  (the closed-loop
transition of chunk k, needed to first order only), dg_{k+1} = T_k dg_k + T_k dS_k m_k + rg_{k+1}, m_k = eta_k - J_k g^_k.

Run:  python tests/tools/proto/scan_newton_math.py   (CPU only; oracle for the coefficients; test infrastructure)
"""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import workloads as wl
from oracle import oracle as orc


def rows_of(a, b, c, d):
    """term -> rows (real term: one row, complex: two)."""
    rows = []
    for m in range(len(a)):
        rows.append(1 if (b[m] == 0.0 and d[m] == 0.0) else 2)
    return rows


def step_vectors(a, b, c, d, t, dtype=np.float64):
    a, b, c, d, t = (np.asarray(x, dtype=dtype) for x in (a, b, c, d, t))
    N = len(t)
    nrow = rows_of(a, b, c, d)
    R = sum(nrow)
    U = np.zeros((N, R), dtype); V = np.zeros((N, R), dtype); P = np.zeros((N, R), dtype)
    dt = np.append(np.diff(t), dtype(0))
    r = 0
    for m in range(len(a)):
        ph = np.exp(-c[m] * dt); ph[-1] = 0
        if nrow[m] == 1:
            U[:, r] = a[m]; V[:, r] = 1; P[:, r] = ph; r += 1
        else:
            co, si = np.cos(d[m] * t), np.sin(d[m] * t)
            U[:, r] = a[m] * co + b[m] * si; U[:, r + 1] = a[m] * si - b[m] * co
            V[:, r] = co; V[:, r + 1] = si; P[:, r] = ph; P[:, r + 1] = ph; r += 2
    return U, V, P


def sweep(U, V, P, A, y, n0, n1, S, g, bounds=()):
    """sequential recursion over [n0, n1) from state (S, g) -> (sum log D, chi2, S, g, states at `bounds`)."""
    S = S.copy(); g = g.copy()
    ld = S.dtype.type(0); chi = S.dtype.type(0)
    snap = {}
    for n in range(n0, n1):
        if n in bounds:
            snap[n] = (S.copy(), g.copy())
        u = U[n]
        Su = S @ u
        D = A[n] - u @ Su
        w = (V[n] - Su) / D
        z = y[n] - u @ g
        ld += np.log(D); chi += z * z / D
        ph = P[n]
        S = (S + D * np.outer(w, w)) * np.outer(ph, ph)
        g = ph * (g + w * z)
    return ld, chi, S, g, snap


def fold(U, V, P, A, y, n0, n1):
    R = U.shape[1]
    Am = np.eye(R); C = np.zeros((R, R)); Jm = np.zeros((R, R)); b = np.zeros(R); eta = np.zeros(R)
    for n in range(n0, n1):
        u = U[n]
        Cu = C @ u
        D = A[n] - u @ Cu
        w = (V[n] - Cu) / D
        z = y[n] - u @ b
        atu = Am.T @ u
        ph = P[n]
        Jm = Jm - np.outer(atu, atu) / D
        eta = eta - atu * z / D
        Am = ph[:, None] * (Am - np.outer(w, atu))
        C = (C + D * np.outer(w, w)) * np.outer(ph, ph)
        b = ph * (b + w * z)
    return Am, b, C, eta, Jm


def apply(el, S, g):
    Am, b, C, eta, Jm = el
    R = len(b)
    M = np.linalg.inv(np.eye(R) + S @ Jm)
    S2 = Am @ (M @ S) @ Am.T + C
    S2 = 0.5 * (S2 + S2.T)
    g2 = Am @ (M @ (g + S @ eta)) + b
    return S2, g2


def combine(ei, ej):
    Ai, bi, Ci, eti, Ji = ei
    Aj, bj, Cj, etj, Jj = ej
    R = len(bi)
    M = np.linalg.inv(np.eye(R) + Ci @ Jj)
    Ao = Aj @ M @ Ai
    bo = Aj @ (M @ (bi + Ci @ etj)) + bj
    Co = Aj @ (M @ Ci) @ Aj.T + Cj; Co = 0.5 * (Co + Co.T)
    r = etj - Jj @ bi
    eto = Ai.T @ (r - Jj @ (M @ (Ci @ r))) + eti
    Jo = Ai.T @ (Jj @ M @ Ai) + Ji; Jo = 0.5 * (Jo + Jo.T)
    return Ao, bo, Co, eto, Jo


def relerr(X, Y):
    return float(np.linalg.norm(np.asarray(X, np.float64) - np.asarray(Y, np.float64)) / max(np.linalg.norm(np.asarray(Y, np.float64)), 1e-300))


def run(basis, J, N, P_chunks, nworst=3, seed=5, alpha2_max=6.0, verbose=True, picks=None, ks=False):
    t, y, s2, f_min, f_max = wl.make_series_fast(N, seed=seed)
    th = wl.prior_theta(256, f_min, f_max, y.mean(), y.std(), 9, alpha2_max)
    # pick the rows where the FP64 sequential oracle is farthest from the 80-bit one (ill-conditioned), but finite
    devs = []
    coefs = []
    for i in range(len(th)):
        if picks is not None and i not in picks:
            coefs.append(None); devs.append(-1.0); continue
        a, b, c, d = orc.approx("SingleBendingPowerLaw", th[i, :3], f_min, f_max, J, th[i, 3], basis=basis)
        coefs.append((a, b, c, d))
        l64 = orc.celerite_logl(a, b, c, d, t, y - th[i, 5], th[i, 4] * s2)
        l80 = orc.celerite_logl(a, b, c, d, t, y - th[i, 5], th[i, 4] * s2, long_double=True)
        devs.append(abs(l64 - float(l80)) / max(1.0, abs(float(l80))) if np.isfinite(l64) else -1.0)
    devs = np.array(devs)
    order = np.argsort(-devs)
    picks = [i for i in order if devs[i] < 1e-7][:nworst] if picks is None else list(picks)
    bounds = [int(round(k * N / P_chunks)) for k in range(P_chunks + 1)]
    out = []
    for i in picks:
        a, b, c, d = coefs[i]
        mu, nu = th[i, 5], th[i, 4]
        U, V, P = step_vectors(a, b, c, d, t)
        A = a.sum() + nu * s2
        yy = y - mu
        R = U.shape[1]
        # truth (80-bit) and FP64 sequential, with boundary states
        Ul, Vl, Pl = step_vectors(a, b, c, d, t, np.longdouble)
        Al = np.longdouble(a).sum() + np.longdouble(nu) * np.longdouble(s2)
        yl = np.longdouble(y) - np.longdouble(mu)
        ldT, chT, _, _, snapT = sweep(Ul, Vl, Pl, Al, yl, 0, N, np.zeros((R, R), np.longdouble), np.zeros(R, np.longdouble), set(bounds))
        ld6, ch6, _, _, snap6 = sweep(U, V, P, A, yy, 0, N, np.zeros((R, R)), np.zeros(R), set(bounds))
        LT = float(-(ldT + chT) / 2); L6 = float(-(ld6 + ch6) / 2)
        scale = max(1.0, abs(LT - N * 0.9189385332))
        # scan: fold, sequential prefix by combine (order differs from Kogge-Stone; same conditioning), states by apply
        els = [fold(U, V, P, A, yy, bounds[k], bounds[k + 1]) for k in range(P_chunks)]
        if ks:      # Kogge-Stone order (what the GPU path does): long ranges are combined with long ranges
            pref = list(els); dd = 1
            while dd < P_chunks:
                pref = [pref[q] if q < dd else combine(pref[q - dd], pref[q]) for q in range(P_chunks)]
                dd *= 2
        else:
            pref = [els[0]]
            for k in range(1, P_chunks):
                pref.append(combine(pref[-1], els[k]))
        st = [(np.zeros((R, R)), np.zeros(R))]
        for k in range(1, P_chunks):
            st.append(apply(pref[k - 1], np.zeros((R, R)), np.zeros(R)))

        def chunk_sweeps(states):
            ld = ch = 0.0; exits = []
            for k in range(P_chunks):
                l, c2, S2, g2, _ = sweep(U, V, P, A, yy, bounds[k], bounds[k + 1], states[k][0], states[k][1])
                ld += l; ch += c2; exits.append((S2, g2))
            return -(ld + ch) / 2, exits

        L_scan, exits = chunk_sweeps(st)
        st_err = max(relerr(st[k][0], snapT[bounds[k]][0]) for k in range(1, P_chunks))
        sq_err = max(relerr(snap6[bounds[k]][0], snapT[bounds[k]][0]) for k in range(1, P_chunks))
        # Picard (run-up of one chunk): states = exits of the previous chunk
        st_p = [st[0]] + exits[:-1]
        L_pic, _ = chunk_sweeps(st_p)
        # Newton: residuals + linear recurrence with T_k, m_k from the composites (first order)
        dS = np.zeros((R, R)); dg = np.zeros(R)
        st_n = [st[0]]
        for k in range(P_chunks - 1):
            Am, bb, C, eta, Jm = els[k]
            Sk, gk = st[k]
            M = np.linalg.inv(np.eye(R) + Sk @ Jm)
            T = Am @ M
            ghat = M @ (gk + Sk @ eta)
            m = eta - Jm @ ghat
            rS = exits[k][0] - st[k + 1][0]
            rg = exits[k][1] - st[k + 1][1]
            dg = T @ dg + T @ (dS @ m) + rg
            dS = T @ dS @ T.T + rS
            st_n.append((st[k + 1][0] + dS, st[k + 1][1] + dg))
        L_new, exits_n = chunk_sweeps(st_n)
        stn_err = max(relerr(st_n[k][0], snapT[bounds[k]][0]) for k in range(1, P_chunks))
        res_after = max(relerr(exits_n[k][0], st_n[k + 1][0]) for k in range(P_chunks - 1))
        res_before = max(relerr(exits[k][0], st[k + 1][0]) for k in range(P_chunks - 1))
        # conditioning proxy of the sequential sweep: eps * sum_n A_n / D_n
        row = dict(i=int(i), alpha2=float(th[i, 2]), seq_vs_80=abs(L6 - LT) / scale, scan_vs_seq=abs(L_scan - L6) / scale,
                   picard_vs_seq=abs(L_pic - L6) / scale, newton_vs_seq=abs(L_new - L6) / scale,
                   newton_vs_80=abs(L_new - LT) / scale, state_err_scan=st_err, state_err_seq=sq_err, state_err_newton=stn_err,
                   res_before=res_before, res_after=res_after)
        out.append(row)
        if verbose:
            print(basis, J, N, P_chunks, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in row.items()}, flush=True)
    return out


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
    for basis, J in (("DRWCelerite", 5), ("DRWCelerite", 10), ("DRWCelerite", 20)):
        run(basis, J, N, 12)
