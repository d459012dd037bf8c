"""Lane-level emulator of the blocked DMMA celerite kernel (csrc/blocked.cuh): every 'register' is an array over the 32
lanes, mma/shfl are emulated with the PTX fragment layouts of mma.sync.m8n8k4.f64.  Validates the index algebra."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from blocked_math import rows_from_coeffs, blocked_logl

LANE = np.arange(32); G_ = LANE >> 2; T_ = LANE & 3

def mma(c0, c1, a, b):
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    A[G_, T_] = a; B[T_, G_] = b
    C[G_, 2 * T_] = c0; C[G_, 2 * T_ + 1] = c1
    D = A @ B + C
    return D[G_, 2 * T_], D[G_, 2 * T_ + 1]

def shfl(v, src):
    return v[src]

def shfl_xor(v, m):
    return v[LANE ^ m]

def transpose_tile(e0, e1):
    odd = (G_ & 1) == 1
    p1 = np.where(odd, e1, e0)
    src1 = 4 * (2 * T_ + (G_ & 1)) + (G_ >> 1)
    r1 = shfl(p1, src1)
    p2 = np.where(odd, e0, e1)
    src2 = 4 * (2 * T_ + 1 - (G_ & 1)) + (G_ >> 1)
    r2 = shfl(p2, src2)
    return np.where(odd, r2, r1), np.where(odd, r1, r2)

def pair_slot(s, sp):  # s > sp
    return s * (s - 1) // 2 + sp

def build_block_table(NT, Ut, V, phi, y, s2, idx, N):
    """θ-independent record of one block of 8 steps (mirrors blocked_table_kernel)."""
    RP = 8 * NT
    nb = len(idx)
    R = Ut.shape[1]
    def row(x):  # pad row vectors to RP
        out = np.zeros(RP); out[:R] = x; return out
    ph = [row(phi[n]) if n < N else np.ones(RP) for n in range(idx[0], idx[0] + 8)]
    ut = [row(Ut[n]) if n < N else np.zeros(RP) for n in range(idx[0], idx[0] + 8)]
    vv = [row(V[n]) if n < N else np.zeros(RP) for n in range(idx[0], idx[0] + 8)]
    Psi0 = np.cumprod(np.array(ph), axis=0)
    PsiE = np.ones((8, RP))
    for s in range(6, -1, -1): PsiE[s] = PsiE[s + 1] * ph[s + 1]
    tab = {}
    Uh = Psi0 * np.array(ut)                        # [s][row]
    tab["UT"] = np.zeros((NT, 8, 8))                # [K][s][r8]
    for K in range(NT): tab["UT"][K] = Uh[:, 8 * K:8 * K + 8]
    tab["VH"] = (PsiE * np.array(vv)).T.copy()      # [row][s]
    tab["PSI8"] = Psi0[7].copy()
    H = np.zeros((32, RP))
    for s in range(8):
        dec = np.ones(RP)
        for sp in range(s - 1, -1, -1):
            dec = dec * ph[sp + 1]
            H[pair_slot(s, sp)] = ut[s] * dec * vv[sp]
    tab["H"] = H
    yy = np.zeros(8); ss = np.zeros(8); mk = np.zeros(8)
    for s in range(8):
        n = idx[0] + s
        if n < N: yy[s] = y[n]; ss[s] = s2[n]; mk[s] = 1.0
    tab["Y"], tab["S2"], tab["MASK"] = yy, ss, mk
    return tab

def fast_rcp(x): return 1.0 / x

def lanes_logl(NT, amp_rows, suma, mu, nu, tables, N):
    RP = 8 * NT
    amp = np.zeros(RP); amp[:len(amp_rows)] = amp_rows
    # persistent lane state
    x = {}
    for I in range(NT):
        for K in range(I + 1):
            x[(I, K)] = [np.zeros(32), np.zeros(32)]
    gpa = np.zeros(32); gpb = np.zeros(32)             # gp rows 8t+g, 8(t+4)+g
    chi2 = np.zeros(32); logacc = np.zeros(32); dkeep = np.ones(32); dfirst = np.ones(32)
    # amplitudes of the K_blk rows (2l, 2l+1)
    ampk0 = np.array([amp[2 * l] if 2 * l < RP else 0.0 for l in LANE])
    ampk1 = np.array([amp[2 * l + 1] if 2 * l + 1 < RP else 0.0 for l in LANE])
    for b, tab in enumerate(tables):
        UT, VH, PSI8, H = tab["UT"], tab["VH"], tab["PSI8"], tab["H"]
        # ---- phase A: K_blk partials (rows 2l, 2l+1) and reduce-scatter into slots
        v = []
        for p in range(32):
            if p < 28:
                h0 = np.array([H[p][2 * l] if 2 * l < RP else 0.0 for l in LANE])
                h1 = np.array([H[p][2 * l + 1] if 2 * l + 1 < RP else 0.0 for l in LANE])
                v.append(ampk0 * h0 + ampk1 * h1)
            else:
                v.append(np.zeros(32))
        n = 32; mask = 16
        while n > 1:
            half = n // 2
            bit = (LANE & mask) != 0
            nv = []
            for i in range(half):
                send = np.where(bit, v[i], v[i + half])
                keep = np.where(bit, v[i + half], v[i])
                nv.append(keep + shfl_xor(send, mask))
            v = nv; n = half; mask >>= 1
        slot = v[0]                                       # lane l holds the total of slot l
        # C0 in D-layout: lane(g,t): C0[g][2t+e]
        c0 = []
        for e in range(2):
            col = 2 * T_ + e
            hi = np.maximum(G_, col); lo = np.minimum(G_, col)
            src = np.where(hi > lo, hi * (hi - 1) // 2 + lo, 0)
            val = shfl(slot, src)
            diag = (suma + nu * tab["S2"][G_]) * tab["MASK"][G_] + (1.0 - tab["MASK"][G_])
            c0.append(np.where(G_ == col, diag, val))
        # ---- phase B: P0[I] = sum_K xfull[I][K] * Uf[K]
        P0 = [[np.zeros(32), np.zeros(32)] for _ in range(NT)]
        for K in range(NT):
            Uf = [UT[K][G_, 2 * T_ + c] for c in range(2)]      # Û[8K+2t+c][s=g]  (tile-major [K][s][r8])
            for I in range(NT):
                if I >= K: xa = x[(I, K)]
                else:      xa = transpose_tile(*x[(K, I)])
                for c in range(2):
                    P0[I][0], P0[I][1] = mma(P0[I][0], P0[I][1], xa[c], Uf[c])
        # ---- phase C: C2 = Ûᵀ P0
        C2 = [np.zeros(32), np.zeros(32)]
        for J in range(NT):
            Uf = [UT[J][G_, 2 * T_ + c] for c in range(2)]
            PT = transpose_tile(*P0[J])
            for c in range(2):
                C2[0], C2[1] = mma(C2[0], C2[1], Uf[c], PT[c])
        Cm = [c0[0] - C2[0], c0[1] - C2[1]]
        # residual r_g = (y_g − μ)·mask − m_g ;  m_s = Σ_rows Û[row][s] gp[row]
        mpart = []
        for s in range(8):
            ua = np.array([UT[t][s][g] if t < NT else 0.0 for g, t in zip(G_, T_)])
            ub = np.array([UT[t + 4][s][g] if t + 4 < NT else 0.0 for g, t in zip(G_, T_)])
            mpart.append(ua * gpa + ub * gpb)
        n = 8; mask = 16
        while n > 1:
            half = n // 2
            bit = (LANE & mask) != 0
            nv = []
            for i in range(half):
                send = np.where(bit, mpart[i], mpart[i + half])
                keep = np.where(bit, mpart[i + half], mpart[i])
                nv.append(keep + shfl_xor(send, mask))
            mpart = nv; n = half; mask >>= 1
        m = mpart[0]
        m = m + shfl_xor(m, 2); m = m + shfl_xor(m, 1)          # lanes (g,*) hold m_g
        r = (tab["Y"][G_] - mu) * tab["MASK"][G_] - m
        # ---- Bm = amp∘V̂ − ψ8∘P0 (in place of P0)
        Bm = []
        for I in range(NT):
            rowi = 8 * I + G_
            Bm.append([amp[rowi] * VH[rowi, 2 * T_ + e] - PSI8[rowi] * P0[I][e] for e in range(2)])
        # ---- phase D: distributed LDLᵀ with the inverse factor riding along
        E = [(G_ == 2 * T_).astype(float), (G_ == 2 * T_ + 1).astype(float)]
        rd = []
        n0 = 8 * b
        for j in range(8):
            tj, ej = j >> 1, j & 1
            dj = shfl(Cm[ej], np.full(32, 4 * j + tj))
            rdj = fast_rcp(dj); rd.append(rdj)
            cgj = shfl(Cm[ej], (LANE & ~3) | tj)
            l = cgj * rdj
            cj = [shfl(Cm[e], 4 * j + T_) for e in range(2)]
            ejr = [shfl(E[e], 4 * j + T_) for e in range(2)]
            below = G_ > j
            for e in range(2):
                Cm[e] = np.where(below, Cm[e] - l * cj[e], Cm[e])
                E[e] = np.where(below, E[e] - l * ejr[e], E[e])
            nstep = n0 + j
            if nstep == 0: dfirst = dj.copy()
            else: dkeep = np.where(LANE == (nstep & 31), dj, dkeep)
        if (b & 3) == 3:
            logacc += np.log(np.abs(dkeep)); dkeep = np.ones(32)
        # z = L⁻¹ r
        r_e = [shfl(r, 4 * (2 * T_ + e)) for e in range(2)]
        z = E[0] * r_e[0] + E[1] * r_e[1]
        z = z + shfl_xor(z, 1); z = z + shfl_xor(z, 2)           # lanes (g,*) hold z_g
        rdg = np.choose(G_, rd)                                    # 1/d_g
        chi2 += np.where(T_ == 0, z * z * rdg, 0.0)
        z_e = [shfl(z, 4 * (2 * T_ + e)) for e in range(2)]
        rd_e = [np.choose(2 * T_ + e, rd) for e in range(2)]
        # ---- phase E: Q̂ = Bm·L⁻ᵀ ; Ŵ = Q̂ D⁻¹
        Q = []
        for I in range(NT):
            q0, q1 = np.zeros(32), np.zeros(32)
            for c in range(2):
                q0, q1 = mma(q0, q1, Bm[I][c], E[c])
            Q.append([q0, q1])
        # ---- phase F: decay + rank-8 update, g update
        gpart = []
        for K in range(NT):
            W = [Q[K][e] * rd_e[e] for e in range(2)]
            gpart.append(W[0] * z_e[0] + W[1] * z_e[1])
            for I in range(K, NT):
                rowf = PSI8[8 * I + G_]
                for e in range(2):
                    x[(I, K)][e] = x[(I, K)][e] * rowf * PSI8[8 * K + 2 * T_ + e]
                for c in range(2):
                    x[(I, K)][0], x[(I, K)][1] = mma(x[(I, K)][0], x[(I, K)][1], Q[I][c], W[c])
        while len(gpart) < 8: gpart.append(np.zeros(32))
        # reduce-scatter over t: lane(g,t) ends with I = t and I = t+4
        bit = (T_ & 2) != 0
        st1 = []
        for base in (0, 1, 4, 5):
            send = np.where(bit, gpart[base], gpart[base + 2]); keep = np.where(bit, gpart[base + 2], gpart[base])
            st1.append(keep + shfl_xor(send, 2))                 # holds I = base (+2 if bit)
        bit0 = (T_ & 1) != 0
        out = []
        for k in (0, 2):                                         # pairs (0|2 , 1|3) and (4|6, 5|7)
            send = np.where(bit0, st1[k], st1[k + 1]); keep = np.where(bit0, st1[k + 1], st1[k])
            out.append(keep + shfl_xor(send, 1))
        ra = 8 * T_ + G_; rb = 8 * (T_ + 4) + G_
        psa = np.where(ra < RP, PSI8[np.minimum(ra, RP - 1)], 0.0)
        psb = np.where(rb < RP, PSI8[np.minimum(rb, RP - 1)], 0.0)
        gpa = psa * gpa + out[0]
        gpb = psb * gpb + out[1]
    la = logacc + np.log(np.abs(dkeep))
    logdet = np.log(dfirst[0]) + la.sum()
    return -0.5 * logdet - 0.5 * chi2.sum() - 0.5 * N * np.log(2 * np.pi)

def run(a, b, c, d, t, y, s2, mu, nu):
    rows = rows_from_coeffs(a, b, c, d)
    R = len(rows); NT = (R + 7) // 8
    amp = np.array([r[0] for r in rows]); cr = np.array([r[1] for r in rows]); dr = np.array([r[2] for r in rows])
    ratio = np.array([r[3] for r in rows]); kind = np.array([r[4] for r in rows])
    N = len(t)
    arg = np.outer(t, dr); co, si = np.cos(arg), np.sin(arg)
    Ut = np.where(kind == 0, co + ratio * si, np.where(kind == 1, si - ratio * co, 1.0))
    V = np.where(kind == 0, co, np.where(kind == 1, si, 1.0))
    phi = np.zeros((N, R)); phi[1:] = np.exp(-np.outer(np.diff(t), cr))
    tables = [build_block_table(NT, Ut, V, phi, y, s2, np.arange(n1, min(n1 + 8, N)), N) for n1 in range(0, N, 8)]
    return lanes_logl(NT, amp, np.sum(a), mu, nu, tables, N)

if __name__ == "__main__":
    ts = np.loadtxt(os.path.join(ROOT, "tests", "golden", "simu_single_subset_time_series.txt"))
    chain = np.load(os.path.join(ROOT, "tests", "golden", "chains.npz"))["simu_single"]
    t, y_raw, yerr = (np.ascontiguousarray(c) for c in ts.T)
    y, s2 = np.log(y_raw), yerr ** 2 / y_raw ** 2
    Nuse = 203
    t, y, s2 = t[:Nuse], y[:Nuse], s2[:Nuse]
    f_min, f_max = 1.0 / (t[-1] - t[0]), 1.0 / np.min(np.diff(t)) / 2.0
    for basis, J in (("SHO", 20), ("DRWCelerite", 20), ("SHO", 7), ("DRWCelerite", 5)):
        for k in (5, 100):
            th = chain[k, 2:8].copy()
            if basis == "DRWCelerite": th[2] += 1.0
            a, b, c, d = orc.approx("SBPL", th[:3], f_min, f_max, J, th[3], basis=basis)
            ref = orc.celerite_logl(a, b, c, d, t, y - th[5], th[4] * s2)
            got = run(a, b, c, d, t, y, s2, th[5], th[4])
            print(basis, J, k, ref, got, abs(got - ref) / max(1, abs(ref)))
