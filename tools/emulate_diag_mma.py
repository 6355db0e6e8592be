"""Lane-level emulation (numpy) of diag64_mma_kernel's index arithmetic: DMMA.8x8x4 fragment layouts, tile dealing,
the parked G^T tiles and the row-wise inverse.  Checks L, W = L^-1 and W^T against numpy.  No GPU needed; run after any
change to the fragment / tile indexing of csrc/diag.cu:diag64_mma_kernel.  (The arithmetic order is not emulated.)"""
import numpy as np

NB, PB, MLD = 64, 8, 68
TRAIL = [0x11, 0x21, 0x31, 0x41, 0x51, 0x61, 0x71, 0x22, 0x32, 0x42, 0x52, 0x62, 0x72, 0x33, 0x43, 0x53, 0x63, 0x73, 0x44, 0x54,
         0x64, 0x74, 0x55, 0x65, 0x75, 0x66, 0x76, 0x77]
LANES = np.arange(32)
R, Q = LANES >> 2, LANES & 3


def dmma(c0, c1, a, b):
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    A[R, Q] = a
    B[Q, R] = b
    C[R, 2 * Q] = c0
    C[R, 2 * Q + 1] = c1
    D = A @ B + C
    return D[R, 2 * Q].copy(), D[R, 2 * Q + 1].copy()


def shfl(v, src):
    return v[src]


def cfrag_to_neg_afrag(c0, c1):
    src = (LANES & ~3) | (Q >> 1)
    x0, y0 = shfl(c0, src), shfl(c1, src)
    x1, y1 = shfl(c0, src + 2), shfl(c1, src + 2)
    a0 = -np.where(Q & 1, y0, x0)
    a1 = -np.where(Q & 1, y1, x1)
    return a0, a1


def cfrag_to_afrag(c0, c1):
    a0, a1 = cfrag_to_neg_afrag(c0, c1)
    return -a0, -a1


def factor(M, V, rinv, pp):
    T = PB * pp * MLD + PB * pp
    l = np.zeros((8, 8))
    for i in range(8):
        for j in range(0, i + 1, 2):
            l[i, j] = M[T + i * MLD + j]
            if j + 1 <= i:
                l[i, j + 1] = M[T + i * MLD + j + 1]
    rr = np.zeros(8)
    for j in range(8):
        d = l[j, j]
        rj = 1.0 / np.sqrt(d)
        rr[j] = rj
        l[j, j] = d * rj
        for i in range(j + 1, 8):
            l[i, j] *= rj
        for i in range(j + 1, 8):
            for c in range(j + 1, i + 1):
                l[i, c] -= l[i, j] * l[c, j]
    for lane in range(32):
        vc = lane & 7
        x = np.zeros(8)
        for i in range(8):
            s = 1.0 if i == vc else 0.0
            for k in range(i):
                s -= l[i, k] * x[k]
            x[i] = s * rr[i]
        for i in range(8):
            V[pp * 64 + i * 8 + vc] = x[i]
    for i in range(8):
        for j in range(i + 1):
            M[T + i * MLD + j] = l[i, j]
        rinv[PB * pp + i] = rr[i]


def run(A):
    M = np.full(NB * MLD, np.nan)
    V = np.full(8 * 64, np.nan)
    rinv = np.full(NB, np.nan)
    W = np.full(NB * NB, np.nan)
    WT = np.full(NB * NB, np.nan)
    Aout = np.full((NB, NB), np.nan)
    for idx in range(NB * NB // 2):
        row, c2 = idx >> 5, idx & 31
        M[row * MLD + 2 * c2: row * MLD + 2 * c2 + 2] = A[row, 2 * c2: 2 * c2 + 2] if 2 * c2 <= (row | 7) else 0.0
    z = np.zeros(32)
    for p in range(-1, 7):
        if p >= 0:
            Vp = p * 64
            b0, b1 = V[Vp + R * 8 + Q], V[Vp + R * 8 + 4 + Q]
            g0, g1 = V[Vp + Q * 8 + R], V[Vp + (4 + Q) * 8 + R]
            seen = []
            for warp in range(4):
                i0 = p + 1 + ((warp - (p + 1)) & 3)
                for i in (i0, i0 + 4):
                    if i >= 8:
                        continue
                    seen.append(i)
                    assert (i & 3) == warp
                    Ti = (PB * i + R) * MLD + PB * p
                    a0, a1 = M[Ti + Q].copy(), M[Ti + 4 + Q].copy()
                    c0, c1 = dmma(z, z, a0, b0)
                    c0, c1 = dmma(c0, c1, a1, b1)
                    M[Ti + 2 * Q] = c0
                    M[Ti + 2 * Q + 1] = c1
            assert sorted(seen) == list(range(p + 1, 8))
        fw = (p + 1) & 3
        Mnew = M.copy()
        if p >= 0:
            # look-ahead tile by the factor warp
            T = PB * (p + 1) * MLD + PB * (p + 1)
            Lr = (PB * (p + 1) + R) * MLD + PB * p
            C = T + R * MLD + 2 * Q
            c0, c1 = M[C].copy(), M[C + 1].copy()
            a0, a1 = M[Lr + Q], M[Lr + 4 + Q]
            c0, c1 = dmma(c0, c1, -a0, a0)
            c0, c1 = dmma(c0, c1, -a1, a1)
            Mnew[C] = c0; Mnew[C + 1] = c1
            m = 7 - p
            ntile = m * (m + 1) // 2
            done = {0}
            TB = 5
            for warp in range(4):
                if warp == fw:
                    continue
                w3 = (warp - fw - 1) & 3
                assert 0 <= w3 < 3
                t0 = 1 + w3
                while t0 < ntile:
                    for u in range(TB):
                        t = t0 + 3 * u
                        if t >= ntile:
                            continue
                        assert t not in done
                        done.add(t)
                        e = TRAIL[8 * p - p * (p + 1) // 2 + t]
                        i, j = e >> 4, e & 15
                        assert p < j <= i <= 7 and not (i == p + 1 and j == p + 1)
                        C = (PB * i + R) * MLD + PB * j + 2 * Q
                        Ao = (PB * i + R) * MLD + PB * p
                        Bo = (PB * j + R) * MLD + PB * p
                        c0, c1 = M[C].copy(), M[C + 1].copy()
                        c0, c1 = dmma(c0, c1, -M[Ao + Q], M[Bo + Q])
                        c0, c1 = dmma(c0, c1, -M[Ao + 4 + Q], M[Bo + 4 + Q])
                        Mnew[C] = c0; Mnew[C + 1] = c1
                    t0 += 3 * TB
            assert len(done) == ntile, (p, sorted(done), ntile)
            gseen = []
            for w3 in range(3):
                for u in range(3):
                    ig = p + 1 + w3 + 3 * u
                    if ig >= 8:
                        continue
                    gseen.append(ig)
                    Lr = (PB * ig + R) * MLD + PB * p
                    e0, e1 = dmma(z, z, M[Lr + Q], g0)
                    e0, e1 = dmma(e0, e1, M[Lr + 4 + Q], g1)
                    Mnew[(PB * p + 2 * Q) * MLD + PB * ig + R] = e0
                    Mnew[(PB * p + 2 * Q + 1) * MLD + PB * ig + R] = e1
            assert sorted(gseen) == list(range(p + 1, 8))
        M = Mnew
        factor(M, V, rinv, p + 1)
    def store_l_column(pc):
        v0, v1 = M[(PB * pc + R) * MLD + PB * pc + 2 * Q], M[(PB * pc + R) * MLD + PB * pc + 2 * Q + 1]
        for lane in range(32):
            r, q = R[lane], Q[lane]
            if 2 * q + 1 <= r:
                Aout[PB * pc + r, PB * pc + 2 * q] = v0[lane]; Aout[PB * pc + r, PB * pc + 2 * q + 1] = v1[lane]
            elif 2 * q == r:
                Aout[PB * pc + r, PB * pc + 2 * q] = v0[lane]
        for i in range(pc + 1, 8):
            Aout[PB * i + R, PB * pc + 2 * Q] = M[(PB * i + R) * MLD + PB * pc + 2 * Q]
            Aout[PB * i + R, PB * pc + 2 * Q + 1] = M[(PB * i + R) * MLD + PB * pc + 2 * Q + 1]

    def store_w(i, j, c0, c1):
        W[(PB * i + R) * NB + PB * j + 2 * Q] = c0
        W[(PB * i + R) * NB + PB * j + 2 * Q + 1] = c1
        WT[(PB * j + 2 * Q) * NB + PB * i + R] = c0
        WT[(PB * j + 2 * Q + 1) * NB + PB * i + R] = c1

    def inverse_row(I):
        n = {}
        store_w(I, I, V[I * 64 + R * 8 + 2 * Q], V[I * 64 + R * 8 + 2 * Q + 1])
        n[I] = (-V[I * 64 + R * 8 + Q], -V[I * 64 + R * 8 + 4 + Q])
        for t in range(1, I + 1):
            j = I - t
            so0, so1, sp0, sp1 = z, z, z, z
            for k in range(I, j + 1, -1):
                gt = (PB * j + R) * MLD + PB * k
                so0, so1 = dmma(so0, so1, n[k][0], M[gt + Q])
                sp0, sp1 = dmma(sp0, sp1, n[k][1], M[gt + 4 + Q])
            gt = (PB * j + R) * MLD + PB * (j + 1)
            s0, s1 = dmma(z, z, n[j + 1][0], M[gt + Q])
            s0, s1 = dmma(s0, s1, n[j + 1][1], M[gt + 4 + Q])
            s0, s1 = s0 + so0 + sp0, s1 + so1 + sp1
            if j > 0:
                n[j] = cfrag_to_neg_afrag(s0, s1)
            store_w(I, j, s0, s1)

    # (the kernel issues column / row p inside step p; the data they read is final by then -- here simply afterwards)
    for pc in range(8):
        store_l_column(pc)
        inverse_row(pc)
        for j in range(pc + 1, 8):
            store_w(pc, j, z, z)
    # statistics as the kernel forms them
    lg = sum(-np.log((rinv[4 * l] * rinv[4 * l + 1]) * (rinv[4 * l + 2] * rinv[4 * l + 3])) for l in range(16))
    return Aout, W.reshape(NB, NB), WT.reshape(NB, NB), rinv, 2.0 * lg


def main():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((NB, 3 * NB))
    A = X @ X.T / (3 * NB) + 0.1 * np.eye(NB)
    Aout, W, WT, rinv, logdet = run(A)
    L = np.linalg.cholesky(A)
    low = np.tril(np.ones((NB, NB), bool))
    assert np.all(np.isnan(Aout[~low])), "wrote above the diagonal"
    eL = np.abs(Aout[low] - L[low]).max()
    Wref = np.linalg.inv(L)
    eW = np.abs(W - Wref).max()
    assert not np.isnan(W).any() and not np.isnan(WT).any()
    assert np.all(W[~low] == 0.0)
    assert np.array_equal(WT, W.T)
    eR = np.abs(rinv - 1.0 / np.diag(L)).max()
    assert abs(logdet - np.linalg.slogdet(A)[1]) < 1e-12 * abs(logdet) + 1e-12
    print(f"max |L - chol| = {eL:.2e}   max |W - inv(L)| = {eW:.2e}   max |rinv - 1/diag| = {eR:.2e}")
    assert eL < 1e-13 and eW < 1e-11 and eR < 1e-12
    print("ok")


if __name__ == "__main__":
    main()
