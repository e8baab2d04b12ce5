"""TEST INFRASTRUCTURE: rkFDQPSolveASM (reference src/rkfd_opt_qp.c:43-181) restated in arbitrary precision (mpmath).

The active-set loop of the reference decides with ABSOLUTE 1e-12 thresholds (`zEqual(.., zTOL)`, rkfd_opt_qp.c:33,110,148)
on numbers of magnitude 1e2..1e3, and with the relaxation 1e-4 of example/model/contactinfo.ztk the KKT matrices it hands
to zLESolveMP have condition ~1e7: in plain double precision the path (and whether the loop leaves through the
anti-cycling exit :152-171 before reaching the minimiser) is decided by rounding noise.  This file evaluates the SAME
algorithm with 50+ digits: the path exact arithmetic takes.  It pins two things (tests/test_oracle_physics.py):
  * oracle `ork_le_solve_mp_sym` ([EXT A-14] zLESolveMP, with its long-double iterative refinement) against the exact
    minimum-norm solution,
  * oracle `ork_qp_solve_asm` against the exact path on QPs taken from BASELINE config C5 (tests/golden/vert_qp_c5.npz,
    written by tests/golden/make_vert_qp_golden.py).
"""
import mpmath as mp

TOL = mp.mpf("1e-12")      # zTOL
ASM_TOL = mp.mpf("1e-8")   # RKFD_OPT_QP_ASM_TOL


def pinv_solve_sym(K, b, cut=mp.mpf("1e-30")):
    """Minimum-norm least-squares solution of the symmetric system K x = b (eigen-decomposition, exact rank cut)."""
    n = K.rows
    E, V = mp.eigsy(K)
    lmax = max(abs(e) for e in E)
    x = mp.zeros(n, 1)
    for k in range(n):
        if abs(E[k]) <= cut * lmax:
            continue
        s = sum(V[i, k] * b[i] for i in range(n)) / E[k]
        for i in range(n):
            x[i] += s * V[i, k]
    return x


def kkt(Q, A, act):
    n, ma = len(Q), len(act)
    K = mp.zeros(n + ma, n + ma)
    for i in range(n):
        for j in range(n):
            K[i, j] = -mp.mpf(Q[i][j])
    for k, r in enumerate(act):
        for j in range(n):
            K[j, n + k] = mp.mpf(A[r][j])
            K[n + k, j] = mp.mpf(A[r][j])
    return K


def qp_solve_asm(Q, c, A, init, dps=50, max_iter=500):
    """min 1/2 x^T Q x + c^T x  s.t.  A x >= 0 by the reference's active-set loop.  Returns (x, idx, iterations, term);
    term 0: optimal (rkfd_opt_qp.c:113), 1: anti-cycling exit (:165), 2: iteration cap."""
    mp.mp.dps = dps
    n, m = len(c), len(A)
    Q = [[mp.mpf(v) for v in r] for r in Q]
    c = [mp.mpf(v) for v in c]
    A = [[mp.mpf(v) for v in r] for r in A]
    dot = lambda a, b: sum(a[j] * b[j] for j in range(n))
    ans = [mp.mpf(v) for v in init]
    idx = [1 if abs(dot(A[i], ans)) < TOL else 0 for i in range(m)]
    hist = []
    for it in range(1, max_iter + 1):
        act = [i for i in range(m) if idx[i]]
        ma = len(act)
        xy = pinv_solve_sym(kkt(Q, A, act), mp.matrix(c + [mp.mpf(0)] * ma))
        if all(abs(xy[i] - ans[i]) < TOL for i in range(n)):
            ans = [xy[i] for i in range(n)]
            lam = [xy[n + i] for i in range(ma)]
            if not any(l < 0 for l in lam):
                return ans, idx, it, 0
            lmin = min(lam)
            for k, i in enumerate(act):
                if abs(lam[k] - lmin) < ASM_TOL:
                    idx[i] = 0
            continue
        d = [xy[i] - ans[i] for i in range(n)]
        alpha = mp.mpf(1)
        for i in range(m):
            ad = dot(A[i], d)
            if idx[i] == 0 and ad < 0:
                t = (0 - dot(A[i], ans)) / ad
                if t < alpha:
                    alpha = t
        ans = [ans[i] + alpha * d[i] for i in range(n)]
        for i in range(m):
            if idx[i] == 0 and abs(dot(A[i], ans)) < TOL:
                idx[i] = 1
        objv = sum(mp.mpf("0.5") * ans[i] * sum(Q[i][j] * ans[j] for j in range(n)) + c[i] * ans[i] for i in range(n))
        for hidx, hobj in hist:
            if hidx == idx and not (abs(hobj / objv - 1) > ASM_TOL):
                return ans, idx, it, 1
        hist.append((list(idx), objv))
    return ans, idx, max_iter, 2
