"""Independent numpy cross-check of the oracle's ABA: world-frame recursive Newton-Euler inverse
dynamics + joint-space inertia by unit accelerations, solved densely: (M + Jm) qdd = tau - h.
Deliberately a different formulation from oracle/rkfd_oracle.c (world frame, inverse dynamics,
dense solve) so that agreement is evidence, not tautology.  Test infrastructure only."""
import numpy as np

G = 9.80665
NDOF = {"fixed": 0, "revolute": 1, "prismatic": 1, "spherical": 3, "float": 6, "cylindrical": 2, "hooke": 2}


def skew(p):
    return np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]], float)


def aa_to_mat(aa):
    th = np.linalg.norm(aa)
    if th < 1e-12:
        return np.eye(3) + skew(aa)
    K = skew(aa / th)
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def rnea(links, q, qd, qdd, fext=None, gravity=True):
    """tau = ID(q, qd, qdd) with world-frame vectors. fext: per-link (f_w, n_w about link origin, world)."""
    n = len(links)
    R = [None] * n; p = [None] * n; w = [None] * n; dw = [None] * n; v = [None] * n; a = [None] * n
    axes = [None] * n
    ofs = np.cumsum([0] + [NDOF[l.jtype] for l in links])
    for i, l in enumerate(links):
        qi, vi, ai = q[ofs[i]:ofs[i + 1]], qd[ofs[i]:ofs[i + 1]], qdd[ofs[i]:ofs[i + 1]]
        if l.parent >= 0:
            Rp, pp, wp, dwp, vp, ap = R[l.parent], p[l.parent], w[l.parent], dw[l.parent], v[l.parent], a[l.parent]
        else:
            Rp, pp, wp, dwp, vp, ap = np.eye(3), np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3)
        Ro = Rp @ np.asarray(l.org_R, float)          # org frame in world (fixed to parent)
        RJ, pJ = np.eye(3), np.zeros(3)
        wJ, dwJ, vJ, dvJ = np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3)   # in org frame
        if l.jtype == "revolute":
            c, s = np.cos(qi[0]), np.sin(qi[0])
            RJ = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
            wJ, dwJ = np.array([0, 0, vi[0]]), np.array([0, 0, ai[0]])
        elif l.jtype == "prismatic":
            pJ = np.array([0, 0, qi[0]]); vJ = np.array([0, 0, vi[0]]); dvJ = np.array([0, 0, ai[0]])
        elif l.jtype == "spherical":
            RJ = aa_to_mat(qi); wJ, dwJ = vi.copy(), ai.copy()
        elif l.jtype == "float":
            pJ = qi[:3].copy(); RJ = aa_to_mat(qi[3:]); vJ, dvJ = vi[:3].copy(), ai[:3].copy()
            wJ, dwJ = vi[3:].copy(), ai[3:].copy()
        elif l.jtype == "cylindrical":      # slides along z (q0), turns about z (q1)
            c, s = np.cos(qi[1]), np.sin(qi[1])
            RJ = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]); pJ = np.array([0, 0, qi[0]])
            vJ, dvJ = np.array([0, 0, vi[0]]), np.array([0, 0, ai[0]])
            wJ, dwJ = np.array([0, 0, vi[1]]), np.array([0, 0, ai[1]])
        elif l.jtype == "hooke":            # R = Rz(q0) Ry(q1): angular velocity z q0' + (Rz y) q1' in the org frame
            c0, s0, c1, s1 = np.cos(qi[0]), np.sin(qi[0]), np.cos(qi[1]), np.sin(qi[1])
            Rz = np.array([[c0, -s0, 0], [s0, c0, 0], [0, 0, 1]]); Ry = np.array([[c1, 0, s1], [0, 1, 0], [-s1, 0, c1]])
            RJ = Rz @ Ry
            ez, y1 = np.array([0, 0, 1.0]), Rz @ np.array([0, 1.0, 0])
            wJ = ez * vi[0] + y1 * vi[1]
            dwJ = ez * ai[0] + y1 * ai[1] + np.cross(ez * vi[0], y1 * vi[1])
        r = Rp @ (np.asarray(l.org_p, float)) + Ro @ pJ
        R[i] = Ro @ RJ
        p[i] = pp + r
        w[i] = wp + Ro @ wJ
        dw[i] = dwp + np.cross(wp, Ro @ wJ) + Ro @ dwJ
        v[i] = vp + np.cross(wp, r) + Ro @ vJ
        a[i] = ap + np.cross(dwp, r) + np.cross(wp, np.cross(wp, r)) + 2 * np.cross(wp, Ro @ vJ) + Ro @ dvJ
        axes[i] = Ro
    f = [None] * n; nn = [None] * n
    for i, l in enumerate(links):
        c = R[i] @ np.asarray(l.com, float)
        ac = a[i] + np.cross(dw[i], c) + np.cross(w[i], np.cross(w[i], c))
        Iw = R[i] @ np.asarray(l.inertia, float) @ R[i].T
        F = l.mass * ac
        if gravity:
            F = F + l.mass * np.array([0, 0, G])
        N = Iw @ dw[i] + np.cross(w[i], Iw @ w[i])
        f[i] = F.copy()
        nn[i] = N + np.cross(c, F)               # about link origin
        if fext is not None:
            f[i] -= fext[i][0]; nn[i] -= fext[i][1]
    tau = np.zeros(ofs[-1])
    for i in range(n - 1, -1, -1):
        l = links[i]
        Ro = axes[i]
        if l.jtype == "revolute":
            tau[ofs[i]] = Ro[:, 2] @ nn[i]
        elif l.jtype == "prismatic":
            tau[ofs[i]] = Ro[:, 2] @ f[i]
        elif l.jtype == "spherical":
            tau[ofs[i]:ofs[i + 1]] = Ro.T @ nn[i]
        elif l.jtype == "cylindrical":
            tau[ofs[i]] = Ro[:, 2] @ f[i]; tau[ofs[i] + 1] = Ro[:, 2] @ nn[i]
        elif l.jtype == "hooke":
            c0, s0 = np.cos(q[ofs[i]]), np.sin(q[ofs[i]])
            tau[ofs[i]] = Ro[:, 2] @ nn[i]; tau[ofs[i] + 1] = (Ro @ np.array([-s0, c0, 0.0])) @ nn[i]
        elif l.jtype == "float":
            tau[ofs[i]:ofs[i] + 3] = Ro.T @ f[i]
            tau[ofs[i] + 3:ofs[i] + 6] = Ro.T @ nn[i]
        if l.parent >= 0:
            f[l.parent] += f[i]
            nn[l.parent] += nn[i] + np.cross(p[i] - p[l.parent], f[i])
    return tau


def forward_dynamics(links, q, qd, tau, jm=None, fext=None):
    nq = len(q)
    h = rnea(links, q, qd, np.zeros(nq), fext=fext)
    h0 = rnea(links, q, np.zeros(nq), np.zeros(nq), gravity=False)
    M = np.zeros((nq, nq))
    for j in range(nq):
        e = np.zeros(nq); e[j] = 1.0
        M[:, j] = rnea(links, q, np.zeros(nq), e, gravity=False) - h0
    if jm is not None:
        M = M + np.diag(jm)
    return np.linalg.solve(M, tau - h), M, h


def motor_terms(links, qd, u):
    """(tau_drive, Jm) per dof for 1-DoF joints (A-6)."""
    ofs = np.cumsum([0] + [NDOF[l.jtype] for l in links])
    td, jm = np.zeros(ofs[-1]), np.zeros(ofs[-1])
    for i, l in enumerate(links):
        m = l.motor
        if m is None or NDOF[l.jtype] != 1:
            continue
        e = min(max(u[i], m.min), m.max)
        if m.type == "dc":
            td[ofs[i]] = m.gear * m.k * m.admittance * e - (m.gear * m.k) ** 2 * m.admittance * qd[ofs[i]]
            jm[ofs[i]] = m.gear ** 2 * (m.rotor_inertia + m.gear_inertia)
        else:
            td[ofs[i]] = e
    return td, jm
