"""ctypes binding of the CPU oracle (oracle/librkfd_oracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  The product package
(roki-fd_b200/) never does.  Parity is UNPINNED (see rkfd_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

JOINT = {"fixed": 0, "revolute": 1, "prismatic": 2, "spherical": 3, "float": 4, "cylindrical": 5, "hooke": 6, "breakablefloat": 7}
MOTOR = {None: 0, "none": 0, "dc": 1, "trq": 2}
CONTACT = {"rigid": 0, "elastic": 1}
SOLVER = {"Vert": 0, "MLCP": 1, "Volume": 2}
LINK_ND = 37

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "librkfd_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("rkfd_oracle.c", "rkfd_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ork_world_new.restype = C.c_void_p
        L.ork_world_new.argtypes = [C.c_int, _ip, _dp]
        L.ork_world_free.argtypes = [C.c_void_p]
        L.ork_world_add_cell.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
        L.ork_world_add_box.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int]
        L.ork_world_add_contact_info.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_double] * 6
        L.ork_world_set_prp.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_int]
        L.ork_world_set_solver.argtypes = [C.c_void_p, C.c_int]
        L.ork_world_finalize.argtypes = [C.c_void_p]
        for f in ("ork_world_nq", "ork_world_nslot", "ork_world_nl"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.ork_env_new.restype = C.c_void_p
        L.ork_env_new.argtypes = [C.c_void_p]
        L.ork_env_free.argtypes = [C.c_void_p]
        L.ork_env_set_state.argtypes = [C.c_void_p, _dp, _dp]
        L.ork_env_get_state.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.ork_env_set_motor_input.argtypes = [C.c_void_p, _dp]
        L.ork_env_get_pivot.argtypes = [C.c_void_p, _ip, _dp]
        L.ork_env_set_pivot.argtypes = [C.c_void_p, _ip, _dp]
        L.ork_env_get_contact.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp]
        L.ork_env_set_contact.argtypes = [C.c_void_p, _ip, _ip, _dp]
        L.ork_env_time.restype = C.c_double
        L.ork_env_time.argtypes = [C.c_void_p]
        L.ork_env_update_init.argtypes = [C.c_void_p]
        L.ork_env_update.argtypes = [C.c_void_p]
        L.ork_env_eval.argtypes = [C.c_void_p, C.c_int]
        L.ork_env_get_link_frames.argtypes = [C.c_void_p, _dp]
        L.ork_env_get_link_vel.argtypes = [C.c_void_p, _dp]
        L.ork_env_get_link_acc.argtypes = [C.c_void_p, _dp]
        L.ork_env_energy.restype = C.c_double
        L.ork_env_energy.argtypes = [C.c_void_p]
        L.ork_env_get_rigid_system.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int]
        L.ork_batch_run.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, _dp]
        L.ork_batch_run_state.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, _dp, _ip, _ip, _dp, _ip]
        L.ork_qp_solve_asm.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _ip]
        L.ork_env_get_qp.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _ip, _ip, C.c_int, C.c_int]
        L.ork_le_solve_mp_sym.argtypes = [C.c_int, _dp, _dp, _dp]
        L.ork_env_get_volume.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp]
        L.ork_world_npair.argtypes = [C.c_void_p]
        L.ork_env_get_volume_constraint.argtypes = [C.c_void_p, _dp]
        L.ork_lp_solve.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        _LIB = L
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


def pack_links(world, links):
    """Flat link description (layout in rkfd_oracle.h): li (nl x 4 int32), ld (nl x LINK_ND float64)."""
    nl = len(links)
    li = np.zeros((nl, 4), np.int32)
    ld = np.zeros((nl, LINK_ND), np.float64)
    for k, l in enumerate(links):
        m = l.motor
        li[k] = (l.parent, JOINT[l.jtype], MOTOR[m.type if m else None], world.stuff_id(l.stuff))
        ld[k, 0:9] = np.asarray(l.org_R, float).reshape(9)
        ld[k, 9:12] = l.org_p
        ld[k, 12] = l.mass
        ld[k, 13:16] = l.com
        ld[k, 16:25] = np.asarray(l.inertia, float).reshape(9)
        ld[k, 25:29] = (l.stiffness, l.viscosity, l.coulomb, l.sfriction)
        if m:
            ld[k, 29:36] = (m.k, m.admittance, m.gear, m.rotor_inertia, m.gear_inertia, m.min, m.max)
        if l.jtype == "breakablefloat":        # thresholds in spare fields (tests/hostsim reads them there; the oracle gets them by call)
            ld[k, 36], ld[k, 29] = l.break_force, l.break_torque
    return li, ld


class OracleWorld:
    """Builds an oracle world from a rokifd_b200.chains.World description."""

    def __init__(self, world):
        L = lib()
        self.desc = world
        links = world.flat_links()
        nl = len(links)
        li, ld = pack_links(world, links)
        _, pi = _i(li)
        _, pd = _d(ld)
        self.h = L.ork_world_new(nl, pi, pd)
        L.ork_world_set_break.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        for k, l in enumerate(links):
            if l.jtype == "breakablefloat":
                L.ork_world_set_break(self.h, k, l.break_force, l.break_torque)
        L.ork_world_add_link_box.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
        L.ork_world_unreg_self_collision.argtypes = [C.c_void_p, C.c_int]
        L.ork_world_set_slide.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, _dp, C.c_int, _dp, _dp]
        order = world.shape_order()
        mlinks = [(chn, l) for chn in world.moving_chains() for l in chn.links]      # the originals of `links`, same order

        def set_slide(is_box, idx, sl, ordv, lR=None, lp=None):
            _, pa = _d(np.asarray(sl[1], float) if sl else np.zeros(3))
            L.ork_world_set_slide(self.h, is_box, idx, 1 if sl else 0, float(sl[0]) if sl else 0.0, pa, ordv,
                                  _d(np.asarray(lR, float).reshape(9))[1] if lR is not None else None, _d(np.asarray(lp, float))[1] if lp is not None else None)
        for k, l in enumerate(links):
            chn, lo = mlinks[k]
            for ci, verts in enumerate(l.cells()):
                v, pv = _d(verts)
                c_idx = L.ork_world_add_cell(self.h, k, v.shape[0], pv)
                set_slide(0, c_idx, l.slides.get(ci), order[(id(chn), id(lo), ci)])
            for bi, (ctr, d, w_, h_) in enumerate(l.boxes):       # box primitives of a moving link: targets for other links' vertices
                _, pR = _d(np.eye(3).reshape(9)); _, pp = _d(np.asarray(ctr, float)); _, ph = _d(np.array([d / 2, w_ / 2, h_ / 2]))
                b_idx = L.ork_world_add_link_box(self.h, k, pR, pp, ph)
                set_slide(1, b_idx, l.slides.get(len(l.shapes) + bi), order[(id(chn), id(lo), len(l.shapes) + bi)])
        base = 0
        for chn in world.moving_chains():
            if not chn.self_collide:
                L.ork_world_unreg_self_collision(self.h, base)
            base += len(chn.links)
        for b in world.boxes:
            _, pR = _d(np.asarray(b.R, float).reshape(9))
            _, pp = _d(b.p)
            _, ph = _d(b.half)
            b_idx = L.ork_world_add_box(self.h, pR, pp, ph, world.stuff_id(b.stuff))
            set_slide(1, b_idx, b.slide, b.order, b.link_R, b.link_p)
        L.ork_world_set_prp(self.h, world.dt, world.pyramid, world.friction_weight, world.max_iter)
        L.ork_world_set_integrator.argtypes = [C.c_void_p, C.c_int]
        L.ork_world_set_integrator(self.h, {"RKG": 0, "RK4": 1, "Euler": 2, "Heun": 3}[getattr(world, "integrator", "RKG")])
        L.ork_world_set_solver(self.h, SOLVER[world.solver])
        for ci in world.contact_info:
            L.ork_world_add_contact_info(self.h, world.stuff_id(ci.stuff_a), world.stuff_id(ci.stuff_b),
                                         CONTACT[ci.type], ci.K, ci.L, ci.E, ci.V, ci.SF, ci.KF)
        L.ork_world_finalize(self.h)
        self.nq = L.ork_world_nq(self.h)
        self.nl = L.ork_world_nl(self.h)
        self.nslot = L.ork_world_nslot(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ork_world_free(self.h)
            self.h = None

    def env(self):
        return OracleEnv(self)

    def batch_run(self, q, qd, u=None, nsteps=1, nthreads=0):
        """Steps B independent envs `nsteps` times (UpdateInit + nsteps x Update). Returns (q, qd, qdd, threads)."""
        q, pq = _d(np.array(q, dtype=np.float64, copy=True))
        qd, pqd = _d(np.array(qd, dtype=np.float64, copy=True))
        B = q.shape[0]
        qdd = np.zeros_like(q)
        pu = None
        if u is not None:
            u, pu = _d(u)
        used = lib().ork_batch_run(self.h, B, pq, pqd, pu, nsteps, nthreads, qdd.ctypes.data_as(_dp))
        return q, qd, qdd, used

    def batch_run_state(self, q, qd, u=None, nsteps=1, nthreads=0):
        """batch_run plus the final contact / pivot state: (q, qd, qdd, active, type, contact force, pivot type)."""
        q, pq = _d(np.array(q, dtype=np.float64, copy=True))
        qd, pqd = _d(np.array(qd, dtype=np.float64, copy=True))
        B, ns, nq = q.shape[0], max(self.nslot, 1), max(self.nq, 1)
        qdd = np.zeros_like(q)
        pu = None
        if u is not None:
            u, pu = _d(u)
        act, typ, piv = np.zeros((B, ns), np.int32), np.zeros((B, ns), np.int32), np.zeros((B, nq), np.int32)
        cf = np.zeros((B, ns, 3))
        lib().ork_batch_run_state(self.h, B, pq, pqd, pu, nsteps, nthreads, qdd.ctypes.data_as(_dp), act.ctypes.data_as(_ip),
                                  typ.ctypes.data_as(_ip), cf.ctypes.data_as(_dp), piv.ctypes.data_as(_ip))
        n = self.nslot
        return q, qd, qdd, act[:, :n], typ[:, :n], cf[:, :n], piv[:, :self.nq]


class OracleEnv:
    def __init__(self, w):
        self.w = w
        self.h = lib().ork_env_new(w.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ork_env_free(self.h)
            self.h = None

    def set_state(self, q, qd):
        _, pq = _d(q)
        _, pqd = _d(qd)
        lib().ork_env_set_state(self.h, pq, pqd)

    def get_state(self):
        n = max(self.w.nq, 1)
        q, qd, qdd = np.zeros(n), np.zeros(n), np.zeros(n)
        lib().ork_env_get_state(self.h, q.ctypes.data_as(_dp), qd.ctypes.data_as(_dp), qdd.ctypes.data_as(_dp))
        return q[:self.w.nq], qd[:self.w.nq], qdd[:self.w.nq]

    def set_motor_input(self, u):
        _, pu = _d(u)
        lib().ork_env_set_motor_input(self.h, pu)

    def get_pivot(self):
        n = max(self.w.nq, 1)
        t, p = np.zeros(n, np.int32), np.zeros(n)
        lib().ork_env_get_pivot(self.h, t.ctypes.data_as(_ip), p.ctypes.data_as(_dp))
        return t[:self.w.nq], p[:self.w.nq]

    def set_pivot(self, t, p):
        _, pt = _i(t)
        _, pp = _d(p)
        lib().ork_env_set_pivot(self.h, pt, pp)

    def get_contact(self):
        n = max(self.w.nslot, 1)
        a, t = np.zeros(n, np.int32), np.zeros(n, np.int32)
        r, f = np.zeros((n, 3)), np.zeros((n, 3))
        lib().ork_env_get_contact(self.h, a.ctypes.data_as(_ip), t.ctypes.data_as(_ip),
                                  r.ctypes.data_as(_dp), f.ctypes.data_as(_dp))
        ns = self.w.nslot
        return a[:ns], t[:ns], r[:ns], f[:ns]

    def set_contact(self, a, t, r):
        _, pa = _i(a)
        _, pt = _i(t)
        _, pr = _d(r)
        lib().ork_env_set_contact(self.h, pa, pt, pr)

    @property
    def t(self):
        return lib().ork_env_time(self.h)

    def update_init(self):
        lib().ork_env_update_init(self.h)

    def update(self):
        lib().ork_env_update(self.h)

    def eval(self, do_up_ref=False):
        lib().ork_env_eval(self.h, int(do_up_ref))
        return self.get_state()[2]

    def link_frames(self):
        fr = np.zeros((self.w.nl, 12))
        lib().ork_env_get_link_frames(self.h, fr.ctypes.data_as(_dp))
        return fr[:, :9].reshape(-1, 3, 3), fr[:, 9:]

    def link_vel(self):
        v = np.zeros((self.w.nl, 6))
        lib().ork_env_get_link_vel(self.h, v.ctypes.data_as(_dp))
        return v

    def link_acc(self):
        a = np.zeros((self.w.nl, 6))
        lib().ork_env_get_link_acc(self.h, a.ctypes.data_as(_dp))
        return a

    def energy(self):
        return lib().ork_env_energy(self.h)

    def rigid_system(self):
        cap = 3 * max(self.w.nslot, 1)
        A, b, f = np.zeros((cap, cap)), np.zeros(cap), np.zeros(cap)
        n = lib().ork_env_get_rigid_system(self.h, A.ctypes.data_as(_dp), b.ctypes.data_as(_dp),
                                           f.ctypes.data_as(_dp), cap)
        return A.reshape(-1)[:n * n].reshape(n, n), b[:n], f[:n]

    def qp(self):
        """Test hook: the Vert QP of the last evaluation: (Q, c, nf, x, idx, iterations, term); term 0 optimal,
        1 anti-cycling exit (rkfd_opt_qp.c:152-171), 2 iteration cap."""
        info = np.zeros(4, np.int32)
        lib().ork_env_get_qp(self.h, None, None, None, None, None, info.ctypes.data_as(_ip), 0, 0)
        n, m = int(info[0]), int(info[1])
        if n == 0:
            return None
        Q, c, nf, x, idx = np.zeros((n, n)), np.zeros(n), np.zeros((m, n)), np.zeros(n), np.zeros(m, np.int32)
        lib().ork_env_get_qp(self.h, Q.ctypes.data_as(_dp), c.ctypes.data_as(_dp), nf.ctypes.data_as(_dp), x.ctypes.data_as(_dp),
                             idx.ctypes.data_as(_ip), info.ctypes.data_as(_ip), n, m)
        return Q, c, nf, x, idx, int(info[2]), int(info[3])

    def volume(self):
        """Volume solver results of the last evaluation per pair: (np, type, wrench[6], center[3]); np = -1: no contact volume."""
        n = max(lib().ork_world_npair(self.w.h), 1)
        npl, ty = np.zeros(n, np.int32), np.zeros(n, np.int32)
        wr, ce = np.zeros((n, 6)), np.zeros((n, 3))
        lib().ork_env_get_volume(self.h, npl.ctypes.data_as(_ip), ty.ctypes.data_as(_ip), wr.ctypes.data_as(_dp), ce.ctypes.data_as(_dp))
        return npl, ty, wr, ce

    def volume_constraint(self):
        """Per pair (Q6 [6,6], c6 [6], norm [3]) of the last Volume evaluation."""
        n = max(lib().ork_world_npair(self.w.h), 1)
        qc = np.zeros((n, 45))
        lib().ork_env_get_volume_constraint(self.h, qc.ctypes.data_as(_dp))
        return qc[:, :36].reshape(n, 6, 6), qc[:, 36:42], qc[:, 42:45]


def lp_solve(A, b, c=None):
    """min c^T x s.t. A x = b, x >= 0 (c None: feasibility).  Returns (ok, x)."""
    b, pb = _d(b)
    m = b.shape[0]
    A, pA = _d(np.asarray(A, float).reshape(m, -1))
    n = A.shape[1]
    pc = None
    if c is not None:
        c, pc = _d(c)
    x = np.zeros(n)
    ok = lib().ork_lp_solve(m, n, pA, pb, pc, x.ctypes.data_as(_dp))
    return bool(ok), x


def qp_solve_asm(Q, c, A, b, init=None):
    Q, pQ = _d(Q)
    c, pc = _d(c)
    n = c.shape[0]
    A, pA = _d(np.asarray(A, float).reshape(-1, n))
    b, pb = _d(b)
    m = b.shape[0]
    pinit = None
    if init is not None:
        init, pinit = _d(init)
    ans = np.zeros(n)
    idx = np.zeros(max(m, 1), np.int32)
    it = lib().ork_qp_solve_asm(n, m, pQ, pc, pA, pb, pinit, ans.ctypes.data_as(_dp), idx.ctypes.data_as(_ip))
    return ans, idx[:m], it


def le_solve_mp_sym(A, b):
    A, pA = _d(A)
    b, pb = _d(b)
    x = np.zeros(b.shape[0])
    lib().ork_le_solve_mp_sym(b.shape[0], pA, pb, x.ctypes.data_as(_dp))
    return x
