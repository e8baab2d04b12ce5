"""TEST INFRASTRUCTURE (it loads the oracle, hence it lives under tests/): parity protocol of SURVEY.md section 8(d) on the GPU: per-evaluation and free-running comparison of the device path
(through the C-ABI) with the CPU oracle on the same synthetic states.  Prints a markdown table (profiles/r01_parity_report.md).

per-evaluation: one committing evaluation (rkFDUpdateInit) on B random envs: rel. error of q'' per env
                ||dq''||_inf / max(||q''_ref||_inf, 1e-12), contact forces likewise, contact / pivot flags identical?
free-running:   H steps of rkFDUpdate: rel. error of q per env, contact flags and friction types identical?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rokifd_b200  # noqa: F401
from rokifd_b200 import capi, chains as ch
from oracle import oracle as orc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = int(sys.argv[2]) if len(sys.argv) > 2 else 100


def rel(a, b):
    a = a.reshape(a.shape[0], -1); b = b.reshape(b.shape[0], -1)
    return np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1e-12)


CFG = [("C1 box on the soft floor (penalty)", ch.World(chains=[ch.box(), ch.floor_soft()], contact_info=[ch.ContactInfo("soft", "body", "elastic", E=100.0, V=1.0, SF=0.5, KF=0.3)]), "box"),
       ("C2 arm7, no contact", ch.world_c2(), None),
       ("C3 arm7 + penalty contact + joint friction", ch.world_c3(base_z=0.3), None),
       ("C4 legged tree + volume contact", ch.world_c4_volume(), "c4"),
       ("C5 arm7 + rigid floor, MLCP", ch.world_c5(base_z=0.3, solver="MLCP"), None),
       ("C5 arm7 + rigid floor, Vert QP (relaxation 1e-4)", ch.world_c5(base_z=0.3, solver="Vert"), None),
       ("arm7 + rigid floor, Vert QP (solver default contact info)", ch.World(chains=[ch.arm7(base_z=0.3, contact_cube=True), ch.floor()], solver="Vert"), None)]

print("| world | envs | contact at t=0 | q'' rel err: max / 99.9 pct / share < 1e-9 | contact force rel err max | flags equal at t=0 | q rel err after {0} steps: max / 99 pct / share < 1e-6 | contact + friction flags equal after {0} steps | oracle s |".format(H))
print("|---|---|---|---|---|---|---|---|---|")
for name, w, pose in CFG:
    q, qd, u = ch.sample_state(w, B, seed=20260418)
    if pose == "box":
        q[:, 2] = np.linspace(0.0, 0.3, B); q[:, 3:6] *= 0.3
    if pose == "c4":
        q, qd, u = ch.sample_c4_standing(w, B, seed=20260418)
    t0 = time.time()
    ow = orc.OracleWorld(w)
    o0 = ow.batch_run_state(q, qd, u, nsteps=0)
    oH = ow.batch_run_state(q, qd, u, nsteps=H)
    tor = time.time() - t0
    fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
    _, _, gqdd = fd.batch_get_state()
    a0, t0_, _, f0 = fd.batch_get_contact() if w.nslot else (np.zeros((B, 0), int),) * 4
    p0 = fd.batch_get_pivot()[0]
    fd.update_n(H)
    gq, gqd, _ = fd.batch_get_state()
    aH, tH, _, _ = fd.batch_get_contact() if w.nslot else (np.zeros((B, 0), int),) * 4
    pH = fd.batch_get_pivot()[0]
    bad = int((fd.batch_get_status() != 0).sum())
    fd.destroy()
    e0 = rel(gqdd, o0[2])
    volume = w.solver == "Volume" and w.nslot
    if w.nslot and not volume:
        ef = rel(f0 * (o0[3] > 0)[:, :, None], o0[5] * (o0[3] > 0)[:, :, None])
        efs = "%.1e" % ef.max()
    else:
        efs = "(pair wrenches: tests)" if volume else "-"
    if w.nslot:
        tm0 = (t0_ == o0[4]) | (o0[3] == 0) if not volume else np.ones_like(a0, bool)
        tmH = (tH == oH[4]) | (oH[3] == 0) if not volume else np.ones_like(aH, bool)
        fl0 = ((a0 == o0[3]).all(1) & tm0.all(1) & (p0 == o0[6]).all(1)).mean()
        flH = ((aH == oH[3]).all(1) & tmH.all(1) & (pH == oH[6]).all(1)).mean()
        c0 = (o0[3].sum(1) > 0).mean()
    else:
        fl0 = (p0 == o0[6]).all(1).mean(); flH = (pH == oH[6]).all(1).mean(); c0 = 0.0
    fin = np.isfinite(oH[0]).all(1) & (np.abs(oH[0]).max(1) < 1e6)
    eH = rel(gq[fin], oH[0][fin])
    print("| %s | %d | %.1f %% | %.1e / %.1e / %.2f %% | %s | %.2f %% | %.1e / %.1e / %.2f %% (%d diverged in the oracle too, %d flagged) | %.2f %% | %.0f |" % (
        name, B, 100 * c0, e0.max(), np.quantile(e0, 0.999), 100 * (e0 < 1e-9).mean(), efs, 100 * fl0,
        eH.max(), np.quantile(eH, 0.99), 100 * (eH < 1e-6).mean(), int((~fin).sum()), bad, 100 * flH, tor), flush=True)
