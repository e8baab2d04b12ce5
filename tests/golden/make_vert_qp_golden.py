"""Writes tests/golden/vert_qp_c5.npz: Vert QPs (Q, c, nf) of BASELINE config C5 (arm7 + cube on the rigid floor,
contactinfo.ztk's K=1000 / L=1e-4) for the environments on which a plain-double zLESolveMP left the reference's
active-set loop through its anti-cycling exit before the minimiser, together with what
rkFDQPSolveASM returns when evaluated in 50-digit arithmetic (tests/ref_qp_mp.py): x, final active set, iteration
count, termination.  Run from the repo root: python tests/golden/make_vert_qp_golden.py  (a few minutes)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rokifd_b200  # noqa: F401,E402
from rokifd_b200 import chains as ch  # noqa: E402
from oracle import oracle as orc  # noqa: E402
import ref_qp_mp  # noqa: E402

ENVS = [2862, 3172, 1585, 261, 1551, 21, 90, 3344]       # of the 4,096 seed-20260418 environments, base_z = 0.3
w = ch.world_c5(base_z=0.3, solver="Vert")
q, qd, u = ch.sample_state(w, 4096, seed=20260418)
ow = orc.OracleWorld(w)
out = {"envs": np.array(ENVS)}
for e in ENVS:
    env = ow.env(); env.set_state(q[e], qd[e]); env.set_motor_input(u[e]); env.eval(True)
    Q, c, nf, x, idx, it, term = env.qp()
    n = len(c)
    init = [1.0 if i % 3 == 0 else 0.0 for i in range(n)]
    xm, im, itm, tm = ref_qp_mp.qp_solve_asm(Q.tolist(), c.tolist(), nf.tolist(), init)
    print("env %d n=%d: exact path %d iterations, term %d; oracle %d iterations, term %d, |x - x_exact| %.2e" % (
        e, n, itm, tm, it, term, max(abs(float(a) - b) for a, b in zip(xm, x))))
    out["Q%d" % e] = Q; out["c%d" % e] = c; out["nf%d" % e] = nf
    out["x%d" % e] = np.array([float(v) for v in xm]); out["idx%d" % e] = np.array(im, np.int32)
    out["info%d" % e] = np.array([itm, tm], np.int32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vert_qp_c5.npz"), **out)
