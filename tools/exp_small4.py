import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rokifd_b200
from rokifd_b200 import capi, chains as ch
w = ch.world_c3(base_z=0.1)
def run(B, nsteps):
    q, qd, u = ch.sample_state(w, B, seed=3)
    fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
    if nsteps: fd.update_n(nsteps)
    out = fd.batch_get_state(); fd.destroy(); return out
for B in (256, 4096, 262144):
    res = []
    for ns in (0, 1, 2, 3):
        got = run(B, ns)
        nb = (~np.isfinite(got[2]).all(1)).sum()
        res.append("%d:%s" % (ns, "ok" if nb == 0 else "NaN(%d)" % nb))
    print(os.environ.get("TAG"), "B=%d" % B, " ".join(res), flush=True)
