import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rokifd_b200
from rokifd_b200 import capi, chains as ch
w = ch.world_c3(base_z=0.1)
def run(B, nsteps):
    q, qd, u = ch.sample_state(w, B, seed=3)
    fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
    if nsteps: fd.update_n(nsteps)
    out = fd.batch_get_state(); fd.destroy(); return out
res = []
for B in (256, 256, 256, 4096, 4096):
    got = run(B, 2)
    res.append("%d:%s" % (B, "ok" if np.isfinite(got[0]).all() and np.isfinite(got[2]).all() else "NaN(%d)" % (~np.isfinite(got[2]).all(1)).sum()))
print(os.environ.get("TAG"), " ".join(res))
