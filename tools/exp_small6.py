import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rokifd_b200
from rokifd_b200 import capi, chains as ch
np.set_printoptions(precision=4, linewidth=200)
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
w = ch.world_c3(base_z=0.1) if which == "c3" else ch.world_c2()
B = 256
q, qd, u = ch.sample_state(w, B, seed=3)
fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
gq, gqd, gqdd = fd.batch_get_state()
pt, pp = fd.batch_get_pivot()
print(os.environ.get("TAG"), which, "nan qdd envs", (~np.isfinite(gqdd).all(1)).sum(), "nan prev_trq envs", (~np.isfinite(pp).all(1)).sum(), "status!=0", (fd.batch_get_status() != 0).sum())
print(" qdd[0]", gqdd[0]); print(" prev_trq[0]", pp[0], "pivtype[0]", pt[0])
if w.nslot:
    a, t, r, f = fd.batch_get_contact()
    print(" active[0]", a[0], "nan cf envs", (~np.isfinite(f).all((1,2))).sum() if f.ndim == 3 else (~np.isfinite(f).all(1)).sum())
fd.destroy()
