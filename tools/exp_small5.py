import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rokifd_b200
from rokifd_b200 import capi, chains as ch
w = ch.world_c3(base_z=0.1)
B = 256
q, qd, u = ch.sample_state(w, B, seed=3)
fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
got = fd.batch_get_state(); fd.batch_sync()
print("q[0]", q[0]); print("qdd[0]", got[2][0]); print("nan envs", (~np.isfinite(got[2]).all(1)).sum())
fd.destroy()
