"""Debug aid: small batch through a forced kernel variant, compared with the generic kernel."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rokifd_b200
from rokifd_b200 import capi, chains as ch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w = ch.world_c3(base_z=0.1)
q, qd, u = ch.sample_state(w, B, seed=3)
def run():
    fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
    fd.update_n(3); out = fd.batch_get_state(); fd.destroy(); return out
got = run()
os.environ["RKFD_SPEC"] = "0"; os.environ.pop("RKFD_FORCE_BLOCK", None); os.environ.pop("RKFD_FORCE_MINB", None)
ref = run()
bad = [int(i) for i in np.where(~np.isfinite(got[0]).all(1) | (np.abs(got[0]-ref[0]).max(1) > 1e-12))[0]]
print("B=%d nonmatching envs: %d  first: %s" % (B, len(bad), bad[:40]))
