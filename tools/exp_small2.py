import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rokifd_b200
from rokifd_b200 import capi, chains as ch
w = ch.world_c3(base_z=0.1)
def run(B, nsteps):
    q, qd, u = ch.sample_state(w, B, seed=3)
    fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
    if nsteps: fd.update_n(nsteps)
    out = fd.batch_get_state(); fd.destroy(); return out
env = dict(os.environ)
for B in (256, 512, 768, 1024, 2048):
    for ns in (0, 1):
        os.environ.update(env)
        got = run(B, ns)
        os.environ["RKFD_SPEC"] = "0"; os.environ.pop("RKFD_FORCE_BLOCK", None); os.environ.pop("RKFD_FORCE_MINB", None)
        ref = run(B, ns)
        d = [np.abs(g - r).max(1) for g, r in zip(got, ref)]
        bad = np.where(~(d[0] < 1e-12) | ~(d[1] < 1e-12) | ~(d[2] < 1e-9))[0]
        print("B=%d nsteps=%d bad envs %d %s maxdiff q %.2e qd %.2e qdd %.2e" % (B, ns, len(bad), bad[:12].tolist(), np.nanmax(d[0]), np.nanmax(d[1]), np.nanmax(d[2])), flush=True)
