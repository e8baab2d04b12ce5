"""Tuning aid: ms/step of the step kernel on C3 (262144 envs) in the contact-free and the settled regime."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rokifd_b200
from rokifd_b200 import capi, chains as ch
base_z = 0.45
w = ch.world_c3(base_z=base_z)
B = 262144
q, qd, u = ch.sample_state(w, B, seed=20260418)
fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
st = torch.cuda.current_stream(); fd.batch_set_stream(st.cuda_stream)
tsim = 0
for n in [20, 20, 660, 100, 100]:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n): fd.update()
    e1.record(st); torch.cuda.synchronize()
    tsim += n
    a, t, r, f = fd.batch_get_contact()
    print("%s steps %4d: %.3f ms/step  envs in contact %.3f  mean active verts %.2f" % (
        os.environ.get("TAG", ""), tsim, e0.elapsed_time(e1) / n, (a.sum(1) > 0).mean(), a.sum(1).mean()), flush=True)
gq, gqd, _ = fd.batch_get_state()
print("checksum q %.12e qd %.12e" % (np.abs(gq).sum(), np.abs(gqd).sum()))
fd.destroy()
