"""Tuning aid: settled C3 step time (262,144 envs, 700 settle steps, 200 timed steps incl. the re-sorts)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rokifd_b200
from rokifd_b200 import capi, chains as ch
w = ch.world_c3(); B = 262144
q, qd, u = ch.sample_state(w, B, seed=20260418)
fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
st = torch.cuda.current_stream(); fd.batch_set_stream(st.cuda_stream)
fd.update_n(700)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(200): fd.update()
    e1.record(st); torch.cuda.synchronize()
    print("C3 settled: %.4f ms/step" % (e0.elapsed_time(e1) / 200), flush=True)
gq = fd.batch_get_state()[0]; print("checksum %.17g" % float(np.abs(gq).sum()))
fd.destroy()
