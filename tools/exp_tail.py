"""Tuning aid: wave quantisation of the step kernel.  592 blocks of 128 threads are resident on the 148 SMs (4 per SM), so 262,144
environments = 2,048 blocks = 3.46 waves.  Times the settled C3 step for batch sizes that are whole waves and for the bench batch."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rokifd_b200
from rokifd_b200 import capi, chains as ch
w = ch.world_c3()
for B in [int(x) for x in (sys.argv[1:] or [227328, 262144, 303104, 151552, 75776])]:
    q, qd, u = ch.sample_state(w, B, seed=20260418)
    fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
    st = torch.cuda.current_stream(); fd.batch_set_stream(st.cuda_stream)
    fd.update_n(700)
    for _ in range(10): fd.update()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(50): fd.update()
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print("%s B %7d (%.2f waves of 592 blocks): %.4f ms/step  %.3f ns/env-step  %.3e env-steps/s" % (os.environ.get("TAG", ""), B, B / 128 / 592, ms, ms * 1e6 / B, B / ms * 1e3), flush=True)
    fd.destroy()
