"""Tuning aid: environment re-sort interval (rkFDBatchSetResortInterval).  Times 256 consecutive settled steps (the sorts that fall
into them included) for several intervals."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rokifd_b200
from rokifd_b200 import capi, chains as ch
cfgs = {"C3": (ch.world_c3(), 262144, 700), "C5-mlcp": (ch.world_c5(base_z=0.45, solver="MLCP"), 131072, 500), "C5-vert": (ch.world_c5(base_z=0.45, solver="Vert"), 131072, 500)}
for name in (sys.argv[1:] or ["C3"]):
    w, B, settle = cfgs[name]
    q, qd, u = ch.sample_state(w, B, seed=20260418)
    for interval in (0, 8, 16, 32, 64, 128):
        fd, _ = capi.create_world(w, B=B); fd.batch_set_resort_interval(interval)
        fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
        st = torch.cuda.current_stream(); fd.batch_set_stream(st.cuda_stream)
        fd.update_n(settle)
        for _ in range(10): fd.update()
        n0 = fd.resort_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(256): fd.update()
        e1.record(st); torch.cuda.synchronize()
        print("%s interval %3d: %.4f ms/step over 256 steps (%d sorts inside)" % (name, interval, e0.elapsed_time(e1) / 256, fd.resort_count - n0), flush=True)
        fd.destroy()
