"""Tuning aid: ms/step of the rigid-contact path (C5: arm7 + cube on the rigid floor, MLCP / Vert-QP)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rokifd_b200
from rokifd_b200 import capi, chains as ch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for solver in ("MLCP", "Vert"):
    w = ch.world_c5(base_z=0.45, solver=solver)
    q, qd, u = ch.sample_state(w, B, seed=20260418)
    fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
    st = torch.cuda.current_stream(); fd.batch_set_stream(st.cuda_stream)
    tsim = 0
    for n in [20, 280, 100, 100]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(n): fd.update()
        e1.record(st); torch.cuda.synchronize()
        tsim += n
        a, t, r, f = fd.batch_get_contact()
        print("%s B=%d steps %4d: %.3f ms/step  %.3e env-steps/s  envs in contact %.3f  mean active verts %.2f  status!=0 %d" % (
            solver, B, tsim, e0.elapsed_time(e1) / n, B * n / (e0.elapsed_time(e1) * 1e-3), (a.sum(1) > 0).mean(), a.sum(1).mean(),
            (fd.batch_get_status() != 0).sum()), flush=True)
    fd.destroy()
