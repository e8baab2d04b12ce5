"""Throughput of the BASELINE.json configurations that run on one GPU (device-resident, ms per step after settling)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rokifd_b200
from rokifd_b200 import capi, chains as ch
CFG = [("C2 arm7, no contact", ch.world_c2(), 65536, 0),
       ("C2 arm7, no contact", ch.world_c2(), 262144, 0),
       ("C3 arm7 + penalty contact", ch.world_c3(), 262144, 700),
       ("C4 biped tree + volume contact (rkfd_volume)", ch.world_c4_volume(), 131072, 10),
       ("C4-shaped: biped tree (12 DoF) + penalty contact", ch.world_c4_penalty(), 131072, 300),
       ("C5 arm7 + rigid floor, MLCP", ch.world_c5(base_z=0.45, solver="MLCP"), 131072, 500),
       ("C5 arm7 + rigid floor, Vert QP", ch.world_c5(base_z=0.45, solver="Vert"), 131072, 500)]
for name, w, B, settle in CFG:
    q, qd, u = ch.sample_state(w, B, seed=20260418)
    if name.startswith("C4-shaped"):
        q[:, 2] = 0.45; q[:, 3:6] *= 0.1; q[:, 6:] *= 0.3
    elif name.startswith("C4"):
        q, qd, u = ch.sample_c4_standing(w, B, seed=20260418)
    fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
    st = torch.cuda.current_stream(); fd.batch_set_stream(st.cuda_stream)
    if settle: fd.update_n(settle)
    for _ in range(5): fd.update()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10 if w.solver == "Volume" else 50
    e0.record(st)
    for _ in range(n): fd.update()
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    frac = 0.0
    if w.nslot:
        a, _, _, _ = fd.batch_get_contact(); frac = float((a.sum(1) > 0).mean())
    print("%-34s B=%7d  %.3f ms/step  %.3e env-steps/s  envs in contact %.3f  bad %d" % (name, B, ms, B / (ms * 1e-3), frac, int((fd.batch_get_status() != 0).sum())), flush=True)
    fd.destroy()
