"""Tuning aid: the biped with both soles on the RIGID floor (contacts on two links of one tree)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rokifd_b200
from rokifd_b200 import capi, chains as ch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for solver in ("MLCP",):
    w = ch.World(chains=[ch.biped(), ch.floor()], contact_info=[ch.ContactInfo("ground", "body", "rigid", K=1000.0, L=0.001, SF=0.5, KF=0.3)], solver=solver)
    q, qd, u = ch.sample_state(w, B, seed=3); q[:, 2] = 0.44; q[:, 3:6] *= 0.1; q[:, 6:] *= 0.3
    fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
    st = torch.cuda.current_stream(); fd.batch_set_stream(st.cuda_stream)
    for n in (5, 20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(n): fd.update()
        e1.record(st); torch.cuda.synchronize()
        a = fd.batch_get_contact()[0]
        print("%s B=%d: %.3f ms/step  %.3e env-steps/s  envs in contact %.3f mean verts %.1f bad %d" % (solver, B, e0.elapsed_time(e1) / n, B * n / (e0.elapsed_time(e1) * 1e-3), (a.sum(1) > 0).mean(), a.sum(1).mean(), int((fd.batch_get_status() != 0).sum())), flush=True)
    fd.destroy()
