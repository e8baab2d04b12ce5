"""Throughput of C4 (BASELINE.json configs[3]): the legged tree on the rigid floor with volume-based contact."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rokifd_b200
from rokifd_b200 import capi, chains as ch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
w = ch.world_c4_volume()
q, qd, u = ch.sample_c4_standing(w, B, seed=3)
fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
st = torch.cuda.current_stream(); fd.batch_set_stream(st.cuda_stream)
for n in [int(x) for x in os.environ.get('STEPS', '3,10,20').split(',')]:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n): fd.update()
    e1.record(st); torch.cuda.synchronize()
    a = fd.batch_get_contact()[0]
    prs = a.reshape(B, -1, 8).any(2)
    print("C4 volume B=%d: %.3f ms/step  %.3e env-steps/s  envs in contact %.3f mean pairs %.2f bad %d" % (B, e0.elapsed_time(e1) / n, B * n / (e0.elapsed_time(e1) * 1e-3), prs.any(1).mean(), prs.sum(1).mean(), int((fd.batch_get_status() != 0).sum())), flush=True)
fd.destroy()
