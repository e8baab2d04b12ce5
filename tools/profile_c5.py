"""C5 (rigid floor) for N steps so that a profiler can capture the step kernel with rigid contacts active:
ncu -k regex:rkfd_step_kernel -s <N> -c 1 python tools/profile_c5.py <N> [MLCP|Vert] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rokifd_b200
from rokifd_b200 import capi, chains as ch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
solver = sys.argv[2] if len(sys.argv) > 2 else "MLCP"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
w = ch.world_c5(base_z=0.45, solver=solver)
q, qd, u = ch.sample_state(w, B, seed=20260418)
fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
for _ in range(n + 3):
    fd.update()
fd.batch_sync()
a, t, r, f = fd.batch_get_contact()
print("envs in contact %.3f, mean active vertices %.2f" % ((a.sum(1) > 0).mean(), a.sum(1).mean()))
fd.destroy()
