"""Tuning aid: the re-sort key of worlds under the Vert / Volume solver - contact count only (RKFD_NO_WORK_SORT=1) against the
work class of the last rigid solve (default).  Times settled steps with the sorts inside.  Usage: exp_work_sort.py C5-vert|C4 [steps]"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rokifd_b200
from rokifd_b200 import capi, chains as ch
name = sys.argv[1] if len(sys.argv) > 1 else "C5-vert"; nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
for off in (1, 0, 1, 0):
    if off: os.environ["RKFD_NO_WORK_SORT"] = "1"
    else: os.environ.pop("RKFD_NO_WORK_SORT", None)
    if name == "C4":
        w = ch.world_c4_volume(); B, settle = 131072, 48
        q, qd, u = ch.sample_c4_standing(w, B, seed=3)
    else:
        w = ch.world_c5(base_z=0.45, solver="Vert" if name == "C5-vert" else "MLCP"); B, settle = 131072, 500
        q, qd, u = ch.sample_state(w, B, seed=20260418)
    fd, _ = capi.create_world(w, B=B)
    fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
    st = torch.cuda.current_stream(); fd.batch_set_stream(st.cuda_stream)
    fd.update_n(settle)
    for _ in range(4): fd.update()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(nsteps): fd.update()
    e1.record(st); torch.cuda.synchronize()
    gq, gqd, _ = fd.batch_get_state()
    print("%s work-sort %s: %.4f ms/step over %d steps, checksum %.12e" % (name, "off" if off else "on ", e0.elapsed_time(e1) / nsteps, nsteps, float(np.abs(gq).sum() + np.abs(gqd).sum())), flush=True)
    fd.destroy()
