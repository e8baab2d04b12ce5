"""Tuning aid: a fixed-base serial arm with arbitrary constant frames (random 7-joint arm, DC motors): generic
table-driven kernel against the rolled general-frame specialisation."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rokifd_b200
from rokifd_b200 import capi, chains as ch
rng = np.random.default_rng(3)
w = ch.World(chains=[ch.random_chain(rng, 8, jtypes=("revolute",), motors=True)])
B = 262144
q = rng.uniform(-1.5, 1.5, (B, w.nq)); qd = rng.uniform(-2, 2, (B, w.nq)); u = rng.uniform(-6, 6, (B, w.nl))
res = {}
for spec in ("0", "10"):
    os.environ["RKFD_SPEC"] = spec
    fd, _ = capi.create_world(w, B=B); fd.batch_set_state(q, qd); fd.batch_set_motor_input(u); fd.update_init()
    st = torch.cuda.current_stream(); fd.batch_set_stream(st.cuda_stream)
    for _ in range(5): fd.update()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(30): fd.update()
    e1.record(st); torch.cuda.synchronize()
    res[spec] = fd.batch_get_state()
    print("spec %s: %.3f ms/step  %.3e env-steps/s" % (spec, e0.elapsed_time(e1) / 30, B * 30 / (e0.elapsed_time(e1) * 1e-3)), flush=True)
    fd.destroy()
print("max |dq| between the two kernels: %.2e" % np.abs(res["0"][0] - res["10"][0]).max())
