#!/usr/bin/env python
"""bench.py - env-steps/s of the batched FD+contact step (BASELINE.json metric) on N B200s.

Workload (config.workload): C3 of SURVEY.md section 8d - 7-DoF arm with an 8-vertex end-effector on a soft
floor (vertex penalty contact + joint friction + DC motors), 262,144 environments per GPU, synthetic
randomised initial states (numpy default_rng(20260418)), dt = 1e-3, Runge-Kutta-Gill.  One "step" =
one rkFDUpdate for every environment (5 dynamics evaluations + RKG combination, one kernel launch).

  python bench.py --gpus N --steps K --warmup W              the CUDA engine (one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...    the CPU restatement of the reference on all
                                                             host cores (the reference itself is un-buildable
                                                             here: ZEDA/ZM/Zeo/RoKi absent)
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 262144
SETTLE_STEPS = 700      # untimed: the arms fall from the random initial states until ~40-50 % of them touch the floor
METRIC = "env-steps/sec (batched FD+contact step)"
UNIT = "env-steps/s"
WORKLOAD = ("C3: arm7 (7-DoF, DC motors, joint friction) + 8-vertex penalty ground contact, %d envs/GPU, states after "
            "%d settle steps from the random initial states" % (B_PER_GPU, SETTLE_STEPS))
# algorithmic work per env-step (restated in DESIGN.md)
ALG_BYTES = 1054.0
ALG_FLOP = 20471.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the device runs the benchmark load (settle + warm-up + timed steps)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t0, self.t1, self.note = None, None, None

    def mark_begin(self):      # the device is under the benchmark's load from here ...
        self.t0 = time.time()

    def mark_end(self):        # ... to here: only samples in between are reported
        self.t1 = time.time()

    def _loaded(self):
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = self.t1 if self.t1 is not None else float("inf")
        return [r for (t, r) in self.rows if t0 + 0.15 <= t <= t1]      # +0.15 s: a sample reports the preceding interval

    def loaded_samples(self):
        return len(self._loaded())

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self._loaded():
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
               "samples": len(sm)}
        if self.note:
            out["note"] = self.note
        return out


def cpu_settled_state(world, ch, n_envs, seed=20260418, settle=SETTLE_STEPS):
    """Initial states of the CPU sample: the same synthetic states, settled by the oracle itself (untimed)."""
    from oracle import oracle as orc
    ow = orc.OracleWorld(world)
    q, qd, u = ch.sample_state(world, n_envs, seed=seed)
    if settle > 0:
        q, qd, _, _ = ow.batch_run(q, qd, u, nsteps=settle)
    return ow, q, qd, u


def cpu_baseline_run(ow, q, qd, u, n_steps, threads=0):
    """The oracle (CPU restatement of the reference's algorithm) on the host cores: env-steps/s."""
    t0 = time.perf_counter()
    _, _, _, used = ow.batch_run(q, qd, u, nsteps=n_steps, nthreads=threads)
    dt = time.perf_counter() - t0
    # batch_run also performs rkFDUpdateInit's evaluation per env: count it as 1/5 of a step
    return q.shape[0] * (n_steps + 0.2) / dt, used, dt


def run_reference(args, emit):
    """--impl reference: the reference's own CPU implementation of the path.  The reference cannot be
    compiled (its dependencies are absent), so this times the oracle port with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import rokifd_b200  # noqa: F401
    from rokifd_b200 import chains as ch
    world = ch.world_c3()
    cores = os.cpu_count() or 1
    n_envs = 512 * cores
    ow, q, qd, u = cpu_settled_state(world, ch, n_envs)
    for _ in range(args.warmup):
        cpu_baseline_run(ow, q, qd, u, 1)
    t0 = time.perf_counter()
    vals = []
    for _ in range(args.steps):
        v, used, _ = cpu_baseline_run(ow, q, qd, u, 1)
        vals.append(v)
    dt = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = "%d envs x 1 rkFDUpdate (+UpdateInit evaluation) per step, %d timed steps" % (n_envs, args.steps)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD, "cpu_sample": sample},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample,
                            "note": "CPU restatement of RoKi-FD's algorithm; reference un-buildable (ZEDA/ZM/Zeo/RoKi absent)"},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def _json_only_stdout():
    """Everything but the result line goes to stderr: native libraries (NCCL prints its version banner) write to file
    descriptor 1 directly, and the contract is ONE JSON line on stdout.  Returns the function that prints it."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real, (json.dumps(obj) + "\n").encode())
    return emit


def bind_to_gpu_numa_node(torch, device_index):
    """Pins this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host buffers of the
    end-to-end leg are allocated next to it (8 ranks on a two-socket box otherwise push half of their PCIe traffic
    across the socket interconnect).  Best effort: returns the node or None."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def main():
    emit = _json_only_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=B_PER_GPU, help="environments per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-upload-state", action="store_true", help="e2e leg: also re-send (q, q') host->device every step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, emit)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import rokifd_b200  # noqa: F401
    from rokifd_b200 import capi, chains as ch, multi

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank)     # before any pinned allocation: host buffers next to the GPU
    dist = None
    if world_size > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sampler = ClockSampler(local_rank)      # started early: the first nvidia-smi of a fresh box can take seconds to report
    sampler.start()
    world = ch.world_c3()
    B = args.envs
    q, qd, u = multi.rank_problem(world, ch, B, rank)     # weak scaling: every rank its own synthetic states
    fd, _ = capi.create_world(world, B=B, devices=[local_rank])
    fd.batch_set_state(q, qd)
    fd.batch_set_motor_input(u)
    fd.update_init()
    stream = torch.cuda.current_stream()
    fd.batch_set_stream(stream.cuda_stream)       # kernels run on torch's current stream: torch events bracket them

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") --------------------------------------------------------
    # settle (untimed, ~1 s of GPU load: also lets nvidia-smi deliver samples of the loaded device), then warm-up
    sampler.mark_begin()
    fd.update_n(SETTLE_STEPS // 2)
    torch.cuda.synchronize()
    fd.update_n(SETTLE_STEPS - SETTLE_STEPS // 2)
    for _ in range(args.warmup):
        fd.update()
    barrier()
    l0 = fd.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record(stream)
    for s in range(args.steps):
        fd.update()
        ev[s + 1].record(stream)
    barrier()
    launches = fd.launch_count - l0
    sampler.mark_end()
    if sampler.loaded_samples() == 0:
        # nvidia-smi delivered nothing inside the timed region (slow start): keep the same load on the device until it does
        t_wait = time.time()
        while sampler.loaded_samples() == 0 and time.time() - t_wait < 5.0:
            fd.update_n(100); torch.cuda.synchronize()
        sampler.mark_end()
        sampler.note = "no nvidia-smi sample fell inside the loaded region (settle + warm-up + timed steps); sampled under the same load right after it"
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(args.steps)]
    if os.environ.get("RKFD_BENCH_DEBUG"):
        print("rank %d: total %.3f ms, per-launch min/median/max %.3f/%.3f/%.3f ms" % (rank, total_ms, min(per_launch_ms), float(np.median(per_launch_ms)), max(per_launch_ms)), file=sys.stderr)
    total_ms = multi.max_over_ranks(dist, total_ms, device="cuda")      # the job's time is the slowest rank's
    value = multi.job_throughput(B, world_size, args.steps, total_ms)
    assert (fd.batch_get_status() == 0).all(), "non-finite accelerations in the timed run"
    ca, _, _, _ = fd.batch_get_contact()
    contact_frac, mean_active = float((ca.sum(1) > 0).mean()), float(ca.sum(1).mean())
    q, qd, _ = (np.ascontiguousarray(x) for x in fd.batch_get_state())     # settled states: inputs of the e2e leg

    # ---- end to end through the C-ABI with HOST buffers ("e2e") -----------------------------------------
    nq, nl = world.nq, world.nl
    hq = torch.from_numpy(q.copy()).pin_memory(); hqd = torch.from_numpy(qd.copy()).pin_memory()
    hu = torch.from_numpy(u.copy()).pin_memory()
    oq = torch.empty((B, nq), dtype=torch.float64).pin_memory(); oqd = torch.empty_like(oq).pin_memory()
    oqdd = torch.empty_like(oq).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_step():
        # every step: this step's inputs host->device from pinned memory, the step, its result device->host.
        # A step's inputs are what the reference's callers set between two rkFDUpdate calls - the motor inputs
        # (rkJointMotorSetInput, example/chain/arm_box_test.c:21); the simulator owns the state, as the reference's
        # rkFD does.  Its result is the state the caller reads back: q, q', q''.  --e2e-upload-state also re-sends
        # (q, q') every step (rkFDChainSetDis/SetVel before every update).
        # The calls are asynchronous (copy streams + staging ring), so the transfers of neighbouring steps overlap
        # the step kernel; rkFDBatchJoin + the closing event make the timed region cover all of them.
        if args.e2e_upload_state:
            fd.batch_set_state_async(hq.data_ptr(), hqd.data_ptr())
        fd.batch_set_motor_input_async(hu.data_ptr())
        fd.update()
        fd.batch_get_state_async(oq.data_ptr(), oqd.data_ptr(), oqdd.data_ptr())

    for _ in range(3):
        e2e_step()
    fd.batch_sync()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    fd.batch_join()
    e1.record(stream)
    fd.batch_sync()
    barrier()
    assert np.isfinite(oq.numpy()).all()
    e2e_value = multi.job_throughput(B, world_size, e2e_steps, multi.max_over_ranks(dist, e0.elapsed_time(e1), device="cuda"))
    h2d = B * ((2 * nq if args.e2e_upload_state else 0) + nl) * 8
    d2h = B * 3 * nq * 8

    # ---- roofline of the dominant kernel (rkfd_step_kernel: the only kernel of a step) ------------------
    hbm_peak, peak_src = load_peaks()
    k_ms = float(np.mean(per_launch_ms))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    ach_gbs = ALG_BYTES * B / (k_ms * 1e-3) / 1e9
    fp64_peak = capi.measure_fp64_tflops()
    ach_tf = ALG_FLOP * B / (k_ms * 1e-3) / 1e12

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD, "envs_per_gpu": B, "dt": world.dt, "integrator": "RKG", "solver": world.solver,
                      "settle_steps": SETTLE_STEPS, "envs_in_contact": contact_frac, "mean_active_vertices": mean_active,
                      "numa_node": numa, "l2": "per-GPU state (%.0f MB) is larger than the 126 MB L2" % (B * 8 * 140 / 1e6)},
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                   "io": "per step: motor inputs u[B][%d] host->device%s, rkFDUpdate, (q, q', q'')[B][%d] device->host; pinned host buffers, "
                         "asynchronous copies on their own streams" % (nl, " + state (q, q')" if args.e2e_upload_state else "", nq)},
           "gpu_launches": int(launches),
           "clocks": clocks,
           "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                        "traffic": traffic, "peak_source": peak_src, "kernel": "rkfd_step_kernel", "kernel_ms": k_ms,
                        "algorithmic_bytes_per_env_step": ALG_BYTES},
           "roofline_fp64": {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": ach_tf / fp64_peak if fp64_peak else None, "peak_source": "measured in this run (DFMA loop)",
                             "algorithmic_flop_per_env_step": ALG_FLOP}}
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_envs, n_steps = 1024 * cores, 400       # ~10-20 s of CPU work on every host core (settling included)
        ow, cq, cqd, cu = cpu_settled_state(world, ch, n_envs)
        v, used, dt = cpu_baseline_run(ow, cq, cqd, cu, n_steps)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": used, "kind": "port",
                               "sample": "%d envs x %d steps of the same workload, settled %d steps first (timed %.1f s)" % (n_envs, n_steps, SETTLE_STEPS, dt),
                               "note": "CPU restatement of RoKi-FD's algorithm; reference un-buildable (ZEDA/ZM/Zeo/RoKi absent)"}
    fd.destroy()
    if rank == 0:
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
