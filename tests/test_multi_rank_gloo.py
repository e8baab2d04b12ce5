"""The N>1 path on CPU: world_size-2 gloo.  bench.py runs one process per GPU; each rank owns a contiguous block
of environments (rokifd_b200.multi), steps it on its own and the job reports the slowest rank's time.  Here the
ranks step their shard with the host build of the kernel core (tests/hostsim: TEST-ONLY, the product has no CPU
path) and the gathered result must equal the single-process run bit for bit (shard-count invariance)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rokifd_b200  # noqa: F401
from rokifd_b200 import chains as ch, multi


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, total, nsteps, ret):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import rokifd_b200  # noqa: F401
    from rokifd_b200 import chains as ch, multi
    from hostsim_py import HostSim
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    w = ch.world_c3(base_z=0.1)
    q, qd, u = ch.sample_state(w, total, seed=5)
    lo, hi = multi.shard_range(total, rank, world_size)
    hs = HostSim(w, hi - lo)
    hs.set_state(q[lo:hi], qd[lo:hi], u[lo:hi]); hs.eval(ref=True); hs.step(nsteps)
    lq, lqd, lqdd = hs.get_state()
    gq = multi.gather_rows(dist, lq[:, :w.nq], total)
    gqdd = multi.gather_rows(dist, lqdd[:, :w.nq], total)
    # the slowest rank's time is the job's time
    tmax = multi.max_over_ranks(dist, 10.0 + rank)
    # weak-scaling problems differ per rank
    pq, _, _ = multi.rank_problem(w, ch, 4, rank)
    allp = [torch.zeros(pq.shape, dtype=torch.float64) for _ in range(world_size)]
    dist.all_gather(allp, torch.from_numpy(np.ascontiguousarray(pq)))
    dist.barrier()
    if rank == 0:
        ret["q"], ret["qdd"], ret["tmax"] = gq, gqdd, tmax
        ret["distinct"] = not np.array_equal(allp[0].numpy(), allp[1].numpy())
    dist.destroy_process_group()


def test_two_ranks_gloo_shard_invariance():
    from hostsim_py import HostSim, lib
    lib()                                  # build the harness once, before the ranks race for it
    total, nsteps, world_size = 13, 6, 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world_size, _free_port(), total, nsteps, ret), nprocs=world_size, join=True)
        gq, gqdd, tmax, distinct = ret["q"], ret["qdd"], ret["tmax"], ret["distinct"]
    w = ch.world_c3(base_z=0.1)
    q, qd, u = ch.sample_state(w, total, seed=5)
    hs = HostSim(w, total)
    hs.set_state(q, qd, u); hs.eval(ref=True); hs.step(nsteps)
    sq, _, sqdd = hs.get_state()
    assert np.array_equal(gq, sq[:, :w.nq]) and np.array_equal(gqdd, sqdd[:, :w.nq])
    assert tmax == 11.0 and distinct
    assert multi.job_throughput(262144, 8, 50, 25.0) == 262144 * 8 * 50 / 0.025


def test_shard_ranges_cover_the_batch():
    for total in (1, 7, 262144, 1000003):
        for world in (1, 2, 3, 8):
            edges = [multi.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1
