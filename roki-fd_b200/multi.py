"""One process per GPU: how a job of `world_size` ranks divides the environments and reports its time.

The path shards with no exchange step: environments never interact, every rank owns a
contiguous block of them and steps it with its own engine; the only communication is the barrier around the
timed region and the max-over-ranks of the measured time.  `bench.py` uses these helpers under NCCL; the CPU
tests run them over gloo with world_size 2."""
import numpy as np


def shard_range(total_envs, rank, world_size):
    """Contiguous block of a global batch owned by `rank`: env e -> rank floor(e * world_size / total_envs)
    (the same rule the in-process sharding of rkFDBatchSetDevices uses, rkfd_engine.cu)."""
    lo = (total_envs * rank) // world_size
    hi = (total_envs * (rank + 1)) // world_size
    return lo, hi


def rank_problem(world, chains_mod, envs_per_rank, rank, seed=20260418):
    """Weak scaling: every rank draws its own synthetic states (seed + rank), `envs_per_rank` environments."""
    return chains_mod.sample_state(world, envs_per_rank, seed=seed + rank)


def max_over_ranks(dist, value, device="cpu"):
    """The job's time is the slowest rank's: all-reduce(MAX) of a scalar (no-op without a process group)."""
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(envs_per_rank, world_size, steps, max_ms):
    """Whole-job env-steps/s: the units all ranks processed divided by the slowest rank's time."""
    return envs_per_rank * world_size * steps / (max_ms * 1e-3)


def gather_rows(dist, local, total_rows):
    """all_gather of row blocks of unequal height (used by the tests to compare a sharded run with the
    single-process run)."""
    import torch
    world = dist.get_world_size()
    sizes = [shard_range(total_rows, r, world) for r in range(world)]
    hmax = max(hi - lo for lo, hi in sizes)
    pad = np.zeros((hmax,) + local.shape[1:], dtype=np.float64)
    pad[:local.shape[0]] = local
    out = [torch.zeros(pad.shape, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(out, torch.from_numpy(pad))
    return np.concatenate([o.numpy()[:hi - lo] for o, (lo, hi) in zip(out, sizes)], axis=0)
