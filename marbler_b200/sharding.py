"""Multi-GPU layout: environments are independent, so a job of B envs is split into contiguous
ranges, one per rank / GPU, with NO data-path collective.  Rank g owns global env ids
[start, start+count); the ids seed the per-env Philox streams, so a sharded run reproduces the
single-GPU run env for env.  The only communication is the small episode-statistics all-reduce."""
import os

import torch
import torch.distributed as dist

from ._lib import NUM_STATS, STAT_NAMES


def shard_range(total_envs, rank, world_size):
    """Contiguous, balanced split: the first (total % world) ranks get one extra env."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(int(total_envs), int(world_size))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def init_distributed(backend=None):
    """Join the torchrun rendezvous if one is described by the environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def bind_to_gpu_numa(local_rank):
    """Restrict this process to the CPUs NVML reports as local to GPU `local_rank` (its NUMA node).  Call it before
    anything allocates pinned host memory: pages are placed on the node of the thread that first touches them, and a
    device->host copy into the far node crosses the inter-socket link (the e2e path of an 8-GPU box is bound by
    exactly that).  Returns a description for the bench line, or None when NVML / the mask is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[local_rank]) if visible and visible.split(",")[local_rank].isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"gpu": index, "cpus": len(cpus), "first": min(cpus), "last": max(cpus)}
    except Exception:
        return None


def allreduce_stats(stats):
    """SUM all ranks' statistics vectors (f64 [NUM_STATS]); NCCL over NVLink on GPUs, gloo on CPU."""
    assert stats.numel() == NUM_STATS
    out = stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def summarize(stats):
    """Named, derived episode statistics from a (reduced) statistics vector."""
    v = [float(x) for x in stats.detach().cpu().tolist()]
    d = dict(zip(STAT_NAMES, v))
    ep = max(d["episodes"], 1.0)
    d["return_mean"] = d["return_sum"] / ep
    d["length_mean"] = d["length_sum"] / ep
    d["qp_iters_per_solve"] = d["qp_iterations"] / max(d["qp_solves"], 1.0)
    return d
