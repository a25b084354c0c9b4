"""Sample-level data parallelism: one process per GPU, contiguous sample ranges per rank,
no data-path collective; the only exchange is the final all-reduce of the integer counts
and FP64 moments (<= 256 B) -- SURVEY.md Sec. 8e.  Works with the ``nccl`` backend on GPUs
and with ``gloo`` on CPU (tests/test_distributed.py, world_size 2)."""
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process if unset)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def init(backend=None):
    """Initialise torch.distributed from the environment when WORLD_SIZE > 1."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_range(n_total, rank, world):
    """Contiguous sample range [lo, hi) of `rank`: sizes differ by at most one, ranges tile [0, n)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_summary(counts, moments):
    """In-place all-reduce of a screen summary: counts int64[8] SUM (bit-exact, associative);
    moments f64[4] = (n_valid, sum d, sum d^2) SUM and min d MIN."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts, moments
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    head, tail = moments[:3].contiguous(), moments[3:].contiguous()
    dist.all_reduce(head, op=dist.ReduceOp.SUM)
    dist.all_reduce(tail, op=dist.ReduceOp.MIN)
    moments[:3] = head
    moments[3:] = tail
    return counts, moments


def allreduce_counts(counts):
    """SUM all-reduce of an integer count tensor (collision counts): exact."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def allgather_moments(moments):
    """Shifted moments about a shared reference state add across ranks.  For a result that does
    not depend on reduction order, gather the per-rank partials and sum them in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return moments
    parts = [torch.empty_like(moments) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, moments.contiguous())
    out = parts[0].clone()
    for p in parts[1:]:
        out += p
    return out


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device=None):
    """max of a python float over ranks (device timing is reported as the max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
