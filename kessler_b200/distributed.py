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


def _world():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def exchange_packed(ints, sums, mins=None):
    """The one collective of this package: every rank contributes integer counts (summed exactly), FP64 partial
    sums and FP64 minima, packed into ONE int64 buffer (the doubles travel as their bit patterns) and exchanged with a
    single all-gather -- <= 1 KB per rank, latency-bound.  Every rank then reduces the same [world, k] table with the
    same three kernels (sum / sum / min over the rank axis), so the result is bit-identical on all ranks and does not
    depend on a reduction tree or on message arrival order.  Returns new tensors (ints_total, sums_total,
    mins_total | None)."""
    if _world() == 1:
        return ints, sums, mins
    ki, ks = ints.numel(), sums.numel()
    pieces = [ints.reshape(-1).to(torch.int64), sums.reshape(-1).contiguous().view(torch.int64)]
    if mins is not None:
        pieces.append(mins.reshape(-1).contiguous().view(torch.int64))
    buf = torch.cat(pieces)
    table = torch.empty((_world(), buf.numel()), dtype=torch.int64, device=buf.device)
    try:
        dist.all_gather_into_tensor(table, buf)
    except (RuntimeError, NotImplementedError):          # a backend without the flat form
        parts = [torch.empty_like(buf) for _ in range(_world())]
        dist.all_gather(parts, buf)
        table = torch.stack(parts)
    ints_out = table[:, :ki].sum(dim=0)
    sums_out = table[:, ki:ki + ks].contiguous().view(torch.float64).sum(dim=0)
    mins_out = table[:, ki + ks:].contiguous().view(torch.float64).amin(dim=0) if mins is not None else None
    return ints_out.reshape(ints.shape), sums_out.reshape(sums.shape), (mins_out.reshape(mins.shape) if mins is not None else None)


def allreduce_summary(counts, moments):
    """In-place all-reduce of a screen summary: counts int64[8] SUM (bit-exact, associative);
    moments f64[4] = (n_valid, sum d, sum d^2) SUM and min d MIN.  One collective."""
    if _world() == 1:
        return counts, moments
    c, s, m = exchange_packed(counts, moments[:3], moments[3:])
    counts.copy_(c)
    moments[:3] = s
    moments[3:] = m
    return counts, moments


def allreduce_counts(counts):
    """SUM all-reduce of an integer count tensor (collision counts): exact."""
    if _world() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def allgather_moments(moments):
    """Shifted moments about a shared reference state add across ranks (one fixed-order reduction, identical on every rank)."""
    if _world() == 1:
        return moments
    _, s, _ = exchange_packed(torch.zeros(0, dtype=torch.int64, device=moments.device), moments)
    return s


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
