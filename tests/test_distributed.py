"""world_size-2 gloo tests of the N > 1 host logic (sharding + the only collective)."""
import os
import subprocess
import sys

import numpy as np

from kessler_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from kessler_b200 import distributed as D
rank, local, world = D.init(backend="gloo")
assert world == 2 and dist.get_backend() == "gloo"
n_total = 1001
lo, hi = D.shard_range(n_total, rank, world)
rng = np.random.default_rng(7)
d_min = rng.uniform(10.0, 9000.0, n_total); i_conj = np.where(d_min < 5000.0, rng.integers(0, 600000, n_total), -1)
# what ksl_screen_summary would give for this rank's shard
mine = slice(lo, hi)
counts = torch.zeros(8, dtype=torch.int64)
counts[0] = hi - lo; counts[2] = int((i_conj[mine] >= 0).sum()); counts[3] = int((i_conj[mine] > 0).sum()); counts[4] = int((d_min[mine] < 70.0).sum())
moments = torch.tensor([hi - lo, d_min[mine].sum(), (d_min[mine] ** 2).sum(), d_min[mine].min()], dtype=torch.float64)
D.allreduce_summary(counts, moments)
assert counts[0].item() == n_total and counts[2].item() == int((i_conj >= 0).sum()) and counts[4].item() == int((d_min < 70.0).sum())
assert moments[0].item() == n_total and abs(moments[1].item() - d_min.sum()) < 1e-6 and moments[3].item() == d_min.min()
c = D.allreduce_counts(torch.tensor([rank + 5], dtype=torch.int64)); assert c.item() == 11
m = torch.arange(28, dtype=torch.float64) * (rank + 1)
g = D.allgather_moments(m); assert torch.equal(g, torch.arange(28, dtype=torch.float64) * 3)
assert D.max_over_ranks(1.5 + rank) == 2.5
# the packed exchange itself: exact integer sums, rank-order FP64 sums (bit patterns survive the int64 transport), minima
ints = torch.tensor([2 ** 40 + rank, -3 * rank], dtype=torch.int64)
sums = torch.tensor([0.1 * (rank + 1), 1e300 * (1 - rank), -0.0], dtype=torch.float64)
mins = torch.tensor([5.0 - rank, float("inf") if rank else 7.0], dtype=torch.float64)
i2, s2, m2 = D.exchange_packed(ints, sums, mins)
assert i2.tolist() == [2 ** 41 + 1, -3] and s2[0].item() == 0.1 + 0.2 and s2[1].item() == 1e300 and m2.tolist() == [4.0, 7.0]
assert ints.tolist() == [2 ** 40 + rank, -3 * rank]            # inputs untouched
D.barrier()
print("rank", rank, "ok")
'''


def test_shard_range_tiles():
    for n in (0, 1, 7, 1000, 1001, 10 ** 8 + 3):
        for world in (1, 2, 3, 8):
            edges = [D.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_is_identity():
    import torch
    c, m = torch.arange(8), torch.arange(4, dtype=torch.float64)
    c2, m2 = D.allreduce_summary(c.clone(), m.clone())
    assert torch.equal(c, c2) and torch.equal(m, m2)
    assert D.max_over_ranks(3.25) == 3.25


def test_two_rank_gloo_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
