"""Host-side sharding logic and the world_size-2 reduction path on CPU (gloo), as the N>1 bench uses it."""
import os
import socket
import subprocess
import sys

from denseflow_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_pairs_cover_every_pair_once():
    for n_frames in (0, 1, 2, 7, 64, 300):
        for step in (1, -1, 3, -5):
            for world in (1, 2, 3, 8):
                m = max(n_frames - abs(step), 0)
                seen = []
                for r in range(world):
                    first, n, f0, nf = shard.shard_pairs(n_frames, step, r, world)
                    seen += list(range(first, first + n))
                    if n:
                        assert nf == n + abs(step) and f0 + nf <= n_frames  # |step| frames of overlap at the edge
                assert seen == list(range(m))


def test_shard_list_round_robin():
    items = list(range(10))
    got = sorted(sum((shard.shard_list(items, r, 3) for r in range(3)), []))
    assert got == items


_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
from denseflow_b200 import shard
rank, local_rank, world = shard.init("gloo")
assert world == 2
shard.barrier()
t = shard.all_max(1.5 + rank)          # slowest rank wins
s = shard.all_sum(10 * (rank + 1))     # flow counters add up
first, n, f0, nf = shard.shard_pairs(64, 1, rank, world)
tot = shard.all_sum(n)
if rank == 0:
    print("RESULT", t, s, tot)
import torch.distributed as dist
dist.destroy_process_group()
"""


def test_world_size_2_gloo_reduction(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT")][0].split()
    assert float(line[1]) == 2.5 and float(line[2]) == 30.0 and float(line[3]) == 63.0
