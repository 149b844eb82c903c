"""Video-list dispatch, host side (no GPU): the shared work queue hands every list index out exactly once — inside one
process and across two processes (world_size 2, gloo) — and dfb_run_list validates its arguments and fails loudly
without a device (the reference's list handling: /root/reference/tools/denseflow.cpp:54-81)."""
import os
import socket
import subprocess
import sys
import threading

import numpy as np
import pytest

from denseflow_b200 import _lib, listrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_queue_hands_out_each_index_once_across_threads():
    q = listrun.WorkQueue("/dfb_test_q_%d" % os.getpid(), create=True)
    try:
        got = [[] for _ in range(8)]

        def pull(k):
            while True:
                i = q.next()
                if i >= 5000:
                    return
                got[k].append(i)

        ts = [threading.Thread(target=pull, args=(k,)) for k in range(8)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert sorted(sum(got, [])) == list(range(5000))
        q.reset()
        assert q.next() == 0
    finally:
        q.close()


_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from denseflow_b200 import shard, listrun
rank, local_rank, world = shard.init("gloo")
name = "/dfb_test_q_" + os.environ["MASTER_PORT"]
q = listrun.WorkQueue(name, create=(rank == 0)) if rank == 0 else None
shard.barrier()                      # rank 0 has created and zeroed the queue
if q is None:
    q = listrun.WorkQueue(name, create=False)
N = 2000
mine = []
while True:
    i = q.next()
    if i >= N:
        break
    mine.append(i)
flags = torch.zeros(N, dtype=torch.int32)
flags[mine] = 1
dist.all_reduce(flags)               # every index claimed by exactly one rank
shard.barrier()
q.close()
if rank == 0:
    print("RESULT", int(flags.min()), int(flags.max()), int(flags.sum()))
dist.destroy_process_group()
"""


def test_queue_is_shared_by_two_processes_gloo(tmp_path):
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
    assert line[1:] == ["1", "1", "2000"]


def test_run_list_argument_checks_and_no_device():
    clip = [np.zeros((32, 48), np.uint8)] * 3
    with pytest.raises(RuntimeError, match="step must be non-zero"):
        listrun.run_list("tvl1", [0], [clip], step=0)
    with pytest.raises(RuntimeError, match="one size"):
        listrun.run_list("tvl1", [0], [[np.zeros((32, 48), np.uint8), np.zeros((32, 40), np.uint8)]])
    if _lib.load().dfb_device_count() == 0:
        with pytest.raises(RuntimeError, match="no CUDA device"):
            listrun.run_list("tvl1", [0], [clip])
    assert listrun.run_list("tvl1", [0], [])["clips"] == 0  # an empty list is not an error (all videos already .done)
