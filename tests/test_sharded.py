"""Row-sharded database (multi-GPU mode): world_size-2 gloo run of the exchange on CPU, and
the NCCL path itself when the box has at least two GPUs."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(mode, world):
    port = str(free_port())
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=port, OMP_NUM_THREADS="4")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "sharded_worker.py"), mode], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "OK" in o


def test_exchange_and_merge_world2_gloo():
    launch("gloo", 2)


def test_exchange_and_merge_world3_gloo():
    launch("gloo", 3)


@pytest.mark.gpu
def test_knn2_sharded_nccl_two_gpus():
    from sfmlocalization_b200 import _lib
    if _lib.load().hulo_device_count() < 2:
        pytest.skip("needs two GPUs (the single-GPU box covers the merge through hulo_merge_top2)")
    launch("nccl", 2)
