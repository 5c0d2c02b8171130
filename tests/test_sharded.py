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


def test_view_partition_properties():
    """hulo_partition_views: contiguous, complete, balanced to within one view."""
    import numpy as np
    from sfmlocalization_b200.gpu import partition_views
    rng = np.random.default_rng(3)
    for n, world in [(0, 4), (1, 8), (5, 8), (100, 1), (100, 8), (1000, 3), (17, 17)]:
        rows = rng.integers(0, 5000, n)
        b = partition_views(rows, world)
        assert b[0] == 0 and b[-1] == n and (np.diff(b) >= 0).all()
        if n >= 4 * world:
            share = np.array([rows[b[r]:b[r + 1]].sum() for r in range(world)])
            assert np.abs(share - rows.sum() / world).max() <= rows.max()
    assert partition_views([7, 7, 7, 7], 2).tolist() == [0, 2, 4]
    assert partition_views([1, 1, 1, 100], 2).tolist() == [0, 3, 4]


def test_view_sharded_query_world2_gloo():
    launch("views-gloo", 2)


def test_view_sharded_query_world3_gloo():
    launch("views-gloo", 3)


@pytest.mark.gpu
def test_view_sharded_localize_nccl_two_gpus():
    from sfmlocalization_b200 import _lib
    if _lib.load().hulo_device_count() < 2:
        pytest.skip("needs two GPUs")
    launch("views-nccl", 2)
