"""Host-side sharding logic and the gradient all-reduce, world_size 2 over gloo on CPU (the N > 1 path of bench.py
and of a data-parallel training step; SMP_beta.h:697-739 is the reference scheme)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphflow_b200 import shard


def test_contiguous_shard_covers_batch():
    for batch in (0, 1, 7, 512, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard.contiguous_shard(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.contiguous_shard(4, 2, 2)


def test_balanced_shards_ragged():
    rng = np.random.default_rng(0)
    sizes = rng.integers(1, 33, 1000)
    parts = shard.balanced_shards(sizes, 64, 8)
    allidx = np.sort(np.concatenate(parts))
    assert np.array_equal(allidx, np.arange(1000))
    loads = np.array([shard.instance_cost(sizes[p], 64).sum() for p in parts])
    assert loads.max() / loads.mean() < 1.01
    for p in parts:
        assert np.all(np.diff(sizes[p]) >= 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    """Each rank: its contiguous shard of 6 small instances -> oracle forward + mix backward (a K gradient) ->
    all-reduce.  The summed gradient must equal the single-process sum."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    from tests.util import random_instance

    orc = pyoracle.COracle("f64")
    rng = np.random.default_rng(5)
    n, C, P, B = 5, 3, 2, 6
    insts = [random_instance(n, C, rng) for _ in range(B)]
    K = rng.uniform(-1, 1, (18 * C, P))
    gY = rng.uniform(-1, 1, (n * n, P))
    lo, hi = shard.contiguous_shard(B, world, rank)
    gK = torch.zeros((18 * C, P), dtype=torch.float64)
    gb = torch.zeros((P,), dtype=torch.float64)
    for i in range(lo, hi):
        X = orc.contract18_forward(insts[i][0].astype(np.float64), insts[i][1].astype(np.float64)).reshape(n * n, 18 * C)
        _, gW = orc.matmul_backward(X, K, gY)
        gK += torch.from_numpy(gW)
        gb += torch.from_numpy(gY.sum(0))
    shard.allreduce_gradients([gK, gb])
    if rank == 0:
        torch.save({"gK": gK, "gb": gb}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo(tmp_path):
    port = _free_port()
    outs = []
    for world in (1, 2):
        out = str(tmp_path / ("w%d.pt" % world))
        mp.spawn(_worker, args=(world, port + world, out), nprocs=world, join=True)
        outs.append(torch.load(out))
    assert torch.allclose(outs[0]["gK"], outs[1]["gK"], rtol=0, atol=1e-10)
    assert torch.allclose(outs[0]["gb"], outs[1]["gb"], rtol=0, atol=1e-12)
