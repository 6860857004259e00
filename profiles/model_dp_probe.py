#!/usr/bin/env python
"""Data-parallel model step (the reference's Threaded_BatchLearn scheme, SMP_beta.h:697-739, with GPUs as the replicas):
the batch of graphs is sharded contiguously over the ranks, every rank runs SMP_beta forward+backward on its shard and
the parameter gradients are summed with ONE NCCL all-reduce.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/model_dp_probe.py [graphs_per_gpu] [V] [L] [C]
Rank 0 prints one JSON line; time = max over ranks (CUDA events)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphflow_b200 import shard  # noqa: E402
from graphflow_b200.model import SMPBetaB200  # noqa: E402
from tests.util import molecular_adjacency  # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 128
V = int(sys.argv[2]) if len(sys.argv) > 2 else 24
L = int(sys.argv[3]) if len(sys.argv) > 3 else 3
C = int(sys.argv[4]) if len(sys.argv) > 4 else 32
VER8 = "ver8" in sys.argv[5:]  # SMP_2D_ver8 (BASELINE config 4's model) instead of SMP_beta
F, D = 5, 2
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total = per_gpu * world
rng = np.random.default_rng(0)  # every rank generates the same global batch and keeps its shard
graphs = []
for _ in range(total):
    adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
    graphs.append((adj, np.eye(F)[rng.integers(0, F, V)]))
lo, hi = shard.contiguous_shard(total, world, rank)
model = SMPBetaB200(L, C, F, D, device=local, k_transposed=VER8)
model.set_flat_params(np.random.default_rng(1).uniform(-1, 1, model.num_params()) * 0.02)  # identical replicas
tb = model.tables(graphs[lo:hi])
targets = [float(V)] * (hi - lo)


GRAPH = "--graph" in sys.argv  # forward+backward replayed from one CUDA graph; the all-reduce stays outside it
graphed = model.capture_step(tb, targets) if GRAPH else None


def step():
    gf, loss, grads = graphed() if GRAPH else model.forward_backward(tb, targets)
    shard.allreduce_gradients([grads])
    return loss, grads


for _ in range(2):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 5
ev0.record()
for _ in range(steps):
    loss, grads = step()
ev1.record()
torch.cuda.synchronize()
t = torch.tensor([ev0.elapsed_time(ev1) / steps], device="cuda", dtype=torch.float64)
c = torch.tensor([float(tb.contractions)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
if rank == 0:
    ms = t.item()
    print(json.dumps({"workload": "%s data-parallel step, L=%d C=%d, %d graphs x %d vertices per GPU" % ("SMP_2D_ver8" if VER8 else "SMP_beta", L, C, per_gpu, V),
                      "n_gpus": world, "ms_per_step": ms, "graphs_per_s": total / (ms * 1e-3), "contractions_per_s": c.item() / (ms * 1e-3),
                      "allreduce_floats": int(grads.numel()), "grad_checksum": float(grads.double().abs().sum()),
                      "cuda_graph": GRAPH}))
if world > 1:
    dist.destroy_process_group()
