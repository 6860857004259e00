#!/usr/bin/env python
"""`ccn_allreduce_grads` (the C-ABI's gradient all-reduce for C++ callers) on N GPUs: every rank creates a raw NCCL
communicator (ncclGetUniqueId / ncclCommInitRank through ctypes, the id exchanged over a gloo group), fills a buffer with
rank + 1, calls the entry point and checks the sum.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/nccl_abi_probe.py"""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphflow_b200  # noqa: E402

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")
torch.cuda.set_device(local)
nccl = ctypes.CDLL("libnccl.so.2")


class UniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_char * 128)]


uid = UniqueId()
if rank == 0:
    assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
t = torch.tensor(list(bytes(uid)), dtype=torch.uint8)
dist.broadcast(t, 0)
ctypes.memmove(ctypes.byref(uid), bytes(t.tolist()), 128)
comm = ctypes.c_void_p()
nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
assert nccl.ncclCommInitRank(ctypes.byref(comm), world, uid, rank) == 0
ctx = graphflow_b200.Context(local)
count = 73792                                     # K_l + b_l of one level at C = 64
buf = torch.full((count,), float(rank + 1), device="cuda")
stream = torch.cuda.current_stream().cuda_stream
ctx._rc(ctx.lib.ccn_allreduce_grads(ctx.h, comm, buf.data_ptr(), count, stream))
torch.cuda.synchronize()
want = world * (world + 1) / 2
ok = bool((buf == want).all().item())
# no communicator -> a clean error, not a crash
bad = ctx.lib.ccn_allreduce_grads(ctx.h, None, buf.data_ptr(), count, stream)
nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
nccl.ncclCommDestroy(comm)
flags = torch.tensor([int(ok), int(bad != 0)])
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"n_gpus": world, "allreduce_sum_ok": bool(flags[0].item()), "null_comm_rejected": bool(flags[1].item()),
                      "count": count}))
dist.destroy_process_group()
sys.exit(0 if flags.min().item() == 1 else 1)
