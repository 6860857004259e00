import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import graphflow_b200
from bench import make_inputs
n, C = 32, 64
for Be in (96, 256):
    ctx = graphflow_b200.Context(0)
    T, adj, gout = make_inputs(Be, n, C, 1, torch.device("cuda", 0))
    hT, hA, hG = T.cpu().pin_memory(), adj.cpu().pin_memory(), gout.cpu().pin_memory()
    hO = torch.empty((Be, n, n, 18 * C)).pin_memory(); hGT = torch.empty((Be, n, n, n, C)).pin_memory()
    ctx.contract18_forward_backward_host(hT, hA, hG, hO, hGT)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): ctx.contract18_forward_backward_host(hT, hA, hG, hO, hGT)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("stage_mib", os.environ.get("CCN_STAGE_MIB"), "batch", Be, "e2e contractions/s", 3 * Be / dt, flush=True)
    ctx.close(); del T, adj, gout, hT, hG, hO, hGT
