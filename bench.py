#!/usr/bin/env python
"""bench.py -- contractions/sec of the CCN hot path (StackTensor3D + RisiContraction_18 forward + backward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json north_star / SURVEY.md section 8d, op-level): a batch of independent synthetic instances with
N = 32 vertices and C = 64 channels, neighbour tensors ~ U[-1,1) handed over as a table of per-vertex slab pointers
(RisiContraction_18::add_tensor: the StackTensor3D copy is fused into the read), molecular-like adjacency I + S (nnz ~ 102),
gout ~ U[-1,1).  One *step* = forward then backward (beta = 0) of the whole per-GPU batch.  `value` = instances that
completed forward+backward per second over all ranks, inputs resident in HBM.  Ranks are independent (no data-path
collective): weak scaling.

`e2e` is the same unit of work measured through the host-buffer C-ABI call a model with host-resident inputs makes per
batch, ccn_gather_levels_readout_forward_backward_host with 4 levels (BASELINE config 3's depth): per contraction instance it
ALSO does the promotion (MatTensorMul + TensorMatMul as a gather), the feature mix (MatMul . K + bias + leaky-ReLU), all
their backward passes (SMP_beta.h:588-616) and, once per batch, the read-out head + loss (SMP_beta.h:620-639), i.e. strictly
more work than the op the reference arm times; everything between the first level's input and the loss stays on the device,
so only the first level's input tensors (n^2 C per vertex, not the n^3 C stack), the index tables, parameters and targets
cross PCIe, and the losses, the input gradient and the parameter gradients come back.
`e2e_levels_only` is the same stack without the read-out (the last Z comes down, its gradient goes up);
`e2e_one_level` is the single-level call (every level's activations cross PCIe).  `e2e_op` keeps round 1's figure (the
stacked T itself crossing PCIe through ccn_contract18_forward_backward_host) with pinned, pageable and cudaHostRegister'ed
caller arrays; `host_copy_ceiling` is what plain pinned copies of the same byte volumes achieve on this box; `contract50` is BASELINE config 5
(RisiContraction_50, N=48, C=128, batch 256) and `family` RisiContraction_4 / _10 at the headline shape.

--impl reference times the reference's own CPU implementation (oracle/_ref: the unmodified GraphFlow_32bit
RisiContraction_18 behind oracle/ref_shim.cpp, replica-parallel over all host cores like SMP_beta.h:697-739) on a
bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# stdout must carry exactly ONE JSON line, but native libraries print there too (NCCL's "NCCL version ..." banner when the
# box sets NCCL_DEBUG=VERSION).  Keep a private handle on the real stdout for the result line and point file descriptor 1
# at stderr for everything else.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_VERT, CHANNELS = 32, 64
METRIC = "CCN-2D RisiContraction_18 fwd+bwd contractions/sec"
UNIT = "contractions/s"


def algorithmic_bytes(n, C):
    """8*(N^3 C + 18 N^2 C + N^2) bytes per contraction forward+backward (fp32)."""
    return 8 * (n ** 3 * C + 18 * n * n * C + n * n)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_bytes(kernel, n, C, batch):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum per instance at N=32, C=64), scaled to the instances of one launch."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    if d.get("N") != n or d.get("C") != C or kernel not in d:
        return None
    return d[kernel]["bytes_per_instance"] * batch


# ---- synthetic inputs (numpy only: also used by the CPU arm; mirrors tests/util.py so that bench.py imports nothing of tests/) ----
def molecular_adjacency(n, rng):
    """adj = I + S, S symmetric 0/1 'molecular-like': random spanning tree with max degree 4 plus floor(n/8) ring-closing
    edges (the reference builds I + adjacency, SMP_beta.h:505-526)."""
    import numpy as np

    A = np.zeros((n, n), np.float32)
    deg = np.zeros(n, np.int64)
    order = rng.permutation(n)
    for k in range(1, n):
        v = order[k]
        cands = [u for u in order[:k] if deg[u] < 4]
        u = cands[rng.integers(len(cands))] if cands else order[rng.integers(k)]
        A[u, v] = A[v, u] = 1
        deg[u] += 1
        deg[v] += 1
    extra, tries = n // 8, 0
    while extra > 0 and tries < 1000:
        tries += 1
        u, v = rng.integers(n), rng.integers(n)
        if u != v and A[u, v] == 0 and deg[u] < 4 and deg[v] < 4:
            A[u, v] = A[v, u] = 1
            deg[u] += 1
            deg[v] += 1
            extra -= 1
    return A + np.eye(n, dtype=np.float32)


def level_workload(graphs, V, C, seed):
    """`graphs` synthetic graphs of V vertices in which every receptive field, at level l-1 and at level l, is the whole vertex
    set in its own random order (n = V for every instance, the stacked T completely dense): the index tables of
    ccn_promote_forward for the V instances of each graph, the reduced adjacencies, and the group pointers of the host call."""
    import numpy as np

    rng = np.random.default_rng(seed)
    uniq = min(graphs, 8)
    per_graph_f = V * V * V * C
    f_off = np.zeros((graphs, V, V), np.int64)
    pos = np.zeros((graphs, V, V, V), np.int32)
    adj = np.zeros((graphs, V, V, V), np.float32)
    for g in range(uniq):
        prev = np.stack([rng.permutation(V) for _ in range(V)])           # phi_{l-1}(w), ordered
        cur = np.stack([rng.permutation(V) for _ in range(V)])            # phi_l(v), ordered
        where = np.argsort(prev, axis=1)                                   # where[w][u] = position of u inside phi_{l-1}(w)
        A = molecular_adjacency(V, rng)
        for v in range(V):
            f_off[g, v] = cur[v] * (V * V * C)                             # slab a comes from vertex w = phi_l(v)[a]
            pos[g, v] = where[cur[v]][:, cur[v]]                           # pos[a][i] = position of phi_l(v)[i] inside phi_{l-1}(w)
            adj[g, v] = A[np.ix_(cur[v], cur[v])]
    for g in range(uniq, graphs):
        f_off[g], pos[g], adj[g] = f_off[g % uniq], pos[g % uniq], adj[g % uniq]
    f_off += (np.arange(graphs, dtype=np.int64) * per_graph_f)[:, None, None]
    return {"f_off": f_off.reshape(-1), "m": np.full(graphs * V * V, V, np.int32), "pos": pos.reshape(-1),
            "adj": adj.reshape(graphs * V, V, V), "f_group_ptr": np.arange(graphs + 1, dtype=np.int64) * per_graph_f,
            "inst_group_ptr": np.arange(graphs + 1, dtype=np.int64) * V, "f_size": graphs * per_graph_f, "instances": graphs * V}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons for one GPU while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for j, nm in enumerate(names) if any(s[3 + j].lower().startswith("active") for s in self.samples)]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Pins this process to the CPUs next to its GPU (NVML's ideal affinity) BEFORE any pinned host memory is allocated, so
    that the pages land on the GPU's own NUMA node: with 8 ranks on one host the copies then spread over all memory
    controllers instead of meeting on the node the launcher happened to start on."""
    try:
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def random_host_instance(n, C, seed):
    import numpy as np

    rng = np.random.default_rng(seed)
    T = rng.uniform(-1, 1, (n, n, n, C)).astype(np.float32)
    adj = molecular_adjacency(n, rng)
    gout = rng.uniform(-1, 1, (n, n, 18 * C)).astype(np.float32)
    return T, adj, gout


def cpu_reference_rate(threads, n, C, reps=1, seed=1, lib="f32"):
    """contractions/s of the reference CPU code: `threads` private replicas x reps x (forward + backward)."""
    from oracle import pyoracle

    T, adj, gout = random_host_instance(n, C, seed)
    if pyoracle.ref_available(lib):
        ref, kind = pyoracle.RefOracle(lib), "reference"
        secs = ref.time_replicas(T, adj, gout, threads, reps)
        done = threads * reps
    else:  # the plain-C restatement, one thread
        orc, kind, threads = pyoracle.COracle("f32"), "port", 1
        t0 = time.perf_counter()
        orc.contract18_forward(T, adj)
        orc.contract18_backward(gout, adj)
        secs, done = time.perf_counter() - t0, 1
    return done / secs, kind, threads, secs, int((adj > 0).sum())


def run_reference(args):
    """The reference's own CPU implementation on all host cores: each step = one forward+backward per host thread on a
    private replica.  The reference's cost is linear in the channel count (its outermost loop is `for f`,
    RisiContraction_18.h:86), so when K full-size steps would not fit a few minutes the per-step sample is cut to the
    first C_s of the C = 64 channels and counted as C_s / C of a contraction (stated in `sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    budget_s = 150.0
    rate, kind, used, secs, nnz = cpu_reference_rate(cores, N_VERT, CHANNELS)  # warm-up at full size, also the estimate
    Cs = CHANNELS
    while Cs > 1 and (args.steps + min(args.warmup, 1)) * secs * Cs / CHANNELS > budget_s:
        Cs //= 2
    t_total, done = 0.0, 0.0
    for _ in range(args.steps):
        rate, kind, used, secs, nnz = cpu_reference_rate(cores, N_VERT, Cs)
        t_total += secs
        done += used * Cs / CHANNELS
    value = done / t_total
    sample = ("%d replicas (one per host thread) x 1 forward+backward of one N=%d instance with %d of the C=%d channels "
              "(= %.3f contraction each; the loop cost is linear in C%s), nnz(adj)=%d, per step"
              % (used, N_VERT, Cs, CHANNELS, Cs / CHANNELS,
                 "" if Cs == CHANNELS else "; the smaller sample is cache-friendlier, so this is an upper bound on the reference's rate", nnz))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "RisiContraction_18 fwd+bwd, N=%d C=%d, CPU replicas" % (N_VERT, CHANNELS)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import graphflow_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpus = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    n, C, B = N_VERT, CHANNELS, args.batch
    ctx = graphflow_b200.Context(local)
    if args.workspace_mib:
        ctx.set_workspace_limit(args.workspace_mib << 20)
    peak, peak_src = measured_peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def device_time(fn, steps, warm=3):
        """ms per call: CUDA events on the current stream around `steps` calls, barrier + synchronize on both sides,
        max over ranks."""
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    def wall_time(fn, steps, warm=1):
        """seconds per call of a synchronous host-buffer entry point, max over ranks."""
        for _ in range(warm):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return max_over_ranks(dt) / steps

    # ================= headline: device-resident StackTensor3D + RisiContraction_18 forward + backward =================
    rng = np.random.default_rng(1234 + rank)
    uniq = [molecular_adjacency(n, rng) for _ in range(min(B, 64))]
    adj = torch.from_numpy(np.stack([uniq[i % len(uniq)] for i in range(B)])).to(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(1234 + rank)
    T = torch.rand((B, n, n, n, C), device=device, generator=gen).mul_(2).sub_(1)
    gout = torch.rand((B, n, n, 18 * C), device=device, generator=gen).mul_(2).sub_(1)
    out = torch.empty((B, n, n, 18 * C), device=device)
    gT = torch.empty((B, n, n, n, C), device=device)
    # The n neighbour tensors [n, n, C] of every instance as a table of slab pointers (RisiContraction_18::add_tensor,
    # RisiContraction_18.h:49-55): the kernels read the slabs where they lie, which IS the StackTensor3D step.
    slab = n * n * C * 4
    idx = torch.arange(B * n, device=device, dtype=torch.int64) * slab
    slabs, gslabs = idx + T.data_ptr(), idx + gT.data_ptr()

    def step():
        ctx.contract18_forward(None, adj, out=out, slabs=slabs, n_max=n, C=C, batch=B)
        ctx.contract18_backward(gout, adj, gslabs=gslabs)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ctx.set_kernel_timing(True)
    launches0 = ctx.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = ctx.kernel_launches - launches0
    ktimes = ctx.kernel_timing()
    ctx.set_kernel_timing(False)
    clocks = sampler.stop() if sampler else None
    if ctx.fused_error_flag() != 0:
        raise RuntimeError("a fused-path tile timed out: results invalid")
    value = world * B * args.steps / (ms * 1e-3)
    extras = {}

    if not args.headline_only:
        # the same step with the stacked Tensor4D input of RisiContraction_18_gpu::setParameter (one base pointer)
        t_st = device_time(lambda: (ctx.contract18_forward(T, adj, out=out), ctx.contract18_backward(gout, adj, gT=gT)), 10)
        extras["stacked_tensor4d_input"] = {"value": world * B / (t_st * 1e-3), "unit": UNIT, "ms_per_step": t_st}
        # dense positive adjacency (the Coulomb-matrix mode of SMP_beta.h:521-524): the five [n x n] . A^T products per channel
        # then run as dense FFMA tiles out of shared memory instead of ~3-entry lists
        adj_dense = torch.rand((B, n, n), device=device, generator=gen) + 0.1
        ctx.set_kernel_timing(True)
        t_dn = device_time(lambda: (ctx.contract18_forward(None, adj_dense, out=out, slabs=slabs, n_max=n, C=C, batch=B),
                                    ctx.contract18_backward(gout, adj_dense, gslabs=gslabs)), 10, warm=2)
        kd = ctx.kernel_timing()
        ctx.set_kernel_timing(False)
        per_dir = 4 * (n ** 3 * C + n * n + 18 * n * n * C) * B
        extras["dense_adjacency"] = {
            "what": "same step with a dense positive adjacency (nnz = n^2 = 1024 instead of ~102)",
            "value": world * B / (t_dn * 1e-3), "unit": UNIT, "ms_per_step": t_dn,
            "roofline_frac": algorithmic_bytes(n, C) * B / (t_dn * 1e-3) / 1e9 / peak,
            "fwd_frac": per_dir / (kd["fwd_fused"][0] / kd["fwd_fused"][1] * 1e-3) / 1e9 / peak if "fwd_fused" in kd else None,
            "bwd_frac": per_dir / (kd["bwd_fused"][0] / kd["bwd_fused"][1] * 1e-3) / 1e9 / peak if "bwd_fused" in kd else None,
            "tensor_pipe": "0 % by design: 21 MFLOP of dense products per instance against 26 MB of compulsory HBM traffic"}
        del adj_dense
    del T, gT, out, gout, slabs, gslabs
    torch.cuda.empty_cache()

    # ================= the level: f_{l-1} -> promotion -> stack -> contraction -> mix (+bias, lrelu) and back ================
    level = e2e = None
    G, V, Co = args.level_graphs, n, C
    w = level_workload(G, V, C, 77 + rank)
    Bl = w["instances"]
    d = lambda x: torch.from_numpy(x).to(device)  # noqa: E731
    f = torch.rand(w["f_size"], device=device, generator=gen).mul_(2).sub_(1)
    Kw = (torch.rand((18 * C, Co), device=device, generator=gen) - 0.5) * 0.1
    bias = torch.rand((Co,), device=device, generator=gen) - 0.5
    gZ = torch.rand((Bl * n * n, Co), device=device, generator=gen) - 0.5
    f_off_d, m_d, pos_d, adj_d = d(w["f_off"]), d(w["m"]), d(w["pos"]), d(w["adj"])
    if not args.headline_only:
        from graphflow_b200 import shard

        X = torch.zeros((Bl, n * n, 18 * C), device=device)
        gX = torch.empty_like(X)
        gf = torch.zeros(w["f_size"], device=device)
        gflat = torch.zeros(18 * C * Co + Co, device=device)             # gK and gbias in ONE buffer: one all-reduce
        gK, gb = gflat[:18 * C * Co].view(18 * C, Co), gflat[18 * C * Co:]
        side = torch.cuda.Stream(device=device)
        done_mix = torch.cuda.Event()

        def level_step():
            _, Y, _ = ctx.gather_level_forward(f, f_off_d, m_d, pos_d, adj_d, Kw, bias, n, X=X)
            gflat.zero_()
            gf.zero_()
            ctx.mix_backward(X.view(Bl * n * n, 18 * C), Kw, gZ, bias=bias, Y=Y, gX=gX.view(Bl * n * n, 18 * C), gW=gK, gbias=gb)
            # the parameter-gradient all-reduce (SMP_beta.h:731-733 add_gradient) runs on a side stream UNDER the contraction backward
            done_mix.record()
            side.wait_event(done_mix)
            with torch.cuda.stream(side):
                shard.allreduce_gradients([gflat])
            ctx.gather_contract18_backward(gX.view(Bl, n, n, 18 * C), adj_d, f_off_d, m_d, pos_d, gf)
            torch.cuda.current_stream().wait_stream(side)

        ctx.set_kernel_timing(True)
        lms = device_time(level_step, 10)
        kt = ctx.kernel_timing()
        ctx.set_kernel_timing(False)
        per512 = lms * 512.0 / Bl
        level = {"what": "ONE fused-promotion level, device resident: f_{l-1} -> [gather+stack+contract18] -> mix fwd (+bias, lrelu) | "
                         "mix bwd (gX, gK, gb) -> [contract18 bwd + scatter into gf] with the all-reduce(gK, gb) overlapped",
                 "value": world * Bl / (lms * 1e-3), "unit": "level instances/s (fwd+bwd)", "ms_per_step": lms,
                 "instances_per_gpu": Bl, "ms_per_512_instances": per512, "allreduce_floats": int(gflat.numel()),
                 "kernels_ms_per_launch": {k: v[0] / v[1] for k, v in kt.items()}}
        if "mix_forward_tc" in kt:
            kms = kt["mix_forward_tc"][0] / kt["mix_forward_tc"][1]
            M = Bl * n * n
            extras["feature_mix"] = {"kernel": "mix_forward_tc", "ms": kms, "rows": M, "K": 18 * C, "P": Co,
                                     "achieved_gbs": 4.0 * M * (18 * C + 2 * Co) / (kms * 1e-3) / 1e9,
                                     "useful_fp32_tflops": 2.0 * M * 18 * C * Co / (kms * 1e-3) / 1e12,
                                     "issued_tf32_tflops": 3 * 2.0 * M * 18 * C * Co / (kms * 1e-3) / 1e12,
                                     "precision": "3xTF32 split (fp32-accurate)"}
        del X, gX, gf
        torch.cuda.empty_cache()

    # ================= end to end: the host-buffer level call, pinned host arrays, copies inside the timed region =================
    pin = lambda t: t.pin_memory()  # noqa: E731
    th = lambda x: torch.from_numpy(x)  # noqa: E731
    hf, hgZ = pin(f.cpu()), pin(gZ.cpu())
    hZ, hgf = pin(torch.empty((Bl * n * n, Co))), pin(torch.empty(w["f_size"]))
    hK, hb, hgK, hgb = pin(Kw.cpu()), pin(bias.cpu()), pin(torch.empty((18 * C, Co))), pin(torch.empty(Co))
    hfo, hm, hpos, hadj = pin(th(w["f_off"])), pin(th(w["m"])), pin(th(w["pos"])), pin(th(w["adj"]))
    hargs = (hf, th(w["f_group_ptr"]), th(w["inst_group_ptr"]), hfo, hm, hpos, hadj, hK, hb, hgZ, hZ, hgf, hgK, hgb, n)
    del f, gZ
    tab_bytes = (8 + 4) * Bl * n + 4 * Bl * n * n + 4 * Bl * n * n           # f_off + m, pos, adj of one level
    par_bytes = 4 * (18 * C * Co + Co)
    e2e_one = None
    if not args.headline_only:
        t1 = wall_time(lambda: ctx.gather_level_forward_backward_host(*hargs), args.e2e_steps)
        e2e_one = {"value": world * Bl / t1, "unit": UNIT, "call": "ccn_gather_level_forward_backward_host (one level per call)",
                   "h2d_bytes_per_step": 4 * (w["f_size"] + Bl * n * n * Co) + tab_bytes + par_bytes,
                   "d2h_bytes_per_step": 4 * (w["f_size"] + Bl * n * n * Co) + par_bytes, "instances_per_step": Bl, "s_per_step": t1}
    # L levels per call: in this synthetic graph family every level has the same (full) receptive fields, so the level l > 1
    # tables are the level-1 tables with the offsets pointing into the previous level's output array instead of f
    Lv = args.e2e_levels
    inst_off = torch.arange(Bl, dtype=torch.int64) * (n * n * Co)
    src_inst = (th(w["f_off"]) // (n * n * C))                            # which instance's tensor slab a of instance i reads
    hfo2 = pin((inst_off[src_inst]).contiguous())
    Kl = [hK] + [pin((Kw.cpu() * (1.0 + 0.01 * l)).contiguous()) for l in range(1, Lv)]
    gKl, gbl = [pin(torch.empty((18 * C, Co))) for _ in range(Lv)], [pin(torch.empty(Co)) for _ in range(Lv)]
    largs = (hf, th(w["f_group_ptr"]), th(w["inst_group_ptr"]), [hfo] + [hfo2] * (Lv - 1), [hm] * Lv, [hpos] * Lv, [hadj] * Lv, Kl,
             [hb] * Lv, hgZ, hZ, hgf, gKl, gbl, n)
    h2d_z = 4 * (w["f_size"] + Bl * n * n * Co) + Lv * (tab_bytes + par_bytes)
    d2h_z = 4 * (w["f_size"] + Bl * n * n * Co) + Lv * par_bytes
    e2e_levels = None
    if not args.headline_only:
        tz = wall_time(lambda: ctx.gather_levels_forward_backward_host(*largs), args.e2e_steps)
        e2e_levels = {"value": world * Bl * Lv / tz, "unit": UNIT, "h2d_bytes_per_step": h2d_z, "d2h_bytes_per_step": d2h_z,
                      "s_per_step": tz, "call": "ccn_gather_levels_forward_backward_host (the last level's Z comes down, its gradient goes up)"}
    # the headline end-to-end call: the same stack PLUS the read-out head and loss on the device -- a training step of the
    # model minus level 0; the last level's output and its gradient stay in HBM, targets go up, predictions / losses / gW come down
    hW, hgW = pin((torch.rand(Co, generator=torch.Generator().manual_seed(5)) - 0.5) * 1e-10), pin(torch.empty(Co))
    htgt, hpred, hloss = pin(torch.rand(G) - 0.5), pin(torch.empty(G)), pin(torch.empty(G))
    rargs = largs[:9] + (hW, htgt, hpred, hloss, hgf, gKl, gbl, hgW, n)
    te = wall_time(lambda: ctx.gather_levels_readout_forward_backward_host(*rargs), args.e2e_steps)
    e2e_value = world * Bl * Lv / te
    checksum = float(hloss.sum()) + float(hgf[:4].sum())                   # device->host read of the step's result
    checksum = checksum if checksum == checksum and abs(checksum) != float("inf") else None
    h2d = 4 * w["f_size"] + Lv * (tab_bytes + par_bytes) + 4 * (Co + G)
    d2h = 4 * w["f_size"] + Lv * par_bytes + 4 * (Co + 2 * G)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "contractions_per_step": Bl * Lv, "instances_per_level": Bl, "levels": Lv, "steps": args.e2e_steps, "s_per_step": te,
           "checksum": checksum,
           "call": "ccn_gather_levels_readout_forward_backward_host: %d levels + read-out + loss, device resident in between; per "
                   "contraction instance promotion + contraction + feature mix, forward and backward" % Lv,
           "host_memory": "pinned", "bytes_per_contraction": (h2d + d2h) / (Bl * Lv)}
    del hf, hgZ, hZ, hgf, hargs, largs, rargs

    # plain pinned copies of the same byte volumes, both directions at once, all ranks at once: the ceiling for ANY host-buffer API
    ceiling = None
    if not args.headline_only:
        nb = 1 << 29
        hs, hd = pin(torch.empty(nb, dtype=torch.uint8)), pin(torch.empty(nb, dtype=torch.uint8))
        ds, dd = torch.empty(nb, dtype=torch.uint8, device=device), torch.empty(nb, dtype=torch.uint8, device=device)
        s1, s2 = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)

        def both():
            with torch.cuda.stream(s1):
                dd.copy_(hs, non_blocking=True)
            with torch.cuda.stream(s2):
                hd.copy_(ds, non_blocking=True)

        tc = wall_time(both, 8)
        gbs = nb / tc / 1e9
        ceiling = {"what": "cudaMemcpyAsync H2D + D2H concurrently, 512 MiB each, pinned, every rank at the same time",
                   "gbs_each_way_per_gpu": gbs, "aggregate_gbs": 2 * gbs * world,
                   "e2e_ceiling_contractions_per_s": world * gbs * 1e9 / (max(h2d, d2h) / (Bl * Lv)),
                   "e2e_frac_of_ceiling": e2e_value / (world * gbs * 1e9 / (max(h2d, d2h) / (Bl * Lv))),
                   "e2e_one_level_frac_of_ceiling": (e2e_one["value"] / (world * gbs * 1e9 / (e2e_one["h2d_bytes_per_step"] / Bl)))
                   if e2e_one else None}
        del hs, hd, ds, dd

        # round 1's e2e: the stacked T itself crosses PCIe (ccn_contract18_forward_backward_host), three kinds of caller memory
        Be = args.e2e_op_batch
        hT, hA, hG = random_host_batch(Be, n, C, 5 + rank)
        hO, hGT = torch.empty((Be, n, n, 18 * C)), torch.empty((Be, n, n, n, C))
        op = {"call": "ccn_contract18_forward_backward_host (the stacked T crosses PCIe: 26.2 MB per instance)", "instances_per_step": Be}
        if world == 1:
            tp = wall_time(lambda: ctx.contract18_forward_backward_host(hT, hA, hG, hO, hGT), 1)
            op["pageable"] = {"value": Be / tp, "note": "plain new[]-style arrays, as the reference's Vector.h:24-25 allocates them"}
        for t in (hT, hA, hG, hO, hGT):
            ctx.host_register(t)
        tr = wall_time(lambda: ctx.contract18_forward_backward_host(hT, hA, hG, hO, hGT), 2)
        op["registered"] = {"value": world * Be / tr, "note": "the same arrays after ccn_host_register (cudaHostRegister in place)"}
        for t in (hT, hA, hG, hO, hGT):
            ctx.host_unregister(t)
        pT, pA, pG, pO, pGT = (pin(t) for t in (hT, hA, hG, hO, hGT))
        del hT, hG, hO, hGT
        tpn = wall_time(lambda: ctx.contract18_forward_backward_host(pT, pA, pG, pO, pGT), 2)
        op["pinned"] = {"value": world * Be / tpn}
        op["unit"] = UNIT
        extras["e2e_op"] = op
        del pT, pG, pO, pGT

    # ================= RisiContraction_50 at BASELINE config 5 (one GPU only): N=48, C=128, batch 256 =================
    if world == 1 and not args.headline_only:
        torch.cuda.empty_cache()
        n5, C5, B5 = 48, 128, args.r50_batch
        gen5 = torch.Generator(device=device)
        gen5.manual_seed(5)
        T5 = torch.rand((B5, n5, n5, n5, C5), device=device, generator=gen5) * 2 - 1
        a5 = (torch.rand((B5, n5, n5), device=device, generator=gen5) < 0.08).float()
        a5 = ((a5 + a5.transpose(1, 2) + torch.eye(n5, device=device)) > 0).float()
        o5 = torch.empty((B5, n5, n5, 50 * C5), device=device)
        g5 = torch.rand((B5, n5, n5, 50 * C5), device=device, generator=gen5) * 2 - 1
        gT5 = torch.empty_like(T5)
        r50_step = lambda: (ctx.contract50_forward(T5, a5, out=o5), ctx.contract50_backward(g5, a5, gT=gT5))  # noqa: E731
        rms = device_time(r50_step, 3, warm=2)
        ctx.set_kernel_timing(True)            # a separate pass for the per-kernel split: the events serialise the launches
        device_time(r50_step, 5, warm=0)
        k5 = ctx.kernel_timing()
        ctx.set_kernel_timing(False)
        b5 = 8 * (n5 ** 3 * C5 + 50 * n5 * n5 * C5 + n5 * n5)
        extras["contract50"] = {
            "workload": "StackTensor3D+RisiContraction_50 fwd+bwd, N=48 C=128, %d instances (BASELINE config 5)" % B5,
            "value": B5 / (rms * 1e-3), "unit": UNIT, "ms_per_step": rms, "algorithmic_bytes_per_instance": b5,
            "achieved_gbs": B5 * b5 / (rms * 1e-3) / 1e9, "roofline_frac": B5 * b5 / (rms * 1e-3) / 1e9 / peak,
            "kernels_ms_per_step": {k: v[0] / 5 for k, v in k5.items()}}
        del T5, a5, o5, g5, gT5
        torch.cuda.empty_cache()
        # the rest of the contraction family on the plan-driven 50-way kernels (RisiContraction_4 / _10), headline shape, 256 instances
        Bf = 256
        Tf = torch.rand((Bf, n, n, n, C), device=device, generator=gen5) * 2 - 1
        af = (torch.rand((Bf, n, n), device=device, generator=gen5) < 0.06).float()
        af = ((af + af.transpose(1, 2) + torch.eye(n, device=device)) > 0).float()
        gTf = torch.empty_like(Tf)
        fam = {}
        for S, name in ((4, "RisiContraction_4"), (10, "RisiContraction_10")):
            of = torch.empty((Bf, n, n, S * C), device=device)
            gof = torch.rand((Bf, n, n, S * C), device=device, generator=gen5) * 2 - 1
            adj_f = None if S == 4 else af
            fms = device_time(lambda: (ctx.contract_family_forward(S, Tf, adj_f, out=of), ctx.contract_family_backward(S, gof, adj_f, gT=gTf)), 5, warm=2)
            bf = 8 * (n ** 3 * C + S * n * n * C + n * n)
            fam[name] = {"value": Bf / (fms * 1e-3), "unit": UNIT, "ms_per_step": fms, "slabs": S, "algorithmic_bytes_per_instance": bf,
                         "roofline_frac": Bf * bf / (fms * 1e-3) / 1e9 / peak}
            del of, gof
        extras["family"] = dict(fam, workload="contraction family fwd+bwd through ccn_contract_family_*, N=%d C=%d, %d instances" % (n, C, Bf))
        del Tf, af, gTf
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    step_bytes = algorithmic_bytes(n, C) * B
    kern = {}
    for name, (kms, cnt) in ktimes.items():
        kern[name] = {"ms_per_step": kms / args.steps, "launches_per_step": cnt / args.steps}
    per_inst = {
        "fwd_fused": 4 * (n ** 3 * C + n * n + 18 * n * n * C),    # reads T + adj once, writes all 18 slabs once
        "bwd_fused": 4 * (18 * n * n * C + n * n + n ** 3 * C),    # reads gout + adj once, writes gT once
    }
    dom = max((k for k in kern if k in per_inst), key=lambda k: kern[k]["ms_per_step"], default=None)
    roofline = None
    if dom:
        achieved = per_inst[dom] * B / (kern[dom]["ms_per_step"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic_bytes(dom, n, C, B), "peak_source": peak_src,
                    "step_achieved": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                    "step_frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                    "kernels": {k: dict(v, achieved_gbs=(per_inst[k] * B / (v["ms_per_step"] * 1e-3) / 1e9)
                                        if k in per_inst else None) for k, v in kern.items()}}

    cores = os.cpu_count() or 1
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, kind, used, secs, nnz = cpu_reference_rate(cores, n, C)
        cpu = {"value": rate, "unit": UNIT, "cores": used, "kind": kind,
               "sample": "%d replicas (one per host thread) x 1 forward+backward of one N=%d C=%d instance "
                         "(nnz(adj)=%d), %.1f s wall, unmodified GraphFlow_32bit RisiContraction_18, g++ -O2" % (used, n, C, nnz, secs)}
        if not args.headline_only:
            try:  # the two other CPU figures SURVEY 8d asks for: best compiler flags, and the 6-threads-per-op _thread variant
                from oracle import pyoracle

                if pyoracle.ref_available("f32_o3"):
                    r3, _, u3, s3, _ = cpu_reference_rate(cores, n, C, lib="f32_o3")
                    cpu["o3_avx2"] = {"value": r3, "cores": u3, "flags": "g++ -O3 -march=x86-64-v3", "secs": s3}
                Tt, at, gt = random_host_instance(12, 16, 3)
                ref = pyoracle.RefOracle("f32")
                st = ref.time_thread_variant(Tt, at, gt, 1)
                ss = ref.time_replicas(Tt, at, gt, 1, 1)
                cpu["thread_variant"] = {"what": "RisiContraction_18_thread (6 threads inside one op, N^6 loops) at the reduced size "
                                                 "N=12 C=16 it can finish, next to the serial op at the same size",
                                         "secs_thread_op": st, "secs_serial_op": ss,
                                         "note": "cost grows ~N^6 for _thread vs nnz*N^3 for the serial op; at N=32 it is not runnable in minutes"}
            except Exception as e:  # noqa: BLE001
                cpu["extras_error"] = repr(e)

    ref_gpu = None
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_bench")
    if world == 1 and not args.headline_only and os.path.exists(exe):
        try:
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local)))
            o = subprocess.run([exe, str(n), str(C), "3"], capture_output=True, text=True, timeout=600, env=env).stdout.strip().splitlines()[-1]
            ref_gpu = json.loads(o)
            ref_gpu["what"] = ("the reference's OWN CUDA kernels and ops (unmodified GraphFlow_gpu_32bit RisiContraction_18_gpu.h / "
                               "MatMul_gpu.h rebuilt with nvcc -arch=sm_100a) on this GPU, one instance per call as the reference runs them")
        except Exception as e:  # noqa: BLE001
            ref_gpu = {"error": repr(e)}

    facade = None
    fexe = os.path.join(ROOT, "tests", "cpp", "_build", "test_facade_f32")
    if world == 1 and not args.headline_only and os.path.exists(fexe):
        try:
            o = subprocess.run([fexe, "bench", str(n), str(C), str(C)], capture_output=True, text=True, timeout=600)
            lines = [ln for ln in o.stdout.splitlines() if ln.startswith("{")]
            facade = json.loads(lines[-1]) if lines else {"error": (o.stdout + o.stderr)[-400:]}
        except Exception as e:  # noqa: BLE001
            facade = {"error": repr(e)}

    facade_model = None
    mexe = os.path.join(ROOT, "tests", "cpp", "_build", "test_model_f64")
    if world == 1 and not args.headline_only and os.path.exists(mexe):
        try:
            o = subprocess.run([mexe, "bench", "3", "32", "24", "128"], capture_output=True, text=True, timeout=600)
            lines = [ln for ln in o.stdout.splitlines() if ln.startswith("{")]
            facade_model = json.loads(lines[-1]) if lines else {"error": (o.stdout + o.stderr)[-400:]}
        except Exception as e:  # noqa: BLE001
            facade_model = {"error": repr(e)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "StackTensor3D+RisiContraction_18 fwd+bwd, N=%d C=%d, %d instances per GPU given as slab-pointer "
                                   "tables (the stack is fused into the read), molecular adjacency (nnz~102)" % (n, C, B),
                       "instances_per_gpu": B, "N": n, "C": C,
                       "l2_policy": "inputs larger than L2 (%.1f GiB streamed per step)" % (step_bytes / 2 ** 30),
                       "e2e_workload": "%d graphs x %d vertices per GPU, every receptive field full (n = %d, dense T), C_in = C_out = %d"
                                       % (G, V, V, C)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "level_step": level, "e2e_levels_only": e2e_levels, "e2e_one_level": e2e_one, "host_copy_ceiling": ceiling, "ref_gpu_kernels": ref_gpu, "facade": facade, "facade_model": facade_model,
            "numa_cpus": (len(cpus) if cpus else None)}
    line.update(extras)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def random_host_batch(Be, n, C, seed):
    import numpy as np
    import torch

    rng = np.random.default_rng(seed)
    T = torch.from_numpy(rng.uniform(-1, 1, (Be, n, n, n, C)).astype(np.float32))
    A = torch.from_numpy(np.stack([molecular_adjacency(n, rng) for _ in range(Be)]))
    G = torch.from_numpy(rng.uniform(-1, 1, (Be, n, n, 18 * C)).astype(np.float32))
    return T, A, G


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 100 (b200), 5 (reference)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=2048, help="instances per GPU per step (T and gT are 16 GiB each at 2048)")
    ap.add_argument("--level-graphs", type=int, default=64, help="graphs of 32 vertices per GPU for the level / e2e figures")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-levels", type=int, default=4, help="levels per host call of the e2e figure (BASELINE config 3: 4 layers)")
    ap.add_argument("--e2e-op-batch", type=int, default=128)
    ap.add_argument("--r50-batch", type=int, default=256)
    ap.add_argument("--workspace-mib", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip every secondary figure (profiling runs)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 5 if args.impl == "reference" else 100
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517"] + sys.argv
        sys.exit(subprocess.call(cmd))
    run_b200(args)


if __name__ == "__main__":
    main()
