#!/usr/bin/env python
"""bench.py -- contractions/sec of the CCN hot path (StackTensor3D + RisiContraction_18 forward + backward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json north_star / SURVEY.md section 8d, op-level): a batch of independent synthetic instances with
N = 32 vertices and C = 64 channels, T ~ U[-1,1), molecular-like adjacency I + S (nnz ~ 102), gout ~ U[-1,1).
One *step* = forward then backward (beta = 0) of the whole per-GPU batch.  `value` = instances that completed
forward+backward per second over all ranks, inputs resident in HBM.  Ranks are independent (no data-path
collective): weak scaling.  `e2e` is the same metric through the host-buffer C-ABI entry point with pinned host
arrays (H2D of T, adj, gout and D2H of out, gT inside the timed region).

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


def make_inputs(batch, n, C, seed, device):
    import numpy as np
    import torch
    from tests.util import molecular_adjacency

    rng = np.random.default_rng(seed)
    uniq = [molecular_adjacency(n, rng) for _ in range(min(batch, 64))]
    adj = torch.from_numpy(np.stack([uniq[i % len(uniq)] for i in range(batch)])).to(device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    T = torch.rand((batch, n, n, n, C), device=device, generator=g).mul_(2).sub_(1)
    gout = torch.rand((batch, n, n, 18 * C), device=device, generator=g).mul_(2).sub_(1)
    return T, adj, gout


def cpu_reference_rate(threads, n, C, reps=1, seed=1):
    """contractions/s of the reference CPU code: `threads` private replicas x reps x (forward + backward)."""
    import numpy as np
    from oracle import pyoracle
    from tests.util import random_instance

    T, adj, gout = random_instance(n, C, np.random.default_rng(seed))
    if pyoracle.ref_available("f32"):
        ref, kind = pyoracle.RefOracle("f32"), "reference"
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
    import torch
    import torch.distributed as dist

    import graphflow_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    n, C, B = N_VERT, CHANNELS, args.batch
    ctx = graphflow_b200.Context(local)
    if args.workspace_mib:
        ctx.set_workspace_limit(args.workspace_mib << 20)
    T, adj, gout = make_inputs(B, n, C, 1234 + rank, device)
    out = torch.empty((B, n, n, 18 * C), device=device)
    gT = torch.empty((B, n, n, n, C), device=device)

    def step():
        ctx.contract18_forward(T, adj, out=out)
        ctx.contract18_backward(gout, adj, gT=gT)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

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
    ms = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches - launches0
    ktimes = ctx.kernel_timing()
    ctx.set_kernel_timing(False)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer entry point (pinned host arrays, copies inside the timed region) ----
    Be = args.e2e_batch
    hT = T[:Be].cpu().pin_memory()
    hA = adj[:Be].cpu().pin_memory()
    hG = gout[:Be].cpu().pin_memory()
    hO = torch.empty((Be, n, n, 18 * C)).pin_memory()
    hGT = torch.empty((Be, n, n, n, C)).pin_memory()
    ctx.contract18_forward_backward_host(hT, hA, hG, hO, hGT)  # warm-up (allocates the staging ring)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        ctx.contract18_forward_backward_host(hT, hA, hG, hO, hGT)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * Be * args.e2e_steps / te.item()
    checksum = float(hO[0, 0, 0, :4].sum())  # device->host read of the step's result
    h2d = 4 * Be * (n ** 3 * C + n * n + 18 * n * n * C)
    d2h = 4 * Be * (18 * n * n * C + n ** 3 * C)

    # ---- one CCN level, forward + backward (secondary figure): contraction -> feature mix (+bias, leaky-ReLU) and back,
    #      with the parameter-gradient all-reduce (NCCL) that a data-parallel training step adds (SMP_beta.h:731-733) ----
    mix = None
    level = None
    if not args.no_mix:
        from graphflow_b200 import shard

        X = out.reshape(B * n * n, 18 * C)
        gen = torch.Generator(device=device)
        gen.manual_seed(99)
        Wm = (torch.rand((18 * C, C), device=device, generator=gen) - 0.5) * 0.1
        bias = torch.rand((C,), device=device, generator=gen) - 0.5
        gZ = torch.rand((B * n * n, C), device=device, generator=gen) - 0.5
        gK = torch.zeros_like(Wm)
        gb = torch.zeros_like(bias)
        gX = gout.reshape(B * n * n, 18 * C)  # the mix backward writes the contraction's output gradient in place of gout

        def level_step():
            ctx.contract18_forward(T, adj, out=out)
            Y, Z = ctx.mix_forward(X, Wm, bias)
            gK.zero_()
            gb.zero_()
            ctx.mix_backward(X, Wm, gZ, bias=bias, Y=Y, gX=gX, gW=gK, gbias=gb)
            ctx.contract18_backward(gout, adj, gT=gT)
            shard.allreduce_gradients([gK, gb])

        for _ in range(3):
            level_step()
        barrier()
        ctx.set_kernel_timing(True)
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lsteps = 20
        l0.record()
        for _ in range(lsteps):
            level_step()
        l1.record()
        barrier()
        lt = torch.tensor([l0.elapsed_time(l1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(lt, op=dist.ReduceOp.MAX)
        kt = ctx.kernel_timing()
        ctx.set_kernel_timing(False)
        lms = lt.item() / lsteps
        level = {"what": "contract18 fwd -> mix fwd (+bias, lrelu) -> mix bwd (gX, gK, gb) -> contract18 bwd -> all-reduce(gK, gb)",
                 "value": world * B / (lms * 1e-3), "unit": "level instances/s (fwd+bwd)", "ms_per_step": lms,
                 "allreduce_floats": int(gK.numel() + gb.numel()),
                 "kernels_ms_per_step": {k: v[0] / lsteps for k, v in kt.items()}}
        if "mix_forward_tc" in kt:
            kms = kt["mix_forward_tc"][0] / kt["mix_forward_tc"][1]
            M = B * n * n
            mix = {"kernel": "mix_forward_tc", "ms": kms, "rows": M, "K": 18 * C, "P": C,
                   "achieved_gbs": 4.0 * M * (18 * C + 2 * C) / (kms * 1e-3) / 1e9,
                   "useful_fp32_tflops": 2.0 * M * 18 * C * C / (kms * 1e-3) / 1e12,
                   "issued_tf32_tflops": 3 * 2.0 * M * 18 * C * C / (kms * 1e-3) / 1e12,
                   "precision": "3xTF32 split (fp32-accurate)"}
        del X, Wm, gZ

    # ---- RisiContraction_50 at BASELINE config 5's shape (secondary figure, one GPU only): N=48, C=128, 64 instances ----
    r50 = None
    if world == 1 and not args.no_mix:
        n5, C5, B5 = 48, 128, 64
        gen = torch.Generator(device=device)
        gen.manual_seed(5)
        T5 = torch.rand((B5, n5, n5, n5, C5), device=device, generator=gen) * 2 - 1
        a5 = (torch.rand((B5, n5, n5), device=device, generator=gen) < 0.08).float()
        a5 = ((a5 + a5.transpose(1, 2) + torch.eye(n5, device=device)) > 0).float()
        o5 = torch.empty((B5, n5, n5, 50 * C5), device=device)
        g5 = torch.rand((B5, n5, n5, 50 * C5), device=device, generator=gen) * 2 - 1
        gT5 = torch.empty_like(T5)
        for _ in range(3):
            ctx.contract50_forward(T5, a5, out=o5)
            ctx.contract50_backward(g5, a5, gT=gT5)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rsteps = 5
        r0.record()
        for _ in range(rsteps):
            ctx.contract50_forward(T5, a5, out=o5)
            ctx.contract50_backward(g5, a5, gT=gT5)
        r1.record()
        torch.cuda.synchronize()
        rms = r0.elapsed_time(r1) / rsteps
        b5 = 8 * (n5 ** 3 * C5 + 50 * n5 * n5 * C5 + n5 * n5)
        pk = measured_peaks()[0]
        r50 = {"workload": "StackTensor3D+RisiContraction_50 fwd+bwd, N=48 C=128, 64 instances (13.8 GiB streamed per step: larger than L2)",
               "value": B5 / (rms * 1e-3), "unit": UNIT, "ms_per_step": rms, "algorithmic_bytes_per_instance": b5,
               "achieved_gbs": B5 * b5 / (rms * 1e-3) / 1e9, "roofline_frac": B5 * b5 / (rms * 1e-3) / 1e9 / pk}
        del T5, a5, o5, g5, gT5

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    step_bytes = algorithmic_bytes(n, C) * B
    kern = {}
    for name, (kms, cnt) in ktimes.items():
        kern[name] = {"ms_per_step": kms / args.steps, "launches_per_step": cnt / args.steps}
    # dominant kernel = largest share of the step; its algorithmic bytes per step
    per_inst = {
        "fwd_stream": 4 * (n ** 3 * C + n * n + 8 * n * n * C),   # reads T + adj, writes slabs 1,3,4,6,7,10,11,13
        "fwd_finish": 4 * (10 * n * n * C),                        # writes slabs 2,5,8,9,12,14,15,16,17,18
        "bwd_planes": 4 * (16 * n * n * C + n * n),                # reads 16 of the 18 gout slabs
        "bwd_stream": 4 * (n ** 3 * C + 2 * n * n * C),            # writes gT, reads slabs 6 and 10
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
                         "(nnz(adj)=%d), %.1f s wall, unmodified GraphFlow_32bit RisiContraction_18" % (used, n, C, nnz, secs)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "StackTensor3D+RisiContraction_18 fwd+bwd, N=%d C=%d, %d instances per GPU, "
                                   "molecular adjacency (nnz~102)" % (n, C, B),
                       "instances_per_gpu": B, "N": n, "C": C,
                       "l2_policy": "inputs larger than L2 (%.1f GiB streamed per step)" % (step_bytes / 2 ** 30)},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "instances_per_step": Be, "steps": args.e2e_steps, "checksum": checksum},
            "gpu_launches": launches, "clocks": clocks, "feature_mix": mix, "level_step": level, "contract50": r50}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 400 (b200), 5 (reference)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="instances per GPU per step")
    ap.add_argument("--e2e-batch", type=int, default=256)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--workspace-mib", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mix", action="store_true", help="skip the secondary feature-mix (tensor core) measurement")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 5 if args.impl == "reference" else 400
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
