#!/usr/bin/env python
"""Benchmark of the meta-transfer hot path (BASELINE.json metric: meta-step utterances/sec, enc2/dec4/d512, k=8).

One "step" = one full meta-step of trainer/asr/transient_trainer.py:150-255 on synthetic cfg-2 data:
N tasks x (k_train=8 inner fwd/bwd + SGD step + k_valid=8 outer fwd/bwd at the adapted weights) + copy-grad
accumulation + one Adam step.  Utterance passes per step U = N * (k_train + k_valid).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--gemm-mode 0|1|2]

N > 1 (torchrun): every rank runs TASKS_PER_GPU tasks of the meta-batch (weak scaling: the meta-batch grows with
N), one NCCL all-reduce of the flat copy_grad arena per step, identical Adam step on every rank.
Prints ONE JSON line on rank 0 (see README / DESIGN.md for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "meta-transfer-learning_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch

K_TRAIN = K_VALID = 8
T_FRAMES, L_TOKENS = 101, 32
TASKS_PER_GPU = 3
LR, META_LR, DROPOUT = 1e-4, 1e-4, 0.1
METRIC = "meta-step utterances/sec (enc2/dec4/d512, k=8)"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE conv.2-forward launch (conv3x3_kw_kernel<64>) from the committed
# `ncu --set full` captures, per gemm mode; None until a capture of that mode exists
ROOFLINE_TRAFFIC_BYTES = {1: 33474816 + 765184, 2: 33627648 + 661760}   # profiles/r01_f_ncu_full_conv2_kw_{tf32,3xtf32}.txt
# sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active of the same captures: how busy the tcgen05 pipe is with
# the instruction mix the mode issues (3xTF32 issues 3 MMA-products per useful product)
ROOFLINE_TENSOR_PIPE_PCT = {1: 30.7, 2: 50.4}
UNIT = "utterance-passes/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]), bf16_sus=float(p["bf16_tflops_sustained"]),
                    src="measured")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


def synth_task(spec, k, seed):
    """SURVEY 8d synthetic batch: x ~ N(0,1) (k,1,161,T) fp32, targets U[4,V) (k,L) int64, full lengths."""
    rng = np.random.default_rng(seed)
    x = torch.from_numpy(rng.standard_normal((k, 1, spec.n_freq, T_FRAMES), dtype=np.float32))
    y = torch.from_numpy(rng.integers(4, spec.vocab, size=(k, L_TOKENS), dtype=np.int64))
    lens = torch.full((k,), T_FRAMES, dtype=torch.int32)
    return x, lens, y


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's algorithm for this path on the host cores: oracle/ref_meta.meta_step (the CPU
    restatement pinned against the live reference) with all host threads, same workload."""
    if rank != 0:
        return
    from oracle import ref_asr, ref_meta
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ref_asr.ModelConfig(dropout=DROPOUT)
    p = ref_asr.init_params(cfg, 0)
    adam = ref_meta.AdamState()
    bufs = ref_asr.buffers(cfg)
    n_tasks = TASKS_PER_GPU

    def step(i):
        tasks = [ref_meta.synth_batch(cfg, K_TRAIN, T_FRAMES, L_TOKENS, 1000 * i + t) for t in range(n_tasks)]
        val = ref_meta.synth_batch(cfg, K_VALID, T_FRAMES, L_TOKENS, 1000 * i + 999)
        ref_meta.meta_step(p, adam, cfg, tasks, val, lr=LR, meta_lr=META_LR, train=True, bufs=bufs)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    u = n_tasks * (K_TRAIN + K_VALID)
    val = u * args.steps / dt
    sample = f"{args.steps} full meta-steps ({n_tasks} tasks x ({K_TRAIN}+{K_VALID}) utterances, T={T_FRAMES}, L={L_TOKENS}, dropout {DROPOUT})"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: meta_transfer_train.py 3 synthetic tasks k-train=8 enc2/dec4 d512 --copy-grad",
                   "tasks": n_tasks, "k_train": K_TRAIN, "k_valid": K_VALID, "frames": T_FRAMES, "tokens": L_TOKENS},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(gpu_index), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import mtl_b200
    from mtl_b200 import lib as L
    from mtl_b200.shard import exchange_copy_grad
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    spec = mtl_b200.ModelSpec()
    s = mtl_b200.Session(spec, dev, gemm_mode=args.gemm_mode)
    lib = L.get_lib()
    n_local = TASKS_PER_GPU
    n_total = n_local * world
    u_total = n_total * (K_TRAIN + K_VALID)

    # parameters: reference init distributions (xavier_uniform on >=2-D), same on every rank
    g = torch.Generator().manual_seed(0)
    theta, theta0, grad, cg, m, v = (s.new_arena() for _ in range(6))
    adam_state = s.new_adam_state()
    views = s.views(theta)
    for name, shape, off, n in s.table:
        if len(shape) >= 2:
            rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
            a = (6.0 / (shape[1] * rf + shape[0] * rf)) ** 0.5
            views[name].copy_((torch.rand(shape, generator=g) * 2 - 1) * a)
        elif "layer_norm" in name and name.endswith("weight"):
            views[name].fill_(1.0)
        elif "layer_norm" in name:
            views[name].zero_()
        else:
            views[name].copy_((torch.rand(shape, generator=g) * 2 - 1) * 0.05)

    total_steps = args.warmup + args.steps
    # host batches (pinned) for every step; device copies for the device-resident measurement
    host = []
    for i in range(total_steps):
        tasks = [synth_task(spec, K_TRAIN, 1000 * i + rank * n_local + t) for t in range(n_local)]
        val = synth_task(spec, K_VALID, 1000 * i + 999)
        host.append(([tuple(t.pin_memory() for t in b) for b in tasks], tuple(t.pin_memory() for t in val)))
    resident = [([tuple(t.to(dev) for t in b) for b in tasks], tuple(t.to(dev) for t in val)) for tasks, val in host]
    n_tok = L_TOKENS + 1
    stepper = mtl_b200.MetaStepper(s, n_local, n_lanes=args.lanes or n_local, use_graph=not args.no_graph)
    all_results = torch.zeros(total_steps, n_local, 16, device=dev)
    torch.cuda.synchronize()

    def meta_step(i, tasks, val):
        # snapshot / reset of the weights (transient_trainer.py:160,237) are folded into the per-lane
        # theta copies inside mtl_meta_tasks; theta itself only changes in the Adam step below
        for t, b in enumerate(tasks):
            stepper.load_task(t, *b, n=n_tok)                   # H2D (e2e) or D2D (resident) into the static slots
        stepper.load_val(*val, n=n_tok)
        res = stepper.run(theta, cg, LR, 1.0 / n_total, dropout=DROPOUT, seed=i * 64 + rank)
        all_results[i].copy_(res, non_blocking=True)
        exchange_copy_grad(cg, dist)                            # the one exchange step
        s.meta_finish(theta, grad, cg, m, v, adam_state, META_LR)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput
    # the clock sampler starts before the warm-up steps (nvidia-smi needs ~0.2 s to produce its first line and the
    # timed region is only ~0.15 s long): its samples cover warm-up + timed region, all under load
    clocks = ClockSampler(local_rank) if rank == 0 else None
    for i in range(args.warmup):
        meta_step(i, *resident[i])
    barrier()
    l0 = lib.mtl_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, total_steps):
        meta_step(i, *resident[i])
    e1.record()
    barrier()
    launches = lib.mtl_launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    clk = clocks.stop() if clocks else None
    losses = all_results[args.warmup:, :, 8].mean(dim=1).tolist()

    # ---------------- end to end: host (pinned) batches in, losses out, every step
    h2d = sum(t.numel() * t.element_size() for b in host[0][0] for t in b) + \
        sum(t.numel() * t.element_size() for t in host[0][1])
    d2h = n_local * 16 * 4
    host_res = torch.zeros(n_local, 16).pin_memory()

    def e2e_step(i):
        tasks, val = host[i]
        meta_step(i, tasks, val)
        host_res.copy_(stepper.results, non_blocking=True)
        torch.cuda.current_stream().synchronize()               # the trainer prints the loss every step
        return float(host_res[:, 8].mean())

    for i in range(min(2, args.warmup)):
        e2e_step(i)
    barrier()
    e0.record()
    for i in range(args.warmup, total_steps):
        e2e_step(i)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist is not None:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms2)
    graph_caps, graph_reps = s.graph_stats()

    if args.timeline and rank == 0:
        dump_timeline(args.timeline, lambda: [meta_step(i, *resident[i]) for i in range(total_steps - 2, total_steps)])

    # ---------------- roofline of the dominant kernel: conv.2 forward contraction (130088 x 64 x 576)
    roof = None
    if rank == 0:
        roof = roofline_conv_gemm(s, lib, dev, args.gemm_mode)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()

    if rank == 0:
        val = u_total * args.steps / (ms_total / 1e3)
        out = {
            "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "f32", 1: "tf32", 2: "3xtf32"}[args.gemm_mode], "data": "synthetic",
            "config": {"workload": "cfg2: meta_transfer_train.py 3 synthetic tasks k-train=8 enc2/dec4 d512 --copy-grad"
                       + (f", {n_local} tasks per GPU x {world} GPUs, one all-reduce of copy_grad" if world > 1 else ""),
                       "tasks": n_total, "k_train": K_TRAIN, "k_valid": K_VALID, "frames": T_FRAMES,
                       "tokens": L_TOKENS, "dropout": DROPOUT, "gemm_mode": args.gemm_mode,
                       "l2": "per-pass working set ~1.1 GB of activations + 56 MB x 6 arenas >> 126 MB L2"},
            "e2e": {"value": u_total * args.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "cuda_graph": {"captures": graph_caps, "replays": graph_reps, "lanes": stepper.n_lanes},
            "clocks": clk,
            "roofline": roof,
            "cpu_baseline": cpu,
            "loss_first_last": [losses[0], losses[-1]],
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def dump_timeline(path, fn):
    """Diagnostic (never part of a reported number): per-kernel start / duration / stream of the calls made by fn."""
    import gzip
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    tmp = path + ".trace.json"
    prof.export_chrome_trace(tmp)
    with open(tmp) as f:
        ev = json.load(f)["traceEvents"]
    os.remove(tmp)
    with gzip.open(path, "wt") as f:
        f.write("name,stream,start_us,dur_us\n")
        for e in ev:
            if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset"):
                name = e["name"].replace(",", ";")[:120]
                f.write("%s,%s,%.3f,%.3f\n" % (name, e.get("args", {}).get("stream", e.get("tid")), e["ts"], e["dur"]))


def roofline_conv_gemm(s, lib, dev, mode):
    """Dominant (FLOP-wise) kernel of the step: the conv.2 forward implicit GEMM, 130088 pixels x 64 x 576
    (models/asr/transformer.py:49), timed through the same entry point the engine launches
    (mtl_conv3x3_relu_fwd -> conv3x3_kw_kernel<64,...>; mode 0 = im2col + fp32 GEMM).
    Inputs are rotated over 5 buffers (5 x 67 MB read+write > 126 MB L2) so no launch re-reads a hot input.
    Algorithmic FLOPs per launch = 2 * B*F*T * Cout * 9*Cin; algorithmic bytes = x + y + w."""
    import ctypes as C
    from mtl_b200 import lib as L
    B, F, T, Cin, Cout = K_TRAIN, 161, T_FRAMES, 64, 64
    M, N, Kd = B * F * T, Cout, 9 * Cin
    NB = 5
    xs = [torch.randn(B, F, T, Cin, device=dev) for _ in range(NB)]
    outs = [torch.empty(B, F, T, Cout, device=dev) for _ in range(NB)]
    w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
    bias = torch.zeros(N, device=dev)
    wg = torch.empty(2, N, Kd, device=dev)                 # [hi | lo] tf32 halves of the weights in 3xTF32
    col = torch.empty(M, Kd, device=dev) if mode == 0 else None
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    pv = lambda t: None if t is None else C.c_void_p(t.data_ptr())

    def run(i):
        L.check(lib.mtl_conv3x3_relu_fwd(mode, pv(xs[i % NB]), pv(w), pv(bias), pv(col), pv(wg), pv(outs[i % NB]),
                                         B, F, T, Cin, Cout, st))
    for i in range(NB):
        run(i)
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pk = _peaks()
    flops = 2.0 * M * N * Kd
    ach = flops / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": ach / pk["bf16"],
            "traffic": ROOFLINE_TRAFFIC_BYTES.get(mode),
            "tensor_pipe_active_pct_ncu": ROOFLINE_TENSOR_PIPE_PCT.get(mode),
            "algorithmic_bytes": 4 * (M * Cin + M * Cout + N * Kd), "flops_per_launch": flops,
            "kernel": "conv.2 forward implicit GEMM %dx%dx%d (%s), incl. its weight re-layout launch" % (
                M, N, Kd, {0: "im2col + gemm_simt_kernel fp32 CUDA cores", 1: "conv3x3_kw_kernel<64> tcgen05 tf32",
                           2: "conv3x3_kw_kernel<64> tcgen05 3xtf32"}[mode]),
            "ms_per_launch": ms,
            "peak_source": pk["src"] + " dense bf16 cuBLAS burst (no fp32-input tensor peak is measured; tf32 nominal"
                           " = 1/2 of bf16, 3xtf32 issues 3 MMAs per product => attainable <= 1/6 of this peak)"}


def cpu_baseline():
    from oracle import ref_asr, ref_meta
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ref_asr.ModelConfig(dropout=DROPOUT)
    p = ref_asr.init_params(cfg, 0)
    adam = ref_meta.AdamState()
    bufs = ref_asr.buffers(cfg)

    def step(i):
        tasks = [ref_meta.synth_batch(cfg, K_TRAIN, T_FRAMES, L_TOKENS, 1000 * i + t) for t in range(TASKS_PER_GPU)]
        val = ref_meta.synth_batch(cfg, K_VALID, T_FRAMES, L_TOKENS, 1000 * i + 999)
        ref_meta.meta_step(p, adam, cfg, tasks, val, lr=LR, meta_lr=META_LR, train=True, bufs=bufs)
    step(0)
    n = 3
    t0 = time.perf_counter()
    for i in range(n):
        step(1 + i)
    dt = time.perf_counter() - t0
    u = TASKS_PER_GPU * (K_TRAIN + K_VALID)
    return {"value": u * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} full cfg-2 meta-steps after 1 warm-up ({u} utterance passes each), torch CPU, {cores} threads"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gemm-mode", type=int, default=int(os.environ.get("MTL_GEMM_MODE", "2")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the meta-step eagerly (no CUDA graph replay)")
    ap.add_argument("--lanes", type=int, default=0, help="concurrent task lanes (default: one per task)")
    ap.add_argument("--timeline", default="", help="after the measurements, trace 2 more steps with torch.profiler "
                    "(CUPTI kernel records) and write name/stream/start/duration rows to this .csv.gz (diagnostic)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # python bench.py --gpus N without torchrun: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
