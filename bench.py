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
UNIT = "utterance-passes/s"
# algorithmic FLOPs of one cfg-2 meta-step with 3 tasks (48 utterance passes, forward + backward = 3 x the
# forward MACs x 2): all dense contractions, and the attention + FFN + projection share the north_star names
STEP_TFLOP_3TASKS, ATTN_FFN_TFLOP_3TASKS = 0.533, 0.0653
MODEL = dict()                      # ModelSpec overrides (cfg 2 = the defaults: enc2 / dec4 / d512 / 8 heads)
WORKLOAD = "cfg2: meta_transfer_train.py 3 synthetic tasks k-train=8 enc2/dec4 d512 --copy-grad"


def select_config(name):
    """cfg2 (default, the BASELINE metric) or cfg4 = BASELINE configs[3]: enc4/dec6, d_model 768, 12 heads, --src-max-len
    5000 (T = 5000 frames, T' = 1250), L = 256 target tokens and d_inner = 768 (not given by BASELINE.json,
    assumed), 8 tasks on 8 GPUs = ONE task per GPU.  Arithmetic stays 3xTF32 / TF32 on fp32 storage (>= the bf16 the
    config names); SURVEY 8d FLOPs: 82.73 TFLOP per 8-task meta-step, 21.26 of them attention + FFN + projections."""
    global T_FRAMES, L_TOKENS, TASKS_PER_GPU, STEP_TFLOP_3TASKS, ATTN_FFN_TFLOP_3TASKS, MODEL, WORKLOAD
    if name == "cfg4":
        T_FRAMES, L_TOKENS, TASKS_PER_GPU = 5000, 256, 1
        STEP_TFLOP_3TASKS, ATTN_FFN_TFLOP_3TASKS = 82.73 * 3 / 8, 21.26 * 3 / 8       # per 3 tasks, like the cfg-2 constants
        MODEL = dict(n_enc=4, n_dec=6, d_model=768, n_heads=12, d_k=64, d_v=64, d_inner=768)
        global METRIC
        METRIC = "meta-step utterances/sec (enc4/dec6/d768/h12, k=8, T=5000)"
        WORKLOAD = ("cfg4: meta_transfer_train.py enc4/dec6 d768 h12 --src-max-len 5000 (T=5000, L=256, d_inner=768 assumed), "
                    "1 task per GPU, --copy-grad")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]), bf16_sus=float(p["bf16_tflops_sustained"]),
                    src="measured")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


def _ncu(role, mode):
    """Metrics of the committed `ncu --set full` capture that documents `role` (profiles/ncu_metrics.json, written by
    tools/ncu_extract.py from profiles/*_ncu_full_*.txt); {} when there is no capture for this engine mode."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_metrics.json")) as f:
            caps = json.load(f)
    except Exception:
        return {}
    best = {}
    for v in caps.values():
        if v.get("role") == role and v.get("gemm_mode") == mode and v.get("source", "") > best.get("source", ""):
            best = v
    return best


def workload_config(n_total, world, scaling):
    """The `config` object, identical in both arms (the reference arm times a bounded sample of the same workload)."""
    w = WORKLOAD
    if world > 1:
        w += (f", {n_total} tasks over {world} GPUs ({scaling} scaling), one exchange of copy_grad per step")
    return {"workload": w, "tasks": n_total, "k_train": K_TRAIN, "k_valid": K_VALID, "frames": T_FRAMES,
            "tokens": L_TOKENS, "dropout": DROPOUT,
            "l2": "per-pass working set ~1.1 GB of activations + 56 MB x 6 arenas >> 126 MB L2"}


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
    sample = (f"{args.steps} full meta-steps of {n_tasks} tasks x ({K_TRAIN}+{K_VALID}) utterances (T={T_FRAMES}, L={L_TOKENS}, "
              f"dropout {DROPOUT}) on {cores} host threads: the CPU port of the reference step (oracle/ref_meta.meta_step; "
              "in the build container it runs 15 % faster than the live TransientTrainer, 1.99 vs 2.36 s/step on 8 cores)")
    n_total = TASKS_PER_GPU * max(1, args.gpus)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n_total, max(1, args.gpus), "weak"),
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
    from mtl_b200.shard import MetaExchange, task_shard
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    spec = mtl_b200.ModelSpec(**MODEL)
    s = mtl_b200.Session(spec, dev, gemm_mode=args.gemm_mode)
    lib = L.get_lib()
    n_local = TASKS_PER_GPU
    n_total = n_local * world
    u_total = n_total * (K_TRAIN + K_VALID)

    # parameters: reference init distributions (xavier_uniform on >=2-D), same on every rank
    g = torch.Generator().manual_seed(0)
    theta, grad, cg, m, v = (s.new_arena() for _ in range(5))
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
    theta_init = theta.clone()
    exchange = MetaExchange(s, dist, mode=args.exchange, overlap=not args.no_overlap)

    total_steps = args.warmup + args.steps
    n_tok = L_TOKENS + 1

    def make_data(task_ids):
        """pinned host batches for every step (task `t` of step i has seed 1000 i + t) + device copies"""
        host = []
        for i in range(total_steps):
            tasks = [synth_task(spec, K_TRAIN, 1000 * i + t) for t in task_ids]
            val = synth_task(spec, K_VALID, 1000 * i + 999)
            host.append(([tuple(t.pin_memory() for t in b) for b in tasks], tuple(t.pin_memory() for t in val)))
        resident = [([tuple(t.to(dev) for t in b) for b in tasks], tuple(t.to(dev) for t in val)) for tasks, val in host]
        return host, resident

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, first, last):
        """device time of fn(first..last-1): barrier + synchronize on both sides, max over ranks"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(first, last):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    class Leg:
        """One measured configuration: `task_ids` are this rank's tasks out of `n_tasks_total`."""

        def __init__(self, task_ids, n_tasks_total):
            self.ids, self.n_total = list(task_ids), n_tasks_total
            self.host, self.resident = make_data(self.ids)
            self.stepper = (mtl_b200.MetaStepper(s, len(self.ids), n_lanes=args.lanes or len(self.ids),
                                                 use_graph=not args.no_graph) if self.ids else None)
            self.results = torch.zeros(total_steps, max(1, len(self.ids)), 16, device=dev)
            self.host_res = torch.zeros(max(1, len(self.ids)), 16).pin_memory()

        def step(self, i, data):
            # snapshot / reset of the weights (transient_trainer.py:160,237) are folded into the per-lane theta copies
            # inside mtl_meta_tasks; theta itself only changes in the Adam step at the end
            tasks, val = data[i]
            if self.stepper is not None:
                for t, b in enumerate(tasks):
                    self.stepper.load_task(t, *b, n=n_tok)      # H2D (e2e) or D2D (resident) into the static slots
                self.stepper.load_val(*val, n=n_tok)
                res = self.stepper.run(theta, cg, LR, 1.0 / self.n_total, dropout=DROPOUT, seed=i * 64 + rank)
                self.results[i].copy_(res, non_blocking=True)
            else:
                s.zero(cg)                                      # a rank without a task still joins the exchange
            exchange.ran_tasks = self.stepper is not None
            exchange.finish(theta, grad, cg, m, v, adam_state, META_LR)   # the one exchange step + Adam

        def e2e_step(self, i):
            self.step(i, self.host)
            if self.stepper is not None:
                self.host_res.copy_(self.stepper.results, non_blocking=True)
            torch.cuda.current_stream().synchronize()           # the trainer prints the loss every step

        def measure(self):
            for i in range(args.warmup):
                self.step(i, self.resident)
            l0 = lib.mtl_launch_count()
            ms = timed(lambda i: self.step(i, self.resident), args.warmup, total_steps)
            launches = lib.mtl_launch_count() - l0
            for i in range(min(2, args.warmup)):
                self.e2e_step(i)
            ms_e2e = timed(self.e2e_step, args.warmup, total_steps)
            return ms, ms_e2e, launches

    # ---------------- weak scaling (the headline): TASKS_PER_GPU tasks on every rank
    # the clock sampler starts before the warm-up steps (nvidia-smi needs ~0.2 s to produce its first line and the
    # timed region is only ~0.15 s long): its samples cover warm-up + timed region, all under load
    clocks = ClockSampler(local_rank) if rank == 0 else None
    weak = Leg(range(rank * n_local, (rank + 1) * n_local), n_total)
    ms_total, ms_e2e, launches = weak.measure()
    clk = clocks.stop() if clocks else None
    losses = weak.results[args.warmup:, :, 8].mean(dim=1).tolist()
    h2d = sum(t.numel() * t.element_size() for b in weak.host[0][0] for t in b) + \
        sum(t.numel() * t.element_size() for t in weak.host[0][1])
    d2h = n_local * 16 * 4
    graph_caps, graph_reps = s.graph_stats()

    # replicas must hold bit-identical weights after the same sequence of exchanged steps
    chk = torch.stack([theta.double().sum(), theta.view(torch.int32).sum(dtype=torch.int64).double()])
    allchk = [chk.clone() for _ in range(world)]
    if dist is not None:
        dist.all_gather(allchk, chk)
    replicas_identical = all(torch.equal(c, allchk[0]) for c in allchk)
    assert replicas_identical, "replicas diverged: " + str([c.tolist() for c in allchk])

    # ---------------- strong scaling (BASELINE configs[2]: the SAME 3-task meta-batch, one task per GPU)
    strong = None
    if world > 1 and not args.no_strong:
        theta.copy_(theta_init); m.zero_(); v.zero_(); adam_state.zero_()
        leg = Leg(task_shard(TASKS_PER_GPU, rank, world), TASKS_PER_GPU)
        ms_s, ms_s_e2e, _ = leg.measure()
        u3 = TASKS_PER_GPU * (K_TRAIN + K_VALID)
        strong = {"tasks": TASKS_PER_GPU, "value": u3 * args.steps / (ms_s / 1e3), "ms_per_step": ms_s / args.steps,
                  "e2e_value": u3 * args.steps / (ms_s_e2e / 1e3), "unit": UNIT,
                  "gpus_with_a_task": min(world, TASKS_PER_GPU),
                  "note": "total work fixed at the cfg-2 meta-batch (3 tasks): rank r runs tasks r, r + N, ...; one task is one "
                          "dependent chain of 2 passes, so a rank with a single task is bound by that chain's latency"}

    if args.timeline and rank == 0:
        dump_timeline(args.timeline, lambda: [weak.step(i, weak.resident) for i in range(total_steps - 2, total_steps)])

    # ---------------- roofline: kernel families by TIME (CUPTI pass), FLOP-dominant kernel timed alone, whole step
    roof = None
    fam = None
    if not args.no_roofline:
        # every rank runs the two traced steps (they contain the collective); only rank 0 records them
        run2 = lambda: [weak.step(i, weak.resident) for i in range(total_steps - 2, total_steps)]
        if rank == 0:
            fam = kernel_families(run2, 2)
        else:
            run2()
        barrier()
    if rank == 0 and not args.no_roofline:
        roof = roofline_conv_gemm(s, lib, dev, args.gemm_mode)
        pk = _peaks()
        tf_step = STEP_TFLOP_3TASKS * n_local / 3.0                 # this rank's share of the step
        af_step = ATTN_FFN_TFLOP_3TASKS * n_local / 3.0
        ach = tf_step / (ms_total / args.steps / 1e3)
        roof["whole_step"] = {"tflop_per_step_per_gpu": tf_step, "achieved": ach, "peak": pk["bf16_sus"], "unit": "TFLOP/s",
                              "frac": ach / pk["bf16_sus"], "peak_source": pk["src"] + " dense bf16 cuBLAS sustained"}
        if fam:
            a = fam["families"]["attn_ffn"]
            ach_a = af_step / (a["busy_ms_per_step"] / 1e3) if a["busy_ms_per_step"] > 0 else 0.0
            roof["attn_ffn"] = {"tflop_per_step_per_gpu": af_step, "busy_ms_per_step": a["busy_ms_per_step"],
                                "sum_kernel_ms_per_step": a["sum_ms_per_step"], "launches_per_step": a["launches_per_step"],
                                "achieved": ach_a, "peak": pk["bf16_sus"], "unit": "TFLOP/s", "frac": ach_a / pk["bf16_sus"],
                                "how": "attention + FFN + projection FLOPs (SURVEY 8d) over the time at least one kernel of "
                                       "the family (block nn.Linear GEMMs, attention, LayerNorm, bias column sums) is running: "
                                       "union of CUPTI kernel intervals, traced outside the timed region"}
            roof["time_dominant"] = fam["top"]
            roof["kernel_time_shares"] = {k: round(f["share"], 4) for k, f in fam["families"].items()}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()
    gpu_base = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        gpu_base = gpu_eager_baseline()

    if rank == 0:
        val = u_total * args.steps / (ms_total / 1e3)
        out = {
            "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "f32", 1: "tf32", 2: "3xtf32"}[args.gemm_mode], "data": "synthetic",
            "config": workload_config(n_total, world, "weak"),
            "engine": {"gemm_mode": args.gemm_mode, "op_modes": os.environ.get("MTL_OP_MODES", "default: 3xTF32 everywhere "
                       "except the VGG input / weight gradients (single-pass TF32)") if args.gemm_mode == 2 else "session mode",
                       "exchange": args.exchange, "lanes": weak.stepper.n_lanes},
            "e2e": {"value": u_total * args.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "cuda_graph": {"captures": graph_caps, "replays": graph_reps, "lanes": weak.stepper.n_lanes},
            "clocks": clk,
            "replicas": {"world": world, "bit_identical": replicas_identical, "theta_checksum": allchk[0].tolist()},
            "strong": strong,
            "roofline": roof,
            "cpu_baseline": cpu,
            "gpu_eager_baseline": gpu_base,
            "loss_first_last": [losses[0], losses[-1]],
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


FAMILIES = (
    # (family, predicate on (name, grid)) -- first match wins.  The tcgen05 GEMM kernel serves the block nn.Linear layers,
    # the stem / vocabulary projections (grids with >= 30 tiles along M or N) and, under its own name, the implicit convs.
    ("vgg", lambda n, g: n.startswith(("conv3x3", "conv_gemm_tc", "conv1_", "conv_w", "maxpool", "feat_transpose", "im2col"))),
    ("stem_vocab", lambda n, g: n.startswith("gemm_") and max(g[:2] or [0]) >= 30),
    ("attn_ffn", lambda n, g: n.startswith(("gemm_", "attn_", "ln_", "colsum"))),
    ("loss_embed", lambda n, g: n.startswith(("ce_", "embed_", "dec_preprocess", "enc_masks"))),
    ("arena", lambda n, g: n.startswith(("ew_kernel", "sumsq", "clip_", "adam_", "set_u64", "scale_"))),
    ("exchange", lambda n, g: n.startswith("nccl")),
)


def kernel_families(fn, n_steps):
    """CUPTI pass (torch.profiler) over `n_steps` steps OUTSIDE any timed region: kernel time by family, the union of the
    busy intervals of each family, and the time-dominant kernel.  Durations of kernels launched with programmatic
    dependent launch include their wait for the predecessor, so the union (wall time with >= 1 kernel of the family in
    flight) is the meaningful denominator."""
    import collections
    from torch.profiler import ProfilerActivity, profile
    try:
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        path = tempfile.mktemp(suffix=".json")
        prof.export_chrome_trace(path)
        with open(path) as f:
            ev = json.load(f)["traceEvents"]
        os.remove(path)
    except Exception as e:
        return {"error": str(e)[:200]} and None
    per = collections.defaultdict(lambda: [0, 0.0])
    fam_iv = collections.defaultdict(list)
    fam_sum = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for e in ev:
        if e.get("cat") != "kernel":
            continue
        name = e["name"].replace("(anonymous namespace)::", "").replace("void ", "")
        grid = list(e.get("args", {}).get("grid", []))
        short = name.split("(")[0]
        fam = next((f for f, pred in FAMILIES if pred(short, grid)), "other")
        per[short][0] += 1
        per[short][1] += e["dur"]
        fam_iv[fam].append((e["ts"], e["ts"] + e["dur"]))
        fam_sum[fam][0] += 1
        fam_sum[fam][1] += e["dur"]
        total += e["dur"]
    if not total:
        return None

    def union(iv):
        iv.sort()
        busy, end = 0.0, None
        for a, b in iv:
            if end is None or a > end:
                busy += b - a
                end = b
            elif b > end:
                busy += b - end
                end = b
        return busy
    fams = {f: {"share": fam_sum[f][1] / total, "sum_ms_per_step": fam_sum[f][1] / 1e3 / n_steps,
                "busy_ms_per_step": union(fam_iv[f]) / 1e3 / n_steps, "launches_per_step": fam_sum[f][0] // n_steps}
            for f in fam_sum}
    for f, _ in FAMILIES:
        fams.setdefault(f, {"share": 0.0, "sum_ms_per_step": 0.0, "busy_ms_per_step": 0.0, "launches_per_step": 0})
    k, (n, us) = max(per.items(), key=lambda kv: kv[1][1])
    top = {"kernel": k, "share_of_summed_kernel_time": us / total, "launches_per_step": n // n_steps, "avg_us": us / n,
           "note": "time-dominant kernel of the step (CUPTI, traced outside the timed region)"}
    return {"families": fams, "top": top}


def gpu_eager_baseline():
    """Secondary baseline (SURVEY 8d / BASELINE.md 3.5): the reference model with PyTorch's own GPU kernels on this B200
    (tools/gpu_eager_baseline.py, run in its own process so that its allocator / cuDNN state never touches the engine)."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_eager_baseline.py")], capture_output=True,
                           text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line)
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}


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
        f.write("name,grid,block,smem,stream,start_us,dur_us\n")
        for e in ev:
            if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset"):
                name = e["name"].replace(",", ";")[:120]
                a = e.get("args", {})
                grid = "x".join(str(v) for v in a.get("grid", [])) or "-"
                block = "x".join(str(v) for v in a.get("block", [])) or "-"
                f.write("%s,%s,%s,%s,%s,%.3f,%.3f\n" % (name, grid, block, a.get("shared memory", 0),
                                                       a.get("stream", e.get("tid")), e["ts"], e["dur"]))


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
            "traffic": _ncu("conv2_fwd", mode).get("dram_bytes"),
            "tensor_pipe_active_pct_ncu": _ncu("conv2_fwd", mode).get("tensor_pipe_active_pct"),
            "ncu_capture": _ncu("conv2_fwd", mode).get("source"),
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
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-kernel roofline leg (A/B runs)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the PyTorch-on-GPU baseline leg (N = 1 only)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling leg (3 tasks in total)")
    ap.add_argument("--exchange", default="allreduce", choices=["allreduce", "sharded"],
                    help="N > 1: all-reduce + full Adam, or reduce-scatter -> Adam on the 1/N slice -> all-gather")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: one blocking all-reduce after the step instead of "
                    "exchanging region A under the convolution backward")
    ap.add_argument("--no-graph", action="store_true", help="run the meta-step eagerly (no CUDA graph replay)")
    ap.add_argument("--lanes", type=int, default=0, help="concurrent task lanes (default: one per task)")
    ap.add_argument("--timeline", default="", help="after the measurements, trace 2 more steps with torch.profiler "
                    "(CUPTI kernel records) and write name/stream/start/duration rows to this .csv.gz (diagnostic)")
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg4"], help="cfg2 = the BASELINE metric; cfg4 = "
                    "BASELINE configs[3] (enc4/dec6 d768, T = 5000), one task per GPU")
    args = ap.parse_args()
    select_config(args.config)
    if args.config != "cfg2":
        args.no_cpu_baseline = args.no_gpu_baseline = args.no_strong = True     # hour-long on the host / not defined
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
