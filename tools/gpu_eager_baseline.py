#!/usr/bin/env python
"""Secondary baseline (BASELINE.md section 3.5 / SURVEY 8d "the real kernel to beat"): the reference model's meta-step
with PyTorch's own GPU kernels on the same B200 -- the oracle port (oracle/ref_asr.py, the CPU restatement pinned
against the live reference) with every tensor on cuda:0, i.e. cuDNN convolutions + cuBLAS GEMMs + ATen elementwise
kernels, fp32 with TF32 off (the reference's arithmetic) and again with TF32 on, (a) eager, exactly as the reference
would run (Python mask loops, per-parameter optimizer ops, host syncs in Decoder.preprocess), and (b) the whole
meta-step captured in ONE torch.cuda.graph (masks / decoder inputs precomputed on the host, fused foreach optimizer
math), which removes every launch gap and host sync -- the strongest PyTorch-library baseline available here.

This is measurement infrastructure (bench.py's `gpu_eager_baseline` leg); nothing in the product imports it.
Prints one JSON object."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

K_TRAIN = K_VALID = 8
T_FRAMES, L_TOKENS, N_TASKS = 101, 32, 3
LR, META_LR, DROPOUT = 1e-4, 1e-4, 0.1
UNIT = "utterance-passes/s"


def _setup(tf32, benchmark=True):
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cudnn.benchmark = bool(benchmark)


def _batches(cfg, ref_meta, i, dev):
    tasks = [ref_meta.synth_batch(cfg, K_TRAIN, T_FRAMES, L_TOKENS, 1000 * i + t) for t in range(N_TASKS)]
    val = ref_meta.synth_batch(cfg, K_VALID, T_FRAMES, L_TOKENS, 1000 * i + 999)
    mv = lambda b: (b[0].to(dev), [int(v) for v in b[1].tolist()], b[2].to(dev))   # lengths: host ints (Python mask loops read them)
    return [mv(b) for b in tasks], mv(val)


def eager(steps=5, warmup=2, tf32=False):
    """ref_meta.meta_step as is, tensors on the GPU."""
    from oracle import ref_asr, ref_meta
    _setup(tf32)
    dev = torch.device("cuda:0")
    cfg = ref_asr.ModelConfig(dropout=DROPOUT)
    with torch.device(dev):
        p = {k: v.to(dev) for k, v in ref_asr.init_params(cfg, 0).items()}
        bufs = {k: v.to(dev) for k, v in ref_asr.buffers(cfg).items()}
        adam = ref_meta.AdamState()
        data = [_batches(cfg, ref_meta, i, dev) for i in range(steps + warmup)]

        def step(i):
            tasks, val = data[i]
            ref_meta.meta_step(p, adam, cfg, tasks, val, lr=LR, meta_lr=META_LR, train=True, bufs=bufs)
        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(warmup, warmup + steps):
            step(i)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
    return dt


def graphed(steps=20, warmup=3, tf32=False):
    """The same meta-step (3 x [train fwd/bwd at theta0, SGD step, val fwd/bwd at the adapted weights], copy-grad sum,
    Adam) captured once in a torch.cuda.graph and replayed: static inputs, decoder inputs / masks from the host."""
    from oracle import ref_asr, ref_meta
    _setup(tf32, benchmark=os.environ.get("MTL_EAGER_CUDNN_BENCHMARK", "1") == "1")
    dev = torch.device("cuda:0")
    cfg = ref_asr.ModelConfig(dropout=DROPOUT)
    orig_pre = ref_asr.decoder_preprocess
    with torch.device(dev):
        p = {k: v.to(dev) for k, v in ref_asr.init_params(cfg, 0).items()}
        names = list(p)
        bufs = {k: v.to(dev) for k, v in ref_asr.buffers(cfg).items()}
        m = {k: torch.zeros_like(v) for k, v in p.items()}
        v2 = {k: torch.zeros_like(v) for k, v in p.items()}
        tasks, val = _batches(cfg, ref_meta, 0, dev)
        pre = {}
        for b in tasks + [val]:
            si, so = orig_pre(b[2].cpu())
            pre[b[2].data_ptr()] = (si.to(dev), so.to(dev))
        ref_asr.decoder_preprocess = lambda trg: pre[trg.data_ptr()]
        # the row masks are built by Python loops with host scalars (common_layers.py:43-48): cached from the warm-up
        orig_mask, masks = ref_asr.length_row_mask, {}

        def cached_mask(n_rows, lengths, like):
            key = (n_rows, tuple(lengths), like.dtype)
            if key not in masks:
                masks[key] = orig_mask(n_rows, lengths, like)
            return masks[key]
        ref_asr.length_row_mask = cached_mask
        try:
            def fwd_bwd(params, batch, scale):
                leaves = [params[k].detach().requires_grad_(True) for k in names]
                pred, gold, _ = ref_asr.forward(dict(zip(names, leaves)), cfg, batch[0], batch[1], batch[2], bufs=bufs, train=True)
                loss = ref_asr.ce_loss(pred, gold) * scale
                return torch.autograd.grad(loss, leaves, allow_unused=True)

            def meta_step():
                theta0 = [p[k] for k in names]
                cg = [torch.zeros_like(t) for t in theta0]
                for tr in tasks:
                    g = [x if x is not None else torch.zeros_like(t) for x, t in zip(fwd_bwd(p, tr, 1.0), theta0)]
                    adapted = dict(zip(names, torch._foreach_add(theta0, g, alpha=-LR)))
                    gv = [x if x is not None else torch.zeros_like(t) for x, t in zip(fwd_bwd(adapted, val, 1.0 / N_TASKS), theta0)]
                    torch._foreach_add_(cg, g)
                    torch._foreach_add_(cg, gv)
                ml, vl = [m[k] for k in names], [v2[k] for k in names]
                torch._foreach_lerp_(ml, cg, 0.1)
                torch._foreach_mul_(vl, 0.999)
                torch._foreach_addcmul_(vl, cg, cg, value=0.001)
                den = torch._foreach_sqrt(vl)
                torch._foreach_add_(den, 1e-8)
                torch._foreach_addcdiv_(theta0, ml, den, value=-META_LR)

            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    meta_step()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                meta_step()
            for _ in range(warmup):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps / 1e3
        finally:
            ref_asr.decoder_preprocess = orig_pre
            ref_asr.length_row_mask = orig_mask


def run(quick=False):
    u = N_TASKS * (K_TRAIN + K_VALID)
    out = {"unit": UNIT, "what": "oracle port of the reference meta-step on cuda:0 with PyTorch's own kernels (cuDNN convs, cuBLAS "
           "GEMMs, ATen elementwise), cfg-2 synthetic workload, dropout 0.1", "torch": torch.__version__}
    for name, fn, kw in (("eager_fp32", eager, dict(tf32=False)), ("eager_tf32", eager, dict(tf32=True)),
                         ("graph_fp32", graphed, dict(tf32=False)), ("graph_tf32", graphed, dict(tf32=True))):
        if quick and name in ("eager_tf32",):
            continue
        try:
            dt = fn(**kw)
            out[name] = {"ms_per_step": 1e3 * dt, "value": u / dt}
        except Exception as e:                                   # a baseline that cannot run is reported, not hidden
            import traceback
            out[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300]),
                         "where": [l.strip() for l in traceback.format_exc().splitlines() if "File" in l][-4:]}
            if os.environ.get("MTL_EAGER_TRACE"):
                traceback.print_exc()
        try:
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
        except Exception:
            pass
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "graph":          # only the captured variants (fresh process: debugging)
        u = N_TASKS * (K_TRAIN + K_VALID)
        os.environ["MTL_EAGER_TRACE"] = "1"
        try:
            dt = graphed(tf32=False)
            print(json.dumps({"graph_fp32": {"ms_per_step": 1e3 * dt, "value": u / dt}}))
        except Exception:
            import traceback
            traceback.print_exc()
    else:
        print(json.dumps(run()))
