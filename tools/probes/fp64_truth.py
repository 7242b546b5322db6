"""Probe: how reproducible is the REFERENCE arithmetic itself at cfg 2?  Runs the CPU oracle's forward+backward in fp32
(all threads and 1 thread) and in fp64 on the same ragged cfg-2 batch and prints, per tensor, rel_err = max|a-b|/max|b|
of fp32 against fp64 -- the floor below which "parity with the fp32 CPU path" is not a property of any implementation.
With a GPU it adds this library's gradients (engine modes 2 and 0) against both.
  python tools/probes/fp64_truth.py [--gpu]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch

from oracle import make_golden as mg
from oracle import ref_asr, ref_meta


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)


def main():
    cfg = ref_asr.CFG2
    p = ref_asr.init_params(cfg, 31)
    batch = mg.cfg2_batch(3100, ragged=True)
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    _, g32, gold, hyp, pred32 = ref_meta.loss_and_grads(p, cfg, batch)
    t1 = time.time()
    torch.set_num_threads(1)
    _, g32s, _, _, pred32s = ref_meta.loss_and_grads(p, cfg, batch)
    torch.set_num_threads(os.cpu_count() or 1)
    t2 = time.time()
    p64 = {k: v.double() for k, v in p.items()}
    bufs64 = {k: v.double() for k, v in ref_asr.buffers(cfg).items()}
    _, g64, _, _, pred64 = ref_meta.loss_and_grads(p64, cfg, (batch[0].double(), batch[1], batch[2]), bufs=bufs64)
    print("fp32 %.1fs, fp32 1 thread %.1fs, fp64 %.1fs" % (t1 - t0, t2 - t1, time.time() - t2))
    names = [k for k in g64 if float(g64[k].abs().max()) > 1e-7]
    rows = {"fp32_vs_fp64": {k: rel(g32[k], g64[k]) for k in names},
            "fp32_1thread_vs_fp32": {k: rel(g32s[k], g32[k]) for k in names}}
    print("pred: fp32 vs fp64 %.2e, fp32 1 thread vs fp32 %.2e" % (rel(pred32, pred64), rel(pred32s, pred32)))
    if "--gpu" in sys.argv:
        import mtl_b200
        from gpu_util import spec_of, to_batch
        for mode in (2, 0):
            s = mtl_b200.Session(spec_of(cfg), gemm_mode=mode)
            theta, grad = s.new_arena(), s.new_arena()
            s.load(theta, p)
            out = s.forward(theta, to_batch(batch))
            pred = out["pred"].clone()
            s.backward(theta, grad, 1.0)
            torch.cuda.synchronize()
            gv = s.views(grad)
            rows["mode%d_vs_fp64" % mode] = {k: rel(gv[k], g64[k]) for k in names}
            rows["mode%d_vs_fp32" % mode] = {k: rel(gv[k], g32[k]) for k in names}
            print("pred: mode %d vs fp64 %.2e vs fp32 %.2e" % (mode, rel(pred, pred64), rel(pred, pred32)))
    for title, r in rows.items():
        top = sorted(r.items(), key=lambda kv: -kv[1])[:6]
        print("%-22s worst %.2e  >1e-3: %d  >5e-4: %d   %s" % (title, top[0][1], sum(v > 1e-3 for v in r.values()),
                                                            sum(v > 5e-4 for v in r.values()),
                                                            "  ".join("%s %.1e" % (k, v) for k, v in top)))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fp64_truth.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
