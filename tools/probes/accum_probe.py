"""Ad-hoc (not pytest): the two-batch accumulation case of tests/test_gpu_parity.py with every conv-gradient
error printed, for A/B runs under MTL_CONV_KW / MTL_BRANCHES."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch
import mtl_b200
from gpu_util import rel_err, spec_of, to_batch
from oracle import ref_asr, ref_meta

cfg = ref_asr.SMALL
p = ref_asr.init_params(cfg, 5)
s = mtl_b200.Session(spec_of(cfg), gemm_mode=int(os.environ.get("MTL_GEMM_MODE", "2")))
b1, b2 = ref_meta.synth_batch(cfg, 4, 41, 7, 1), ref_meta.synth_batch(cfg, 3, 30, 5, 2)
_, g1, *_ = ref_meta.loss_and_grads(p, cfg, b1)
_, g2, *_ = ref_meta.loss_and_grads(p, cfg, b2, 1.0 / 3)
which = sys.argv[1] if len(sys.argv) > 1 else "both"
for rep in range(2):
    theta, grad = s.new_arena(), s.new_arena()
    s.load(theta, p)
    ref = {k: torch.zeros_like(g1[k]) for k in g1}
    if which in ("both", "b1"):
        s.forward(theta, to_batch(b1)); s.backward(theta, grad, 1.0)
        ref = {k: ref[k] + g1[k] for k in g1}
    if which in ("both", "b2"):
        s.forward(theta, to_batch(b2)); s.backward(theta, grad, 1.0 / 3)
        ref = {k: ref[k] + g2[k] for k in g1}
    torch.cuda.synchronize()
    v = s.views(grad)
    errs = {k: rel_err(v[k], ref[k]) for k in g1 if float(ref[k].abs().max()) > 1e-7}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    for k in ("conv.2.weight", "conv.0.weight"):
        d = (v[k].cpu() - ref[k]).abs() / ref[k].abs().max()
        print("  ", k, "err by tap (kh,kw):", [["%.1e" % float(d[:, :, a, b].max()) for b in range(3)] for a in range(3)],
              "by co block of 16:", ["%.1e" % float(d[i:i + 16].max()) for i in range(0, d.shape[0], 16)],
              "by ci block of 16:", ["%.1e" % float(d[:, i:i + 16].max()) for i in range(0, d.shape[1], 16)])
    print(which, "rep", rep, "KW", os.environ.get("MTL_CONV_KW", "-"), "BR", os.environ.get("MTL_BRANCHES", "-"),
          " ".join(f"{k}={e:.2e}" for k, e in worst))
