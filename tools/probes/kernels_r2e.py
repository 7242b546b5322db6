"""Round-2e kernels ONCE at their cfg-2 / cfg-5 shapes through the operator-level C ABI (for `ncu --set full`): the fused
low-rank projection pair (forward q|k|v group, forward single, backward group), conv.0 forward with register weights,
the tiled feature transposes, the kw-box weight gradients with the four-stage TF32 pipeline (conv.2 / conv.5 / conv.7
shapes), and -- with `lm` as argv[1] -- one LSTM language-model pass (recurrent step kernels)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch

import mtl_b200
from gpu_util import P, dev, lib, ok, stream

L = lib()
d = dev()
r = lambda *s: torch.randn(*s, device=d)


def ptrs(ts):
    return (C.c_void_p * len(ts))(*[None if t is None else t.data_ptr() for t in ts])


if len(sys.argv) > 1 and sys.argv[1] == "lm":
    from oracle import ref_lm
    cfg = ref_lm.LM_CFG5
    s = mtl_b200.LmSession(mtl_b200.LmSpec(cfg.vocab, cfg.ninp, cfg.nhid, cfg.nlayers), d)
    theta, grad = s.new_arena(), s.new_arena()
    s.load(theta, ref_lm.init_params(cfg, 1))
    (blk,), _ = ref_lm.synth_blocks(cfg, 1, 35, 20, 5)
    s.run(theta, blk[0], blk[1], grad=grad, dropout=0.2, seed=3)
    torch.cuda.synchronize()
    print("ok lm")
    sys.exit(0)

M, K1, rk, N2 = 264, 512, 100, 512
for (G, bwd, ctas) in ((3, 0, 148), (1, 0, 148), (3, 1, 148)):
    xs = [r(M, K1) for _ in range(G)]
    w1 = [(r(K1, rk) if bwd else r(rk, K1)) * 0.1 for _ in range(G)]
    w2 = [(r(rk, N2) if bwd else r(N2, rk)) * 0.1 for _ in range(G)]
    bias = [None if bwd else r(N2) for _ in range(G)]
    a = [torch.zeros(M, rk, device=d) for _ in range(G)]
    y = [torch.zeros(M, N2, device=d) for _ in range(G)]
    ok(L.mtl_lowrank_pair(2, bwd, G, M, K1, rk, N2, ptrs(xs), K1, ptrs(w1), ptrs(w2), ptrs(bias), ptrs(a), ptrs(y), N2, ctas, stream()))
B, Fq, T = 8, 161, 101
x0, w0, b0 = r(B, 1, Fq, T), r(64, 1, 3, 3) * 0.3, r(64) * 0.1
c1 = torch.empty(B, Fq, T, 64, device=d)
ok(L.mtl_conv1_fwd(P(x0), P(w0), P(b0), P(c1), B, Fq, T, 64, stream()))
p4, feat = r(B, 40, 25, 128), torch.empty(B, 25, 5120, device=d)
ok(L.mtl_feat_transpose(P(p4), P(feat), B, 40, 25, 128, 0, stream()))
ok(L.mtl_feat_transpose(P(feat), P(p4), B, 40, 25, 128, 1, stream()))
for (cin, cout, Fd, Td) in ((128, 128, 80, 50), (64, 128, 80, 50), (64, 64, Fq, T)):
    x, w, dyc = torch.relu(r(B, Fd, Td, cin)), r(cout, cin, 3, 3) * 0.05, r(B, Fd, Td, cout)
    dw, dbc = torch.zeros_like(w), torch.zeros(cout, device=d)
    n = int(L.mtl_conv3x3_bwd_scratch_floats(1, B, Fd, Td, cin, cout))
    scr = torch.zeros(n + 1024, device=d)
    ok(L.mtl_conv3x3_bwd(1, P(x), P(w), P(dyc), P(x), P(dw), P(dbc), None, P(scr), B, Fd, Td, cin, cout, stream()))
torch.cuda.synchronize()
print("ok")
