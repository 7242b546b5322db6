"""Every backward / optimizer / feature kernel of the hot path ONCE at its cfg-2 shape, through the operator-level C ABI
and on small allocations (ncu's kernel replay saves and restores the process's device memory around every pass: the
full-pass capture in profiles/r02_c_* took 8 s per launch for that reason).  For `ncu --set full`."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch

from gpu_util import P, dev, lib, ok, stream

L = lib()
d = dev()
r = lambda *s: torch.randn(*s, device=d)
B, Fq, T, M, dm, H, dk, V = 8, 161, 101, 264, 512, 8, 64, 3765
mode = 2
# LayerNorm backward (+ parameter gradients), dropout 0.1
xhat, rstd, gam, dout = r(M, dm), torch.rand(M, device=d) + 0.5, r(dm), r(M, dm)
dy, dres, dg, db = torch.empty(M, dm, device=d), torch.empty(M, dm, device=d), torch.zeros(dm, device=d), torch.zeros(dm, device=d)
ok(L.mtl_ln_bwd(P(dout), P(xhat), P(rstd), P(gam), None, 0.1, 1, 3, P(dy), P(dres), 0, P(dg), P(db), M, dm, stream()))
# short-sequence attention backward (tensor-core fragments), causal + dropout; tiled kernels at T' = 100
for Tq, Tk in ((33, 33), (100, 100)):
    q, k, v = r(B * Tq, H * dk), r(B * Tk, H * dk), r(B * Tk, H * dk)
    keypad = torch.zeros(B, Tk, dtype=torch.uint8, device=d)
    o, lse = torch.empty(B * Tq, H * dk, device=d), torch.empty(B * H * Tq, device=d)
    ok(L.mtl_attn_fwd(P(q), P(k), P(v), P(keypad), B, H, Tq, Tk, dk, 1, 0.1, 5, 2, P(o), P(lse), stream()))
    do = r(B * Tq, H * dk)
    delta = torch.empty(B * H * Tq, device=d)
    dq, dkk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ok(L.mtl_attn_bwd(P(q), P(k), P(v), P(keypad), P(o), P(lse), P(do), B, H, Tq, Tk, dk, 1, 0.1, 5, 2, P(delta), P(dq), P(dkk),
                      P(dv), stream()))
# VGG backward: conv.7 (128 -> 128), conv.5 (64 -> 128), conv.2 (64 -> 64); TF32 gradients = the default policy, and 3xTF32
for (cin, cout, Fd, Td) in ((128, 128, 80, 50), (64, 128, 80, 50), (64, 64, Fq, T)):
    for m in (1, 2):
        x, w, dyc = torch.relu(r(B, Fd, Td, cin)), r(cout, cin, 3, 3) * 0.05, r(B, Fd, Td, cout)
        dw, dbc, dx = torch.zeros_like(w), torch.zeros(cout, device=d), torch.empty(B, Fd, Td, cin, device=d)
        n = int(L.mtl_conv3x3_bwd_scratch_floats(m, B, Fd, Td, cin, cout))
        scr = torch.zeros(n + 1024, device=d)
        ok(L.mtl_conv3x3_bwd(m, P(x), P(w), P(dyc), P(x), P(dw), P(dbc), P(dx), P(scr), B, Fd, Td, cin, cout, stream()))
# pooling backward (both resolutions), conv.0 weight gradient, feature transpose backward, embedding backward
for (Fd, Td, Cc) in ((Fq, T, 64), (80, 50, 128)):
    x, dp = torch.relu(r(B, Fd, Td, Cc)), r(B, Fd // 2, Td // 2, Cc)
    dx = torch.empty_like(x)
    ok(L.mtl_maxpool2_relu_bwd(P(x), P(dp), P(dx), B, Fd, Td, Cc, stream()))
x0, dc1 = r(B, 1, Fq, T), r(B, Fq, T, 64)
dw0, db0 = torch.zeros(64, 1, 3, 3, device=d), torch.zeros(64, device=d)
ok(L.mtl_conv1_wgrad(P(x0), P(dc1), P(dw0), P(db0), B, Fq, T, 64, stream()))
dfeat, dp4 = r(B * 25, 5120), torch.empty(B, 40, 25, 128, device=d)
ok(L.mtl_feat_transpose(P(dfeat), P(dp4), B, 40, 25, 128, 1, stream()))
tok = torch.randint(1, V, (B * 33,), dtype=torch.int32, device=d)
E, pe, dE = r(V, dm), r(2500, dm), torch.zeros(V, dm, device=d)
ok(L.mtl_embed(P(tok), P(E), P(pe), 0.1, 7, 1, None, P(r(B * 33, dm)), P(dE), B, 33, dm, stream()))
# arena operations of the meta-step on the 14 M-float arenas
n = 14022080
a, b2, mm, vv = r(n), r(n), torch.zeros(n, device=d), torch.zeros(n, device=d)
scratch = torch.zeros(1040, device=d)
state = torch.zeros(4, dtype=torch.int32, device=d)
ok(L.mtl_arena_axpy(P(a), P(b2), 1.0, n, stream()))
ok(L.mtl_arena_sgd(P(a), P(b2), 1e-4, n, stream()))
ok(L.mtl_arena_clip(P(b2), n, 400.0, P(scratch), stream()))
ok(L.mtl_arena_adam(P(a), P(b2), P(mm), P(vv), P(state), 1e-4, 0.9, 0.999, 1e-8, n, stream()))
# input features: a 1 s and a 10 s utterance
import scipy.signal.windows
win = torch.from_numpy(scipy.signal.windows.hamming(320)).float().to(d)
for ns in (16000, 160000):
    wav = r(ns) * 0.1
    Tt = 1 + ns // 160
    out = torch.empty(161, Tt, device=d)
    stat = torch.zeros(2, dtype=torch.float64, device=d)
    ok(L.mtl_spectrogram(P(wav), ns, 320, 160, P(win), P(out), Tt, 1, P(stat), stream()))
torch.cuda.synchronize()
print("ok")
