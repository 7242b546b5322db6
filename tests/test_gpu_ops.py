"""-m gpu: every single-operator C-ABI entry point against the same op in PyTorch (CPU, fp32/fp64).
These are the building blocks of the hot path; shapes cover cfg-2 sizes, ragged tails and odd strides."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_util import GEMM_MODE, P, dev, lib, ok, rel_err, stream
from oracle import ref_asr

pytestmark = pytest.mark.gpu


def _r(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


# ----------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("tA,tB", [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(200, 512, 5120), (264, 100, 512), (264, 3765, 512), (130, 64, 576),
                                   (512, 100, 264), (7, 5, 3), (128, 128, 32), (257, 131, 100)])
def test_gemm_layouts(mode, tA, tB, M, N, K):
    A = _r(K, M, seed=1) if tA else _r(M, K, seed=1)
    B = _r(N, K, seed=2) if tB else _r(K, N, seed=2)
    ref = (A.t() if tA else A).double() @ (B.t() if tB else B).double()
    ldc = (N + 3) // 4 * 4
    Cd = torch.full((M, ldc), 7.0, device=dev())
    Ad, Bd = A.to(dev()), B.to(dev())
    ok(lib().mtl_gemm(mode, tA, tB, M, N, K, 1.0, P(Ad), A.shape[1], P(Bd), B.shape[1], 0.0, P(Cd), ldc,
                      None, 0, None, 1, stream()))
    out = Cd.cpu()
    tol = 2e-5 if mode == 0 else (2e-3 if mode == 1 else 5e-5)
    assert rel_err(out[:, :N], ref) < tol
    if ldc > N:
        assert torch.all(out[:, N:] == 7.0), "padding columns were written"


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gemm_epilogues(mode):
    M, N, K = 264, 512, 512
    A, W, b, C0, aux = _r(M, K, seed=3), _r(N, K, seed=4), _r(N, seed=5), _r(M, N, seed=6), _r(M, N, seed=7)
    Ad, Wd, bd, auxd = A.to(dev()), W.to(dev()), b.to(dev()), aux.to(dev())
    tol = {0: 2e-5, 1: 2e-3, 2: 5e-5}[mode]
    # bias + relu
    Cd = torch.empty(M, N, device=dev())
    ok(lib().mtl_gemm(mode, 0, 1, M, N, K, 1.0, P(Ad), K, P(Wd), K, 0.0, P(Cd), N, P(bd), 1, None, 1, stream()))
    assert rel_err(Cd, F.relu(F.linear(A.double(), W.double(), b.double()))) < tol
    # alpha/beta accumulate
    Cd = C0.to(dev())
    ok(lib().mtl_gemm(mode, 0, 1, M, N, K, 0.5, P(Ad), K, P(Wd), K, 1.0, P(Cd), N, None, 0, None, 1, stream()))
    assert rel_err(Cd, 0.5 * F.linear(A.double(), W.double()) + C0.double()) < tol
    # relu-backward mask
    Cd = torch.empty(M, N, device=dev())
    ok(lib().mtl_gemm(mode, 0, 1, M, N, K, 1.0, P(Ad), K, P(Wd), K, 0.0, P(Cd), N, None, 2, P(auxd), 1, stream()))
    assert rel_err(Cd, F.linear(A.double(), W.double()) * (aux > 0).double()) < tol
    # K slabs with the ReLU-backward mask (a linear epilogue: every slab masks its partial; the masked FFN dgrad of the engine)
    if mode != 0:
        Cd = torch.zeros(M, N, device=dev())
        ok(lib().mtl_gemm(mode, 0, 1, M, N, K, 1.0, P(Ad), K, P(Wd), K, 1.0, P(Cd), N, None, 2, P(auxd), 4, stream()))
        assert rel_err(Cd, F.linear(A.double(), W.double()) * (aux > 0).double()) < tol
    # split-K wgrad-style: C[N,K] += A^T[M,N]^T . X[M,K]
    Mr = 4000
    dy, x = _r(Mr, 64, seed=8), _r(Mr, 576, seed=9)
    Cd = torch.ones(64, 576, device=dev())
    dyd, xd = dy.to(dev()), x.to(dev())
    ok(lib().mtl_gemm(mode, 1, 0, 64, 576, Mr, 1.0, P(dyd), 64, P(xd), 576, 1.0, P(Cd), 576,
                      None, 0, None, 8, stream()))
    assert rel_err(Cd, dy.double().t() @ x.double() + 1.0) < tol


# ----------------------------------------------------------------------------------------- fused low-rank pair
def _ptrs(ts):
    import ctypes as C
    arr = (C.c_void_p * len(ts))(*[None if t is None else t.data_ptr() for t in ts])
    return arr


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("bwd", [0, 1])
@pytest.mark.parametrize("G,M,K1,r,N2,ctas", [(1, 264, 512, 100, 512, 0), (3, 264, 512, 100, 512, 148),
                                              (2, 200, 512, 100, 512, 64), (1, 66, 768, 100, 768, 96),
                                              (1, 7, 64, 20, 36, 0), (3, 130, 100, 128, 68, 1000), (1, 129, 36, 4, 512, 7)])
def test_lowrank_pair(mode, bwd, G, M, K1, r, N2, ctas):
    """y = linear_b(linear_a(x)) (common_layers.py:287-289,303) and its input gradient as ONE kernel, vs fp64."""
    xs = [_r(M, K1, seed=10 + g) for g in range(G)]
    # forward: w1 = A [r, K1], w2 = Bw [N2, r]; backward: w1 = Bw [K1, r], w2 = A [r, N2]
    w1 = [(_r(K1, r, seed=20 + g) if bwd else _r(r, K1, seed=20 + g)) * 0.1 for g in range(G)]
    w2 = [(_r(r, N2, seed=30 + g) if bwd else _r(N2, r, seed=30 + g)) * 0.1 for g in range(G)]
    bias = [None if bwd else _r(N2, seed=40 + g) for g in range(G)]
    y0 = [_r(M, N2, seed=50 + g) if bwd else torch.zeros(M, N2) for g in range(G)]       # backward accumulates into dx
    xd, w1d, w2d = [[t.to(dev()) for t in ts] for ts in (xs, w1, w2)]
    bd = [None if b is None else b.to(dev()) for b in bias]
    ad = [torch.zeros(M, r, device=dev()) for _ in range(G)]
    yd = [t.to(dev()) for t in y0]
    ok(lib().mtl_lowrank_pair(mode, bwd, G, M, K1, r, N2, _ptrs(xd), K1, _ptrs(w1d), _ptrs(w2d), _ptrs(bd), _ptrs(ad),
                              _ptrs(yd), N2, ctas, stream()))
    tol = 2e-3 if mode == 1 else 5e-5
    for g in range(G):
        a_ref = xs[g].double() @ (w1[g].double() if bwd else w1[g].double().t())
        y_ref = a_ref @ (w2[g].double() if bwd else w2[g].double().t()) + y0[g].double()
        if not bwd:
            y_ref = y_ref + bias[g].double()
        assert rel_err(ad[g], a_ref) < tol, ("a", g)
        assert rel_err(yd[g], y_ref) < tol, ("y", g)


def test_lowrank_pair_shared_output_accumulates():
    """q | k | v input gradients of a self-attention land in ONE dx (the three problems reduce-add into it)."""
    M, K1, r, N2 = 264, 512, 100, 512
    xs = [_r(M, K1, seed=60 + g) for g in range(3)]
    w1 = [_r(K1, r, seed=70 + g) * 0.1 for g in range(3)]
    w2 = [_r(r, N2, seed=80 + g) * 0.1 for g in range(3)]
    dx0 = _r(M, N2, seed=90)
    xd, w1d, w2d = [[t.to(dev()) for t in ts] for ts in (xs, w1, w2)]
    ad = [torch.zeros(M, r, device=dev()) for _ in range(3)]
    dx = dx0.to(dev())
    ok(lib().mtl_lowrank_pair(2, 1, 3, M, K1, r, N2, _ptrs(xd), K1, _ptrs(w1d), _ptrs(w2d), None, _ptrs(ad),
                              _ptrs([dx, dx, dx]), N2, 148, stream()))
    ref = dx0.double() + sum(xs[g].double() @ w1[g].double() @ w2[g].double() for g in range(3))
    assert rel_err(dx, ref) < 5e-5


# ----------------------------------------------------------------------------------------- LayerNorm block
@pytest.mark.parametrize("M,d", [(200, 512), (33, 64), (5, 768)])
def test_ln_fwd_bwd(M, d):
    y, res = _r(M, d, seed=1), _r(M, d, seed=2)
    gam, bet = _r(d, seed=3) * 0.1 + 1.0, _r(d, seed=4) * 0.1
    rm = (torch.arange(M) % 5 != 0).float()
    T = 11
    pe = _r(T, d, seed=5)
    dout = _r(M, d, seed=6)
    yd, resd, gd, bd, rmd, ped, doutd = [t.to(dev()) for t in (y, res, gam, bet, rm, pe, dout)]
    out, xhat, rstd = torch.empty(M, d, device=dev()), torch.empty(M, d, device=dev()), torch.empty(M, device=dev())
    ok(lib().mtl_ln_fwd(P(yd), P(resd), P(gd), P(bd), P(rmd), P(ped), T, 0.0, 0, 0, P(out), P(xhat), P(rstd), M, d,
                        stream()))
    yr, rr, gr, br = [t.double().requires_grad_(True) for t in (y, res, gam, bet)]
    ref = (F.layer_norm(yr + rr, (d,), gr, br) + pe.double()[torch.arange(M) % T]) * rm.double().unsqueeze(1)
    assert rel_err(out, ref) < 1e-5
    ref.backward(dout.double())
    dy, dres = torch.empty(M, d, device=dev()), torch.full((M, d), 1.0, device=dev())
    dg, db = torch.zeros(d, device=dev()), torch.zeros(d, device=dev())
    ok(lib().mtl_ln_bwd(P(doutd), P(xhat), P(rstd), P(gd), P(rmd), 0.0, 0, 0, P(dy), P(dres), 1, P(dg), P(db), M, d,
                        stream()))
    assert rel_err(dy, yr.grad) < 2e-5
    assert rel_err(dres, rr.grad + 1.0) < 2e-5
    assert rel_err(dg, gr.grad) < 2e-5
    assert rel_err(db, br.grad) < 2e-5


def test_ln_dropout_mask_is_consistent_between_fwd_and_bwd():
    M, d, p = 64, 512, 0.3
    y = torch.ones(M, d, device=dev())
    gam, bet = torch.ones(d, device=dev()), torch.zeros(d, device=dev())
    out, xhat, rstd = torch.empty(M, d, device=dev()), torch.empty(M, d, device=dev()), torch.empty(M, device=dev())
    ok(lib().mtl_ln_fwd(P(y), None, P(gam), P(bet), None, None, 1, p, 1234, 7, P(out), P(xhat), P(rstd), M, d, stream()))
    # z = mask/(1-p): after LN the dropped elements are the negative ones
    dropped = (xhat < 0)
    frac = float(dropped.float().mean())
    assert abs(frac - p) < 0.02, frac
    dout = torch.randn(M, d, device=dev())
    dy = torch.empty(M, d, device=dev())
    dg, db = torch.zeros(d, device=dev()), torch.zeros(d, device=dev())
    ok(lib().mtl_ln_bwd(P(dout), P(xhat), P(rstd), P(gam), None, p, 1234, 7, P(dy), None, 0, P(dg), P(db), M, d, stream()))
    assert torch.all(dy[dropped] == 0)
    assert float((dy[~dropped] != 0).float().mean()) > 0.99
    # different site -> different mask
    xh2 = torch.empty(M, d, device=dev())
    ok(lib().mtl_ln_fwd(P(y), None, P(gam), P(bet), None, None, 1, p, 1234, 8, P(out), P(xh2), P(rstd), M, d, stream()))
    assert float(((xh2 < 0) != dropped).float().mean()) > 0.2


# ----------------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, keypad, B, H, Tq, Tk, dk, causal):
    q4 = q.view(B, Tq, H, dk).permute(0, 2, 1, 3)
    k4 = k.view(B, Tk, H, dk).permute(0, 2, 1, 3)
    v4 = v.view(B, Tk, H, dk).permute(0, 2, 1, 3)
    s = q4 @ k4.transpose(-1, -2) / math.sqrt(dk)
    mask = keypad.bool().view(B, 1, 1, Tk).expand(B, H, Tq, Tk).clone()
    if causal:
        mask |= torch.triu(torch.ones(Tq, Tk, dtype=torch.bool), diagonal=1)
    s = s.masked_fill(mask, float("-inf"))
    o = torch.softmax(s, dim=-1) @ v4
    return o.permute(0, 2, 1, 3).reshape(B * Tq, H * dk)


@pytest.mark.parametrize("B,H,Tq,Tk,dk,causal", [(8, 8, 25, 25, 64, 0), (8, 8, 33, 33, 64, 1), (8, 8, 33, 25, 64, 0),
                                                 (2, 2, 70, 70, 32, 1), (2, 3, 45, 100, 64, 0), (1, 2, 130, 130, 64, 1),
                                                 (2, 2, 64, 64, 32, 1), (3, 2, 1, 7, 64, 0), (2, 4, 48, 17, 64, 0),
                                                 (2, 2, 17, 64, 32, 0), (1, 2, 1250, 1250, 64, 0), (1, 2, 257, 1250, 64, 0),
                                                 (1, 2, 301, 301, 64, 1), (2, 2, 65, 200, 32, 0)])
def test_attention_fwd_bwd(B, H, Tq, Tk, dk, causal):
    q, k, v = _r(B * Tq, H * dk, seed=1), _r(B * Tk, H * dk, seed=2), _r(B * Tk, H * dk, seed=3)
    keypad = torch.zeros(B, Tk, dtype=torch.uint8)
    for b in range(B):
        keypad[b, Tk - (b % 4) * 3:] = 1 if b % 4 else 0
    d_o = _r(B * Tq, H * dk, seed=4)
    qd, kd, vd, kpd, dod = [t.to(dev()) for t in (q, k, v, keypad, d_o)]
    o = torch.empty(B * Tq, H * dk, device=dev())
    lse = torch.empty(B * H * Tq, device=dev())
    ok(lib().mtl_attn_fwd(P(qd), P(kd), P(vd), P(kpd), B, H, Tq, Tk, dk, causal, 0.0, 0, 0, P(o), P(lse), stream()))
    qr, kr, vr = [t.double().requires_grad_(True) for t in (q, k, v)]
    ref = _attn_ref(qr, kr, vr, keypad, B, H, Tq, Tk, dk, causal)
    assert rel_err(o, ref) < (5e-5 if max(Tq, Tk) > 256 else 2e-5)      # 1250-key softmax rows: 5e-5
    ref.backward(d_o.double())
    dq, dk_, dv = torch.empty_like(qd), torch.empty_like(kd), torch.empty_like(vd)
    delta = torch.empty(B * H * Tq, device=dev())
    ok(lib().mtl_attn_bwd(P(qd), P(kd), P(vd), P(kpd), P(o), P(lse), P(dod), B, H, Tq, Tk, dk, causal, 0.0, 0, 0,
                          P(delta), P(dq), P(dk_), P(dv), stream()))
    assert rel_err(dq, qr.grad) < 5e-5
    assert rel_err(dk_, kr.grad) < 5e-5
    assert rel_err(dv, vr.grad) < 5e-5


def test_attention_dropout_is_a_consistent_linear_map():
    """With dropout the op is o = (P*mask/keep) V for a FIXED mask: check fwd/bwd agree via <dO, o> adjointness
    and that E[o] ~ undropped o."""
    _attention_dropout_check(33)


def test_long_attention_dropout_is_a_consistent_linear_map():
    """Same properties for the flash kernels (T > 64), whose forward / dQ / dKV kernels each rebuild the Philox mask."""
    _attention_dropout_check(100)


def _attention_dropout_check(T):
    B, H, dk, p = 2, 4, 64, 0.25
    q, k, v = [_r(B * T, H * dk, seed=s).to(dev()) for s in (1, 2, 3)]
    kp = torch.zeros(B, T, dtype=torch.uint8, device=dev())
    o0, o1, lse = torch.empty_like(q), torch.empty_like(q), torch.empty(B * H * T, device=dev())
    ok(lib().mtl_attn_fwd(P(q), P(k), P(v), P(kp), B, H, T, T, dk, 1, 0.0, 0, 0, P(o0), P(lse), stream()))
    acc = torch.zeros_like(q)
    n_rep = 64
    for s in range(n_rep):
        ok(lib().mtl_attn_fwd(P(q), P(k), P(v), P(kp), B, H, T, T, dk, 1, p, 99, s, P(o1), P(lse), stream()))
        acc += o1
    assert rel_err(acc / n_rep, o0) < 0.25
    # adjoint test on V (o is linear in V for fixed mask): <dO, o(V)> == <dV, V>
    d_o = _r(B * T, H * dk, seed=5).to(dev())
    dq, dk_, dv, delta = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q), torch.empty(B * H * T, device=dev())
    ok(lib().mtl_attn_fwd(P(q), P(k), P(v), P(kp), B, H, T, T, dk, 1, p, 99, 3, P(o1), P(lse), stream()))
    ok(lib().mtl_attn_bwd(P(q), P(k), P(v), P(kp), P(o1), P(lse), P(d_o), B, H, T, T, dk, 1, p, 99, 3, P(delta), P(dq),
                          P(dk_), P(dv), stream()))
    lhs, rhs = float((d_o.double() * o1.double()).sum()), float((dv.double() * v.double()).sum())
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))
    # dQ / dK of the dropped map against finite differences of <dO, o> along random directions (fixed mask = fixed seed / site)
    for which, grad in (("q", dq), ("k", dk_)):
        dirn = _r(B * T, H * dk, seed=11).to(dev())
        eps = 1e-2
        vals = []
        for sgn in (1.0, -1.0):
            qq = q + sgn * eps * dirn if which == "q" else q
            kk = k + sgn * eps * dirn if which == "k" else k
            ok(lib().mtl_attn_fwd(P(qq), P(kk), P(v), P(kp), B, H, T, T, dk, 1, p, 99, 3, P(o1), P(lse), stream()))
            vals.append(float((d_o.double() * o1.double()).sum()))
        fd = (vals[0] - vals[1]) / (2 * eps)
        an = float((grad.double() * dirn.double()).sum())
        assert abs(fd - an) < 2e-3 * max(1.0, abs(an)), (which, fd, an)


# ----------------------------------------------------------------------------------------- CE + argmax
@pytest.mark.parametrize("smoothing", [0.0, 0.1])
def test_ce_fwd_bwd(smoothing):
    M, V = 264, 3765
    ld = (V + 3) // 4 * 4
    logits = _r(M, V, seed=1, scale=3.0)
    gold = torch.randint(4, V, (M,), generator=torch.Generator().manual_seed(2))
    gold[::7] = 0
    lg = torch.zeros(M, ld)
    lg[:, :V] = logits
    lgd, goldd = lg.to(dev()), gold.int().to(dev())
    row_lse, row_loss = torch.empty(M, device=dev()), torch.empty(M, device=dev())
    hyp, out8 = torch.empty(M, dtype=torch.int32, device=dev()), torch.empty(8, device=dev())
    ok(lib().mtl_ce_fwd(P(lgd), ld, P(goldd), M, V, smoothing, P(row_lse), P(row_loss), P(hyp), P(out8), stream()))
    lr = logits.double().requires_grad_(True)
    ref = ref_asr.ce_loss(lr.view(1, M, V), gold.view(1, M), smoothing)
    assert abs(float(out8[0]) - float(ref)) < 1e-5 * abs(float(ref))
    assert int(out8[1]) == int((gold != 0).sum())
    assert torch.equal(hyp.cpu().long(), logits.argmax(dim=1))
    assert int(out8[2]) == ref_asr.num_correct(logits.view(1, M, V), gold.view(1, M))
    (ref * 0.5).backward()
    dl = torch.full((M, ld), 9.0, device=dev())
    ok(lib().mtl_ce_bwd(P(lgd), ld, P(goldd), P(row_lse), P(out8), 0.5, smoothing, P(dl), M, V, stream()))
    assert rel_err(dl[:, :V], lr.grad) < 1e-5
    assert torch.all(dl[:, V:] == 0)


# ----------------------------------------------------------------------------------------- VGG pieces
def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


CONV_TOL = {0: 2e-5, 1: 2e-3, 2: 5e-5}


@pytest.mark.parametrize("B,F4,T4,C", [(8, 40, 25, 128), (2, 5, 3, 128), (1, 40, 312, 128), (3, 7, 9, 64)])
def test_feat_transpose_roundtrip(B, F4, T4, C):
    """(B, C, F', T') -> view(B, C*F', T').transpose(1, 2) (models/asr/transformer.py:136-138) on the NHWC activation,
    and its gradient: exact copies, compared bit for bit."""
    p4 = _r(B, F4, T4, C, seed=4)                                     # NHWC: [B, F', T', C]
    ref = p4.permute(0, 3, 1, 2).reshape(B, C * F4, T4).transpose(1, 2).contiguous()   # [B, T', C*F' ] index c*F'+f
    p4d = p4.to(dev())
    feat = torch.empty(B, T4, C * F4, device=dev())
    ok(lib().mtl_feat_transpose(P(p4d), P(feat), B, F4, T4, C, 0, stream()))
    assert torch.equal(feat.cpu(), ref)
    back = torch.empty(B, F4, T4, C, device=dev())
    ok(lib().mtl_feat_transpose(P(feat), P(back), B, F4, T4, C, 1, stream()))
    assert torch.equal(back.cpu(), p4)


@pytest.mark.parametrize("B,Fq,T", [(2, 21, 19), (2, 161, 101), (1, 4, 7)])
def test_conv1_fwd(B, Fq, T):
    x = _r(B, 1, Fq, T, seed=1)
    w1, b1 = _r(64, 1, 3, 3, seed=2) * 0.3, _r(64, seed=3) * 0.1
    out = torch.empty(B, Fq, T, 64, device=dev())
    xd, w1d, b1d = x.to(dev()), w1.to(dev()), b1.to(dev())
    ok(lib().mtl_conv1_fwd(P(xd), P(w1d), P(b1d), P(out), B, Fq, T, 64, stream()))
    ref1 = F.relu(F.conv2d(x.double(), w1.double(), b1.double(), padding=1))
    assert rel_err(out, _nhwc(ref1)) < 1e-5


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("masked", [True, False])
@pytest.mark.parametrize("B,Fq,T,cin,cout", [(2, 21, 19, 64, 64), (2, 21, 19, 64, 128), (1, 9, 140, 128, 128),
                                             (2, 161, 101, 64, 64), (2, 80, 50, 128, 128), (3, 5, 4, 64, 128),
                                             (4, 21, 41, 64, 64), (4, 10, 20, 64, 128), (4, 10, 20, 128, 128)])
def test_conv3x3_fwd_bwd(mode, B, Fq, T, cin, cout, masked):
    """conv.2 / conv.5 / conv.7 (models/asr/transformer.py:51-57): forward with fused bias+ReLU, and the three
    backward contractions (weight grad, bias grad, input grad with the upstream ReLU mask) against autograd.
    Modes 1/2 are the tcgen05 implicit GEMMs (tap-shifted 4-D TMA boxes, zero padding = TMA OOB fill)."""
    seed = 100 + Fq + T + cin + cout
    xin = F.relu(_r(B, cin, Fq, T, seed=seed))                     # post-ReLU activation of the previous layer
    w, b = _r(cout, cin, 3, 3, seed=seed + 1) * 0.05, _r(cout, seed=seed + 2) * 0.1
    xr, wr, br = xin.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    pre = F.conv2d(xr, wr, br, padding=1)
    dy = _r(B, cout, Fq, T, seed=seed + 3)
    pre.backward(dy.double())
    tol = CONV_TOL[mode]
    # forward
    Pn = B * Fq * T
    col = torch.empty(Pn, 9 * cin, device=dev()) if mode == 0 else None
    wg = torch.empty(2, cout, 9 * cin, device=dev())              # [hi | lo] halves in 3xTF32
    o = torch.full((B, Fq, T, cout), 7.0, device=dev())
    xind, wd_, bd_ = _nhwc(xin).to(dev()), w.to(dev()), b.to(dev())
    ok(lib().mtl_conv3x3_relu_fwd(mode, P(xind), P(wd_), P(bd_), P(col), P(wg), P(o), B, Fq, T, cin, cout, stream()))
    assert rel_err(o, _nhwc(F.relu(pre.detach()))) < tol
    # backward (dx masked by the ReLU of the layer below: aux = its post-ReLU output = xin)
    n_scr = int(lib().mtl_conv3x3_bwd_scratch_floats(mode, B, Fq, T, cin, cout))
    scr = torch.empty(n_scr, device=dev())
    dw = torch.ones(cout, cin, 3, 3, device=dev())                 # accumulation semantics: += on top of ones
    db = torch.ones(cout, device=dev())
    dx = torch.full((B, Fq, T, cin), 7.0, device=dev())
    dyd = _nhwc(dy).to(dev())
    ok(lib().mtl_conv3x3_bwd(mode, P(xind), P(wd_), P(dyd), P(xind) if masked else None, P(dw), P(db), P(dx), P(scr),
                             B, Fq, T, cin, cout, stream()))
    assert rel_err(dw - 1.0, wr.grad) < tol
    assert rel_err(db - 1.0, br.grad) < max(tol, 2e-5)
    assert rel_err(dx, _nhwc(xr.grad * (xin > 0) if masked else xr.grad)) < tol


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("B,Fq,T,cin,cout", [(2, 21, 19, 64, 64), (2, 161, 101, 64, 64), (2, 80, 50, 128, 128), (1, 9, 140, 128, 128),
                                             (3, 5, 4, 64, 128), (2, 32, 16, 64, 64), (1, 33, 17, 64, 128)])
def test_conv3x3_relu_pool_fused(mode, B, Fq, T, cin, cout):
    """conv -> ReLU -> MaxPool2d(2, 2) (models/asr/transformer.py:49-51,56-58) from ONE kernel: the pooled tensor must be
    bit-identical to pooling the kernel's own un-pooled output (floor pooling: odd last rows / columns dropped)."""
    x = torch.relu(_r(B, cin, Fq, T, seed=1))
    w, b = _r(cout, cin, 3, 3, seed=2) * 0.05, _r(cout, seed=3) * 0.1
    xd, wd, bd = _nhwc(x).contiguous().to(dev()), w.to(dev()), b.to(dev())
    wg = torch.empty(2 * cout * 9 * cin + 64, device=dev())
    out = torch.full((B, Fq, T, cout), 7.0, device=dev())
    pool = torch.full((B, Fq // 2, T // 2, cout), 7.0, device=dev())
    ok(lib().mtl_conv3x3_relu_pool_fwd(mode, P(xd), P(wd), P(bd), P(wg), P(out), P(pool), B, Fq, T, cin, cout, stream()))
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1))
    assert rel_err(out, _nhwc(ref)) < CONV_TOL[mode]
    own = F.max_pool2d(out.permute(0, 3, 1, 2), 2, stride=2).permute(0, 2, 3, 1).contiguous()
    assert torch.equal(pool, own)


@pytest.mark.parametrize("Fq,T", [(21, 19), (20, 18), (161, 101)])
def test_maxpool_fwd_and_relu_pool_bwd(Fq, T):
    B, Cc = 2, 64
    pre = _r(B, Cc, Fq, T, seed=1)
    pr = pre.double().requires_grad_(True)
    act = F.relu(pr)
    pooled = F.max_pool2d(act, 2, stride=2)
    g = _r(*pooled.shape, seed=2)
    pooled.backward(g.double())
    xd = _nhwc(act.detach().float()).to(dev())
    out = torch.empty(B, Fq // 2, T // 2, Cc, device=dev())
    ok(lib().mtl_maxpool2_fwd(P(xd), P(out), B, Fq, T, Cc, stream()))
    assert torch.equal(out.cpu(), _nhwc(pooled.detach().float()))
    dx = torch.empty(B, Fq, T, Cc, device=dev())
    gd = _nhwc(g).to(dev())
    ok(lib().mtl_maxpool2_relu_bwd(P(xd), P(gd), P(dx), B, Fq, T, Cc, stream()))
    assert rel_err(dx, _nhwc(pr.grad)) < 1e-6


def test_decoder_preprocess_matches_reference_semantics():
    trg = torch.tensor([[5, 6, 7, 0, 0], [9, 0, 8, 4, 11], [0, 0, 0, 0, 12], [4, 5, 6, 7, 8]], dtype=torch.int64)
    si_ref, so_ref = ref_asr.decoder_preprocess(trg)
    B, L = trg.shape
    n = si_ref.shape[1]
    si, so = torch.empty(B, n, dtype=torch.int32, device=dev()), torch.empty(B, n, dtype=torch.int32, device=dev())
    rm, kp = torch.empty(B, n, device=dev()), torch.empty(B, n, dtype=torch.uint8, device=dev())
    trgd = trg.to(dev())
    ok(lib().mtl_dec_preprocess(P(trgd), B, L, n, P(si), P(so), P(rm), P(kp), stream()))
    assert torch.equal(si.cpu().long(), si_ref)
    assert torch.equal(so.cpu().long(), so_ref)
    assert torch.equal(rm.cpu(), (si_ref != 2).float())
    assert torch.equal(kp.cpu().bool(), si_ref == 2)


# ----------------------------------------------------------------------------------------- arena / optimizers
def test_arena_ops_match_torch_optim():
    n = 1_000_003 // 4 * 4 + 4
    s = __import__("mtl_b200").Session(__import__("gpu_util").spec_of(ref_asr.SMALL))
    g = torch.Generator().manual_seed(0)
    w = torch.randn(n, generator=g)
    p_ref = torch.nn.Parameter(w.clone())
    opt = torch.optim.Adam([p_ref], lr=3e-3)
    p, m, v = w.to(dev()), torch.zeros(n, device=dev()), torch.zeros(n, device=dev())
    st = s.new_adam_state()
    for it in range(4):
        gr = torch.randn(n, generator=g) * (10.0 ** -it)
        p_ref.grad = gr.clone()
        opt.step()
        grd = gr.to(dev())
        s.adam(p, grd, m, v, st, 3e-3)
        assert torch.allclose(p.cpu(), p_ref.detach(), rtol=1e-6, atol=1e-7), (it, float((p.cpu() - p_ref.detach()).abs().max()))
    assert int(st[0]) == 4
    # SGD, axpy, copy, zero
    gr = torch.randn(n, generator=g)
    grd = gr.to(dev())
    s.sgd(p, grd, 0.1)
    assert torch.allclose(p.cpu(), p_ref.detach() - 0.1 * gr, rtol=0, atol=1e-6)
    y = torch.ones(n, device=dev())
    s.axpy(y, grd, 2.0)
    assert torch.allclose(y.cpu(), 1 + 2 * gr, rtol=0, atol=1e-6)
    z = torch.empty(n, device=dev())
    s.copy(z, y)
    assert torch.equal(z, y)
    s.zero(z)
    assert float(z.abs().max()) == 0.0
    # clip_grad_norm_
    gd = gr.to(dev()).clone()
    out = s.clip(gd, 0.01)
    pr = torch.nn.Parameter(torch.zeros(n))
    pr.grad = gr.clone()
    total = torch.nn.utils.clip_grad_norm_([pr], 0.01)
    assert abs(float(out[0]) - float(total)) < 1e-5 * float(total)
    assert torch.allclose(gd.cpu(), pr.grad, rtol=1e-5, atol=1e-10)
