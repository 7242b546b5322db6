"""GPU probe: which operator of the VGG backward loses precision on REAL cfg-2 data (not on random op-level inputs)?
The oracle (fp64, CPU) provides every intermediate of the front-end and its gradient; each of our operators is then run
alone through the C ABI on the oracle's fp32-rounded inputs and compared with the oracle's result for that operator."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.nn.functional as F

from gpu_util import P, dev, lib, ok, rel_err, stream
from oracle import make_golden as mg
from oracle import ref_asr


def nhwc(t):   # (B,C,F,T) -> (B,F,T,C) contiguous fp32 on the GPU
    return t.detach().permute(0, 2, 3, 1).contiguous().float().to(dev())


def main():
    cfg = ref_asr.CFG2
    p = {k: v.double().requires_grad_(True) for k, v in ref_asr.init_params(cfg, 31).items()}
    x, lens, trg = mg.cfg2_batch(3100, ragged=True)
    torch.set_num_threads(os.cpu_count() or 1)
    bufs = {k: v.double() for k, v in ref_asr.buffers(cfg).items()}
    xs = x.double()
    c1 = F.relu(F.conv2d(xs, p["conv.0.weight"], p["conv.0.bias"], padding=1)); c1.retain_grad()
    z2 = F.conv2d(c1, p["conv.2.weight"], p["conv.2.bias"], padding=1); z2.retain_grad()
    c2 = F.relu(z2); c2.retain_grad()
    p2 = F.max_pool2d(c2, 2, stride=2); p2.retain_grad()
    z3 = F.conv2d(p2, p["conv.5.weight"], p["conv.5.bias"], padding=1); z3.retain_grad()
    c3 = F.relu(z3); c3.retain_grad()
    z4 = F.conv2d(c3, p["conv.7.weight"], p["conv.7.bias"], padding=1); z4.retain_grad()
    c4 = F.relu(z4); c4.retain_grad()
    p4 = F.max_pool2d(c4, 2, stride=2); p4.retain_grad()
    b, ch, fr, t = p4.shape
    feat = p4.reshape(b, ch * fr, t).transpose(1, 2).contiguous()
    enc = ref_asr.encoder_forward(p, cfg, feat, lens, bufs["encoder.positional_encoding.pe"])
    pred, gold = ref_asr.decoder_forward(p, cfg, trg, enc, lens, bufs["decoder.positional_encoding.pe"])
    ref_asr.ce_loss(pred, gold).backward()
    L = lib()
    B, Fq, T = x.shape[0], x.shape[2], x.shape[3]
    F2, T2 = Fq // 2, T // 2

    def conv_bwd(name, xin, dz, w, Cin, Cout, Fd, Td, relu_aux=None, mode=2):
        xg, dyg = nhwc(xin), nhwc(dz)
        wg = w.detach().float().to(dev()).contiguous()
        dw = torch.zeros_like(wg); db = torch.zeros(Cout, device=dev()); dx = torch.empty_like(xg)
        n = int(L.mtl_conv3x3_bwd_scratch_floats(mode, B, Fd, Td, Cin, Cout))
        scratch = torch.zeros(n + 1024, device=dev())
        ok(L.mtl_conv3x3_bwd(mode, P(xg), P(wg), P(dyg), None if relu_aux is None else P(nhwc(relu_aux)), P(dw), P(db), P(dx),
                             P(scratch), B, Fd, Td, Cin, Cout, stream()))
        torch.cuda.synchronize()
        return dw.cpu(), db.cpu(), dx.cpu()

    # conv.7 backward alone: inputs c3, d(z4)
    for mode in (2, 0):
        dw, db, dx = conv_bwd("conv.7", c3, z4.grad, p["conv.7.weight"], 128, 128, F2, T2, relu_aux=c3, mode=mode)
        print("mode %d conv.7: dw %.2e  db %.2e  dx(masked by relu(c3)) %.2e" % (
            mode, rel_err(dw, p["conv.7.weight"].grad), rel_err(db, p["conv.7.bias"].grad),
            rel_err(dx, z3.grad.permute(0, 2, 3, 1))))
        dw, db, dx = conv_bwd("conv.5", p2, z3.grad, p["conv.5.weight"], 64, 128, F2, T2, mode=mode)
        print("mode %d conv.5: dw %.2e  db %.2e  dx %.2e" % (mode, rel_err(dw, p["conv.5.weight"].grad),
                                                          rel_err(db, p["conv.5.bias"].grad), rel_err(dx, p2.grad.permute(0, 2, 3, 1))))
        dw, db, dx = conv_bwd("conv.2", c1, z2.grad, p["conv.2.weight"], 64, 64, Fq, T, relu_aux=c1, mode=mode)
        z1g = (c1.grad * (c1 > 0)).permute(0, 2, 3, 1)
        print("mode %d conv.2: dw %.2e  db %.2e  dx %.2e" % (mode, rel_err(dw, p["conv.2.weight"].grad),
                                                          rel_err(db, p["conv.2.bias"].grad), rel_err(dx, z1g)))
    # pool backward alone
    for name, cc, pg, zg, Fd, Td, Cc in (("pool2", c4, p4.grad, z4.grad, F2, T2, 128), ("pool1", c2, p2.grad, z2.grad, Fq, T, 64)):
        xg, dpg = nhwc(cc), nhwc(pg)
        dxg = torch.empty_like(xg)
        ok(L.mtl_maxpool2_relu_bwd(P(xg), P(dpg), P(dxg), B, Fd, Td, Cc, stream()))
        torch.cuda.synchronize()
        print("%s bwd: %.2e" % (name, rel_err(dxg.cpu(), zg.permute(0, 2, 3, 1))))
    # magnitudes: how much cancellation do the bias sums see?
    for nm, zg in (("conv.7", z4.grad), ("conv.5", z3.grad), ("conv.2", z2.grad)):
        s = zg.sum(dim=(0, 2, 3)); a = zg.abs().sum(dim=(0, 2, 3))
        print("%s bias grad: max|sum| %.3e, max sum|terms| %.3e, ratio %.1f" % (nm, float(s.abs().max()), float(a.max()), float(a.max() / s.abs().max())))


if __name__ == "__main__":
    main()
