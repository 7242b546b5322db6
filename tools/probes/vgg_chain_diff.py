"""GPU probe: element-wise diff of every VGG front-end intermediate (forward and gradient) of OUR full pass against the
fp64 oracle on the ragged cfg-2 batch -- where do the conv gradients pick up their 1e-3?"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.nn.functional as F

import mtl_b200
from gpu_util import rel_err, spec_of, to_batch
from oracle import make_golden as mg
from oracle import ref_asr


def main():
    mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    ragged = (sys.argv[2] != "full") if len(sys.argv) > 2 else True
    cfg = ref_asr.CFG2
    p0 = ref_asr.init_params(cfg, 31)
    p = {k: v.double().requires_grad_(True) for k, v in p0.items()}
    batch = mg.cfg2_batch(3100, ragged=ragged)
    x, lens, trg = batch
    torch.set_num_threads(os.cpu_count() or 1)
    bufs = {k: v.double() for k, v in ref_asr.buffers(cfg).items()}
    c1 = F.relu(F.conv2d(x.double(), p["conv.0.weight"], p["conv.0.bias"], padding=1)); c1.retain_grad()
    z2 = F.conv2d(c1, p["conv.2.weight"], p["conv.2.bias"], padding=1); z2.retain_grad()
    c2 = F.relu(z2)
    p2 = F.max_pool2d(c2, 2, stride=2); p2.retain_grad()
    z3 = F.conv2d(p2, p["conv.5.weight"], p["conv.5.bias"], padding=1); z3.retain_grad()
    c3 = F.relu(z3)
    z4 = F.conv2d(c3, p["conv.7.weight"], p["conv.7.bias"], padding=1); z4.retain_grad()
    c4 = F.relu(z4)
    p4 = F.max_pool2d(c4, 2, stride=2); p4.retain_grad()
    b, ch, fr, t = p4.shape
    feat = p4.reshape(b, ch * fr, t).transpose(1, 2).contiguous(); feat.retain_grad()
    enc = ref_asr.encoder_forward(p, cfg, feat, lens, bufs["encoder.positional_encoding.pe"])
    pred, gold = ref_asr.decoder_forward(p, cfg, trg, enc, lens, bufs["decoder.positional_encoding.pe"])
    ref_asr.ce_loss(pred, gold).backward()
    z1g = c1.grad * (c1 > 0)
    nh = lambda v: v.detach().permute(0, 2, 3, 1).contiguous()
    ref = {"c1": nh(c1), "c2": nh(c2), "p2": nh(p2), "c3": nh(c3), "c4": nh(c4), "p4": nh(p4), "feat": feat.detach(),
           "dfeat": feat.grad, "dp4": nh(p4.grad), "dc4": nh(z4.grad), "dc3": nh(z3.grad), "dp2": nh(p2.grad),
           "dc2": nh(z2.grad), "dc1": nh(z1g)}
    s = mtl_b200.Session(spec_of(cfg), gemm_mode=mode)
    theta, grad = s.new_arena(), s.new_arena()
    s.load(theta, p0)
    s.forward(theta, to_batch(batch))
    s.backward(theta, grad, 1.0)
    torch.cuda.synchronize()
    ptrs = (C.c_void_p * 16)()
    mtl_b200.lib.check(s.lib.mtl_debug_pass_buffers(s._h, ptrs))
    ws = s._ws
    names = ["c1", "c2", "p2", "c3", "c4", "p4", "feat", "dfeat", "dp4", "dc4", "dc3", "dp2", "dc2", "dc1"]
    print("mode", mode, "ragged", ragged)
    for i, nm in enumerate(names):
        r = ref[nm]
        off = ptrs[i] - ws.data_ptr()
        ours = ws[off:off + r.numel() * 4].view(torch.float32).view(r.shape).cpu().double()
        d = (ours - r).abs()
        mx = float(r.abs().max())
        bad = d > 1e-4 * mx
        print("%-6s rel_err %.2e   elements off by > 1e-4 max: %d of %d   sum|diff|/sum|ref| %.2e" % (
            nm, float(d.max()) / mx, int(bad.sum()), r.numel(), float(d.sum() / r.abs().sum())))
        if nm in ("dc4", "dc2", "dp4") and int(bad.sum()):
            idx = bad.nonzero()[:12]
            for j in idx:
                j = tuple(int(v) for v in j)
                print("        at (b,f,t,c)=%s ours %.4e ref %.4e" % (j, float(ours[j]), float(r[j])))
    gv = s.views(grad)
    for k in ("conv.0.weight", "conv.0.bias", "conv.2.weight", "conv.5.weight", "conv.7.weight", "conv.7.bias"):
        print("%-14s %.2e" % (k, rel_err(gv[k], p[k].grad)))


if __name__ == "__main__":
    main()
