"""Ad-hoc: phase stamps + back-to-back timing of the short-sequence attention kernels at cfg-2 shapes."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch
from gpu_util import P, dev, lib, ok, stream

B, H, dk = 8, 8, 64
for (Tq, Tk, causal, p) in [(33, 33, 1, 0.1), (33, 25, 0, 0.1), (25, 25, 0, 0.0)]:
    q, k, v, d_o = [torch.randn(B * T, H * dk, device=dev()) for T in (Tq, Tk, Tk, Tq)]
    kp = torch.zeros(B, Tk, dtype=torch.uint8, device=dev())
    o = torch.empty_like(q); lse = torch.empty(B * H * Tq, device=dev()); delta = torch.empty_like(lse)
    dq, dk_, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    fwd = lambda: ok(lib().mtl_attn_fwd(P(q), P(k), P(v), P(kp), B, H, Tq, Tk, dk, causal, p, 99, 3, P(o), P(lse), stream()))
    bwd = lambda: ok(lib().mtl_attn_bwd(P(q), P(k), P(v), P(kp), P(o), P(lse), P(d_o), B, H, Tq, Tk, dk, causal, p, 99, 3,
                                        P(delta), P(dq), P(dk_), P(dv), stream()))
    for f, name in ((fwd, "fwd"), (bwd, "bwd")):
        for _ in range(5):
            f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            f()
        e1.record()
        torch.cuda.synchronize()
        print(f"Tq={Tq} Tk={Tk} causal={causal} p={p} {name}: {e0.elapsed_time(e1) * 5:.2f} us/launch (stream, back to back)")
    buf = (C.c_longlong * 32)()
    ok(lib().mtl_debug_attn_stamps(buf))
    t = list(buf)
    print("   fwd phases (cycles): load %d  scores %d  softmax %d  pv+store %d" % (t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3]))
    print("   bwd phases (cycles): load %d  delta %d  scores+dp %d  dq/dk/dv+store %d" % (t[17] - t[16], t[18] - t[17], t[19] - t[18], t[20] - t[19]))
