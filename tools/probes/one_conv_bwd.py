"""Ad-hoc (not pytest): launches of one VGG conv backward (weight gradient + bias gradient + input gradient) through the C
ABI, for `ncu --set full -k regex:gemm_tc_kernel` captures of the weight-gradient kernel and A/B timing.
usage: one_conv_bwd.py [mode] [reps] [layer: 1 (161x101, 64->64) | 2 (80x50, 64->128) | 3 (80x50, 128->128)]"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT]
import torch
from mtl_b200 import lib as L

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
layer = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lib = L.get_lib()
dev = torch.device("cuda:0")
B = 8
F, T, Cin, Cout = {1: (161, 101, 64, 64), 2: (80, 50, 64, 128), 3: (80, 50, 128, 128)}[layer]
pv = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
NB = 4
xs = [torch.randn(B, F, T, Cin, device=dev).relu_() for _ in range(NB)]
dys = [torch.randn(B, F, T, Cout, device=dev) for _ in range(NB)]
dxs = [torch.empty(B, F, T, Cin, device=dev) for _ in range(NB)]
dw = torch.zeros(Cout, Cin, 3, 3, device=dev)
db = torch.zeros(Cout, device=dev)
scr = torch.empty(int(lib.mtl_conv3x3_bwd_scratch_floats(mode, B, F, T, Cin, Cout)), device=dev)
def run(i, with_dx=False):
    L.check(lib.mtl_conv3x3_bwd(mode, pv(xs[i % NB]), pv(w), pv(dys[i % NB]), pv(xs[i % NB]), pv(dw), pv(db),
                                pv(dxs[i % NB]) if with_dx else None, pv(scr), B, F, T, Cin, Cout, st))
for i in range(3):
    run(i)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(reps):
    run(i)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
fl = 2.0 * B * F * T * Cout * 9 * Cin
print(f"conv layer {layer} mode {mode}: wgrad + bias grad (zero, wgrad GEMM, scatter, colsum) {us:.1f} us/call, {fl / us / 1e6:.1f} TFLOP/s of wgrad")
