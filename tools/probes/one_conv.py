"""Ad-hoc (not pytest): a few launches of the conv.2 forward implicit GEMM at cfg-2 size through the C ABI,
for `ncu --set full -k regex:gemm_tc_kernel` captures.  usage: one_conv.py [mode] [reps]"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT]
import torch
from mtl_b200 import lib as L

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
lib = L.get_lib()
dev = torch.device("cuda:0")
B, F, T, Cin, Cout = 8, 161, 101, 64, 64
pv = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
bias = torch.zeros(Cout, device=dev)
wg = torch.empty(Cout, 9 * Cin, device=dev)
for i in range(reps):
    x = torch.randn(B, F, T, Cin, device=dev)
    y = torch.empty(B, F, T, Cout, device=dev)
    L.check(lib.mtl_conv3x3_relu_fwd(mode, pv(x), pv(w), pv(bias), None, pv(wg), pv(y), B, F, T, Cin, Cout, st))
torch.cuda.synchronize()
print("ok", float(y.sum()))
