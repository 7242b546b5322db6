"""Ad-hoc (not pytest): launches of one VGG conv forward implicit GEMM through the C ABI, for
`ncu --set full -k regex:gemm_tc_kernel` captures and A/B timing.
usage: one_conv.py [mode] [reps] [layer: 1 (161x101, 64->64) | 2 (80x50, 64->128) | 3 (80x50, 128->128)]"""
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
pv = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
bias = torch.zeros(Cout, device=dev)
wg = torch.empty(2, Cout, 9 * Cin, device=dev)
NB = 6
xs = [torch.randn(B, F, T, Cin, device=dev) for _ in range(NB)]
ys = [torch.empty(B, F, T, Cout, device=dev) for _ in range(NB)]
def run(i):
    L.check(lib.mtl_conv3x3_relu_fwd(mode, pv(xs[i % NB]), pv(w), pv(bias), None, pv(wg), pv(ys[i % NB]), B, F, T, Cin, Cout, st))
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
print(f"conv layer {layer} mode {mode} stages_env={os.environ.get('MTL_CONV_STAGES','-')}: {us:.1f} us/launch (incl. weight re-layout), "
      f"{fl / us / 1e6:.1f} TFLOP/s, checksum {float(ys[0].sum()):.3f}")
if os.environ.get("MTL_GEMM_DBG", "0") != "0":
    buf = (C.c_longlong * 160)()
    L.check(lib.mtl_debug_gemm_stamps(buf))
    t = list(buf)
    t0 = min(v for v in t[32:] if v > 0)
    print("k-block  producer-issued  splitter-saw-full  mma-saw-ready  mma-committed   (cycles since first stamp)")
    for it in range(24):
        row = t[32 + 4 * it: 36 + 4 * it]
        if not any(row):
            break
        print(f"{it:7d}  " + "  ".join(f"{(v - t0) if v else -1:15d}" for v in row))
    names = {0: "entry", 1: "setup done", 4: "mma committed all", 5: "epilogue sees tmem_full", 9: "phase 1 done", 10: "staging barrier", 11: "w2 loop start", 12: "w3 loop start", 13: "w4 loop start", 14: "w5 loop start", 15: "w2 loop done", 16: "w3 loop done", 17: "w4 loop done", 18: "w5 loop done", 6: "epilogue done", 7: "final sync"}
    if int(os.environ["MTL_GEMM_DBG"]) == 1:
        for k, nme in names.items():
            print(f"{nme:28s} {t[k] - t0:8d}")
