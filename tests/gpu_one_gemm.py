"""Ad-hoc (not pytest): a handful of launches of ONE GEMM through mtl_gemm_repeat, for ncu captures.
usage: gpu_one_gemm.py mode tA tB M N K [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch
from gpu_util import P, dev, lib, ok, stream
mode, tA, tB, M, N, K = map(int, sys.argv[1:7])
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 8
beta = float(sys.argv[8]) if len(sys.argv) > 8 else 0.0
split = int(sys.argv[9]) if len(sys.argv) > 9 else 1
A = torch.randn((K, M) if tA else (M, K), device=dev())
B = torch.randn((N, K) if tB else (K, N), device=dev())
C = torch.zeros(M, N, device=dev())
ok(lib().mtl_gemm_repeat(reps, mode, tA, tB, M, N, K, P(A), A.shape[1], P(B), B.shape[1], beta, P(C), N, split, stream()))
torch.cuda.synchronize()
print("ok", float(C.abs().sum()))
if os.environ.get("MTL_GEMM_DBG", "0") != "0":
    import ctypes
    buf = (ctypes.c_longlong * 160)()
    ok(lib().mtl_debug_gemm_stamps(buf))
    t = list(buf)[:19]
    names = ["entry", "setup done", "producer 1st issue", "mma sees full[0]", "mma committed all", "epilogue sees tmem_full",
             "epilogue done", "after final sync", "first tmem_ld done", "phase 1 done", "staging barrier passed",
             "w2 loop start", "w3 loop start", "w4 loop start", "w5 loop start", "w2 loop done", "w3 loop done", "w4 loop done", "w5 loop done"]
    for n, v in zip(names, t):
        print(f"{n:28s} +{v - t[0]:8d} cycles")
