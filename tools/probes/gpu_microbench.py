"""Ad-hoc GPU timing (not pytest): device-side time per launch of single GEMMs (mtl_gemm_repeat: the launch loop
runs inside the library, so Python/ctypes overhead is out of the picture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch
from gpu_util import P, dev, lib, ok, stream


def gemm_us(mode, tA, tB, M, N, K, reps=300):
    A = torch.randn((K, M) if tA else (M, K), device=dev())
    B = torch.randn((N, K) if tB else (K, N), device=dev())
    C = torch.zeros(M, N, device=dev())
    st = stream()
    run = lambda r: ok(lib().mtl_gemm_repeat(r, mode, tA, tB, M, N, K, P(A), A.shape[1], P(B), B.shape[1], 0.0, P(C), N, 1, st))
    run(20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); run(reps); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print("cluster split-K:", os.environ.get("MTL_CLUSTER_SPLITK", "1"))
print("us per launch, back to back in one stream      mode1(tf32)  mode2(3xtf32)")
for (tA, tB, M, N) in [(0, 1, 264, 512), (0, 1, 128, 128), (0, 1, 2048, 512), (1, 0, 512, 512)]:
    for K in (32, 128, 512, 2048, 5120):
        r = [gemm_us(m, tA, tB, M, N, K) for m in (1, 2)]
        print(f"tA={tA} tB={tB} M={M:5d} N={N:5d} K={K:5d}: {r[0]:9.2f} {r[1]:9.2f}")
