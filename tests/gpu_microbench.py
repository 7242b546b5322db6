"""Ad-hoc GPU timing (not pytest): back-to-back launches of single ops through the C ABI, CUDA-event timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch
from gpu_util import P, dev, lib, ok, stream


def timeit(fn, reps=200):
    for _ in range(10):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def gemm_case(mode, tA, tB, M, N, K, split=1):
    A = torch.randn((K, M) if tA else (M, K), device=dev())
    B = torch.randn((N, K) if tB else (K, N), device=dev())
    C = torch.zeros(M, N, device=dev())
    st = stream()
    fn = lambda: ok(lib().mtl_gemm(mode, tA, tB, M, N, K, 1.0, P(A), A.shape[1], P(B), B.shape[1], 1.0 if split > 1 else 0.0,
                                   P(C), N, None, 0, None, split, st))
    return timeit(fn)


print("GEMM back-to-back us/launch  (mode 0 simt fp32 / 1 tf32 / 2 3xtf32)")
for (tA, tB, M, N, K, split) in [(0, 1, 264, 100, 512, 1), (0, 1, 264, 512, 100, 1), (0, 1, 264, 512, 512, 1),
                                 (0, 1, 200, 512, 5120, 1), (0, 1, 264, 3765, 512, 1), (1, 0, 512, 512, 264, 1),
                                 (1, 0, 512, 5120, 200, 1), (0, 0, 264, 512, 3765, 1), (0, 1, 792, 512, 512, 1),
                                 (0, 1, 4096, 4096, 4096, 1)]:
    r = [gemm_case(m, tA, tB, M, N, K, split) for m in (0, 1, 2)]
    fl = 2.0 * M * N * K
    print(f"tA={tA} tB={tB} M={M:5d} N={N:5d} K={K:5d}: " + "  ".join(f"{t:8.1f}" for t in r) +
          f"   TF/s: " + " ".join(f"{fl / t / 1e6:7.2f}" for t in r))

print("conv3x3 implicit GEMM us/launch (fwd, bwd) modes 1, 2")
for (B, F, T, ci, co) in [(8, 161, 101, 64, 64), (8, 80, 50, 64, 128), (8, 80, 50, 128, 128)]:
    x = torch.relu(torch.randn(B, F, T, ci, device=dev()))
    w = torch.randn(co, ci, 3, 3, device=dev()) * 0.05
    b = torch.zeros(co, device=dev())
    y = torch.empty(B, F, T, co, device=dev())
    dy = torch.randn(B, F, T, co, device=dev())
    wg = torch.empty(co, 9 * ci, device=dev())
    dw, db, dx = torch.zeros_like(w), torch.zeros_like(b), torch.empty_like(x)
    st = stream()
    for mode in (1, 2):
        scr = torch.empty(int(lib().mtl_conv3x3_bwd_scratch_floats(mode, B, F, T, ci, co)), device=dev())
        tf = timeit(lambda: ok(lib().mtl_conv3x3_relu_fwd(mode, P(x), P(w), P(b), None, P(wg), P(y), B, F, T, ci, co, st)), 50)
        tb = timeit(lambda: ok(lib().mtl_conv3x3_bwd(mode, P(x), P(w), P(dy), P(x), P(dw), P(db), P(dx), P(scr), B, F, T, ci, co, st)), 50)
        fl = 2.0 * B * F * T * 9 * ci * co
        print(f"B={B} F={F} T={T} {ci}->{co} mode {mode}: fwd {tf:7.1f} us ({fl / tf / 1e6:6.1f} TF/s)  bwd(wgrad+bias+dgrad) {tb:7.1f} us "
              f"({2 * fl / tb / 1e6:6.1f} TF/s)")
