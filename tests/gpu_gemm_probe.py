"""Ad-hoc GPU probe (not pytest): one tcgen05 GEMM case per process so a trap cannot poison later cases.
usage: gpu_gemm_probe.py tA tB M N K [split]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch
from gpu_util import P, dev, lib, ok, stream

tA, tB, M, N, K = map(int, sys.argv[1:6])
split = int(sys.argv[6]) if len(sys.argv) > 6 else 1
g = torch.Generator().manual_seed(1)
A = torch.randn((K, M) if tA else (M, K), generator=g)
B = torch.randn((N, K) if tB else (K, N), generator=g)
# positive operands expose a truncation bias (every product would be under-estimated)
Ap, Bp = A.abs() + 0.5, B.abs() + 0.5
for name, (a, b) in {"randn": (A, B), "positive": (Ap, Bp)}.items():
    ref = (a.t() if tA else a).double() @ (b.t() if tB else b).double()
    ldc = (N + 3) // 4 * 4
    Cd = torch.zeros(M, ldc, device=dev()) if split > 1 else torch.full((M, ldc), 7.0, device=dev())
    ad, bd = a.to(dev()), b.to(dev())
    ok(lib().mtl_gemm(1, tA, tB, M, N, K, 1.0, P(ad), a.shape[1], P(bd), b.shape[1], 1.0 if split > 1 else 0.0, P(Cd), ldc,
                      None, 0, None, split, stream()))
    torch.cuda.synchronize()
    out = Cd.cpu()[:, :N].double()
    err = (out - ref)
    rel = float(err.abs().max() / ref.abs().max())
    bias = float((err / ref).mean()) if name == "positive" else float("nan")
    pad_ok = bool(torch.all(Cd.cpu()[:, N:] == (0.0 if split > 1 else 7.0))) if ldc > N else True
    print(f"tA={tA} tB={tB} M={M} N={N} K={K} split={split} {name}: rel_err={rel:.3e} mean_signed_rel={bias:.3e} pad_ok={pad_ok}")
