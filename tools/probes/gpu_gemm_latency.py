"""Ad-hoc GPU timing (not pytest): latency of ONE small GEMM as a node of a dependent chain inside a CUDA graph
(what the meta-step graph pays per node), plus the per-CTA %globaltimer span of the last launch (MTL_GEMM_DBG=99)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch
from gpu_util import P, dev, lib, ok

SHAPES = [(0, 1, 264, 100, 512, 0.0, 1), (0, 1, 264, 512, 100, 0.0, 1), (0, 1, 264, 512, 512, 0.0, 1), (0, 1, 264, 1536, 512, 0.0, 1),
          (0, 0, 264, 512, 512, 0.0, 1), (0, 0, 264, 512, 1536, 1.0, 25), (1, 0, 512, 512, 264, 1.0, 9), (1, 0, 512, 100, 264, 1.0, 9),
          (0, 0, 512, 512, 100, 0.0, 1)]
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N_NODES = 100
print("mode", mode, "cluster", os.environ.get("MTL_CLUSTER_SPLITK", "1"), "dbg", os.environ.get("MTL_GEMM_DBG", "0"))
for (tA, tB, M, N, K, beta, split) in SHAPES:
    A = torch.randn((K, M) if tA else (M, K), device=dev())
    B = torch.randn((N, K) if tB else (K, N), device=dev())
    C = torch.zeros(M, N, device=dev())
    st = torch.cuda.Stream()
    run = lambda r, s: ok(lib().mtl_gemm_repeat(r, mode, tA, tB, M, N, K, P(A), A.shape[1], P(B), B.shape[1], beta, P(C), N, split,
                                                ctypes.c_void_p(s.cuda_stream)))
    run(3, st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        run(N_NODES, st)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (10 * N_NODES)
    line = f"tA={tA} tB={tB} M={M:5d} N={N:5d} K={K:5d} beta={beta} split={split:3d}: {us:6.2f} us/node"
    if os.environ.get("MTL_GEMM_DBG") == "99":
        buf = (ctypes.c_ulonglong * 512)()
        ok(lib().mtl_debug_gemm_span(buf))
        v = [(buf[2 * i], buf[2 * i + 1]) for i in range(256) if buf[2 * i + 1] > buf[2 * i] > 0]
        # only the CTAs of the last launch: entries within 50 us of the latest exit
        last = max(e for _, e in v)
        v = [(s, e) for s, e in v if last - s < 50000]
        s0 = min(s for s, _ in v)
        line += f" | ctas {len(v):3d} start skew {max(s for s, _ in v) - s0:5d} ns, span {last - s0:6d} ns, mean cta {sum(e - s for s, e in v) / len(v):7.0f} ns"
    if os.environ.get("MTL_GEMM_DBG") == "1":
        st_ = (ctypes.c_longlong * 160)()
        ok(lib().mtl_debug_gemm_stamps(st_))
        t0 = st_[0]
        names = {1: "pdl_wait done", 2: "first TMA", 3: "MMA sees operands", 4: "MMAs issued", 5: "tmem_full seen", 9: "staged",
                 10: "cluster sync", 6: "stored", 7: "end"}
        line += " | cycles from entry: " + ", ".join("%s %d" % (names[i], st_[i] - t0) for i in (1, 2, 3, 4, 5, 9, 10, 6, 7) if st_[i] > t0)
    print(line)
