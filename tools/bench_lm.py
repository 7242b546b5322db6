"""Measures the LM meta-iteration (BASELINE configs[4]: lm/main_meta_transfer.py defaults -- LSTM 2 x 200, emsize 200,
bptt 35, batch 20, 3 tasks, dropout 0.2, clip 0.25; synthetic token ids, vocabulary 10 000 because the corpora are not
shipped) on one GPU through LmSession.meta_step, next to the CPU oracle (oracle/ref_lm.meta_step, dropout off) on the
host cores.  One JSON line.  The iteration does not shard over tasks (every train forward starts from the previous
task's hidden state): more GPUs = independent replicas."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import mtl_b200  # noqa: E402
from oracle import ref_lm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--vocab", type=int, default=10000)
    ap.add_argument("--gemm-mode", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--kernels", action="store_true", help="after the measurement: per-kernel time table of 2 traced steps (stderr)")
    a = ap.parse_args()
    cfg = ref_lm.LmConfig(vocab=a.vocab)
    T, B, n_tasks = 35, 20, 3
    dev = torch.device("cuda:0")
    s = mtl_b200.LmSession(mtl_b200.LmSpec(cfg.vocab, cfg.ninp, cfg.nhid, cfg.nlayers), dev, gemm_mode=a.gemm_mode)
    p = ref_lm.init_params(cfg, 1)
    theta, work, grad, meta = (s.new_arena() for _ in range(4))
    s.load(theta, p)
    hidden = s.new_hidden(B)
    data = [ref_lm.synth_blocks(cfg, n_tasks, T, B, 100 + i) for i in range(4)]
    dev_data = [([(x.to(dev), y.to(dev)) for x, y in tr], (va[0].to(dev), va[1].to(dev))) for tr, va in data]
    res = torch.zeros(n_tasks, 16, device=dev)
    w = [0.1, 0.1, 0.8]
    lib = mtl_b200.get_lib()

    def step(i):
        tr, va = dev_data[i % len(dev_data)]
        s.meta_step(theta, work, grad, meta, hidden, tr, va, w, lr=20.0 / 1000, meta_lr_factor=3.0, clip=0.25, dropout=0.2,
                    seed=i, results=res, graph=not a.no_graph)

    l0 = lib.mtl_launch_count()
    step(0)                                        # eager (also the first sighting of the graph path): counts the kernels of one step
    launches = lib.mtl_launch_count() - l0
    for i in range(1, max(a.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(a.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    tokens = n_tasks * 2 * T * B
    out = {"metric": "LM meta-iteration token-passes/sec (LSTM 2x200, bptt 35, batch 20, 3 tasks)", "value": tokens / ms * 1e3,
           "unit": "token-passes/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
           "higher_is_better": True, "dtype": {0: "f32", 1: "tf32", 2: "3xtf32"}[a.gemm_mode], "data": "synthetic",
           "config": {"workload": "cfg5: lm/main_meta_transfer.py defaults, synthetic token ids", "vocab": cfg.vocab,
                      "tasks": n_tasks, "bptt": T, "batch": B, "dropout": 0.2},
           "cuda_graph": not a.no_graph, "gpu_launches": launches, "loss_last": [float(v) for v in res[:, 8].cpu()]}
    if not a.no_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        pp, hid = p, None
        ts = []
        for i in range(3):
            tr, va = data[i % len(data)]
            t0 = time.perf_counter()
            pp, hid, *_ = ref_lm.meta_step(pp, cfg, tr, va, w, hid, lr=20.0 / 1000, meta_lr_factor=3.0, clip=0.25)
            ts.append(time.perf_counter() - t0)
        cpu_ms = min(ts[1:]) * 1e3
        out["cpu_baseline"] = {"value": tokens / cpu_ms * 1e3, "unit": "token-passes/s", "cores": os.cpu_count(), "kind": "port",
                               "ms_per_step": cpu_ms, "sample": "2 meta-iterations after 1 warm-up, oracle/ref_lm.meta_step (autograd, dropout off)"}
    print(json.dumps(out))
    if a.kernels:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(2):
                step(1000 + i)
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:25]
        for e in rows:
            print(f"{e.device_time_total / 2:10.1f} us/step  n/step {e.count / 2:7.1f}  avg {e.device_time_total / max(e.count, 1):8.2f} us  {e.key[:90]}",
                  file=sys.stderr)


if __name__ == "__main__":
    main()
