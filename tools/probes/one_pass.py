"""One cfg-2 forward + backward + the arena operations of a meta-step, eagerly on one stream (no CUDA graph, no task
lanes): the launch list `ncu` walks for the per-kernel captures under profiles/ (every kernel of the hot path once)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

import mtl_b200


def main():
    spec = mtl_b200.ModelSpec()
    s = mtl_b200.Session(spec, "cuda:0", gemm_mode=2)
    rng = np.random.default_rng(0)
    theta, theta0, grad, cg, m, v = (s.new_arena() for _ in range(6))
    theta.copy_(torch.from_numpy(rng.uniform(-0.05, 0.05, s.n_floats).astype(np.float32)))
    for name, shape, off, n in s.table:
        if "layer_norm" in name and name.endswith("weight"):
            s.views(theta)[name].fill_(1.0)
    x = torch.from_numpy(rng.standard_normal((8, 1, 161, 101), dtype=np.float32))
    y = torch.from_numpy(rng.integers(4, spec.vocab, size=(8, 32), dtype=np.int64))
    lens = torch.tensor([101, 101, 80, 80, 40, 40, 20, 20], dtype=torch.int32)
    b = mtl_b200.Batch.from_host(x, lens, y, s.device)
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    for _ in range(reps):
        s.copy(theta0, theta)
        s.zero(grad)
        out = s.forward(theta, b, dropout=0.1, seed=1)
        s.backward(theta, grad, 1.0)
        s.clip(grad, 400.0)
        s.sgd(theta, grad, 1e-4)
        s.axpy(cg, grad, 1.0)
        s.meta_finish(theta, grad, cg, m, v, s.new_adam_state(), 1e-4, clip=True)
    enc = s.encode(theta, b.x, b.lens)
    s.greedy(theta, enc, 1, 2)
    torch.cuda.synchronize()
    print("loss", float(out["ce"][0]))


if __name__ == "__main__":
    main()
