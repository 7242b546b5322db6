"""Ad-hoc: dumps the workspace + gradient arena after forward / after backward to /tmp/<tag>_{fwd,bwd}.pt"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch
import mtl_b200
from gpu_util import spec_of, to_batch
from oracle import ref_asr, ref_meta
tag = sys.argv[1]
cfg = ref_asr.SMALL
p = ref_asr.init_params(cfg, 5)
s = mtl_b200.Session(spec_of(cfg), gemm_mode=2)
b1 = ref_meta.synth_batch(cfg, 4, 41, 7, 1)
theta, grad = s.new_arena(), s.new_arena()
s.load(theta, p)
bt = to_batch(b1)
s._ws = None
need = int(s.lib.mtl_workspace_bytes(s._h, 4, 41, 8))
s._ws = torch.zeros(need + (64 << 20), dtype=torch.uint8, device=s.device)
s.forward(theta, bt)
torch.cuda.synchronize()
torch.save(s._ws[:need].view(torch.float32).cpu(), f"/tmp/{tag}_fwd.pt")
s.backward(theta, grad, 1.0)
torch.cuda.synchronize()
torch.save(s._ws[:need].view(torch.float32).cpu(), f"/tmp/{tag}_bwd.pt")
torch.save(grad.cpu(), f"/tmp/{tag}_grad.pt")
print(tag, "need", need)
