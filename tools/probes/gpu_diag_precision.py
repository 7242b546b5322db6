"""Ad-hoc GPU diagnostics (not pytest): per-engine gradient accuracy on the cfg-2 golden forward/backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import numpy as np
import torch
import mtl_b200
from gpu_util import rel_err, spec_of, to_batch
from oracle import ref_asr, make_golden as mg

g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cfg2_fwd_bwd.npz"))
cfg = ref_asr.CFG2
p = ref_asr.init_params(cfg, 31)
batch = to_batch(mg.cfg2_batch(3100, ragged=True))
for mode in (0, 2, 1):
    s = mtl_b200.Session(spec_of(cfg), gemm_mode=mode)
    theta, grad = s.new_arena(), s.new_arena()
    s.load(theta, p)
    out = s.forward(theta, batch)
    pred = out["pred"].clone()
    s.backward(theta, grad, 1.0)
    torch.cuda.synchronize()
    grads = s.views(grad)
    flat = pred.reshape(-1).cpu()
    ps = flat[torch.from_numpy(mg.pred_sample_idx(flat.numel()))]
    errs = []
    for name, _ in ref_asr.param_specs(cfg):
        gn = float(g["gnorm/" + name])
        if gn <= 1e-6:
            continue
        v = grads[name].cpu()
        en = abs(float(v.double().norm()) - gn) / gn
        sm = v.reshape(-1)[torch.from_numpy(mg.sample_idx(v.numel()))]
        es = float((sm - torch.from_numpy(g["gsamp/" + name])).abs().max()) / gn
        errs.append((max(en, es), en, es, name))
    errs.sort(reverse=True)
    print(f"mode {mode}: loss {float(out['ce'][0]):.7f} (ref {float(g['loss']):.7f}) pred sample rel err "
          f"{rel_err(ps, torch.from_numpy(g['pred_samples'])):.2e}")
    for e in errs[:4]:
        print(f"    worst grad: {e[3]:55s} norm err {e[1]:.2e} sample err {e[2]:.2e}")
