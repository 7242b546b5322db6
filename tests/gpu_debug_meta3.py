"""Ad-hoc GPU diagnostics: the golden SMALL meta scenario (val batch shape != train batch shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch
import mtl_b200
from gpu_util import dev, rel_err, spec_of, to_batch
from oracle import ref_asr, ref_meta, make_golden as mg

cfg, m = ref_asr.SMALL, mg.SMALL_META
p = ref_asr.init_params(cfg, m["seed"])
s = mtl_b200.Session(spec_of(cfg))
tasks, val = mg.small_tasks(0)
theta, theta0, grad, cg = (s.new_arena() for _ in range(4))
s.load(theta, p); s.copy(theta0, theta)
vb = to_batch(val)
print("val shapes", val[0].shape, val[2].shape, vb.n, "train", tasks[0][0].shape, tasks[0][2].shape)
# plain forward of val at theta0 vs oracle
o = s.forward(theta, vb)
lv, gv, gold, hyp, pred = ref_meta.loss_and_grads(p, cfg, val)
print("val@theta0 loss", float(o["ce"][0]), lv, "pred err", rel_err(o["pred"], pred))
g = s.new_arena(); s.backward(theta, g, 1.0)
e = {k: rel_err(s.views(g)[k], gv[k]) for k in gv if float(gv[k].abs().max()) > 1e-7}
k = max(e, key=e.get); print("val grads worst", k, e[k])
for i, tr in enumerate(tasks):
    o = s.forward(theta, to_batch(tr))
    lt, gt, _, _, predt = ref_meta.loss_and_grads(p, cfg, tr)
    print("train", i, "loss", float(o["ce"][0]), lt, "pred err", rel_err(o["pred"], predt))
    res = torch.zeros(16, device=dev())
    s.zero(cg)
    s.meta_task(theta, theta0, grad, cg, to_batch(tr), vb, m["lr"], 1.0 / 3, results=res)
    pt = {k: t.clone() for k, t in p.items()}
    ref_meta.sgd_step_(pt, gt, m["lr"])
    lv2, gv2, *_ = ref_meta.loss_and_grads(pt, cfg, val, 1.0 / 3)
    print("   meta_task tr/val", float(res[0]), float(res[8]), "oracle", lt, lv2)
