"""Ad-hoc GPU diagnostics (not a pytest file): stage-by-stage comparison of one SMALL meta task vs the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch
import mtl_b200
from gpu_util import dev, rel_err, spec_of, to_batch
from oracle import ref_asr, ref_meta

cfg = ref_asr.SMALL
p = ref_asr.init_params(cfg, 3)
s = mtl_b200.Session(spec_of(cfg))
tr = ref_meta.synth_batch(cfg, 4, 41, 7, 0)
va = ref_meta.synth_batch(cfg, 4, 41, 7, 50)
lr = 1e-2

def worst(views, ref):
    e = {k: rel_err(views[k], ref[k]) for k in ref if float(ref[k].abs().max()) > 1e-7}
    k = max(e, key=e.get)
    return k, e[k]

theta, theta0, grad, cg = (s.new_arena() for _ in range(4))
s.load(theta, p); s.copy(theta0, theta)
btr, bva = to_batch(tr), to_batch(va)
# stage 1: g_tr
s.zero(grad)
o = s.forward(theta, btr); s.backward(theta, grad, 1.0)
l_o, g_o, *_ = ref_meta.loss_and_grads(p, cfg, tr)
print("train loss", float(o["ce"][0]), l_o, "worst g_tr", worst(s.views(grad), g_o))
# stage 2: sgd
s.sgd(theta, grad, lr)
po = {k: v.clone() for k, v in p.items()}
ref_meta.sgd_step_(po, g_o, lr)
print("theta' worst", worst(s.views(theta), po))
# stage 3: val at theta'
o = s.forward(theta, bva)
lv, gv, *_ = ref_meta.loss_and_grads(po, cfg, va, 1.0 / 3)
print("val loss", float(o["ce"][0]), lv)
g2 = s.new_arena()
s.backward(theta, g2, 1.0 / 3)
print("worst g_val", worst(s.views(g2), gv))
# stage 4: full meta_task
theta.copy_(theta0)
res = torch.zeros(16, device=dev())
s.zero(cg)
s.meta_task(theta, theta0, grad, cg, btr, bva, lr, 1.0 / 3, results=res)
print("meta_task losses", float(res[0]), float(res[8]), "oracle", l_o, lv)
ref_cg = {k: g_o[k] + gv[k] for k in g_o}
print("worst cg", worst(s.views(cg), ref_cg))
# adam
m, v, st = s.new_arena(), s.new_arena(), s.new_adam_state()
adam = ref_meta.AdamState()
pa = {k: v_.clone() for k, v_ in p.items()}
th = s.new_arena(); s.load(th, p)
for it in range(3):
    gr = {k: torch.randn_like(v_) * (10.0 ** -it) for k, v_ in p.items()}
    ga = s.new_arena(); s.load(ga, gr)
    s.adam(th, ga, m, v, st, 1e-3)
    ref_meta.adam_step_(pa, gr, adam, 1e-3)
    d = max(float((s.views(th)[k].cpu() - pa[k]).abs().max()) for k in pa)
    print("adam step", it, "max abs diff", d, "state", st.tolist())
