"""Ad-hoc GPU diagnostics: the 3-step SMALL meta scenario, step by step vs the oracle."""
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
steps = []
for st in range(3):
    tasks = [ref_meta.synth_batch(cfg, 4, 41, 7, 100 * st + i,
                                  lengths=[41, 30, 9, 5] if (st == 1 and i == 0) else None,
                                  tgt_lengths=[7, 5, 3, 1] if (st == 1 and i == 0) else None) for i in range(3)]
    steps.append((tasks, ref_meta.synth_batch(cfg, 4, 41, 7, 100 * st + 50)))

def errs(views, ref, top=4):
    e = {k: rel_err(views[k], ref[k]) for k in ref if float(ref[k].abs().max()) > 1e-7}
    ks = sorted(e, key=e.get, reverse=True)[:top]
    return [(k, float("%.2e" % e[k])) for k in ks]

po = {k: v.clone() for k, v in p.items()}
adam = ref_meta.AdamState()
theta, theta0, grad, cg = (s.new_arena() for _ in range(4))
m, v, st_ = s.new_arena(), s.new_arena(), s.new_adam_state()
s.load(theta, p)
for si, (tasks, val) in enumerate(steps):
    n = len(tasks)
    res = torch.zeros(n, 16, device=dev())
    s.copy(theta0, theta); s.zero(cg)
    vb = to_batch(val)
    cgs = []
    for i, tr in enumerate(tasks):
        s.meta_task(theta, theta0, grad, cg, to_batch(tr), vb, 1e-2, 1.0 / n, results=res[i])
        cgs.append({k: t.clone() for k, t in s.views(cg).items()})
    s.meta_finish(theta, grad, cg, m, v, st_, 1e-3)
    # oracle, instrumented per task
    th0 = {k: t.clone() for k, t in po.items()}
    r = ref_meta.meta_step(po, adam, cfg, tasks, val, lr=1e-2, meta_lr=1e-3)
    print("step", si, "tr", [round(float(x), 6) for x in res[:, 0]], "oracle", [round(float(x), 6) for x in r["tr_losses"]])
    print("       val", [round(float(x), 6) for x in res[:, 8]], "oracle", [round(float(x), 6) for x in r["val_losses"]])
    # per task cumulative cg from oracle
    acc = {k: torch.zeros_like(t) for k, t in th0.items()}
    for i, tr in enumerate(tasks):
        pt = {k: t.clone() for k, t in th0.items()}
        _, g, *_ = ref_meta.loss_and_grads(pt, cfg, tr)
        ref_meta.sgd_step_(pt, g, 1e-2)
        _, gv, *_ = ref_meta.loss_and_grads(pt, cfg, val, 1.0 / n)
        for k in acc:
            acc[k] += g[k] + gv[k]
        print("   task", i, "cum cg errs", errs(cgs[i], acc))
    print("   final cg errs vs meta_step", errs(s.views(cg), r["copy_grad"]))
    d = {k: float((s.views(theta)[k].cpu() - po[k]).abs().max()) for k in po}
    ks = sorted(d, key=d.get, reverse=True)[:4]
    print("   theta max abs diff", [(k, float("%.2e" % d[k])) for k in ks])
