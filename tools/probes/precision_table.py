"""GPU probe: per-tensor gradient error of the cfg-2 forward+backward against the CPU oracle for a list of
per-operation-class precision policies (Session.set_op_mode).  Writes gpurun_out/precision_table.json and prints,
per policy, the worst tensors.  rel_err = max|a-b| / max|b| per tensor (north_star's tolerance: 1e-3).

  python tools/probes/precision_table.py [policy ...]      policy = comma list class=mode, e.g. conv_fwd=1,conv_dgrad=1
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch

import mtl_b200
from gpu_util import rel_err, spec_of, to_batch
from oracle import make_golden as mg
from oracle import ref_asr, ref_meta

CLASSES = mtl_b200.Session.OP_CLASSES


def main():
    cfg = ref_asr.CFG2
    p = ref_asr.init_params(cfg, 31)
    batch = mg.cfg2_batch(3100, ragged=True)
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    loss_o, g_o, gold_o, hyp_o, pred_o = ref_meta.loss_and_grads(p, cfg, batch)
    print("oracle fwd+bwd %.2f s on %d threads" % (time.time() - t0, torch.get_num_threads()), flush=True)
    policies = sys.argv[1:] or (
        [""] + ["%s=1" % c for c in CLASSES] +
        ["conv_fwd=1,conv_dgrad=1", "conv_fwd=1,conv_dgrad=1,conv_wgrad=1", "lin_fwd=1,lin_dgrad=1,lin_wgrad=1,stem=1,vocab=1",
         "lin_fwd=1,lin_dgrad=1", "lin_wgrad=1,conv_wgrad=1", "conv_dgrad=1,conv_wgrad=1", "conv_dgrad=1,conv_wgrad=1,vocab=1",
         "conv_dgrad=1,conv_wgrad=1,vocab=1,lin_wgrad=1",
         ",".join("%s=1" % c for c in CLASSES)])
    table = {}
    for pol in policies:
        s = mtl_b200.Session(spec_of(cfg), gemm_mode=2)
        for kv in [x for x in pol.split(",") if x]:
            k, v = kv.split("=")
            s.set_op_mode(k, int(v))
        theta, grad = s.new_arena(), s.new_arena()
        s.load(theta, p)
        out = s.forward(theta, to_batch(batch))
        pred = out["pred"].clone()
        s.backward(theta, grad, 1.0)
        torch.cuda.synchronize()
        gv = s.views(grad)
        errs = {k: rel_err(gv[k], g_o[k]) for k in g_o if float(g_o[k].abs().max()) > 1e-7}
        l2 = {k: float((gv[k].cpu().double() - g_o[k].double()).norm() / g_o[k].double().norm()) for k in errs}
        keep = gold_o != 0
        hyp_ok = bool(torch.equal(out["hyp"].cpu().long()[keep], hyp_o[keep]))
        top = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
        row = {"pred": rel_err(pred, pred_o), "loss": abs(float(out["ce"][0]) - loss_o) / abs(loss_o), "hyp_exact": hyp_ok,
               "worst": top[0][1], "n_over_1e-3": sum(1 for v in errs.values() if v > 1e-3),
               "n_over_5e-4": sum(1 for v in errs.values() if v > 5e-4), "top": top, "all": errs,
               "l2_worst": max(l2.values()), "l2_n_over_1e-3": sum(1 for v in l2.values() if v > 1e-3),
               "l2_top": sorted(l2.items(), key=lambda kv: -kv[1])[:4], "l2_all": l2}
        table[pol or "all=2"] = row
        print("%-60s pred %.1e loss %.1e hyp %s worst %.2e  >1e-3: %d  >5e-4: %d" % (
            pol or "all=2", row["pred"], row["loss"], hyp_ok, row["worst"], row["n_over_1e-3"], row["n_over_5e-4"]))
        for k, v in top[:4]:
            print("      %-55s %.2e" % (k, v))
        print("      L2-relative: worst %.2e (%s), > 1e-3: %d" % (row["l2_worst"], row["l2_top"][0][0], row["l2_n_over_1e-3"]))
        sys.stdout.flush()
        del s
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "precision_table.json"), "w") as f:
        json.dump(table, f, indent=1)


if __name__ == "__main__":
    main()
