"""Oracle restatement vs the committed golden vectors produced by the live reference
(oracle/make_golden.py).  Runs anywhere (no /root/reference, no GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import make_golden as mg
from oracle import ref_asr, ref_meta

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name))


def test_small_fwd_bwd_golden():
    g = _load("small_fwd_bwd.npz")
    cfg = ref_asr.SMALL
    p = ref_asr.init_params(cfg, 21)
    batch = ref_meta.synth_batch(cfg, 4, 41, 7, 2100, lengths=[41, 30, 9, 5], tgt_lengths=[7, 5, 3, 1])
    loss, grads, gold, hyp, pred = ref_meta.loss_and_grads(p, cfg, batch)
    assert np.array_equal(gold.numpy(), g["gold"])
    keep = g["gold"] != 0
    assert np.array_equal(hyp.numpy()[keep], g["hyp"][keep])
    np.testing.assert_allclose(pred.numpy(), g["pred"], rtol=0, atol=1e-5)
    assert abs(loss - float(g["loss"])) < 1e-5
    assert ref_asr.num_correct(pred, gold) == int(g["num_correct"])
    for k, v in grads.items():
        np.testing.assert_allclose(v.numpy(), g["grad/" + k], rtol=0, atol=1e-6, err_msg=k)


def test_small_meta_golden():
    g = _load("small_meta.npz")
    cfg, m = ref_asr.SMALL, mg.SMALL_META
    p = ref_asr.init_params(cfg, m["seed"])
    adam = ref_meta.AdamState()
    for s in range(m["n_steps"]):
        tasks, val = mg.small_tasks(s)
        r = ref_meta.meta_step(p, adam, cfg, tasks, val, lr=m["lr"], meta_lr=m["meta_lr"])
        assert abs(r["loss"] - g["losses"][s]) < 1e-4
    for k in p:
        np.testing.assert_allclose(r["copy_grad"][k].numpy(), g["copy_grad/" + k], rtol=0, atol=2e-6, err_msg=k)
        atol = 2.1 * m["meta_lr"] * m["n_steps"] if k.endswith("key_linear_b.bias") else 1e-2 * m["meta_lr"]
        np.testing.assert_allclose(p[k].numpy(), g["theta/" + k], rtol=0, atol=atol, err_msg=k)


@pytest.mark.timeout(600)
def test_cfg2_fwd_bwd_golden():
    g = _load("cfg2_fwd_bwd.npz")
    cfg = ref_asr.CFG2
    p = ref_asr.init_params(cfg, 31)
    loss, grads, gold, hyp, pred = ref_meta.loss_and_grads(p, cfg, mg.cfg2_batch(3100, ragged=True))
    assert np.array_equal(gold.numpy(), g["gold"])
    keep = g["gold"] != 0
    assert np.array_equal(hyp.numpy()[keep], g["hyp"][keep])
    assert abs(loss - float(g["loss"])) < 1e-5
    flat = pred.reshape(-1)
    np.testing.assert_allclose(flat[torch.from_numpy(mg.pred_sample_idx(flat.numel()))].numpy(),
                               g["pred_samples"], rtol=0, atol=1e-5)
    for k, v in grads.items():
        gn = float(g["gnorm/" + k])
        assert abs(float(v.double().norm()) - gn) <= 1e-5 * gn + 1e-9, k
        s = v.reshape(-1)[torch.from_numpy(mg.sample_idx(v.numel()))].numpy()
        np.testing.assert_allclose(s, g["gsamp/" + k], rtol=0, atol=1e-5 * max(gn, 1e-6), err_msg=k)


def test_adam_and_sgd_formulas_match_torch_optim():
    torch.manual_seed(0)
    w = {"a": torch.randn(7, 5), "b": torch.randn(11)}
    mods = {k: torch.nn.Parameter(v.clone()) for k, v in w.items()}
    opt = torch.optim.Adam(list(mods.values()), lr=3e-3)
    st = ref_meta.AdamState()
    for it in range(4):
        grads = {k: torch.randn_like(v) * (10.0 ** -it) for k, v in w.items()}
        for k in mods:
            mods[k].grad = grads[k].clone()
        opt.step()
        ref_meta.adam_step_(w, grads, st, 3e-3)
        for k in w:
            assert torch.allclose(w[k], mods[k].detach(), rtol=0, atol=1e-7)
    sgd = torch.optim.SGD(list(mods.values()), lr=0.1)
    sgd.step()
    ref_meta.sgd_step_(w, grads, 0.1)
    for k in w:
        assert torch.allclose(w[k], mods[k].detach(), rtol=0, atol=1e-7)
    total = torch.nn.utils.clip_grad_norm_(list(mods.values()), 0.01)
    t2 = ref_meta.clip_grad_norm_(grads, 0.01)
    assert abs(float(total) - t2) < 1e-6 * t2
    for k in w:
        assert torch.allclose(grads[k], mods[k].grad, rtol=1e-5, atol=1e-9)
