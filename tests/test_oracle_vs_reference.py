"""Pins the oracle restatement (oracle/ref_asr.py, oracle/ref_meta.py) against the LIVE
reference imported from /root/reference (build container only; skipped on the GPU box)."""
import copy
import re

import pytest
import torch

from oracle import live_reference as live
from oracle import ref_asr, ref_meta

pytestmark = pytest.mark.skipif(not live.available(), reason="/root/reference not present")

CFG = ref_asr.SMALL


def _assert_params_close(model, po, meta_lr, n_steps):
    """The key-projection bias has a mathematically ZERO gradient (softmax is invariant to a
    per-query constant q.b_k), so its Adam-normalised update is +-lr*sign(rounding noise):
    bound it by the largest possible drift instead of comparing values."""
    for n, t in model.named_parameters():
        # other tensors: 1% of one Adam update (|update| <= lr); rounding noise on elements
        # whose gradient is ~eps (dead ReLUs) moves g/(|g|+eps) by more than float epsilon
        d = (t.detach() - po[n]).abs()
        if n.endswith("key_linear_b.bias"):
            assert float(d.max()) <= 2.1 * meta_lr * n_steps, n
            continue
        # conv/GEMM backward summation order depends on the thread count, so a handful of
        # elements with |g| ~ sqrt(v) noise may move by a few % of one update: allow <=0.01% of
        # a tensor's elements beyond 1% of an update, none beyond 10% of an update.
        assert float((d > 1e-2 * meta_lr).float().mean()) <= 1e-4, n
        assert float(d.max()) <= 1e-1 * meta_lr, n


def _batches(seed0, n_tasks, ragged=False):
    tasks = []
    for i in range(n_tasks):
        if ragged and i == 0:
            tasks.append(ref_meta.synth_batch(CFG, 4, 41, 7, seed0 + i, lengths=[41, 30, 9, 5],
                                              tgt_lengths=[7, 5, 3, 1]))
        else:
            tasks.append(ref_meta.synth_batch(CFG, 4, 41, 7, seed0 + i))
    return tasks


def test_param_inventory_matches_reference():
    p = ref_asr.init_params(CFG, 1)
    model, _, _ = live.build_model(CFG, p)
    ref_names = [(n, tuple(t.shape)) for n, t in model.named_parameters()]
    assert ref_names == [(n, tuple(s)) for n, s in ref_asr.param_specs(CFG)]
    bufs = ref_asr.buffers(CFG)
    for n, b in model.named_buffers():
        assert torch.equal(b, bufs[n]), n


def test_cfg2_inventory_counts():
    assert len(ref_asr.param_specs(ref_asr.CFG2)) == 190
    assert ref_asr.num_params(ref_asr.CFG2) == 14022080


@pytest.mark.parametrize("ragged", [False, True])
def test_forward_loss_grads_match_reference(ragged):
    p = ref_asr.init_params(CFG, 2)
    model, vocab, _ = live.build_model(CFG, p)
    model.train()
    x, lens, y = _batches(10, 1, ragged)[0]
    pred_r, gold_r, hyp_r = model(x, lens, y)
    with live.reference_imports():
        from utils.metrics import calculate_metrics
        loss_r, _ = calculate_metrics(pred_r, gold_r, 0, smoothing=0.0, loss_type="ce")
    loss_r.backward()
    loss_o, g_o, gold_o, hyp_o, pred_o = ref_meta.loss_and_grads(p, CFG, (x, lens, y))
    assert torch.equal(gold_r, gold_o)
    assert torch.equal(pred_r.detach(), pred_o)          # same torch ops, same order -> bit-exact
    keep = gold_o != 0
    assert torch.equal(hyp_r[keep], hyp_o[keep])
    assert float(loss_r.detach()) == loss_o
    for n, t in model.named_parameters():
        assert torch.allclose(t.grad, g_o[n], rtol=0, atol=1e-7), n


def test_label_smoothing_matches_reference():
    p = ref_asr.init_params(CFG, 2)
    x, lens, y = _batches(11, 1, True)[0]
    pred, gold, _ = ref_asr.forward(p, CFG, x, lens, y)
    with live.reference_imports():
        from utils.metrics import calculate_loss
        # NB: through calculate_metrics the reference passes a (B,T) mask against the flattened
        # gold (utils/metrics.py:79,117) and raises for B>1; the formula itself is checked here
        # with the mask flattened.
        l_ref = calculate_loss(pred, gold, 0, non_pad_mask=gold.ne(0).view(-1), smoothing=0.1,
                               loss_type="ce")
    assert float(l_ref) == float(ref_asr.ce_loss(pred, gold, 0.1))


@pytest.mark.parametrize("clip", [False, True])
def test_meta_step_matches_reference_trainer(clip):
    """Three unchanged TransientTrainer iterations == three oracle meta_steps (theta and the
    printed loss), including the train-gradient leak and the shared val batch."""
    n_steps, n_tasks = 3, 3
    p = ref_asr.init_params(CFG, 3)
    model, vocab, args = live.build_model(CFG, p, lr=1e-2, meta_lr=1e-3, clip=clip, max_norm=0.5)
    steps_tasks = [_batches(100 * s, n_tasks, ragged=(s == 1)) for s in range(n_steps)]
    steps_val = [ref_meta.synth_batch(CFG, 4, 41, 7, 100 * s + 50) for s in range(n_steps)]
    out = live.run_transient(model, vocab, args, steps_tasks, steps_val, n_steps)
    ref_losses = [float(m) for m in re.findall(r"TRAIN LOSS:([0-9.]+)", out)]
    adam = ref_meta.AdamState()
    po = {k: v.clone() for k, v in p.items()}
    losses = []
    for s in range(n_steps):
        r = ref_meta.meta_step(po, adam, CFG, steps_tasks[s], steps_val[s], lr=1e-2, meta_lr=1e-3,
                               clip=clip, max_norm=0.5)
        losses.append(r["loss"])
    assert len(ref_losses) == n_steps
    for a, b in zip(ref_losses, losses):
        assert abs(a - b) < 1e-4
    _assert_params_close(model, po, 1e-3, n_steps)


def test_copy_grad_is_train_plus_val_over_n():
    """accumulated outer grad = sum_i [g_tr_i(theta0) + g_val(theta_i')/N]."""
    p = ref_asr.init_params(CFG, 4)
    model, vocab, args = live.build_model(CFG, p, lr=1e-2, meta_lr=1e-3)
    tasks = _batches(7, 2)
    val = ref_meta.synth_batch(CFG, 4, 41, 7, 99)
    live.run_transient(model, vocab, args, [tasks], [val], 1)
    po = {k: v.clone() for k, v in p.items()}
    r = ref_meta.meta_step(po, ref_meta.AdamState(), CFG, tasks, val, lr=1e-2, meta_lr=1e-3)
    for (n, _), cg in zip(model.named_parameters(), model.copy_grad):
        assert torch.allclose(cg, r["copy_grad"][n], rtol=0, atol=1e-6), n


def test_joint_step_matches_reference_trainer():
    n_steps, n_tasks = 2, 2
    p = ref_asr.init_params(CFG, 5)
    model, vocab, args = live.build_model(CFG, p, lr=1e-3)
    steps_tasks = [_batches(300 + 10 * s, n_tasks) for s in range(n_steps)]
    live.run_joint(model, vocab, args, steps_tasks, n_steps)
    adam = ref_meta.AdamState()
    po = {k: v.clone() for k, v in p.items()}
    for s in range(n_steps):
        ref_meta.joint_step(po, adam, CFG, steps_tasks[s], lr=1e-3)
    _assert_params_close(model, po, 1e-3, n_steps)


def test_greedy_search_and_encode_match_live_reference():
    """Transformer.encode (models/asr/transformer.py:78-98) and Decoder.greedy_search (modules/decoder.py:131-184: 300
    hard-coded steps, strings cut at the first EOS) of the live reference vs the restatement."""
    import dataclasses
    cfg = dataclasses.replace(CFG, tgt_max_len=320)              # the reference indexes the PE table up to 300 positions
    p = ref_asr.init_params(cfg, 4)
    model, vocab, args = live.build_model(cfg, p)
    model.eval()
    x, lens, y = ref_meta.synth_batch(cfg, 3, 41, 7, 77, lengths=[41, 25, 9], tgt_lengths=[7, 4, 2])
    args.cuda = False
    with torch.no_grad():
        enc_ref = model.encode(x, lens)
        strs_ref = model.decoder.greedy_search(enc_ref, args, start_token=vocab.SOS_ID)
        enc_o = ref_asr.encode(p, cfg, x, lens)
        ids = ref_asr.greedy_search(p, cfg, enc_o, start_token=vocab.SOS_ID, max_steps=300)
    assert torch.equal(enc_o, enc_ref)
    for b in range(3):
        st = ""
        for t in ids[b].tolist():
            if t == vocab.EOS_ID:
                break
            st += vocab.id2label[t]
        assert st == strs_ref[b], b
