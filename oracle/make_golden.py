"""Generate tests/golden/*.npz from the LIVE reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden
Every number written here comes out of the UNMODIFIED reference code
(models/asr/transformer.py, trainer/asr/transient_trainer.py, utils/metrics.py)
driven through oracle/live_reference.py; inputs/weights are regenerated from numpy
PCG64 seeds by oracle.ref_asr.init_params / oracle.ref_meta.synth_batch.

  small_fwd_bwd.npz   SMALL cfg, ragged batch: pred, gold, hyp, loss, every grad tensor
  small_meta.npz      SMALL cfg, 2 TransientTrainer iterations (3 tasks): printed losses,
                      copy_grad after the last step, all parameters after the last Adam step
  cfg2_fwd_bwd.npz    BASELINE cfg 2 (enc2/dec4/d512, B=8, T=101, L=32): loss, gold, hyp,
                      pred samples, per-tensor grad L2 norms + 32 strided samples per tensor
  ref_checkpoint_small.th  a checkpoint written by the reference's save_meta_model (SMALL cfg, after 2 Adam steps)
  cfg2_meta.npz       cfg 2, ONE TransientTrainer iteration (3 tasks, k=8): printed loss,
                      per-tensor L2 norm + samples of copy_grad and of (theta_after - theta_before)
"""
from __future__ import annotations

import os
import re
import sys

import numpy as np
import torch

from . import live_reference as live
from . import ref_asr, ref_meta

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
N_SAMPLES = 32


def sample_idx(numel: int) -> np.ndarray:
    """Deterministic strided sample positions shared by generator and tests."""
    n = min(N_SAMPLES, numel)
    return (np.arange(n, dtype=np.int64) * max(1, numel // n)) % numel


def pred_sample_idx(numel: int) -> np.ndarray:
    return np.arange(0, numel, max(1, numel // 4096), dtype=np.int64)[:4096]


def _ref_fwd_bwd(cfg, params, batch):
    model, vocab, _ = live.build_model(cfg, params)
    model.train()
    x, lens, y = batch
    pred, gold, hyp = model(x, lens, y)
    with live.reference_imports():
        from utils.metrics import calculate_metrics
        loss, ncorrect = calculate_metrics(pred, gold, 0, smoothing=0.0, loss_type="ce")
    loss.backward()
    grads = {n: t.grad.detach().clone() for n, t in model.named_parameters()}
    return pred.detach(), gold, hyp, float(loss.detach()), ncorrect, grads


def small_fwd_bwd():
    cfg = ref_asr.SMALL
    params = ref_asr.init_params(cfg, 21)
    batch = ref_meta.synth_batch(cfg, 4, 41, 7, 2100, lengths=[41, 30, 9, 5], tgt_lengths=[7, 5, 3, 1])
    pred, gold, hyp, loss, ncorrect, grads = _ref_fwd_bwd(cfg, params, batch)
    d = dict(pred=pred.numpy(), gold=gold.numpy(), hyp=hyp.numpy(), loss=np.float64(loss),
             num_correct=np.int64(ncorrect))
    d.update({"grad/" + k: v.numpy() for k, v in grads.items()})
    np.savez_compressed(os.path.join(OUT, "small_fwd_bwd.npz"), **d)


def small_tasks(step):
    cfg = ref_asr.SMALL
    tasks = [ref_meta.synth_batch(cfg, 4, 41, 7, 2200 + 10 * step + i,
                                  lengths=[41, 33, 20, 8] if i == 1 else None,
                                  tgt_lengths=[7, 6, 2, 4] if i == 1 else None) for i in range(3)]
    val = ref_meta.synth_batch(cfg, 4, 37, 6, 2290 + step)
    return tasks, val


SMALL_META = dict(lr=1e-2, meta_lr=1e-3, n_steps=2, seed=22)


def small_meta():
    cfg = ref_asr.SMALL
    m = SMALL_META
    params = ref_asr.init_params(cfg, m["seed"])
    model, vocab, args = live.build_model(cfg, params, lr=m["lr"], meta_lr=m["meta_lr"])
    st = [small_tasks(s) for s in range(m["n_steps"])]
    out = live.run_transient(model, vocab, args, [t for t, _ in st], [v for _, v in st], m["n_steps"])
    losses = [float(x) for x in re.findall(r"TRAIN LOSS:([0-9.]+)", out)]
    d = dict(losses=np.array(losses))
    for (n, t), cg in zip(model.named_parameters(), model.copy_grad):
        d["theta/" + n] = t.detach().numpy()
        d["copy_grad/" + n] = cg.numpy()
    np.savez_compressed(os.path.join(OUT, "small_meta.npz"), **d)


def cfg2_batch(seed, ragged=False):
    cfg = ref_asr.CFG2
    if ragged:
        return ref_meta.synth_batch(cfg, 8, 101, 32, seed, lengths=[101, 101, 80, 80, 40, 40, 20, 20],
                                    tgt_lengths=[32, 30, 25, 32, 12, 7, 3, 1])
    return ref_meta.synth_batch(cfg, 8, 101, 32, seed)


def cfg2_fwd_bwd():
    cfg = ref_asr.CFG2
    params = ref_asr.init_params(cfg, 31)
    pred, gold, hyp, loss, ncorrect, grads = _ref_fwd_bwd(cfg, params, cfg2_batch(3100, ragged=True))
    flat = pred.reshape(-1)
    d = dict(gold=gold.numpy(), hyp=hyp.numpy(), loss=np.float64(loss), num_correct=np.int64(ncorrect),
             pred_samples=flat[torch.from_numpy(pred_sample_idx(flat.numel()))].numpy(),
             pred_norm=np.float64(pred.double().norm()))
    for k, v in grads.items():
        d["gnorm/" + k] = np.float64(v.double().norm())
        d["gsamp/" + k] = v.reshape(-1)[torch.from_numpy(sample_idx(v.numel()))].numpy()
    np.savez_compressed(os.path.join(OUT, "cfg2_fwd_bwd.npz"), **d)


CFG2_META = dict(lr=1e-4, meta_lr=1e-4, seed=32)


def cfg2_tasks():
    return [cfg2_batch(3200 + i) for i in range(3)], cfg2_batch(3290)


def cfg2_meta():
    cfg = ref_asr.CFG2
    m = CFG2_META
    params = ref_asr.init_params(cfg, m["seed"])
    model, vocab, args = live.build_model(cfg, params, lr=m["lr"], meta_lr=m["meta_lr"], k_train=8, k_valid=8)
    tasks, val = cfg2_tasks()
    out = live.run_transient(model, vocab, args, [tasks], [val], 1)
    losses = [float(x) for x in re.findall(r"TRAIN LOSS:([0-9.]+)", out)]
    d = dict(losses=np.array(losses))
    for (n, t), cg in zip(model.named_parameters(), model.copy_grad):
        delta = t.detach() - params[n]
        idx = torch.from_numpy(sample_idx(t.numel()))
        d["cgnorm/" + n] = np.float64(cg.double().norm())
        d["cgsamp/" + n] = cg.reshape(-1)[idx].numpy()
        d["dsamp/" + n] = delta.reshape(-1)[idx].numpy()
    np.savez_compressed(os.path.join(OUT, "cfg2_meta.npz"), **d)


REF_CKPT = dict(seed=23, lr=1e-2, meta_lr=1e-3, epoch=7)


def ref_checkpoint():
    """A checkpoint written by the REFERENCE's own save_meta_model (utils/functions.py:101-126) after two real Adam steps:
    pickled Vocab + argparse Namespace + model_state_dict + the live torch.optim.SGD / Adam objects.  The -m gpu test
    loads it through OUR load_meta_model and resumes training (checkpoint wire format)."""
    import shutil
    import tempfile
    cfg = ref_asr.SMALL
    m = REF_CKPT
    params = ref_asr.init_params(cfg, m["seed"])
    tmp = tempfile.mkdtemp()
    model, vocab, args = live.build_model(cfg, params, lr=m["lr"], meta_lr=m["meta_lr"], save_folder=tmp, name="ref",
                                          is_factorized=False, r=cfg.rank, model="TRFS")
    inner = torch.optim.SGD(model.parameters(), lr=m["lr"])
    outer = torch.optim.Adam(model.parameters(), lr=m["meta_lr"])
    model.train()
    for step in range(2):
        x, lens, y = ref_meta.synth_batch(cfg, 4, 41, 7, 2300 + step)
        outer.zero_grad()
        pred, gold, _ = model(x, lens, y)
        torch.nn.functional.cross_entropy(pred.view(-1, pred.size(2)), gold.view(-1), ignore_index=0).backward()
        outer.step()
    with live.reference_imports():
        from utils.data import Vocab
        from utils.functions import save_meta_model
        import contextlib, io
        v2 = Vocab()                         # same class object as the one pickle will look up in this import context
        v2.__dict__.update(vocab.__dict__)
        with contextlib.redirect_stdout(io.StringIO()):
            save_meta_model(model, v2, m["epoch"], inner, outer, {"avg_valid_loss": 1.25}, args, best_model=False)
    shutil.copyfile(os.path.join(tmp, "ref", "epoch_%d.th" % m["epoch"]), os.path.join(OUT, "ref_checkpoint_small.th"))
    shutil.rmtree(tmp)


LM_SMALL_GOLD = dict(seed=41, data_seed=4100, T=9, B=5)


def lm_small():
    """Reference RNNModel forward + CE + backward on a small LSTM LM, two chained blocks (the second starts from the first
    one's hidden state, as forward_one_batch threads it)."""
    from oracle import ref_lm
    cfg, m = ref_lm.LM_SMALL, LM_SMALL_GOLD
    p = ref_lm.init_params(cfg, m["seed"])
    (b0, b1), _ = ref_lm.synth_blocks(cfg, 2, m["T"], m["B"], m["data_seed"])
    d = {}
    hidden = None
    for i, (tok, trg) in enumerate((b0, b1)):
        loss, grads, logits, hidden = live.lm_fwd_bwd(cfg, p, tok, trg, hidden)
        d[f"loss{i}"] = np.float32(loss)
        d[f"logits{i}"] = logits.numpy()
        d[f"h{i}"], d[f"c{i}"] = hidden[0].numpy(), hidden[1].numpy()
        for k, v in grads.items():
            d[f"grad{i}/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "lm_small.npz"), **d)


def main():
    if not live.available():
        sys.exit("needs the reference at /root/reference")
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    only = sys.argv[1:]
    for fn in (small_fwd_bwd, small_meta, cfg2_fwd_bwd, cfg2_meta, ref_checkpoint, lm_small):
        if only and fn.__name__ not in only:
            continue
        fn()
        print("wrote", fn.__name__)


if __name__ == "__main__":
    main()
