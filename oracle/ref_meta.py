"""Functional CPU restatement of the reference meta-transfer / joint training step
(TEST INFRASTRUCTURE; see oracle/__init__.py).

Reference citations (relative to /root/reference):
  trainer/asr/transient_trainer.py:150-255   meta-step body (snapshot, inner SGD, shared
                                             val batch, copy-grad accumulation, reset, Adam)
  trainer/asr/transient_trainer.py:198-199,226-229  NO zero_grad between the inner step and the
                                             val backward -> accumulated grad = g_tr + g_val/N
  trainer/asr/joint_trainer.py:178-271       joint step: sum_i grad(L_i / N), one Adam step
  models/asr/transformer.py:204-240          copy_grad buffer API
  torch.optim.SGD / Adam / clip_grad_norm_   third-party (torch 2.11): formulas restated below and
                                             pinned against torch.optim in tests/test_oracle_*.py
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import ref_asr


@dataclass
class AdamState:
    """torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8, no weight decay/amsgrad."""
    step: int = 0
    m: dict = field(default_factory=dict)
    v: dict = field(default_factory=dict)


def clip_grad_norm_(grads: dict, max_norm: float) -> float:
    """torch.nn.utils.clip_grad_norm_ (L2): coef = max_norm / (total + 1e-6), clamped to 1."""
    total = torch.sqrt(sum((g.detach().double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads.values():
        g.mul_(coef)
    return float(total)


def sgd_step_(params: dict, grads: dict, lr: float):
    """torch.optim.SGD(lr) without momentum: p <- p - lr*g."""
    for k in params:
        params[k].add_(grads[k], alpha=-lr)


def adam_step_(params: dict, grads: dict, st: AdamState, lr: float,
               b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam single-tensor update:
    m=b1 m+(1-b1) g; v=b2 v+(1-b2) g^2; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)."""
    st.step += 1
    bc1 = 1.0 - b1 ** st.step
    bc2 = 1.0 - b2 ** st.step
    for k, p in params.items():
        g = grads[k]
        if k not in st.m:
            st.m[k] = torch.zeros_like(p)
            st.v[k] = torch.zeros_like(p)
        st.m[k].lerp_(g, 1 - b1)
        st.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (st.v[k].sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(st.m[k], denom, value=-(lr / bc1))


def loss_and_grads(params: dict, cfg, batch, scale: float = 1.0, smoothing: float = 0.0,
                   train: bool = False, bufs=None):
    """One forward_one_batch + backward (transient_trainer.py:25-46,198-199).
    batch = (x (B,1,F,T) f32, lengths (B,), trg (B,L) i64).  Returns (loss, grads of
    scale*loss, gold, hyp)."""
    x, lengths, trg = batch
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    pred, gold, hyp = ref_asr.forward(leaves, cfg, x, lengths, trg, bufs=bufs, train=train)
    loss = ref_asr.ce_loss(pred, gold, smoothing)
    (loss * scale).backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return float(loss.detach()), grads, gold, hyp, pred.detach()


def meta_step(params: dict, adam: AdamState, cfg, tasks, val, lr: float, meta_lr: float,
              clip: bool = False, max_norm: float = 400.0, smoothing: float = 0.0,
              train: bool = False, bufs=None):
    """TransientTrainer.train loop body, is_copy_grad=True (transient_trainer.py:150-255).

    tasks: list of N train batches; val: the single shared val batch (:168-169).
    Mutates ``params`` / ``adam`` in place.  Returns dict(loss=mean val loss as printed at :268,
    copy_grad=accumulated outer gradient, val_losses, hyps)."""
    n = len(tasks)
    theta0 = {k: v.detach().clone() for k, v in params.items()}             # :160
    cg = {k: torch.zeros_like(v) for k, v in params.items()}                 # :165
    val_losses, tr_losses, hyps = [], [], []
    for tr in tasks:
        lt, g, _, hyp_tr, _ = loss_and_grads(params, cfg, tr, 1.0, smoothing, train, bufs)  # :188-199
        tr_losses.append(lt)
        if clip:
            clip_grad_norm_(g, max_norm)                                     # :205-206
        sgd_step_(params, g, lr)                                             # :207
        lv, gv, _, hyp_v, _ = loss_and_grads(params, cfg, val, 1.0 / n, smoothing, train, bufs)  # :215-227
        val_losses.append(lv)
        hyps.append((hyp_tr, hyp_v))
        for k in cg:                                                         # :229 (+ leak of g_tr)
            cg[k] += g[k] + gv[k]
        for k in params:                                                     # :237
            params[k].copy_(theta0[k])
    grads = {k: v.clone() for k, v in cg.items()}                            # :248
    if clip:
        clip_grad_norm_(grads, max_norm)                                     # :253-254
    adam_step_(params, grads, adam, meta_lr)                                 # :255
    return dict(loss=sum(val_losses) / n, val_losses=val_losses, tr_losses=tr_losses,
                copy_grad=cg, hyps=hyps)


def joint_step(params: dict, adam: AdamState, cfg, tasks, lr: float,
               clip: bool = False, max_norm: float = 400.0, smoothing: float = 0.0,
               train: bool = False, bufs=None):
    """JointTrainer.train loop body without discriminator (joint_trainer.py:178-271):
    grads of sum_i L_i/N, optional clip, Adam(lr)."""
    n = len(tasks)
    acc = {k: torch.zeros_like(v) for k, v in params.items()}
    losses = []
    for tr in tasks:
        l, g, *_ = loss_and_grads(params, cfg, tr, 1.0 / n, smoothing, train, bufs)
        losses.append(l)
        for k in acc:
            acc[k] += g[k]
    if clip:
        clip_grad_norm_(acc, max_norm)
    adam_step_(params, acc, adam, lr)
    return dict(loss=sum(losses) / n, losses=losses, grads=acc)


# --------------------------------------------------------------------------- synthetic workload

def synth_batch(cfg, k: int, t_frames: int, l_tokens: int, seed: int, lengths=None, tgt_lengths=None):
    """SURVEY.md section 8d synthetic data: x ~ N(0,1) (k,1,F,T) f32, targets U[4,V) (k,L) i64.
    Optional ragged lengths: frames beyond a row's length are zeroed (zero-padded collate,
    utils/data_loader.py:284-300), targets beyond tgt_lengths are PAD."""
    rng = np.random.default_rng(seed)          # PCG64: stream is stable across numpy versions
    x = torch.from_numpy(rng.standard_normal((k, 1, cfg.n_freq, t_frames), dtype=np.float32))
    y = torch.from_numpy(rng.integers(4, cfg.vocab, size=(k, l_tokens), dtype=np.int64))
    lens = torch.full((k,), t_frames, dtype=torch.int32)
    if lengths is not None:
        lens = torch.tensor(lengths, dtype=torch.int32)
        for i, ln in enumerate(lengths):
            x[i, :, :, ln:] = 0
    if tgt_lengths is not None:
        for i, ln in enumerate(tgt_lengths):
            y[i, ln:] = ref_asr.PAD_ID
    return x, lens, y
