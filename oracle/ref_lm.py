"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the LSTM language-model path.

Functional restatement (PyTorch CPU ops, autograd) of
  * RNNModel.forward             lm/model/rnn_model.py:54-62  (Embedding -> Dropout -> nn.LSTM -> Dropout -> Linear)
  * the meta-iteration           lm/main_meta_transfer.py:293-372, as the FIRST-ORDER meta-gradient (each task's validation
                                 gradient at its adapted weights): the reference's own `batch_loss.backward()` after
                                 `load_state_dict` raises on torch >= 1.5, so the reference cannot be run for this step.
Pin: `forward` / `loss_and_grads` are checked against the live reference RNNModel + nn.CrossEntropyLoss
(tests/test_oracle_vs_reference.py::test_lm_*) and against tests/golden/lm_small.npz (oracle/make_golden.py).  The
meta-step has no runnable reference: PARITY UNPINNED for `meta_step` beyond its building blocks (stated in DESIGN.md).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class LmConfig:
    vocab: int = 10000
    ninp: int = 200
    nhid: int = 200
    nlayers: int = 2


LM_SMALL = LmConfig(vocab=53, ninp=16, nhid=24, nlayers=2)
LM_CFG5 = LmConfig()                 # lm/main_meta_transfer.py:25-36 defaults; vocabulary size synthetic (data not shipped)


def param_names(cfg: LmConfig) -> List[str]:
    out = ["encoder.weight"]
    for l in range(cfg.nlayers):
        out += [f"rnn.weight_ih_l{l}", f"rnn.weight_hh_l{l}", f"rnn.bias_ih_l{l}", f"rnn.bias_hh_l{l}"]
    return out + ["decoder.weight", "decoder.bias"]


def init_params(cfg: LmConfig, seed: int) -> Dict[str, torch.Tensor]:
    """Same distributions as RNNModel.__init__ + init_weights (uniform(-0.1, 0.1) tables, nn.LSTM's U(-1/sqrt(H), 1/sqrt(H)))."""
    g = torch.Generator().manual_seed(seed)
    u = lambda *s, a: (torch.rand(*s, generator=g) * 2 - 1) * a
    k = cfg.nhid ** -0.5
    p = {"encoder.weight": u(cfg.vocab, cfg.ninp, a=0.1)}
    for l in range(cfg.nlayers):
        n_in = cfg.ninp if l == 0 else cfg.nhid
        p[f"rnn.weight_ih_l{l}"] = u(4 * cfg.nhid, n_in, a=k)
        p[f"rnn.weight_hh_l{l}"] = u(4 * cfg.nhid, cfg.nhid, a=k)
        p[f"rnn.bias_ih_l{l}"] = u(4 * cfg.nhid, a=k)
        p[f"rnn.bias_hh_l{l}"] = u(4 * cfg.nhid, a=k)
    p["decoder.weight"] = u(cfg.vocab, cfg.nhid, a=0.1)
    p["decoder.bias"] = torch.zeros(cfg.vocab)
    return p


def forward(p: Dict[str, torch.Tensor], cfg: LmConfig, tokens: torch.Tensor, hidden=None):
    """tokens (T, B) -> logits (T, B, V), (h, c) each (L, B, H).  Eval-mode (no dropout): lm/model/rnn_model.py:54-62 with
    nn.LSTM's cell written out (gate order i, f, g, o; gates = x W_ih^T + b_ih + h W_hh^T + b_hh)."""
    T, B = tokens.shape
    H = cfg.nhid
    x = F.embedding(tokens, p["encoder.weight"])
    hs, cs = [], []
    for l in range(cfg.nlayers):
        h = hidden[0][l] if hidden is not None else x.new_zeros(B, H)
        c = hidden[1][l] if hidden is not None else x.new_zeros(B, H)
        xg = x @ p[f"rnn.weight_ih_l{l}"].t() + p[f"rnn.bias_ih_l{l}"]
        ys = []
        for t in range(T):
            g = xg[t] + h @ p[f"rnn.weight_hh_l{l}"].t() + p[f"rnn.bias_hh_l{l}"]
            i, f, gg, o = g.split(H, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            ys.append(h)
        x = torch.stack(ys)
        hs.append(h)
        cs.append(c)
    logits = x.reshape(T * B, H) @ p["decoder.weight"].t() + p["decoder.bias"]
    return logits.view(T, B, cfg.vocab), (torch.stack(hs), torch.stack(cs))


def loss_and_grads(p, cfg, tokens, targets, hidden=None):
    """nn.CrossEntropyLoss()(output.view(-1, ntokens), targets) and its gradient w.r.t. every parameter
    (lm/main_meta_transfer.py:316-319).  Returns (loss, grads, logits, hidden_out detached)."""
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    logits, hid = forward(q, cfg, tokens, hidden)
    loss = F.cross_entropy(logits.view(-1, cfg.vocab), targets.view(-1))
    grads = torch.autograd.grad(loss, [q[k] for k in param_names(cfg)])
    return float(loss.detach()), dict(zip(param_names(cfg), grads)), logits.detach(), (hid[0].detach(), hid[1].detach())


def clip_(grads: Dict[str, torch.Tensor], max_norm: float):
    """torch.nn.utils.clip_grad_norm_ on a dict of gradients (in place); returns the total norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads.values():
        g.mul_(coef)
    return float(total)


def meta_step(p, cfg, train: Sequence[Tuple[torch.Tensor, torch.Tensor]], val, weights, hidden, lr, meta_lr_factor, clip):
    """One meta-iteration (first-order), dropout off.  Returns (new params, new hidden, train losses, val losses, meta grad)."""
    names = param_names(cfg)
    meta = {k: torch.zeros_like(p[k]) for k in names}
    tr_losses, val_losses = [], []
    for i, (tok, trg) in enumerate(train):
        loss, g, _, hidden = loss_and_grads(p, cfg, tok, trg, hidden)        # train pass at theta0, from the running hidden
        tr_losses.append(loss)
        if clip:
            clip_(g, clip)
        pi = {k: p[k] - (lr / meta_lr_factor) * g[k] for k in names}         # inner SGD step
        vloss, vg, _, _ = loss_and_grads(pi, cfg, val[0], val[1], hidden)    # val pass at theta_i with the train pass's hidden
        val_losses.append(vloss)
        for k in names:
            meta[k] += weights[i] * vg[k]
    if clip:
        clip_(meta, clip)
    new_p = {k: p[k] - lr * meta[k] for k in names}
    return new_p, hidden, tr_losses, val_losses, meta


def synth_blocks(cfg: LmConfig, n_tasks: int, T: int, B: int, seed: int):
    """Synthetic token blocks of the LMDataset.sample contract: per task (tokens (T, B), targets (T*B,)) + the shared val block."""
    g = torch.Generator().manual_seed(seed)
    def block():
        stream = torch.randint(0, cfg.vocab, (T + 1, B), generator=g)
        return stream[:T].contiguous(), stream[1:].reshape(-1).contiguous()
    return [block() for _ in range(n_tasks)], block()
