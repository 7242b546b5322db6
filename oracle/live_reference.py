"""Drive the UNMODIFIED reference (read-only at /root/reference) as a live oracle
(TEST INFRASTRUCTURE).  Only usable in the build container: the GPU box has no
/root/reference, so nothing here is reachable from ``-m gpu`` tests, smoke() or bench.py.

Procedure = SURVEY.md Appendix C: two stub modules (``Levenshtein``, ``stanfordcorenlp``),
``args.cuda=True`` with ``Tensor.cuda`` patched to identity (works around the unbound
``val_cuda_inputs`` at trainer/asr/transient_trainer.py:210-215).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MTL_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "meta_transfer_train.py"))


def _edit_distance(a: str, b: str) -> int:
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


_SHADOWED = ("utils", "models", "modules", "trainer")


@contextlib.contextmanager
def reference_imports():
    """Context in which ``import utils.*``, ``models.*`` ... resolve to the reference tree.
    Restores sys.path / sys.modules afterwards so our own compat packages of the same
    names can be imported in the same process."""
    saved = {k: v for k, v in sys.modules.items()
             if k.split(".")[0] in _SHADOWED or k in ("Levenshtein", "stanfordcorenlp")}
    for k in saved:
        del sys.modules[k]
    lev = types.ModuleType("Levenshtein"); lev.distance = _edit_distance
    nlp = types.ModuleType("stanfordcorenlp"); nlp.StanfordCoreNLP = object
    sys.modules["Levenshtein"] = lev
    sys.modules["stanfordcorenlp"] = nlp
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            yield
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules
                  if k.split(".")[0] in _SHADOWED or k in ("Levenshtein", "stanfordcorenlp")]:
            del sys.modules[k]
        sys.modules.update(saved)


def make_args(cfg, **over):
    a = argparse.Namespace(
        feat_extractor="vgg_cnn", sample_rate=(cfg.n_freq - 1) * 100, window_size=0.02, feat="spectrogram",
        num_enc_layers=cfg.n_enc, num_dec_layers=cfg.n_dec, num_heads=cfg.n_heads,
        dim_model=cfg.d_model, dim_key=cfg.d_k, dim_value=cfg.d_v, dim_input=cfg.n_freq,
        dim_inner=cfg.d_inner, dim_emb=cfg.d_model, src_max_len=cfg.src_max_len,
        tgt_max_len=cfg.tgt_max_len, dropout=cfg.dropout, emb_trg_sharing=False,
        label_smoothing=0.0, name="oracle", lr=1e-4, meta_lr=1e-4, k_train=2, k_valid=2,
        cuda=True, clip=False, max_norm=400, save_every=10 ** 9, save_folder="/tmp/mtl_oracle_save")
    for k, v in over.items():
        setattr(a, k, v)
    return a


def build_model(cfg, params: dict, **arg_over):
    """Instantiate the reference Transformer for ``cfg`` and load ``params`` into it."""
    with reference_imports():
        from utils.data import Vocab
        from utils.functions import init_transformer_model
        vocab = Vocab()
        for i in range(cfg.vocab - 4):
            vocab.add_label("w%d" % i)
        args = make_args(cfg, **arg_over)
        with contextlib.redirect_stdout(io.StringIO()):
            model = init_transformer_model(args, vocab, is_factorized=False, r=cfg.rank)
    assert args.dim_input == cfg.d_input
    missing = model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=False)
    assert not missing.unexpected_keys and all(k.endswith(".pe") for k in missing.missing_keys)
    return model, vocab, args


class ListSampler:
    """Object honouring SpectrogramDataset.sample (utils/data_loader.py:245-321): returns
    ((x, sizes, pct, y, ysz), (same for val)); pops one pre-built entry per call."""

    def __init__(self, entries):
        self.entries = list(entries)
        self.i = 0

    def sample(self, k_train, k_valid, manifest_id):
        e = self.entries[min(self.i, len(self.entries) - 1)]
        self.i += 1
        return e


def sampler_tuple(batch):
    x, lens, y = batch
    ysz = (y != 0).sum(1).to(torch.int32)
    return (x.clone(), lens.clone(), torch.ones(len(lens)), y.clone(), ysz)


def run_transient(model, vocab, args, steps_tasks, steps_val, n_steps):
    """Run TransientTrainer.train UNCHANGED for n_steps.  steps_tasks[s][i] = train batch of
    task i at step s; steps_val[s] = the shared val batch (delivered through the last
    manifest's sampler, transient_trainer.py:168-169).  Returns captured stdout."""
    n_tasks = len(steps_tasks[0])
    samplers = []
    for i in range(n_tasks):
        # the trainer prefetches one extra step: repeat the last entry
        ent = [(sampler_tuple(steps_tasks[s][i]), sampler_tuple(steps_val[s])) for s in range(n_steps)]
        samplers.append(ListSampler(ent))
    buf = io.StringIO()
    with reference_imports():
        from trainer.asr.transient_trainer import TransientTrainer
        orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            with contextlib.redirect_stdout(buf):
                TransientTrainer().train(model, vocab, samplers, [], "ce", 0, n_steps, args,
                                         evaluate_every=10 ** 9, early_stop="cer,200",
                                         is_copy_grad=True)
        finally:
            torch.Tensor.cuda = orig
    out = buf.getvalue()
    if "Error:" in out:
        raise RuntimeError("reference trainer swallowed an exception:\n" + out[-2000:])
    return out


def run_joint(model, vocab, args, steps_tasks, n_steps):
    n_tasks = len(steps_tasks[0])
    samplers = []
    for i in range(n_tasks):
        ent = [(sampler_tuple(steps_tasks[s][i]), sampler_tuple(steps_tasks[s][i])) for s in range(n_steps)]
        samplers.append(ListSampler(ent))
    buf = io.StringIO()
    args.cuda = False
    with reference_imports():
        from trainer.asr.joint_trainer import JointTrainer
        with contextlib.redirect_stdout(buf):
            JointTrainer().train(model, vocab, samplers, [], "ce", 0, n_steps, args,
                                 evaluate_every=10 ** 9, early_stop="cer,200", is_copy_grad=True)
    out = buf.getvalue()
    if "Error:" in out:
        raise RuntimeError("reference trainer swallowed an exception:\n" + out[-2000:])
    return out


# ----------------------------------------------------------------------------- LM sub-project (lm/model/rnn_model.py)
def build_lm_model(cfg, params):
    """The UNMODIFIED reference RNNModel('LSTM', ...) with `params` loaded (eval mode: dropout off).  The module is loaded
    from its file by path so that neither `model` nor `util` of the reference's lm/ directory shadows anything here."""
    import importlib.util
    path = os.path.join(REFERENCE_ROOT, "lm", "model", "rnn_model.py")
    spec = importlib.util.spec_from_file_location("_ref_lm_rnn_model", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    m = mod.RNNModel('LSTM', cfg.vocab, cfg.ninp, cfg.nhid, cfg.nlayers, dropout=0.2, tie_weights=False)
    m.load_state_dict({k: v.clone() for k, v in params.items()})
    m.eval()
    return m


def lm_fwd_bwd(cfg, params, tokens, targets, hidden=None):
    """Reference forward + nn.CrossEntropyLoss + backward (lm/main_meta_transfer.py:268-319), eval-mode dropout."""
    import torch
    m = build_lm_model(cfg, params)
    hid = m.init_hidden(tokens.shape[1]) if hidden is None else tuple(h.clone() for h in hidden)
    m.zero_grad()
    out, hid = m(tokens, hid)
    loss = torch.nn.CrossEntropyLoss()(out.view(-1, cfg.vocab), targets)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    return float(loss.detach()), grads, out.detach(), (hid[0].detach(), hid[1].detach())
