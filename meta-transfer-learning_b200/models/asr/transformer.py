"""``Transformer`` with the reference's API (models/asr/transformer.py:15-240) on the B200 engine.

What is the same as the reference: constructor signature, sub-module / parameter / buffer names and
order (so ``state_dict`` interchanges and ``model.parameters()`` has the reference's copy_grad order),
the model-wide xavier_uniform re-initialisation, ``forward -> (pred, gold, hyp)``, ``encode`` /
``decode`` names, and the copy-grad buffer API (``zero_copy_grad`` / ``add_copy_grad`` /
``to_copy_grad`` / ``from_copy_grad``).

What is different underneath: after ``.cuda()`` every parameter is a VIEW into one flat fp32 arena and
every ``.grad`` a view into a sibling arena; ``forward`` is a single call into libmtl_b200 (VGG
front-end, encoder, decoder, vocabulary projection, CE and top-1 fused over a bump workspace) and
``loss.backward()`` is a single call that accumulates into the gradient arena.  ``pred`` is a real
autograd tensor, so reference-style code (``calculate_metrics`` -> ``loss.backward()`` ->
``torch.optim`` steps on ``model.parameters()`` -> ``load_state_dict``) keeps working unchanged; the
trainers in ``trainer/asr`` bypass autograd altogether and drive the arenas directly.

There is no CPU arithmetic path: calling the model before ``.cuda()`` raises."""
from __future__ import annotations

import itertools

import weakref

import torch
import torch.nn as nn

import mtl_b200
from mtl_b200.session import Batch


class _FusedForward(torch.autograd.Function):
    """pred = engine.forward(theta, batch); backward pushes d(loss)/d(pred) through the engine, which
    accumulates parameter gradients straight into the gradient arena (the ``.grad`` views)."""

    @staticmethod
    def forward(ctx, hook, model, batch, dropout, seed):
        out = model._session.forward(model._theta, batch, dropout=dropout, seed=seed, smoothing=model.label_smoothing)
        ctx.model = model
        ctx.ticket = model._ticket = next(model._tickets)
        model._last = out
        pred = out["pred"].clone()          # the workspace copy is overwritten by the backward pass
        ctx.mark_non_differentiable(out["gold"], out["hyp"])
        return pred, out["gold"], out["hyp"]

    @staticmethod
    def backward(ctx, dpred, _g, _h):
        ctx.model._engine_backward(ctx.ticket, dpred=dpred)
        return None, None, None, None, None


class _FusedLoss(torch.autograd.Function):
    """Scalar CE of the last forward (already computed by the fused CE kernel); its backward uses the
    fused softmax-minus-onehot gradient instead of materialising d(pred) on the host side."""

    @staticmethod
    def forward(ctx, hook, model, ticket):
        # hangs off the model's hook tensor, NOT off pred: nothing must flow back into _FusedForward.backward
        ctx.model, ctx.ticket = model, ticket
        return model._last["ce"][0].clone()

    @staticmethod
    def backward(ctx, dloss):
        ctx.model._engine_backward(ctx.ticket, scale=float(dloss))
        return None, None, None


class Transformer(nn.Module):
    """
    args:
        encoder: modules.Encoder, decoder: modules.Decoder, vocab: utils.data.Vocab
    """

    def __init__(self, encoder, decoder, vocab, feat_extractor='vgg_cnn', train=True, is_factorized=False, r=100):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.vocab = vocab
        self.feat_extractor = feat_extractor
        self.is_factorized = is_factorized
        self.r = r
        self.copy_grad = None
        print("feat extractor:", feat_extractor)
        if feat_extractor != 'vgg_cnn':
            raise NotImplementedError(f"feat_extractor={feat_extractor!r}: only the VGG front-end "
                                      "(models/asr/transformer.py:47-59) is on the B200 hot path")
        chans = [(1, 64), (64, 64), None, (64, 128), (128, 128), None]       # None = 2x2 max-pool
        layers = []
        for c in chans:
            layers += [nn.MaxPool2d(2, stride=2)] if c is None else [nn.Conv2d(c[0], c[1], 3, stride=1, padding=1), nn.ReLU()]
        self.conv = nn.Sequential(*layers)                                   # indices 0,2,5,7 hold the convolutions
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        # engine state (bound by .cuda())
        self.label_smoothing = 0.0           # smoothing the fused CE uses (trainers set it from args)
        self._session = None
        self._theta = self._grad = self._cg = None
        self._tickets = itertools.count(1)
        self._ticket = 0
        self._last = None
        self._seed = itertools.count(int(torch.initial_seed()) & 0x7FFFFFFF)
        self._direct_grads = False           # True while a trainer writes the gradient arena directly (no autograd)

    # ------------------------------------------------------------------ engine binding
    def spec(self) -> mtl_b200.ModelSpec:
        e, d = self.encoder, self.decoder
        if not (e.dim_model == d.dim_model and e.num_heads == d.num_heads and e.dim_key == d.dim_key
                and e.dim_value == d.dim_value and e.dim_inner == d.dim_inner and e.r == d.r):
            raise ValueError("encoder and decoder must share dim_model/heads/dim_key/dim_value/dim_inner/r")
        if e.dim_input % 128:
            raise ValueError("dim_input must be 128 * (n_freq // 4) (utils/functions.py:318-321)")
        f4 = e.dim_input // 128
        # n_freq is only known up to the floor-pooling: any F with (F//2)//2 == f4 has the same parameters;
        # the engine is told the exact F of the first batch it sees (see _session_for)
        return mtl_b200.ModelSpec(n_enc=e.num_layers, n_dec=d.num_layers, d_model=e.dim_model, n_heads=e.num_heads,
                                  d_k=e.dim_key, d_v=e.dim_value, d_inner=e.dim_inner, rank=e.r,
                                  vocab=len(self.vocab.label2id), n_freq=4 * f4 + 1,
                                  src_max_len=e.src_max_length, tgt_max_len=d.trg_max_length)

    def _bind(self, device):
        """Moves the parameters into one flat arena on ``device`` and makes every Parameter a view of it."""
        if not torch.cuda.is_available():
            raise mtl_b200.MtlError("models.asr.transformer.Transformer needs a CUDA device: libmtl_b200 has no CPU path")
        spec = self.spec()
        s = mtl_b200.Session(spec, device)
        named = list(self.named_parameters())
        assert [n for n, _ in named] == [n for n, *_ in s.table], "parameter order differs from the engine layout"
        theta, grad = s.new_arena(), s.new_arena()
        tv, gv = s.views(theta), s.views(grad)
        for name, p in named:
            tv[name].copy_(p.data)
            p.data = tv[name]
            p.grad = None
        self._session, self._theta, self._grad, self._cg = s, theta, grad, None
        self._grad_views = [gv[n] for n, _ in named]
        self._params = [p for _, p in named]
        self.copy_grad = None
        # the engine reads the PE buffers of the state_dict (so a loaded checkpoint's tables are honoured)
        s.pe_enc = self.encoder.positional_encoding.pe[0]
        s.pe_dec = self.decoder.positional_encoding.pe[0]
        self._hook = torch.zeros(1, device=device, requires_grad=True)
        self.decoder.__dict__["_owner"] = weakref.ref(self)      # Decoder.greedy_search reaches the engine through the model

    def cuda(self, device=None):
        if not torch.cuda.is_available():
            raise mtl_b200.MtlError("model.cuda(): no CUDA device (sm_100a) is visible and libmtl_b200 has no CPU path")
        return super().cuda(device)

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        p0 = next(self.parameters())
        if p0.is_cuda:
            self._bind(p0.device)
        else:
            self._session = None
        return self

    @property
    def session(self):
        if self._session is None:
            raise mtl_b200.MtlError("the model is not on a CUDA device: call model.cuda() first (no CPU path)")
        return self._session

    def arenas(self):
        """(theta, grad) flat fp32 arenas the parameters / gradients are views of."""
        self.session
        return self._theta, self._grad

    def _attach_grads(self):
        """Makes every ``.grad`` the arena view; gradients dropped by ``zero_grad(set_to_none=True)`` restart at 0."""
        missing = [i for i, p in enumerate(self._params) if p.grad is None]
        if len(missing) == len(self._params):
            self._session.zero(self._grad)
        else:
            for i in missing:
                self._grad_views[i].zero_()
        for i, p in enumerate(self._params):
            g = self._grad_views[i]
            if p.grad is None:
                p.grad = g
            elif p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
                p.grad = g

    def grads_pending(self) -> bool:
        """True when the gradient arena holds gradients an optimizer step should consume."""
        return self._direct_grads or any(p.grad is not None for p in self._params)

    def _adopt_foreign_grads(self):
        """Copies gradients that were assigned from outside (``p.grad = t``) into the arena views."""
        for i, p in enumerate(self._params):
            g = self._grad_views[i]
            if p.grad is not None and p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
                p.grad = g

    def _engine_backward(self, ticket, scale=1.0, dpred=None):
        if ticket != self._ticket:
            raise RuntimeError("backward through a forward pass that is no longer the model's latest: the engine "
                               "keeps one activation record (call backward before the next forward)")
        self._attach_grads()
        self._session.backward(self._theta, self._grad, scale, dpred=dpred)
        self._ticket = 0

    # ------------------------------------------------------------------ reference API
    def _batch(self, padded_input, padded_target, input_lengths):
        s = self.session
        dev = s.device
        x = padded_input.to(device=dev, dtype=torch.float32).contiguous()
        if x.dim() != 4 or x.size(1) != 1:
            raise ValueError("padded_input must be (B, 1, freq, T)")
        if (x.size(2) // 2) // 2 != self.encoder.dim_input // 128:
            raise ValueError(f"{x.size(2)} frequency bins do not match dim_input={self.encoder.dim_input}")
        if x.size(2) != s.spec.n_freq:       # first batch fixes the exact bin count (same parameter shapes)
            self._rebuild_session(x.size(2))
            s = self._session
        lens = torch.as_tensor(input_lengths).to(device=dev, dtype=torch.int32).contiguous()
        trg = padded_target.to(device=dev, dtype=torch.int64).contiguous()
        n = int((padded_target != self.vocab.PAD_ID).sum(dim=1).max().item()) + 1
        return Batch(x, lens, trg, n)

    def _rebuild_session(self, n_freq):
        import dataclasses
        old = self._session
        s = mtl_b200.Session(dataclasses.replace(old.spec, n_freq=n_freq), old.device)
        s.pe_enc, s.pe_dec = old.pe_enc, old.pe_dec
        assert s.n_floats == old.n_floats
        self._session = s

    def forward(self, padded_input, input_lengths, padded_target, verbose=False):
        """(B,1,F,T) spectrograms, (B) raw frame counts, (B,L) PAD-padded targets ->
        pred (B,n,V) logits, gold (B,n), hyp (B,n) with n = longest target + 1 (transformer.py:120-149)."""
        batch = self._batch(padded_input, padded_target, input_lengths)
        drop = float(self.encoder.dropout_rate) if self.training else 0.0
        if torch.is_grad_enabled():
            pred, gold, hyp = _FusedForward.apply(self._hook, self, batch, drop, next(self._seed))
        else:
            out = self._session.forward(self._theta, batch, dropout=drop, seed=next(self._seed),
                                        smoothing=self.label_smoothing)
            self._last, self._ticket = out, 0
            pred, gold, hyp = out["pred"].clone(), out["gold"], out["hyp"]
        pred._mtl_owner = (self, self._ticket)
        return pred, gold.long(), hyp.long()

    def fused_loss(self, pred, smoothing=0.0):
        """CE of ``pred`` if it is this model's latest forward output and was computed with the same label
        smoothing, as an autograd scalar wired to the fused backward; else None (caller falls back to torch)."""
        owner = getattr(pred, "_mtl_owner", None)
        if owner is None or owner[0] is not self or self._last is None or float(smoothing) != float(self.label_smoothing):
            return None
        if owner[1] == 0 or not pred.requires_grad:
            return self._last["ce"][0].clone()
        if owner[1] != self._ticket:
            return None
        return _FusedLoss.apply(self._hook, self, owner[1])

    def encode(self, padded_input, input_lengths):
        """(B,1,F,T) spectrograms, (B) raw frame counts -> encoder_padded_outputs (B, T', H)  (transformer.py:78-98)."""
        s = self.session
        x = padded_input.to(device=s.device, dtype=torch.float32).contiguous()
        if x.size(2) != s.spec.n_freq:
            self._rebuild_session(x.size(2))
            s = self._session
        lens = torch.as_tensor(input_lengths).to(device=s.device, dtype=torch.int32).contiguous()
        return s.encode(self._theta, x, lens)

    def decode(self, encoder_padded_outputs, input_lengths, padded_target):
        raise NotImplementedError("teacher-forced decoding from a detached encoder output is not exposed: the engine runs "
                                  "encoder and decoder as one pass -- call the model (forward) instead")

    def evaluate(self, padded_input, input_lengths, padded_target, args, beam_search=False, beam_width=0, beam_nbest=0,
                 lm=None, lm_rescoring=False, lm_weight=0.1, c_weight=1, start_token=-1, verbose=False, max_steps=300):
        """-> (None, strs_hyps, strs_gold)  (transformer.py:162-202): greedy 1-best strings of the batch and the gold
        strings (every position of <y><EOS><PAD>..., special tokens included, as the reference joins them)."""
        if beam_search or lm is not None or lm_rescoring:
            raise NotImplementedError("beam search / LM rescoring (decoder.py:186-291) are outside the B200 hot path; "
                                      "greedy search (the reference's own fallback, transformer.py:191-199) is implemented")
        enc = self.encode(padded_input, input_lengths)
        _, gold = self.decoder.preprocess(padded_target.cpu())
        strs_gold = ["".join(self.vocab.id2label[int(x)] for x in row) for row in gold]
        strs_hyps = self.decoder.greedy_search(enc, args, start_token=start_token, max_steps=max_steps)
        return None, strs_hyps, strs_gold


    # ------------------------------------------------------------------ copy-grad buffer (transformer.py:204-240)
    def init_copy_grad_(self):
        s = self.session
        self._cg = s.new_arena()
        v = s.views(self._cg)
        self.copy_grad = [v[n] for n, *_ in s.table]

    def zero_copy_grad(self):
        if self._cg is None:
            self.init_copy_grad_()
        else:
            self._session.zero(self._cg)

    def add_copy_grad(self):
        if self._cg is None:
            self.init_copy_grad_()
        self._attach_grads()
        self._session.axpy(self._cg, self._grad, 1.0)

    def to_copy_grad(self):
        if self._cg is None:
            self.init_copy_grad_()
        self._attach_grads()
        self._session.copy(self._cg, self._grad)

    def from_copy_grad(self):
        if self._cg is None:
            self.init_copy_grad_()
        self._attach_grads()
        self._session.copy(self._grad, self._cg)
