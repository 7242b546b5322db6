"""``RNNModel`` with the reference's API (lm/model/rnn_model.py:12-70) on the B200 engine.

Same constructor signature, sub-module / parameter names and order (``encoder``, ``rnn``, ``decoder`` -- so
``state_dict`` interchanges), same ``init_weights`` and ``init_hidden``; ``forward(input, hidden) -> (decoded, hidden)``.
After ``.cuda()`` every parameter is a view into one flat fp32 arena and ``forward`` is one call into libmtl_b200
(embedding + dropout + LSTM stack + dropout + decoder).  The meta-loop of lm/main_meta_transfer.py drives the arenas
directly (``lm.meta.LMMetaTrainer``).  Only the LSTM type is on the device path; there is no CPU arithmetic path."""
from __future__ import annotations

import itertools

import torch
import torch.nn as nn

import mtl_b200


class RNNModel(nn.Module):
    """Container module with an encoder, a recurrent module, and a decoder."""

    def __init__(self, rnn_type, ntoken, ninp, nhid, nlayers, dropout=0.5, tie_weights=False):
        super().__init__()
        if rnn_type != 'LSTM':
            raise NotImplementedError(f"--model {rnn_type}: only the LSTM (the script's default, "
                                      "lm/main_meta_transfer.py:23) is implemented on the device")
        if tie_weights:
            raise NotImplementedError("--tied (decoder.weight = encoder.weight) is not implemented on the device path")
        self.drop = nn.Dropout(dropout)
        self.encoder = nn.Embedding(ntoken, ninp)
        self.rnn = nn.LSTM(ninp, nhid, nlayers, dropout=dropout)
        self.decoder = nn.Linear(nhid, ntoken)
        self.rnn_type = rnn_type
        self.ntoken, self.ninp = ntoken, ninp
        self.nhid = nhid
        self.nlayers = nlayers
        self.dropout = dropout
        self.init_weights()
        self._session = None
        self._theta = None
        self._seed = itertools.count(int(torch.initial_seed()) & 0x7FFFFFFF)

    def init_weights(self):
        initrange = 0.1
        self.encoder.weight.data.uniform_(-initrange, initrange)
        self.decoder.bias.data.fill_(0)
        self.decoder.weight.data.uniform_(-initrange, initrange)

    # ------------------------------------------------------------------ engine binding
    def spec(self) -> mtl_b200.LmSpec:
        return mtl_b200.LmSpec(vocab=self.ntoken, ninp=self.ninp, nhid=self.nhid, nlayers=self.nlayers)

    def _bind(self, device):
        s = mtl_b200.LmSession(self.spec(), device)
        named = list(self.named_parameters())
        assert [n for n, _ in named] == [n for n, *_ in s.table], "parameter order differs from the engine layout"
        theta = s.new_arena()
        tv = s.views(theta)
        for name, p in named:
            tv[name].copy_(p.data)
            p.data = tv[name]
            p.grad = None
        self._session, self._theta = s, theta

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
    def session(self) -> mtl_b200.LmSession:
        if self._session is None:
            raise mtl_b200.MtlError("the model is not on a CUDA device: call model.cuda() first (no CPU path)")
        return self._session

    def arena(self) -> torch.Tensor:
        self.session
        return self._theta

    # ------------------------------------------------------------------ reference API
    def forward(self, input, hidden):
        """input (T, B) token ids, hidden (h, c) -> (decoded (T, B, ntoken), (h, c)).  Inference / evaluation surface:
        the result carries no autograd graph (training goes through ``lm.meta.LMMetaTrainer``)."""
        s = self.session
        p = self.dropout if self.training else 0.0
        out = s.run(self._theta, input, hidden=hidden, dropout=p, seed=next(self._seed), want_logits=True)
        return out["logits"], out["hidden"]

    def init_hidden(self, bsz):
        weight = next(self.parameters()).data
        return (weight.new_zeros(self.nlayers, bsz, self.nhid), weight.new_zeros(self.nlayers, bsz, self.nhid))
