"""Parameter containers and mask helpers with the reference's names (modules/common_layers.py).

On the hot path none of the ``forward`` methods below run: models.asr.transformer.Transformer hands the
whole model (all parameters are views of one flat arena) to the CUDA engine, which fuses the masks
(lengths -> in-kernel predicates), the low-rank attention, the FFN and the layer norms.  What these
classes must get right is therefore the *state*: parameter names, shapes, registration order
(= ``model.parameters()`` order = copy_grad order) and the sequence of RNG draws at construction, so
that ``torch.manual_seed(s); init_transformer_model(...)`` yields the reference's initial weights.
The mask helpers are kept as plain (vectorised) tensor functions for callers that import them."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn


class FusedOnly(nn.Module):
    """Base of the layer containers: their arithmetic only exists fused inside the CUDA engine."""

    def forward(self, *args, **kwargs):
        raise RuntimeError(
            f"{type(self).__name__}.forward is not a standalone op in mtl_b200: the layer runs fused inside "
            "models.asr.transformer.Transformer (libmtl_b200); call the model, not the sub-module")


# ------------------------------------------------------------------ padding / masks (common_layers.py:13-84)
def pad_list(xs, pad_value):
    """List of (L_i, ...) tensors -> (N, max L, ...) filled with pad_value (common_layers.py:13-20)."""
    return torch.nn.utils.rnn.pad_sequence(list(xs), batch_first=True, padding_value=pad_value)


def pad_list_with_mask(xs, pad_value):
    """pad_list plus a bool mask that is True on the padding (common_layers.py:22-32)."""
    padded = pad_list(xs, pad_value)
    lens = torch.tensor([x.size(0) for x in xs])
    mask = torch.arange(padded.size(1)).unsqueeze(0) >= lens.unsqueeze(1)
    return padded, mask


def get_non_pad_mask(padded_input, input_lengths=None, pad_idx=None):
    """(N, T, 1) float mask: 1 on real positions.  Either by lengths or by a pad id (common_layers.py:38-54)."""
    assert input_lengths is not None or pad_idx is not None
    if input_lengths is not None:
        n, t = padded_input.size(0), padded_input.size(1)
        lens = torch.as_tensor(input_lengths, device=padded_input.device).long().view(n, 1)
        mask = (torch.arange(t, device=padded_input.device).view(1, t) < lens).to(padded_input.dtype
                                                                                 if padded_input.is_floating_point()
                                                                                 else torch.float32)
    else:
        assert padded_input.dim() == 2
        mask = padded_input.ne(pad_idx).float()
    return mask.unsqueeze(-1)


def get_attn_key_pad_mask(seq_k, seq_q, pad_idx):
    """(N, Tq, Tk) bool mask, True where the KEY is padding (common_layers.py:56-65)."""
    return seq_k.eq(pad_idx).unsqueeze(1).expand(-1, seq_q.size(1), -1)


def get_attn_pad_mask(padded_input, input_lengths, expand_length):
    """Key-padding mask from lengths, expanded over expand_length queries (common_layers.py:67-74)."""
    non_pad = get_non_pad_mask(padded_input, input_lengths=input_lengths)
    return non_pad.squeeze(-1).lt(1).unsqueeze(1).expand(-1, expand_length, -1)


def get_subsequent_mask(seq):
    """(N, T, T) uint8 upper-triangular "future" mask (common_layers.py:76-83)."""
    n, t = seq.size()
    tri = torch.triu(torch.ones((t, t), device=seq.device, dtype=torch.uint8), diagonal=1)
    return tri.unsqueeze(0).expand(n, -1, -1)


# ------------------------------------------------------------------ layers
class PositionalEncoding(nn.Module):
    """Buffer ``pe`` (1, max_length, dim_model): sin on even, cos on odd features (common_layers.py:86-108)."""

    def __init__(self, dim_model, max_length=2000):
        super().__init__()
        pos = torch.arange(0, max_length).unsqueeze(1).float()
        freq = torch.exp(torch.arange(0, dim_model, 2).float() * -(math.log(10000.0) / dim_model))
        table = torch.zeros(max_length, dim_model, requires_grad=False)
        table[:, 0::2] = torch.sin(pos * freq)
        table[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pe", table.unsqueeze(0))

    def forward(self, input):
        return self.pe[:, :input.size(1)]


class PositionwiseFeedForward(FusedOnly):
    """LN(dropout(W2 relu(W1 x + b1) + b2) + x)  (common_layers.py:110-132)."""

    def __init__(self, dim_model, dim_ff, dropout=0.1):
        super().__init__()
        self.linear_1 = nn.Linear(dim_model, dim_ff)
        self.linear_2 = nn.Linear(dim_ff, dim_model)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(dim_model)


class FactorizedMultiHeadAttention(FusedOnly):
    """Low-rank multi-head attention: q/k/v/out projections are d -> r -> H*dk products
    (common_layers.py:238-306).  Registration order q_a, q_b, k_a, k_b, v_a, v_b, layer_norm, out_a, out_b."""

    def __init__(self, num_heads, dim_model, dim_key, dim_value, dropout=0.1, r=100):
        super().__init__()
        self.num_heads, self.dim_model, self.dim_key, self.dim_value, self.r = num_heads, dim_model, dim_key, dim_value, r
        widths = {"query": dim_key, "key": dim_key, "value": dim_value}
        for name, w in widths.items():
            setattr(self, f"{name}_linear_a", nn.Linear(dim_model, r, bias=False))
            setattr(self, f"{name}_linear_b", nn.Linear(r, num_heads * w))
        for name, w in widths.items():      # same draws as the reference (overwritten later by the model-wide xavier pass)
            std = np.sqrt(2.0 / (dim_model + w))
            nn.init.normal_(getattr(self, f"{name}_linear_a").weight, mean=0, std=std)
            nn.init.normal_(getattr(self, f"{name}_linear_b").weight, mean=0, std=std)
        self.attention = ScaledDotProductAttention(temperature=np.power(dim_key, 0.5), attn_dropout=dropout)
        self.layer_norm = nn.LayerNorm(dim_model)
        self.output_linear_a = nn.Linear(num_heads * dim_value, r, bias=False)
        self.output_linear_b = nn.Linear(r, dim_model)
        nn.init.xavier_normal_(self.output_linear_a.weight)
        nn.init.xavier_normal_(self.output_linear_b.weight)
        self.dropout = nn.Dropout(dropout)


class ScaledDotProductAttention(FusedOnly):
    """softmax(mask(QK^T / temperature)) V with dropout on the probabilities (common_layers.py:308-331)."""

    def __init__(self, temperature, attn_dropout=0.1):
        super().__init__()
        self.temperature = temperature
        self.dropout = nn.Dropout(attn_dropout)
        self.softmax = nn.Softmax(dim=2)
