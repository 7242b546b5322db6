"""Encoder container (modules/encoder.py:15-106): input Linear + LayerNorm + PE, then N x [self-attn, FFN]."""
import torch.nn as nn

from .common_layers import FactorizedMultiHeadAttention, FusedOnly, PositionalEncoding, PositionwiseFeedForward


def _no_factorized(flag):
    if flag:
        raise NotImplementedError("--is-factorized (low-rank FFN / input projection) is outside the B200 hot path; "
                                  "attention is always low-rank, exactly as in the reference")


class EncoderLayer(FusedOnly):
    def __init__(self, num_heads, dim_model, dim_inner, dim_key, dim_value, dropout=0.1, is_factorized=False, r=100):
        super().__init__()
        _no_factorized(is_factorized)
        self.is_factorized, self.r = is_factorized, r
        self.self_attn = FactorizedMultiHeadAttention(num_heads, dim_model, dim_key, dim_value, dropout=dropout, r=r)
        self.pos_ffn = PositionwiseFeedForward(dim_model, dim_inner, dropout=dropout)


class Encoder(FusedOnly):
    def __init__(self, num_layers, num_heads, dim_model, dim_key, dim_value, dim_input, dim_inner, dropout=0.1,
                 src_max_length=2500, is_factorized=False, r=100):
        super().__init__()
        _no_factorized(is_factorized)
        self.dim_input, self.num_layers, self.num_heads = dim_input, num_layers, num_heads
        self.dim_model, self.dim_key, self.dim_value, self.dim_inner = dim_model, dim_key, dim_value, dim_inner
        self.src_max_length, self.is_factorized, self.r = src_max_length, is_factorized, r
        self.dropout = nn.Dropout(dropout)
        self.dropout_rate = dropout
        self.input_linear = nn.Linear(dim_input, dim_model)
        self.layer_norm_input = nn.LayerNorm(dim_model)
        self.positional_encoding = PositionalEncoding(dim_model, src_max_length)
        self.layers = nn.ModuleList(
            EncoderLayer(num_heads, dim_model, dim_inner, dim_key, dim_value, dropout=dropout, is_factorized=False, r=r)
            for _ in range(num_layers))
