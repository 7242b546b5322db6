"""Parameter inventory of the ASR model in ``model.parameters()`` order.

Mirrors the registration order of the reference modules (models/asr/transformer.py:24-59,
modules/encoder.py:40-51, modules/decoder.py:39-53, modules/common_layers.py:117-120,250-270):
encoder stem, encoder layers, decoder embedding, decoder layers, vocab projection, VGG convs."""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import List, Tuple


@dataclass(frozen=True)
class ModelSpec:
    n_enc: int = 2
    n_dec: int = 4
    d_model: int = 512
    n_heads: int = 8
    d_k: int = 64
    d_v: int = 64
    d_inner: int = 512
    rank: int = 100
    vocab: int = 3765
    n_freq: int = 161
    src_max_len: int = 5000
    tgt_max_len: int = 2500

    @property
    def d_input(self) -> int:           # utils/functions.py:318-321
        return 128 * ((self.n_freq // 2) // 2)

    def to_dict(self):
        return asdict(self)


def _attn(prefix: str, c: ModelSpec):
    d, r = c.d_model, c.rank
    hk, hv = c.n_heads * c.d_k, c.n_heads * c.d_v
    return [
        (f"{prefix}.query_linear_a.weight", (r, d)), (f"{prefix}.query_linear_b.weight", (hk, r)),
        (f"{prefix}.query_linear_b.bias", (hk,)),
        (f"{prefix}.key_linear_a.weight", (r, d)), (f"{prefix}.key_linear_b.weight", (hk, r)),
        (f"{prefix}.key_linear_b.bias", (hk,)),
        (f"{prefix}.value_linear_a.weight", (r, d)), (f"{prefix}.value_linear_b.weight", (hv, r)),
        (f"{prefix}.value_linear_b.bias", (hv,)),
        (f"{prefix}.layer_norm.weight", (d,)), (f"{prefix}.layer_norm.bias", (d,)),
        (f"{prefix}.output_linear_a.weight", (r, hv)), (f"{prefix}.output_linear_b.weight", (d, r)),
        (f"{prefix}.output_linear_b.bias", (d,)),
    ]


def _ffn(prefix: str, c: ModelSpec):
    d, f = c.d_model, c.d_inner
    return [
        (f"{prefix}.linear_1.weight", (f, d)), (f"{prefix}.linear_1.bias", (f,)),
        (f"{prefix}.linear_2.weight", (d, f)), (f"{prefix}.linear_2.bias", (d,)),
        (f"{prefix}.layer_norm.weight", (d,)), (f"{prefix}.layer_norm.bias", (d,)),
    ]


def param_specs(c: ModelSpec) -> List[Tuple[str, Tuple[int, ...]]]:
    d = c.d_model
    s = [("encoder.input_linear.weight", (d, c.d_input)), ("encoder.input_linear.bias", (d,)),
         ("encoder.layer_norm_input.weight", (d,)), ("encoder.layer_norm_input.bias", (d,))]
    for l in range(c.n_enc):
        s += _attn(f"encoder.layers.{l}.self_attn", c)
        s += _ffn(f"encoder.layers.{l}.pos_ffn", c)
    s.append(("decoder.trg_embedding.weight", (c.vocab, d)))
    for l in range(c.n_dec):
        s += _attn(f"decoder.layers.{l}.self_attn", c)
        s += _attn(f"decoder.layers.{l}.encoder_attn", c)
        s += _ffn(f"decoder.layers.{l}.pos_ffn", c)
    s.append(("decoder.output_linear.weight", (c.vocab, d)))
    s += [("conv.0.weight", (64, 1, 3, 3)), ("conv.0.bias", (64,)),
          ("conv.2.weight", (64, 64, 3, 3)), ("conv.2.bias", (64,)),
          ("conv.5.weight", (128, 64, 3, 3)), ("conv.5.bias", (128,)),
          ("conv.7.weight", (128, 128, 3, 3)), ("conv.7.bias", (128,))]
    return s
