"""Functional CPU restatement of the reference ASR model (TEST INFRASTRUCTURE).

Restates, as plain functions over a ``{name: tensor}`` dict, what the reference
builds out of ``nn.Module`` objects.  All arithmetic is delegated to the same
third-party library the reference uses (PyTorch CPU ops: ``F.conv2d``,
``F.linear``, ``F.layer_norm``, ``torch.bmm``, ``F.cross_entropy``), so the only
thing restated here is the *wiring* — and that wiring is pinned against the
live reference in ``tests/test_oracle_vs_reference.py``.

Reference citations (relative to /root/reference):
  models/asr/transformer.py:47-59,120-149   VGG front-end, flatten, enc/dec, topk
  modules/encoder.py:53-80,98-106           encoder stem + layers
  modules/decoder.py:55-115,311-323         preprocess, masks, decoder layers, vocab proj
  modules/common_layers.py:38-108           masks, positional encoding
  modules/common_layers.py:122-132          PositionwiseFeedForward
  modules/common_layers.py:276-331          FactorizedMultiHeadAttention + SDPA
  utils/metrics.py:96-126                   CE / label smoothing loss
  utils/functions.py:307-351                model factory (dim_input derivation)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict

import numpy as np
import torch
import torch.nn.functional as F

PAD_ID, SOS_ID, EOS_ID, OOV_ID = 0, 1, 2, 3  # utils/data.py:8


@dataclass(frozen=True)
class ModelConfig:
    """Hyper-parameters that shape the parameter set (utils/functions.py:307-351)."""
    n_enc: int = 2
    n_dec: int = 4
    d_model: int = 512
    n_heads: int = 8
    d_k: int = 64
    d_v: int = 64
    d_inner: int = 512
    rank: int = 100           # --r, always-on low-rank attention (modules/encoder.py:92)
    vocab: int = 3765         # 4 specials + 3761 labels
    n_freq: int = 161         # sample_rate*window_size/2 + 1
    src_max_len: int = 5000
    tgt_max_len: int = 2500
    dropout: float = 0.0

    @property
    def d_input(self) -> int:  # utils/functions.py:318-321
        return 128 * ((self.n_freq // 2) // 2)

    def to_dict(self):
        return asdict(self)


SMALL = ModelConfig(n_enc=1, n_dec=2, d_model=64, n_heads=2, d_k=32, d_v=32, d_inner=96,
                    rank=20, vocab=53, n_freq=21, src_max_len=200, tgt_max_len=100)
CFG2 = ModelConfig()


# --------------------------------------------------------------------------- parameters

def _attn_specs(prefix: str, c: ModelConfig):
    """Registration order inside FactorizedMultiHeadAttention (common_layers.py:250-270)."""
    d, r = c.d_model, c.rank
    hk, hv = c.n_heads * c.d_k, c.n_heads * c.d_v
    return [
        (f"{prefix}.query_linear_a.weight", (r, d)),
        (f"{prefix}.query_linear_b.weight", (hk, r)),
        (f"{prefix}.query_linear_b.bias", (hk,)),
        (f"{prefix}.key_linear_a.weight", (r, d)),
        (f"{prefix}.key_linear_b.weight", (hk, r)),
        (f"{prefix}.key_linear_b.bias", (hk,)),
        (f"{prefix}.value_linear_a.weight", (r, d)),
        (f"{prefix}.value_linear_b.weight", (hv, r)),
        (f"{prefix}.value_linear_b.bias", (hv,)),
        (f"{prefix}.layer_norm.weight", (d,)),
        (f"{prefix}.layer_norm.bias", (d,)),
        (f"{prefix}.output_linear_a.weight", (r, hv)),
        (f"{prefix}.output_linear_b.weight", (d, r)),
        (f"{prefix}.output_linear_b.bias", (d,)),
    ]


def _ffn_specs(prefix: str, c: ModelConfig):
    d, f = c.d_model, c.d_inner
    return [
        (f"{prefix}.linear_1.weight", (f, d)),
        (f"{prefix}.linear_1.bias", (f,)),
        (f"{prefix}.linear_2.weight", (d, f)),
        (f"{prefix}.linear_2.bias", (d,)),
        (f"{prefix}.layer_norm.weight", (d,)),
        (f"{prefix}.layer_norm.bias", (d,)),
    ]


def param_specs(c: ModelConfig):
    """(name, shape) in ``model.parameters()`` order: encoder, decoder, conv
    (Transformer.__init__ registers encoder/decoder before self.conv,
    models/asr/transformer.py:24-59)."""
    d = c.d_model
    s = [
        ("encoder.input_linear.weight", (d, c.d_input)),
        ("encoder.input_linear.bias", (d,)),
        ("encoder.layer_norm_input.weight", (d,)),
        ("encoder.layer_norm_input.bias", (d,)),
    ]
    for l in range(c.n_enc):
        s += _attn_specs(f"encoder.layers.{l}.self_attn", c)
        s += _ffn_specs(f"encoder.layers.{l}.pos_ffn", c)
    s.append(("decoder.trg_embedding.weight", (c.vocab, d)))
    for l in range(c.n_dec):
        s += _attn_specs(f"decoder.layers.{l}.self_attn", c)
        s += _attn_specs(f"decoder.layers.{l}.encoder_attn", c)
        s += _ffn_specs(f"decoder.layers.{l}.pos_ffn", c)
    s.append(("decoder.output_linear.weight", (c.vocab, d)))
    s += [
        ("conv.0.weight", (64, 1, 3, 3)), ("conv.0.bias", (64,)),
        ("conv.2.weight", (64, 64, 3, 3)), ("conv.2.bias", (64,)),
        ("conv.5.weight", (128, 64, 3, 3)), ("conv.5.bias", (128,)),
        ("conv.7.weight", (128, 128, 3, 3)), ("conv.7.bias", (128,)),
    ]
    return s


def num_params(c: ModelConfig) -> int:
    return sum(int(np.prod(sh)) for _, sh in param_specs(c))


def init_params(c: ModelConfig, seed: int = 0):
    """Deterministic (numpy PCG64) initialisation with the reference's *distributions*:
    xavier_uniform on every >=2-D tensor (transformer.py:74-76), LayerNorm (1, 0),
    default nn.Linear / nn.Conv2d bias U(-1/sqrt(fan_in), 1/sqrt(fan_in)).
    Values differ from torch's RNG stream on purpose: parity tests load the same
    dict into both implementations."""
    rng = np.random.default_rng(seed)
    out = {}
    fan_in_of = {}
    for name, shape in param_specs(c):
        if len(shape) >= 2:
            rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
            fan_in, fan_out = shape[1] * rf, shape[0] * rf
            a = math.sqrt(6.0 / (fan_in + fan_out))
            w = rng.uniform(-a, a, size=shape).astype(np.float32)
            fan_in_of[name.rsplit(".", 1)[0]] = fan_in
        elif "layer_norm" in name:
            w = np.ones(shape, np.float32) if name.endswith("weight") else np.zeros(shape, np.float32)
        else:
            b = 1.0 / math.sqrt(fan_in_of[name.rsplit(".", 1)[0]])
            w = rng.uniform(-b, b, size=shape).astype(np.float32)
        out[name] = torch.from_numpy(w)
    return out


def positional_table(max_len: int, d: int) -> torch.Tensor:
    """common_layers.py:93-99 (buffer ``pe`` of shape (1, max_len, d))."""
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len).unsqueeze(1).float()
    div = torch.exp(torch.arange(0, d, 2).float() * -(math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


def buffers(c: ModelConfig):
    return {
        "encoder.positional_encoding.pe": positional_table(c.src_max_len, c.d_model),
        "decoder.positional_encoding.pe": positional_table(c.tgt_max_len, c.d_model),
    }


# --------------------------------------------------------------------------- forward pieces

def vgg_frontend(p, x):
    """transformer.py:47-59 then :136-138.  x (B,1,F,T) -> (B,T',128*F'')."""
    h = F.relu(F.conv2d(x, p["conv.0.weight"], p["conv.0.bias"], padding=1))
    h = F.relu(F.conv2d(h, p["conv.2.weight"], p["conv.2.bias"], padding=1))
    h = F.max_pool2d(h, 2, stride=2)
    h = F.relu(F.conv2d(h, p["conv.5.weight"], p["conv.5.bias"], padding=1))
    h = F.relu(F.conv2d(h, p["conv.7.weight"], p["conv.7.bias"], padding=1))
    h = F.max_pool2d(h, 2, stride=2)
    b, ch, fr, t = h.shape
    return h.reshape(b, ch * fr, t).transpose(1, 2).contiguous()


def length_row_mask(n_rows: int, lengths, like: torch.Tensor) -> torch.Tensor:
    """get_non_pad_mask with input_lengths (common_layers.py:43-48): rows >= len are 0.
    NB lengths are the *raw* (pre-CNN) frame counts (transformer.py:140)."""
    m = like.new_ones(len(lengths), n_rows)
    for i, ln in enumerate(lengths):
        m[i, int(ln):] = 0
    return m


def dropout(x, pdrop, train):
    return F.dropout(x, pdrop, training=train) if (train and pdrop > 0) else x


def lowrank_attention(p, pre, c: ModelConfig, xq, xkv, mask, pdrop=0.0, train=False):
    """FactorizedMultiHeadAttention.forward + ScaledDotProductAttention.forward
    (common_layers.py:276-306, 317-331).  mask: (B,Tq,Tk) bool, True = masked."""
    b, tq, _ = xq.shape
    tk = xkv.shape[1]
    h, dk, dv = c.n_heads, c.d_k, c.d_v
    q = F.linear(F.linear(xq, p[f"{pre}.query_linear_a.weight"]),
                 p[f"{pre}.query_linear_b.weight"], p[f"{pre}.query_linear_b.bias"])
    k = F.linear(F.linear(xkv, p[f"{pre}.key_linear_a.weight"]),
                 p[f"{pre}.key_linear_b.weight"], p[f"{pre}.key_linear_b.bias"])
    v = F.linear(F.linear(xkv, p[f"{pre}.value_linear_a.weight"]),
                 p[f"{pre}.value_linear_b.weight"], p[f"{pre}.value_linear_b.bias"])
    q = q.view(b, tq, h, dk).permute(2, 0, 1, 3).reshape(h * b, tq, dk)
    k = k.view(b, tk, h, dk).permute(2, 0, 1, 3).reshape(h * b, tk, dk)
    v = v.view(b, tk, h, dv).permute(2, 0, 1, 3).reshape(h * b, tk, dv)
    att = torch.bmm(q, k.transpose(1, 2)) / (dk ** 0.5)
    att = att.masked_fill(mask.repeat(h, 1, 1), float("-inf"))
    att = dropout(torch.softmax(att, dim=2), pdrop, train)
    o = torch.bmm(att, v).view(h, b, tq, dv).permute(1, 2, 0, 3).reshape(b, tq, h * dv)
    o = F.linear(F.linear(o, p[f"{pre}.output_linear_a.weight"]),
                 p[f"{pre}.output_linear_b.weight"], p[f"{pre}.output_linear_b.bias"])
    o = dropout(o, pdrop, train)
    return F.layer_norm(o + xq, (c.d_model,), p[f"{pre}.layer_norm.weight"], p[f"{pre}.layer_norm.bias"])


def ffn(p, pre, c: ModelConfig, x, pdrop=0.0, train=False):
    """PositionwiseFeedForward.forward (common_layers.py:122-132)."""
    y = F.linear(F.relu(F.linear(x, p[f"{pre}.linear_1.weight"], p[f"{pre}.linear_1.bias"])),
                 p[f"{pre}.linear_2.weight"], p[f"{pre}.linear_2.bias"])
    y = dropout(y, pdrop, train)
    return F.layer_norm(y + x, (c.d_model,), p[f"{pre}.layer_norm.weight"], p[f"{pre}.layer_norm.bias"])


def encoder_forward(p, c: ModelConfig, feat, lengths, pe, train=False):
    """Encoder.forward (encoder.py:53-80): masks from raw lengths, LN(Linear)+PE, layers."""
    b, t, _ = feat.shape
    rows = length_row_mask(t, lengths, feat)                      # (B,T')
    amask = (rows < 1).unsqueeze(1).expand(-1, t, -1)             # key padding, (B,T',T')
    x = F.layer_norm(F.linear(feat, p["encoder.input_linear.weight"], p["encoder.input_linear.bias"]),
                     (c.d_model,), p["encoder.layer_norm_input.weight"], p["encoder.layer_norm_input.bias"])
    x = x + pe[:, :t]
    rm = rows.unsqueeze(-1)
    for l in range(c.n_enc):
        x = lowrank_attention(p, f"encoder.layers.{l}.self_attn", c, x, x, amask, c.dropout, train) * rm
        x = ffn(p, f"encoder.layers.{l}.pos_ffn", c, x, c.dropout, train) * rm
    return x


def decoder_preprocess(trg):
    """Decoder.preprocess (decoder.py:55-69): strip PAD, <SOS>+y padded with EOS (input),
    y+<EOS> padded with PAD (gold)."""
    seqs = [y[y != PAD_ID] for y in trg]
    n = max(len(s) for s in seqs) + 1
    seq_in = trg.new_full((len(seqs), n), EOS_ID)
    seq_out = trg.new_full((len(seqs), n), PAD_ID)
    for i, s in enumerate(seqs):
        seq_in[i, 0] = SOS_ID
        seq_in[i, 1:1 + len(s)] = s
        seq_out[i, :len(s)] = s
        seq_out[i, len(s)] = EOS_ID
    return seq_in, seq_out


def decoder_forward(p, c: ModelConfig, trg, enc_out, lengths, pe, train=False):
    """Decoder.forward (decoder.py:71-115) + DecoderLayer.forward (:311-323)."""
    seq_in, seq_out = decoder_preprocess(trg)
    b, n = seq_in.shape
    rm = (seq_in != EOS_ID).float().unsqueeze(-1)                                 # decoder.py:86
    causal = torch.triu(torch.ones(n, n, dtype=torch.bool), diagonal=1).unsqueeze(0)
    smask = (seq_in == EOS_ID).unsqueeze(1).expand(-1, n, -1) | causal            # :87-90
    t = enc_out.shape[1]
    cmask = (length_row_mask(t, lengths, enc_out) < 1).unsqueeze(1).expand(-1, n, -1)  # :93-94
    x = dropout(F.embedding(seq_in, p["decoder.trg_embedding.weight"]) * 1.0 + pe[:, :n], c.dropout, train)
    for l in range(c.n_dec):
        pre = f"decoder.layers.{l}"
        x = lowrank_attention(p, f"{pre}.self_attn", c, x, x, smask, c.dropout, train) * rm
        x = lowrank_attention(p, f"{pre}.encoder_attn", c, x, enc_out, cmask, c.dropout, train) * rm
        x = ffn(p, f"{pre}.pos_ffn", c, x, c.dropout, train) * rm
    pred = F.linear(x, p["decoder.output_linear.weight"])                          # :108-112
    return pred, seq_out


def forward(p, c: ModelConfig, x, lengths, trg, bufs=None, train=False):
    """Transformer.forward (transformer.py:120-149) -> (pred, gold, hyp)."""
    bufs = bufs or buffers(c)
    feat = vgg_frontend(p, x)
    enc = encoder_forward(p, c, feat, lengths, bufs["encoder.positional_encoding.pe"], train)
    pred, gold = decoder_forward(p, c, trg, enc, lengths, bufs["decoder.positional_encoding.pe"], train)
    hyp = torch.topk(pred, 1, dim=2)[1].squeeze(2)
    return pred, gold, hyp


def ce_loss(pred, gold, smoothing: float = 0.0):
    """calculate_loss, loss_type == 'ce' (utils/metrics.py:96-126)."""
    v = pred.size(2)
    pr, gd = pred.reshape(-1, v), gold.reshape(-1)
    if smoothing > 0.0:
        npm = gd.ne(PAD_ID)
        one_hot = torch.zeros_like(pr).scatter(1, (npm.long() * gd).view(-1, 1), 1)
        one_hot = one_hot * (1 - smoothing) + (1 - one_hot) * smoothing / v
        lp = F.log_softmax(pr, dim=1)
        return -(one_hot * lp).sum(dim=1).masked_select(npm).sum() / npm.sum().item()
    return F.cross_entropy(pr, gd, ignore_index=PAD_ID, reduction="mean")


def num_correct(pred, gold):
    """calculate_metrics token accuracy (utils/metrics.py:83-89)."""
    v = pred.size(2)
    am = pred.reshape(-1, v).max(1)[1]
    gd = gold.reshape(-1)
    return int(am.eq(gd).masked_select(gd.ne(PAD_ID)).sum())


def greedy_search(p, c: ModelConfig, enc_out, start_token: int = SOS_ID, max_steps: int = 300, bufs=None):
    """Decoder.greedy_search (modules/decoder.py:131-184): ys = [start]; `max_steps` times the whole decoder over ys with an
    all-ones non-pad mask, the subsequent mask only, NO encoder-side mask (dec_enc_attn_mask=None) and eval-mode
    dropout; arg-max of the last position is appended.  Returns the (B, max_steps) token matrix (the reference then
    cuts each row at the first EOS, :175-183)."""
    bufs = bufs or buffers(c)
    pe = bufs["decoder.positional_encoding.pe"]
    b = enc_out.shape[0]
    ys = torch.full((b, 1), start_token, dtype=torch.long)
    out = []
    for _ in range(max_steps):
        n = ys.shape[1]
        causal = torch.triu(torch.ones(n, n, dtype=torch.bool), diagonal=1).unsqueeze(0).expand(b, -1, -1)
        nomask = torch.zeros(b, n, enc_out.shape[1], dtype=torch.bool)
        x = F.embedding(ys, p["decoder.trg_embedding.weight"]) * 1.0 + pe[:, :n]
        for l in range(c.n_dec):
            pre = f"decoder.layers.{l}"
            x = lowrank_attention(p, f"{pre}.self_attn", c, x, x, causal)
            x = lowrank_attention(p, f"{pre}.encoder_attn", c, x, enc_out, nomask)
            x = ffn(p, f"{pre}.pos_ffn", c, x)
        prob = F.linear(x, p["decoder.output_linear.weight"])
        nxt = prob[:, -1].max(dim=1)[1]
        out.append(nxt)
        ys = torch.cat([ys, nxt.unsqueeze(-1)], dim=1)
    return torch.stack(out, dim=1)


def encode(p, c: ModelConfig, x, lengths, bufs=None):
    """Transformer.encode (models/asr/transformer.py:78-98)."""
    bufs = bufs or buffers(c)
    return encoder_forward(p, c, vgg_frontend(p, x), lengths, bufs["encoder.positional_encoding.pe"])
