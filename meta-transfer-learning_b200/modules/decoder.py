"""Decoder container (modules/decoder.py:14-115,293-323): embedding + PE, N x [self-attn, cross-attn, FFN],
bias-free vocabulary projection.  ``preprocess`` is kept as a host-side utility (the engine does the same on
the device, csrc/norm_embed.cu: dec_preprocess_kernel)."""
import torch
import torch.nn as nn

from .common_layers import FactorizedMultiHeadAttention, FusedOnly, PositionalEncoding, PositionwiseFeedForward, pad_list
from .encoder import _no_factorized


class DecoderLayer(FusedOnly):
    def __init__(self, dim_model, dim_inner, num_heads, dim_key, dim_value, dropout=0.1, is_factorized=False, r=100):
        super().__init__()
        _no_factorized(is_factorized)
        self.is_factorized, self.r = is_factorized, r
        self.self_attn = FactorizedMultiHeadAttention(num_heads, dim_model, dim_key, dim_value, dropout=dropout, r=r)
        self.encoder_attn = FactorizedMultiHeadAttention(num_heads, dim_model, dim_key, dim_value, dropout=dropout, r=r)
        self.pos_ffn = PositionwiseFeedForward(dim_model, dim_inner, dropout=dropout)


class Decoder(FusedOnly):
    def __init__(self, vocab, num_layers, num_heads, dim_emb, dim_model, dim_inner, dim_key, dim_value, dropout=0.1,
                 trg_max_length=1000, emb_trg_sharing=False, is_factorized=False, r=100):
        super().__init__()
        _no_factorized(is_factorized)
        if dim_emb != dim_model:
            raise ValueError("dim_emb must equal dim_model (the reference adds the embedding to a dim_model-wide PE)")
        self.vocab, self.num_layers, self.num_heads = vocab, num_layers, num_heads
        self.dim_emb, self.dim_model, self.dim_inner = dim_emb, dim_model, dim_inner
        self.dim_key, self.dim_value = dim_key, dim_value
        self.dropout_rate, self.emb_trg_sharing = dropout, emb_trg_sharing   # sharing flag is stored, never acted on (decoder.py:32)
        self.trg_max_length, self.is_factorized, self.r = trg_max_length, is_factorized, r
        n_labels = len(vocab.label2id)
        self.trg_embedding = nn.Embedding(n_labels, dim_emb, padding_idx=vocab.PAD_ID)
        self.positional_encoding = PositionalEncoding(dim_model, max_length=trg_max_length)
        self.dropout = nn.Dropout(dropout)
        self.layers = nn.ModuleList(
            DecoderLayer(dim_model, dim_inner, num_heads, dim_key, dim_value, dropout=dropout, is_factorized=False, r=r)
            for _ in range(num_layers))
        self.output_linear = nn.Linear(dim_model, n_labels, bias=False)
        nn.init.xavier_normal_(self.output_linear.weight)
        self.x_logit_scale = 1.0

    def preprocess(self, padded_input):
        """(B, L) PAD-padded targets -> decoder input <SOS> y (padded with EOS) and gold y <EOS> (padded with
        PAD), both (B, max len + 1)  (decoder.py:55-69)."""
        v = self.vocab
        seqs = [row[row != v.PAD_ID] for row in padded_input]
        sos, eos = seqs[0].new_tensor([v.SOS_ID]), seqs[0].new_tensor([v.EOS_ID])
        seq_in = pad_list([torch.cat([sos, y]) for y in seqs], v.EOS_ID)
        seq_out = pad_list([torch.cat([y, eos]) for y in seqs], v.PAD_ID)
        return seq_in, seq_out

    def greedy_search(self, encoder_padded_outputs, args=None, beam_width=2, lm_rescoring=False, lm=None, lm_weight=0.1,
                      c_weight=1, start_token=-1, max_steps=300):
        """Greedy 1-best decoding (decoder.py:131-184): start token, `max_steps` (the reference hard-codes 300) arg-max
        steps over the whole prefix, each row cut at its first EOS.  -> list of B strings.  One engine call
        (mtl_asr_greedy) without host syncs instead of 300 Python-driven decoder passes."""
        if lm is not None or lm_rescoring:
            raise NotImplementedError("LM rescoring is outside the B200 hot path")
        model = self._owner() if getattr(self, "_owner", None) is not None else None
        if model is None:
            raise RuntimeError("the decoder is not attached to a Transformer on a CUDA device")
        ids = model.session.greedy(model._theta, encoder_padded_outputs, int(start_token), int(max_steps)).cpu().tolist()
        v = self.vocab
        out = []
        for row in ids:
            st = ""
            for t in row:
                if t == v.EOS_ID:
                    break
                st += v.id2label[t]
            out.append(st)
        return out

    def beam_search(self, *a, **k):
        raise NotImplementedError("beam search (decoder.py:186-291) is outside the B200 hot path")
