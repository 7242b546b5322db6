"""Shared helpers for the -m gpu tests (everything goes through the C ABI via mtl_b200)."""
import ctypes as C
import os

import numpy as np
import torch

import mtl_b200
from mtl_b200 import lib as L
from oracle import ref_asr

GEMM_MODE = int(os.environ.get("MTL_GEMM_MODE", "2"))   # default = the tcgen05 3xTF32 engine


def dev():
    return torch.device("cuda:0")


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def lib():
    return L.get_lib()


def ok(rc):
    L.check(rc)


def spec_of(cfg: ref_asr.ModelConfig) -> mtl_b200.ModelSpec:
    return mtl_b200.ModelSpec(n_enc=cfg.n_enc, n_dec=cfg.n_dec, d_model=cfg.d_model, n_heads=cfg.n_heads,
                              d_k=cfg.d_k, d_v=cfg.d_v, d_inner=cfg.d_inner, rank=cfg.rank, vocab=cfg.vocab,
                              n_freq=cfg.n_freq, src_max_len=cfg.src_max_len, tgt_max_len=cfg.tgt_max_len)


def to_batch(b):
    x, lens, y = b
    return mtl_b200.Batch.from_host(x, lens, y, dev())


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| (the per-tensor relative fp32 tolerance north_star names)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)
