"""Helpers shared by the API tests: a script-like ``args`` namespace, a synthetic label set / vocabulary and
synthetic WAV manifests (the reference's label JSON and corpora are data assets we do not ship)."""
import argparse
import os

import numpy as np
import scipy.io.wavfile as wavfile


def labels(n):
    """n distinct single-character labels (lower-case latin, then CJK code points)."""
    base = list(" abcdefghijklmnopqrstuvwxyz'")
    return base[:n] + [chr(0x4E00 + i) for i in range(max(0, n - len(base)))]


def make_vocab(n_labels):
    from utils.data import Vocab
    v = Vocab()
    for lab in labels(n_labels):
        v.add_token(lab)
        v.add_label(lab)
    return v


def script_args(cfg=None, **over):
    """The fields meta_transfer_train.py / joint_train.py put on ``args`` (their argparse defaults), with the
    model flags of ``cfg`` (an oracle ModelConfig) when given."""
    a = argparse.Namespace(
        model="TRFS", name="apitest", sample_rate=16000, k_train=2, k_valid=2, num_workers=0, label_smoothing=0.0,
        window_size=.02, window_stride=.01, window="hamming", epochs=2, cuda=True, early_stop="loss,10", save_every=1,
        save_folder="/tmp/mtl_api_save", emb_trg_sharing=False, feat_extractor="vgg_cnn", feat="spectrogram",
        verbose=False, continue_from="", augment=False, noise_dir=None, noise_prob=0.4, noise_min=0.0, noise_max=0.5,
        num_enc_layers=2, num_dec_layers=4, num_heads=8, dim_model=512, dim_key=64, dim_value=64, dim_input=161,
        dim_inner=512, dim_emb=512, src_max_len=5000, tgt_max_len=2500, lr=1e-4, meta_lr=1e-4, evaluate_every=1000,
        loss="ce", clip=False, max_norm=400, is_factorized=False, r=100, dropout=0.1, input_type="char",
        copy_grad=True, cpu_state_dict=False, train_partition_list=None)
    if cfg is not None:
        a.sample_rate = (cfg.n_freq - 1) * 100
        a.num_enc_layers, a.num_dec_layers, a.num_heads = cfg.n_enc, cfg.n_dec, cfg.n_heads
        a.dim_model = a.dim_emb = cfg.d_model
        a.dim_key, a.dim_value, a.dim_inner, a.r = cfg.d_k, cfg.d_v, cfg.d_inner, cfg.rank
        a.src_max_len, a.tgt_max_len, a.dropout = cfg.src_max_len, cfg.tgt_max_len, cfg.dropout
    for k, v in over.items():
        setattr(a, k, v)
    return a


def write_manifest(folder, name, n_utts, sr=16000, seed=0, text_labels=None, min_s=0.4, max_s=1.0, txt_files=False):
    """n_utts synthetic int16 WAVs + a ``wav_path,transcript`` CSV (transcript inline or as .txt files)."""
    rng = np.random.default_rng(seed)
    text_labels = text_labels or list("abcdefghij ")
    rows = []
    os.makedirs(folder, exist_ok=True)
    for i in range(n_utts):
        n = int(sr * rng.uniform(min_s, max_s))
        wav = (rng.standard_normal(n) * 3000).astype(np.int16)
        path = os.path.join(folder, f"{name}_{i}.wav")
        wavfile.write(path, sr, wav)
        text = "".join(rng.choice(text_labels, size=int(rng.integers(3, 12))))
        text = text.strip() or "a"
        if txt_files:
            tpath = os.path.join(folder, f"{name}_{i}.txt")
            with open(tpath, "w", encoding="utf8") as f:
                f.write(text + "\n")
            rows.append(f"{path},{tpath}")
        else:
            rows.append(f"{path},{text}")
    manifest = os.path.join(folder, f"{name}.csv")
    with open(manifest, "w", encoding="utf8") as f:
        f.write("\n".join(rows) + "\n")
    return manifest
