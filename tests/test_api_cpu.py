"""CPU tests of the reference-compatible Python surface: imports, constructor / state_dict
parity with the live reference, the input pipeline contracts, and the unchanged reference script running
against our packages up to the point where it needs the GPU."""
import contextlib
import io
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import api_util
from oracle import live_reference as lr
from oracle import ref_asr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_import_surface_of_the_reference_scripts():
    """Every name meta_transfer_train.py:13-18 and joint_train.py:13-18 import."""
    from torchsummary import summary  # noqa: F401
    from trainer.asr.transient_trainer import TransientTrainer
    from trainer.asr.joint_trainer import JointTrainer
    from utils.data import Vocab
    from utils.data_loader import SpectrogramDataset, LogFBankDataset, AudioDataLoader, BucketingSampler  # noqa: F401
    from utils.functions import (load_meta_model, init_transformer_model, init_optimizer, compute_num_params,  # noqa: F401
                                 generate_labels, load_joint_model, load_discriminator, init_discriminator_model)
    import inspect
    sig = inspect.signature(TransientTrainer.train)
    assert list(sig.parameters)[1:] == ["model", "vocab", "train_data_list", "valid_loader_list", "loss_type", "start_it",
                                        "num_it", "args", "inner_opt", "outer_opt", "evaluate_every", "window_size",
                                        "last_summary_every", "last_metrics", "early_stop", "cpu_state_dict",
                                        "is_copy_grad"]
    sig = inspect.signature(JointTrainer.train)
    assert list(sig.parameters)[1:] == ["model", "vocab", "train_data_list", "valid_loader_list", "loss_type", "start_it",
                                        "num_it", "args", "evaluate_every", "window_size", "last_summary_every",
                                        "last_metrics", "early_stop", "cpu_state_dict", "is_copy_grad", "opt_name",
                                        "discriminator"]
    v = Vocab()
    assert (v.PAD_ID, v.SOS_ID, v.EOS_ID, v.OOV_ID) == (0, 1, 2, 3)
    assert v.id2label == ["<PAD>", "<SOS>", "<EOS>", "<OOV>"] and v.label2id["<EOS>"] == 2


def test_text_helpers():
    from utils.functions import generate_labels, post_process
    from utils.metrics import calculate_cer, calculate_wer
    l2i, i2l = generate_labels(["a", "b", "a"], ["<PAD>", "<SOS>"])
    assert l2i == {"<PAD>": 0, "<SOS>": 1, "a": 2, "b": 3} and i2l[3] == "b"
    assert post_process("<SOS>he▁llo<EOS><PAD>", ["<PAD>", "<SOS>", "<EOS>", "<OOV>"]) == "he llo"
    assert calculate_cer("kitten", "sitting") == 3 and calculate_cer("", "abc") == 3
    assert calculate_wer("the cat sat", "the cat sat down") == 1


def _our_model(cfg, seed):
    from utils.functions import init_transformer_model
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        args = api_util.script_args(cfg)
        model = init_transformer_model(args, api_util.make_vocab(cfg.vocab - 4), is_factorized=False, r=cfg.rank)
    return model, args


@pytest.mark.skipif(not lr.available(), reason="live reference not mounted")
def test_init_transformer_model_matches_live_reference_bit_for_bit():
    """Same torch seed -> the reference's parameter names, order, shapes, buffers AND initial values
    (construction consumes the RNG in the same sequence; utils/functions.py:307-351)."""
    cfg = ref_asr.SMALL
    ours, args = _our_model(cfg, 123456)
    assert args.dim_input == cfg.d_input
    with lr.reference_imports():
        from utils.data import Vocab
        from utils.functions import init_transformer_model, compute_num_params
        vocab = Vocab()
        for lab in api_util.labels(cfg.vocab - 4):
            vocab.add_token(lab)
            vocab.add_label(lab)
        torch.manual_seed(123456)
        with contextlib.redirect_stdout(io.StringIO()):
            ref = init_transformer_model(lr.make_args(cfg), vocab, is_factorized=False, r=cfg.rank)
        ref_counts = compute_num_params(ref)
    from utils.functions import compute_num_params as ours_counts
    sd_o, sd_r = ours.state_dict(), ref.state_dict()
    assert list(sd_o.keys()) == list(sd_r.keys())
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
    for k in sd_r:
        assert sd_o[k].shape == sd_r[k].shape and torch.equal(sd_o[k], sd_r[k]), k
    assert tuple(ours_counts(ours)) == tuple(ref_counts)
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref_asr.param_specs(cfg)]


def test_model_refuses_to_run_without_cuda():
    import mtl_b200
    cfg = ref_asr.SMALL
    model, _ = _our_model(cfg, 0)
    x = torch.zeros(2, 1, cfg.n_freq, 21)
    with pytest.raises(mtl_b200.MtlError):
        model(x, torch.tensor([21, 21], dtype=torch.int32), torch.ones(2, 3, dtype=torch.long))
    if not torch.cuda.is_available():
        with pytest.raises(mtl_b200.MtlError):
            model.cuda()
    with pytest.raises(RuntimeError):
        model.encoder(torch.zeros(1, 5, cfg.d_input), [5])        # layers only exist fused


def test_stft_matches_torch_stft():
    from utils.data_loader import stft_magnitude
    import scipy.signal.windows
    rng = np.random.default_rng(0)
    y = rng.standard_normal(16000).astype(np.float32)
    win = scipy.signal.windows.hamming(320)
    mag = stft_magnitude(y, 320, 160, win)
    ref = torch.stft(torch.from_numpy(y), 320, hop_length=160, win_length=320,
                     window=torch.from_numpy(win).float(), center=True, pad_mode="reflect", return_complex=True).abs()
    assert mag.shape == (161, 101) == tuple(ref.shape)
    assert np.allclose(mag, ref.numpy(), rtol=1e-4, atol=1e-3)


def test_dataset_sample_and_loader_contracts(tmp_path):
    """SpectrogramDataset.sample / AudioDataLoader batch layouts (utils/data_loader.py:245-321,401-440)."""
    from utils.data_loader import AudioDataLoader, BucketingSampler, SpectrogramDataset
    vocab = api_util.make_vocab(40)
    args = api_util.script_args(src_max_len=90)
    audio_conf = dict(sample_rate=16000, window_size=.02, window_stride=.01, window="hamming", noise_dir=None,
                      noise_prob=0.4, noise_levels=(0.0, 0.5))
    m0 = api_util.write_manifest(str(tmp_path), "m0", 5, seed=1)
    m1 = api_util.write_manifest(str(tmp_path), "m1", 4, seed=2, txt_files=True)
    ds = SpectrogramDataset(vocab, args, audio_conf, manifest_filepath_list=[m0, m1], normalize=True, is_train=True)
    np.random.seed(3)
    (x, sizes, pct, y, ysz), (vx, vsizes, vpct, vy, vysz) = ds.sample(3, 2, 1)
    assert x.shape[:3] == (3, 1, 161) and x.dtype == torch.float32 and vx.shape[0] == 2
    assert sizes.dtype == torch.int32 and ysz.dtype == torch.int32 and y.dtype == torch.int64 and pct.dtype == torch.float32
    assert int(sizes.max()) == x.shape[3] <= 90                      # truncated to src_max_len frames
    assert torch.allclose(pct, sizes.float() / x.shape[3])
    for i in range(3):
        assert float(x[i, 0, :, int(sizes[i]):].abs().max() if sizes[i] < x.shape[3] else 0.0) == 0.0   # zero padded
        assert int((y[i] != 0).sum()) == int(ysz[i])
        n = int(sizes[i])
        assert abs(float(x[i, 0, :, :n].mean())) < 0.15               # utterance-normalised (before truncation)
    assert y.min() >= 0 and y.max() < len(vocab.label2id) and not ((y > 0) & (y < 4)).any()
    # the reference prepends " " to .txt transcripts and lower-cases
    assert ds.parse_transcript("AB c") == [vocab.label2id[c] for c in "ab c"]
    valid = SpectrogramDataset(vocab, args, audio_conf, manifest_filepath_list=[m0], normalize=True)
    assert len(BucketingSampler(valid, batch_size=2)) == 3
    loader = AudioDataLoader(pad_token_id=vocab.PAD_ID, dataset=valid, num_workers=0, batch_size=3)
    inputs, targets, pcts, in_sizes, tgt_sizes = next(iter(loader))
    assert inputs.shape[:3] == (3, 1, 161) and targets.dtype == torch.int64 and in_sizes.dtype == torch.int32
    assert list(in_sizes) == sorted(in_sizes, reverse=True)            # collate sorts by length, longest first


@pytest.mark.skipif(not lr.available(), reason="live reference not mounted")
@pytest.mark.parametrize("script", ["meta_transfer_train.py", "joint_train.py"])
def test_unchanged_reference_script_runs_against_our_packages_until_it_needs_the_gpu(tmp_path, script):
    """The reference CLI script, byte for byte, with our packages on sys.path: argparse, Vocab, datasets, loaders
    and init_transformer_model all run; without a GPU the first engine call (model.cuda()) refuses loudly."""
    if torch.cuda.is_available():
        pytest.skip("covered by the -m gpu end-to-end test")
    labs = api_util.labels(60)
    labels_path = tmp_path / "labels.json"
    import json
    labels_path.write_text(json.dumps(labs), encoding="utf8")
    tr = api_util.write_manifest(str(tmp_path), "train0", 4, seed=1)
    va = api_util.write_manifest(str(tmp_path), "valid0", 2, seed=2)
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"),
           os.path.join(lr.REFERENCE_ROOT, script), "--train-manifest-list", tr, "--valid-manifest-list", va,
           "--test-manifest-list", va, "--labels-path", str(labels_path), "--sample-rate", "16000", "--k-train", "2",
           "--num-workers", "1", "--num-enc-layers", "1", "--num-dec-layers", "1", "--num-heads", "2",
           "--dim-model", "64", "--dim-key", "32", "--dim-value", "32", "--dim-inner", "64", "--dim-emb", "64",
           "--r", "12", "--cuda", "--copy-grad", "--epochs", "1", "--name", "cpu_probe", "--save-folder", str(tmp_path)]
    if script == "meta_transfer_train.py":
        cmd += ["--k-valid", "2"]
    p = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    out = p.stdout + p.stderr
    assert "TRAINING FROM SCRATCH" in out and "feat extractor: vgg_cnn" in out, out[-2000:]
    assert p.returncode != 0 and "MtlError" in out and "CUDA device" in out, out[-2000:]


def test_checkpoint_written_by_the_reference_unpickles_and_loads_into_our_model():
    """tests/golden/ref_checkpoint_small.th was written by the REFERENCE's save_meta_model (oracle/make_golden.py:
    ref_checkpoint).  With only our packages importable it must unpickle (utils.data.Vocab, argparse.Namespace,
    torch.optim objects) and its model_state_dict must load key for key (utils/functions.py:158-188)."""
    import oracle.make_golden as mg
    path = os.path.join(ROOT, "tests", "golden", "ref_checkpoint_small.th")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"vocab", "args", "epoch", "model_state_dict", "inner_opt", "outer_opt", "metrics"}
    assert ck["epoch"] == mg.REF_CKPT["epoch"] and ck["metrics"]["avg_valid_loss"] == 1.25
    from utils.data import Vocab
    assert type(ck["vocab"]) is Vocab and ck["vocab"].id2label[:4] == ["<PAD>", "<SOS>", "<EOS>", "<OOV>"]
    assert isinstance(ck["outer_opt"], torch.optim.Adam) and isinstance(ck["inner_opt"], torch.optim.SGD)
    sd_opt = ck["outer_opt"].state_dict()
    assert len(sd_opt["state"]) == len(ref_asr.param_specs(ref_asr.SMALL)) and int(sd_opt["state"][0]["step"]) == 2
    from utils.functions import init_transformer_model
    args = ck["args"]
    with contextlib.redirect_stdout(io.StringIO()):
        model = init_transformer_model(args, ck["vocab"], is_factorized=args.is_factorized, r=args.r)
    res = model.load_state_dict(ck["model_state_dict"])
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in model.state_dict().items():
        assert torch.equal(v, ck["model_state_dict"][k]), k
