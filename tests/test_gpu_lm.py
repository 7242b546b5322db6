"""-m gpu: the LSTM language-model path through the C ABI against the CPU oracle (oracle/ref_lm.py, pinned
against the live reference RNNModel) and the committed golden vectors.  Tolerance: 1e-3 relative fp32 per tensor
(north_star), measured far below; argmax indices bit-exact."""
import argparse
import os
import sys

import numpy as np
import pytest
import torch

import mtl_b200
from gpu_util import GEMM_MODE, dev, rel_err
from oracle import make_golden as mg
from oracle import ref_lm

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = {0: 2e-5, 1: 5e-3, 2: 1e-4}[GEMM_MODE]


def _session(cfg):
    return mtl_b200.LmSession(mtl_b200.LmSpec(cfg.vocab, cfg.ninp, cfg.nhid, cfg.nlayers), dev(), gemm_mode=GEMM_MODE)


def test_lm_small_pass_matches_golden_reference_vectors():
    g = np.load(os.path.join(GOLD, "lm_small.npz"))
    cfg, m = ref_lm.LM_SMALL, mg.LM_SMALL_GOLD
    p = ref_lm.init_params(cfg, m["seed"])
    (b0, b1), _ = ref_lm.synth_blocks(cfg, 2, m["T"], m["B"], m["data_seed"])
    s = _session(cfg)
    theta = s.new_arena()
    s.load(theta, p)
    hidden = None
    for i, (tok, trg) in enumerate((b0, b1)):
        grad = s.new_arena()
        out = s.run(theta, tok, trg, hidden=hidden, grad=grad, want_logits=True)
        torch.cuda.synchronize()
        hidden = out["hidden"]
        assert abs(float(out["loss"][0]) - float(g[f"loss{i}"])) < TOL * abs(float(g[f"loss{i}"]))
        assert int(out["loss"][1]) == tok.numel()                          # nn.CrossEntropyLoss(): every token counts
        ref_logits = torch.from_numpy(g[f"logits{i}"])
        assert rel_err(out["logits"], ref_logits) < TOL
        assert torch.equal(out["logits"].cpu().argmax(-1), ref_logits.argmax(-1))
        assert rel_err(hidden[0], torch.from_numpy(g[f"h{i}"])) < TOL and rel_err(hidden[1], torch.from_numpy(g[f"c{i}"])) < TOL
        gv = s.views(grad)
        for k in ref_lm.param_names(cfg):
            assert rel_err(gv[k], torch.from_numpy(g[f"grad{i}/" + k])) < 10 * TOL, k


@pytest.mark.parametrize("T,B", [(35, 20), (7, 3), (1, 1)])
def test_lm_cfg5_pass_matches_oracle(T, B):
    """The script's default model (emsize = nhid = 200, 2 layers, bptt 35, batch 20) on a 10k-word synthetic vocabulary."""
    cfg = ref_lm.LmConfig(vocab=10000 if T > 1 else 37)
    p = ref_lm.init_params(cfg, 11)
    (blk,), _ = ref_lm.synth_blocks(cfg, 1, T, B, 110)
    gen = torch.Generator().manual_seed(3)
    h0 = (torch.randn(cfg.nlayers, B, cfg.nhid, generator=gen) * 0.2, torch.randn(cfg.nlayers, B, cfg.nhid, generator=gen) * 0.2)
    torch.set_num_threads(os.cpu_count() or 1)
    loss_o, g_o, logits_o, hid_o = ref_lm.loss_and_grads(p, cfg, blk[0], blk[1], h0)
    s = _session(cfg)
    theta, grad = s.new_arena(), s.new_arena()
    s.load(theta, p)
    out = s.run(theta, blk[0], blk[1], hidden=h0, grad=grad, scale=0.5, want_logits=True)
    torch.cuda.synchronize()
    assert abs(float(out["loss"][0]) - loss_o) < TOL * abs(loss_o)
    assert rel_err(out["logits"], logits_o) < TOL
    assert rel_err(out["hidden"][0], hid_o[0]) < TOL and rel_err(out["hidden"][1], hid_o[1]) < TOL
    gv = s.views(grad)
    for k in ref_lm.param_names(cfg):
        assert rel_err(gv[k], 0.5 * g_o[k]) < 10 * TOL, k                  # loss_scale multiplies the gradient
    # a second backward ACCUMULATES (no zero_grad inside the pass)
    s.run(theta, blk[0], blk[1], hidden=h0, grad=grad, scale=0.5)
    torch.cuda.synchronize()
    assert rel_err(gv["decoder.weight"], g_o["decoder.weight"]) < 10 * TOL


def test_lm_meta_step_matches_oracle():
    """One and then a second chained meta-iteration (hidden state carried) against ref_lm.meta_step, dropout off."""
    cfg = ref_lm.LmConfig(vocab=1000, ninp=64, nhid=64, nlayers=2)
    p = ref_lm.init_params(cfg, 5)
    s = _session(cfg)
    theta, work, grad, meta = s.new_arena(), s.new_arena(), s.new_arena(), s.new_arena()
    s.load(theta, p)
    hidden = s.new_hidden(4)
    hid_o = None
    w = [0.1, 0.1, 0.8]
    for it in range(2):
        train, val = ref_lm.synth_blocks(cfg, 3, 9, 4, 50 + it)
        res = torch.zeros(3, 16, device=dev())
        s.meta_step(theta, work, grad, meta, hidden, train, val, w, lr=2.0, meta_lr_factor=3.0, clip=0.25, dropout=0.0,
                    seed=it, results=res)
        torch.cuda.synchronize()
        p, hid_o, trl, vall, meta_o = ref_lm.meta_step(p, cfg, train, val, w, hid_o, lr=2.0, meta_lr_factor=3.0, clip=0.25)
        r = res.cpu()
        for i in range(3):
            assert abs(float(r[i, 0]) - trl[i]) < 5 * TOL * abs(trl[i]) and abs(float(r[i, 8]) - vall[i]) < 5 * TOL * abs(vall[i])
        mv, tv = s.views(meta), s.views(theta)
        for k in ref_lm.param_names(cfg):
            assert rel_err(mv[k], meta_o[k]) < 20 * TOL, ("meta grad", it, k)
            assert rel_err(tv[k], p[k]) < 5 * TOL, ("theta", it, k)
        assert rel_err(hidden[0], hid_o[0]) < 5 * TOL and rel_err(hidden[1], hid_o[1]) < 5 * TOL


def test_lm_dropout_is_a_consistent_mask():
    """Train-mode dropout: masks come from Philox keyed by (seed, site), are recomputed in the backward, and scale kept
    units by 1 / (1 - p): the same seed reproduces the pass bit for bit, another seed does not, and the gradient of the
    dropped pass is the gradient of THAT masked network (finite-difference check on one decoder bias)."""
    cfg = ref_lm.LM_SMALL
    p = ref_lm.init_params(cfg, 2)
    (blk,), _ = ref_lm.synth_blocks(cfg, 1, 8, 4, 20)
    s = _session(cfg)
    theta = s.new_arena()
    s.load(theta, p)
    a = s.run(theta, blk[0], blk[1], dropout=0.3, seed=9, want_logits=True)
    b = s.run(theta, blk[0], blk[1], dropout=0.3, seed=9, want_logits=True)
    c = s.run(theta, blk[0], blk[1], dropout=0.3, seed=10, want_logits=True)
    torch.cuda.synchronize()
    assert torch.equal(a["logits"], b["logits"]) and not torch.equal(a["logits"], c["logits"])
    grad = s.new_arena()
    s.run(theta, blk[0], blk[1], dropout=0.3, seed=9, grad=grad)
    tv, gv = s.views(theta), s.views(grad)
    eps = 1e-2
    base = float(a["loss"][0])
    tv["decoder.bias"][3] += eps
    up = float(s.run(theta, blk[0], blk[1], dropout=0.3, seed=9)["loss"][0])
    assert abs((up - base) / eps - float(gv["decoder.bias"][3])) < 2e-3


def test_lm_model_and_trainer_api():
    """RNNModel(...).cuda() + LMMetaTrainer.step on an LMDataset: the reference-facing surface end to end."""
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "meta-transfer-learning_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from lm.meta import LMMetaTrainer
    from lm.model.rnn_model import RNNModel
    from lm.util.data import LMDataset
    torch.manual_seed(4)
    V = 300
    args = argparse.Namespace(bptt=6, batch_size=4, cuda=True, lr=1.0, meta_lr_factor=3.0, clip=0.25, dropout=0.0, ratio=0.8, seed=1)
    gen = torch.Generator().manual_seed(8)
    ds = LMDataset([torch.randint(0, V, (500,), generator=gen) for _ in range(3)], args)
    model = RNNModel('LSTM', V, 32, 32, 2, dropout=0.0).cuda()
    p0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    out, hid = model(ds.sample(0, 0)[0], model.init_hidden(4))
    cfg = ref_lm.LmConfig(vocab=V, ninp=32, nhid=32, nlayers=2)
    logits_o, _ = ref_lm.forward(p0, cfg, ds.sample(0, 0)[0].cpu())
    assert out.shape == (6, 4, V) and rel_err(out, logits_o) < TOL
    tr = LMMetaTrainer(model, args)
    res = tr.step(ds, 0)
    torch.cuda.synchronize()
    train = [(ds.sample(i, 0)[0].cpu(), ds.sample(i, 0)[1].cpu()) for i in range(3)]
    val = (ds.sample(-1, 0)[2].cpu(), ds.sample(-1, 0)[3].cpu())
    new_p, *_ = ref_lm.meta_step(p0, cfg, train, val, [0.1, 0.1, 0.8], None, lr=1.0, meta_lr_factor=3.0, clip=0.25)
    for k, v in model.state_dict().items():                               # the parameters ARE the arena views
        assert rel_err(v, new_p[k]) < 5 * TOL, k
    assert res.shape == (3, 16)


def test_lm_meta_step_graph_replay_equals_eager():
    """graph=True: static token buffers + device seed word + CUDA-graph replay from the third call on; same parameters
    and hidden state as the eager path after four iterations (dropout off: bit-for-bit up to reduce-add order)."""
    cfg = ref_lm.LmConfig(vocab=500, ninp=32, nhid=32, nlayers=2)
    p = ref_lm.init_params(cfg, 6)
    out = []
    for use_graph in (False, True):
        s = _session(cfg)
        theta, work, grad, meta = s.new_arena(), s.new_arena(), s.new_arena(), s.new_arena()
        s.load(theta, p)
        hidden = s.new_hidden(4)
        res = torch.zeros(3, 16, device=dev())
        for it in range(4):
            train, val = ref_lm.synth_blocks(cfg, 3, 6, 4, 80 + it)
            s.meta_step(theta, work, grad, meta, hidden, train, val, [0.1, 0.1, 0.8], lr=1.0, meta_lr_factor=3.0, clip=0.25,
                        dropout=0.0, seed=it, results=res, graph=use_graph)
        torch.cuda.synchronize()
        out.append((theta.cpu().clone(), hidden[0].cpu().clone(), res.cpu().clone()))
    assert rel_err(out[1][0], out[0][0]) < 1e-5 and rel_err(out[1][1], out[0][1]) < 1e-5
    assert rel_err(out[1][2][:, 8], out[0][2][:, 8]) < 1e-5


def test_lm_cli_runs_on_synthetic_corpora(tmp_path):
    """lm/main_meta_transfer.py (the reference script's flags) end to end on nine tiny synthetic corpus files: model
    build, LMDataset sampling, meta-iterations from the CUDA graph, a log line, validation / test evaluation, checkpoint."""
    import random
    import subprocess
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "meta-transfer-learning_b200")
    rnd = random.Random(3)
    words = [f"w{i}" for i in range(60)]
    d = tmp_path / "data"
    d.mkdir()
    for name in ("seame_train", "seame_valid", "seame_test", "cv_train", "cv_valid", "cv_test", "hkust_train", "hkust_dev"):
        n_lines = 120 if name.endswith("train") else 30
        with open(d / f"{name}.txt", "w") as f:
            for _ in range(n_lines):
                f.write(" ".join(rnd.choice(words) for _ in range(rnd.randint(3, 12))) + "\n")
    cmd = [sys.executable, os.path.join(pkg, "lm", "main_meta_transfer.py"), "--cuda", "--data-dir", str(d), "--emsize", "32",
           "--nhid", "32", "--bptt", "5", "--batch_size", "4", "--lr", "1", "--iterations", "13", "--log-interval", "4",
           "--valid-interval", "8", "--save", str(tmp_path / "model"), "--name", "t"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "| it   4 | lr" in r.stdout and "val loss" in r.stdout and "End of training" in r.stdout
    assert os.path.exists(tmp_path / "model" / "t.pt")
    sd = torch.load(tmp_path / "model" / "t.pt", map_location="cpu")
    assert list(sd.keys())[0] == "encoder.weight" and sd["encoder.weight"].shape[1] == 32


def test_lm_meta_step_with_ragged_blocks_matches_oracle():
    """Task blocks of different lengths (the tail of a corpus is shorter than bptt, lm/util/data.py:37): the iteration is
    composed from single passes and arena kernels and still equals the oracle."""
    cfg = ref_lm.LmConfig(vocab=400, ninp=32, nhid=32, nlayers=2)
    p = ref_lm.init_params(cfg, 8)
    g = torch.Generator().manual_seed(12)

    def block(T, B=4):
        stream = torch.randint(0, cfg.vocab, (T + 1, B), generator=g)
        return stream[:T].contiguous(), stream[1:].reshape(-1).contiguous()
    train, val = [block(7), block(3), block(7)], block(5)
    s = _session(cfg)
    theta, work, grad, meta = s.new_arena(), s.new_arena(), s.new_arena(), s.new_arena()
    s.load(theta, p)
    hidden = s.new_hidden(4)
    res = torch.zeros(3, 16, device=dev())
    w = [0.1, 0.1, 0.8]
    s.meta_step(theta, work, grad, meta, hidden, train, val, w, lr=1.5, meta_lr_factor=3.0, clip=0.25, dropout=0.0, seed=0,
                results=res, graph=True)
    torch.cuda.synchronize()
    new_p, hid_o, trl, vall, _ = ref_lm.meta_step(p, cfg, train, val, w, None, lr=1.5, meta_lr_factor=3.0, clip=0.25)
    tv = s.views(theta)
    for k in ref_lm.param_names(cfg):
        assert rel_err(tv[k], new_p[k]) < 5 * TOL, k
    assert rel_err(hidden[0], hid_o[0]) < 5 * TOL
    r = res.cpu()
    assert all(abs(float(r[i, 8]) - vall[i]) < 5 * TOL * abs(vall[i]) for i in range(3))


def test_lm_sequence_kernels_match_step_kernels():
    """MTL_LM_SEQ=1: one persistent kernel per layer and direction (weights in registers, global-memory barrier between
    time steps) instead of one launch per step -- same pass, checked against the oracle in a fresh process (the switch is
    read once per process)."""
    import subprocess
    env = dict(os.environ, MTL_LM_SEQ="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-k",
                        "cfg5_pass_matches_oracle or meta_step_matches_oracle"], capture_output=True, text=True, env=env, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-1500:] + r.stderr[-500:]


def test_lm_trainer_replays_one_graph():
    """LMMetaTrainer.step keeps every pointer of the iteration fixed (static token / loss buffers, device seed word): the
    step is captured ONCE and replayed afterwards, and the replayed iterations keep training (loss blocks change)."""
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "meta-transfer-learning_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from lm.meta import LMMetaTrainer
    from lm.model.rnn_model import RNNModel
    from lm.util.data import LMDataset
    torch.manual_seed(9)
    V = 200
    args = argparse.Namespace(bptt=5, batch_size=4, cuda=True, lr=1.0, meta_lr_factor=3.0, clip=0.25, dropout=0.2, ratio=0.8, seed=3)
    gen = torch.Generator().manual_seed(5)
    ds = LMDataset([torch.randint(0, V, (900,), generator=gen) for _ in range(3)], args)
    model = RNNModel('LSTM', V, 32, 32, 2, dropout=0.2).cuda()
    model.train()
    tr = LMMetaTrainer(model, args)
    graphs, losses = [], []
    for it in range(6):
        r = tr.step(ds, it)
        losses.append(r[:, 8].clone())
        g = getattr(model.session, "_graph", None)
        graphs.append(None if g is None else g["graph"])
    torch.cuda.synchronize()
    assert graphs[0] is None and graphs[1] is not None            # eager first sighting, captured at the second
    assert all(g is graphs[1] for g in graphs[2:])                # ... and replayed ever after
    assert not torch.equal(losses[2], losses[5])
    assert all(torch.isfinite(l).all() for l in losses)
