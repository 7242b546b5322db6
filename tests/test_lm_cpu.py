"""CPU checks of the LM row: the oracle restatement against the live reference RNNModel and against the
committed golden vectors, the LMDataset.sample contract, and the host-side API surface (no GPU needed)."""
import argparse
import os

import numpy as np
import pytest
import torch

from oracle import live_reference as live
from oracle import make_golden as mg
from oracle import ref_lm

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _blocks():
    cfg, m = ref_lm.LM_SMALL, mg.LM_SMALL_GOLD
    p = ref_lm.init_params(cfg, m["seed"])
    (b0, b1), _ = ref_lm.synth_blocks(cfg, 2, m["T"], m["B"], m["data_seed"])
    return cfg, p, (b0, b1)


def test_lm_oracle_matches_golden():
    """ref_lm.loss_and_grads vs outputs of the reference RNNModel (two chained blocks; hidden state threaded)."""
    g = np.load(os.path.join(GOLD, "lm_small.npz"))
    cfg, p, blocks = _blocks()
    hidden = None
    for i, (tok, trg) in enumerate(blocks):
        loss, grads, logits, hidden = ref_lm.loss_and_grads(p, cfg, tok, trg, hidden)
        assert abs(loss - float(g[f"loss{i}"])) < 1e-5
        np.testing.assert_allclose(logits.numpy(), g[f"logits{i}"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(hidden[0].numpy(), g[f"h{i}"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(hidden[1].numpy(), g[f"c{i}"], rtol=0, atol=1e-6)
        for k, v in grads.items():
            np.testing.assert_allclose(v.numpy(), g[f"grad{i}/" + k], rtol=0, atol=1e-6, err_msg=k)


@pytest.mark.skipif(not live.available(), reason="/root/reference not present")
def test_lm_oracle_matches_live_reference():
    cfg = ref_lm.LmConfig(vocab=211, ninp=32, nhid=40, nlayers=2)
    p = ref_lm.init_params(cfg, 7)
    (blk,), _ = ref_lm.synth_blocks(cfg, 1, 12, 6, 70)
    h0 = (torch.randn(2, 6, 40, generator=torch.Generator().manual_seed(1)) * 0.3,
          torch.randn(2, 6, 40, generator=torch.Generator().manual_seed(2)) * 0.3)
    loss_r, g_r, out_r, hid_r = live.lm_fwd_bwd(cfg, p, blk[0], blk[1], h0)
    loss_o, g_o, out_o, hid_o = ref_lm.loss_and_grads(p, cfg, blk[0], blk[1], h0)
    # parameter inventory and order = the reference's model.parameters()
    m = live.build_lm_model(cfg, p)
    assert [n for n, _ in m.named_parameters()] == ref_lm.param_names(cfg)
    assert abs(loss_r - loss_o) < 1e-6
    assert float((out_r - out_o).abs().max()) < 2e-6
    assert float((hid_r[0] - hid_o[0]).abs().max()) < 1e-6 and float((hid_r[1] - hid_o[1]).abs().max()) < 1e-6
    for k in g_r:
        assert float((g_r[k] - g_o[k]).abs().max()) < 1e-6 * max(1.0, float(g_r[k].abs().max())), k


def test_lm_meta_step_is_first_order_and_threads_hidden():
    """Structure of the meta-iteration the oracle states (lm/main_meta_transfer.py:293-372): train passes all at theta0
    from the running hidden state; the update is lr * clip(sum_i w_i * val-gradient at theta_i)."""
    cfg = ref_lm.LM_SMALL
    p = ref_lm.init_params(cfg, 3)
    train, val = ref_lm.synth_blocks(cfg, 3, 7, 4, 30)
    w = [0.1, 0.1, 0.8]
    new_p, hidden, trl, vall, meta = ref_lm.meta_step(p, cfg, train, val, w, None, lr=2.0, meta_lr_factor=3.0, clip=0.25)
    # hidden after the step = three chained train forwards at theta0
    h = None
    for tok, trg in train:
        _, _, _, h = ref_lm.loss_and_grads(p, cfg, tok, trg, h)
    assert torch.equal(h[0], hidden[0]) and torch.equal(h[1], hidden[1])
    total = float(torch.sqrt(sum((g.double() ** 2).sum() for g in meta.values())))
    assert total <= 0.25 + 1e-5                               # clipped
    for k in p:
        torch.testing.assert_close(new_p[k], p[k] - 2.0 * meta[k])


def test_lm_dataset_sample_contract():
    """lm/util/data.py:20-67: batchify layout and the (block i, block i + 1) sampling rule."""
    import sys
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "meta-transfer-learning_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from lm.util.data import LMDataset
    args = argparse.Namespace(bptt=5, batch_size=3, cuda=False)
    streams = [torch.arange(100), torch.arange(1000, 1047)]
    ds = LMDataset(streams, args)
    ids = ds.task_list[0]
    assert ids.shape == (33, 3) and ids[:, 0].tolist() == list(range(33)) and ids[0].tolist() == [0, 33, 66]
    for it in (0, 1, 5, 6, 7, 40):
        tr_x, tr_y, va_x, va_y = ds.sample(0, it)
        pos = (it * 5) % 33
        pos -= pos % 5
        n = min(5, 33 - 1 - pos)
        assert torch.equal(tr_x, ids[pos:pos + n]) and torch.equal(tr_y, ids[pos + 1:pos + 1 + n].reshape(-1))
        vpos = ((it + 1) * 5) % 33
        vpos -= vpos % 5
        assert torch.equal(va_x, ids[vpos:vpos + min(5, 32 - vpos)])
    # manifest -1 = the last task (the shared meta-validation block)
    assert torch.equal(ds.sample(-1, 2)[2], ds.sample(1, 2)[2])


def test_lm_model_api_surface():
    """RNNModel keeps the reference's constructor, sub-module names and state_dict keys; compute needs .cuda()."""
    import sys
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "meta-transfer-learning_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    import mtl_b200
    from lm.model.rnn_model import RNNModel
    cfg = ref_lm.LM_SMALL
    torch.manual_seed(5)
    m = RNNModel('LSTM', cfg.vocab, cfg.ninp, cfg.nhid, cfg.nlayers, dropout=0.2)
    assert list(m.state_dict().keys()) == ref_lm.param_names(cfg)
    assert [n for n, _ in mtl_b200.lm_param_specs(m.spec())] == ref_lm.param_names(cfg)
    assert float(m.decoder.bias.abs().max()) == 0.0 and float(m.encoder.weight.abs().max()) <= 0.1
    with pytest.raises(NotImplementedError):
        RNNModel('GRU', 10, 4, 4, 1)
    with pytest.raises(mtl_b200.MtlError):
        m(torch.zeros(3, 2, dtype=torch.long), m.init_hidden(2))      # no CPU path


def test_lm_corpus_shares_one_dictionary_and_task_weights(tmp_path):
    """The three corpora of lm/main_meta_transfer.py:130-139 extend ONE dictionary in reading order (`<eos>` closes every
    line), and the validation losses are weighted (1 - ratio) / 2, (1 - ratio) / 2, ratio (lm/main_meta_transfer.py:346-349)."""
    import sys
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "meta-transfer-learning_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from lm.meta import task_weights
    from lm.util.data import Corpus
    (tmp_path / "a.txt").write_text("x y\nz  x\n")
    (tmp_path / "b.txt").write_text("y w\n")
    a = Corpus(str(tmp_path / "a.txt"))
    b = Corpus(str(tmp_path / "b.txt"), None, str(tmp_path / "a.txt"), a.dictionary)
    assert a.train.tolist() == [0, 1, 2, 3, 0, 2]                     # x y <eos> z x <eos>
    assert b.train.tolist() == [1, 4, 2] and b.test.tolist() == a.train.tolist() and b.valid is None
    assert b.dictionary is a.dictionary and len(a.dictionary) == 5 and a.dictionary.idx2word[4] == "w"
    w = task_weights(3, 0.8)
    assert w == pytest.approx([0.1, 0.1, 0.8]) and sum(w) == pytest.approx(1.0)
