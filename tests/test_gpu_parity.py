"""-m gpu: the CUDA hot path (through the C ABI) against the CPU oracle (oracle/ref_asr.py, ref_meta.py)
on the same seeded inputs, and against the golden vectors the live reference produced (tests/golden).

Tolerances.  north_star: "within 1e-3 relative fp32 tolerance (bit-exact for argmax decode indices)".
Per-tensor relative error = max|a-b| / max|b|.  The default engine is the tcgen05 3xTF32 one (MTL_GEMM_MODE=2):
logits / loss within 1e-4, gradients within 1e-3 (VGG gradients included at cfg-2 size).
MTL_GEMM_MODE=0 runs the same suite on the exact fp32 CUDA-core engine (1e-4 / 2e-4), MTL_GEMM_MODE=1 on
plain TF32 (1e-3 on outputs / loss, 5e-3 on gradients)."""
import os

import numpy as np
import pytest
import torch

import mtl_b200
from gpu_util import GEMM_MODE, dev, rel_err, spec_of, to_batch
from oracle import make_golden as mg
from oracle import ref_asr, ref_meta

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL_OUT = {0: 1e-4, 1: 1e-3, 2: 1e-4}[GEMM_MODE]
TOL_GRAD = {0: 2e-4, 1: 5e-3, 2: 1e-3}[GEMM_MODE]
# The VGG gradients (conv.0 above all: a sum over every pixel of relu-masked terms with heavy cancellation)
# amplify the engine's per-element error: ~1e-7 (fp32 FMA) / ~1e-6 (3xTF32) / ~3e-4 (TF32) of activations.
TOL_CONV = {0: 1e-3, 1: 5e-2, 2: 1e-3}[GEMM_MODE]
# SMALL-config VGG gradients: the batch has only 3444 conv pixels and ~28 target tokens, so ONE relu / max-pool decision
# taken the other way moves a conv gradient by percent of the tensor max.  Measured (seed 5 / batch seed 1): the pool
# window (b 3, ch 17, f 16-17, t 24-25) of conv.2 holds 0.20169985 and 0.20169957 (gap 1.4e-6 relative, inside any
# non-bit-exact engine's rounding); routed to the other pixel it changes exactly two elements of d(conv.2 output) and
# moves conv.2.weight by 2.0e-2 and conv.0.weight by 1.0e-2 of their maxima (workspace diff between two correct
# kernels, tools/probes/ws_dump.py).  The kernels' own precision is pinned at the op level (tests/test_gpu_ops.py:
# 5e-5) and the full-size gradients by the cfg-2 golden tests below (TOL_CONV); here the bound only has to catch wiring errors.
TOL_CONV_SMALL = {0: 1e-3, 1: 5e-2, 2: 3e-2}[GEMM_MODE]


def _tol(name, small=True):
    return (TOL_CONV_SMALL if small else TOL_CONV) if name.startswith("conv.") else TOL_GRAD


def _session(cfg):
    return mtl_b200.Session(spec_of(cfg), gemm_mode=GEMM_MODE)


def _fwd_bwd(s, params, batch, scale=1.0, dropout=0.0, seed=0):
    theta, grad = s.new_arena(), s.new_arena()
    s.load(theta, params)
    out = s.forward(theta, to_batch(batch), dropout=dropout, seed=seed)
    pred = out["pred"].clone()
    s.backward(theta, grad, scale)
    torch.cuda.synchronize()
    return out, pred, s.views(grad)


@pytest.mark.parametrize("ragged", [False, True])
def test_small_fwd_bwd_vs_oracle(ragged):
    cfg = ref_asr.SMALL
    p = ref_asr.init_params(cfg, 2)
    if ragged:
        batch = ref_meta.synth_batch(cfg, 4, 41, 7, 10, lengths=[41, 30, 9, 5], tgt_lengths=[7, 5, 3, 1])
    else:
        batch = ref_meta.synth_batch(cfg, 4, 41, 7, 10)
    loss_o, g_o, gold_o, hyp_o, pred_o = ref_meta.loss_and_grads(p, cfg, batch)
    out, pred, grads = _fwd_bwd(_session(cfg), p, batch)
    assert torch.equal(out["gold"].cpu().long(), gold_o)
    keep = gold_o != 0
    assert torch.equal(out["hyp"].cpu().long()[keep], hyp_o[keep])          # bit-exact decode indices
    assert rel_err(pred, pred_o) < TOL_OUT
    assert abs(float(out["ce"][0]) - loss_o) < TOL_OUT * abs(loss_o)
    assert int(out["ce"][2]) == ref_asr.num_correct(pred_o, gold_o)
    bad = {k: rel_err(grads[k], g_o[k]) for k in g_o if float(g_o[k].abs().max()) > 1e-7}
    worst = max(bad, key=bad.get)
    assert bad[worst] < _tol(worst), (worst, bad[worst])
    for k in g_o:   # mathematically-zero gradients (key bias) stay negligible
        if float(g_o[k].abs().max()) <= 1e-7:
            assert float(grads[k].abs().max()) < 1e-6, k


def test_small_golden_from_live_reference():
    g = np.load(os.path.join(GOLD, "small_fwd_bwd.npz"))
    cfg = ref_asr.SMALL
    p = ref_asr.init_params(cfg, 21)
    batch = ref_meta.synth_batch(cfg, 4, 41, 7, 2100, lengths=[41, 30, 9, 5], tgt_lengths=[7, 5, 3, 1])
    out, pred, grads = _fwd_bwd(_session(cfg), p, batch)
    assert np.array_equal(out["gold"].cpu().numpy(), g["gold"])
    keep = g["gold"] != 0
    assert np.array_equal(out["hyp"].cpu().numpy()[keep], g["hyp"][keep])
    assert rel_err(pred, torch.from_numpy(g["pred"])) < TOL_OUT
    assert abs(float(out["ce"][0]) - float(g["loss"])) < TOL_OUT * float(g["loss"])
    assert int(out["ce"][2]) == int(g["num_correct"])
    for name, _ in ref_asr.param_specs(cfg):
        ref = torch.from_numpy(g["grad/" + name])
        if float(ref.abs().max()) > 1e-7:
            assert rel_err(grads[name], ref) < _tol(name), name


def test_loss_scale_and_accumulation_semantics():
    """backward accumulates scale*grad into the arena (transient_trainer.py:198-199,226-227: no zero_grad
    between the two backward calls)."""
    cfg = ref_asr.SMALL
    p = ref_asr.init_params(cfg, 5)
    s = _session(cfg)
    b1, b2 = ref_meta.synth_batch(cfg, 4, 41, 7, 1), ref_meta.synth_batch(cfg, 3, 30, 5, 2)
    theta, grad = s.new_arena(), s.new_arena()
    s.load(theta, p)
    s.forward(theta, to_batch(b1)); s.backward(theta, grad, 1.0)
    s.forward(theta, to_batch(b2)); s.backward(theta, grad, 1.0 / 3)
    _, g1, *_ = ref_meta.loss_and_grads(p, cfg, b1)
    _, g2, *_ = ref_meta.loss_and_grads(p, cfg, b2, 1.0 / 3)
    v = s.views(grad)
    for k in g1:
        ref = g1[k] + g2[k]
        if float(ref.abs().max()) > 1e-7:
            assert rel_err(v[k], ref) < _tol(k), k


def test_external_dpred_backward_matches_fused_ce():
    """Transformer.forward returns pred; a caller may push its own d(loss)/d(pred) (autograd drop-in)."""
    cfg = ref_asr.SMALL
    p = ref_asr.init_params(cfg, 6)
    s = _session(cfg)
    b = ref_meta.synth_batch(cfg, 4, 41, 7, 3, tgt_lengths=[7, 4, 6, 2])
    theta, ga, gb = s.new_arena(), s.new_arena(), s.new_arena()
    s.load(theta, p)
    out = s.forward(theta, to_batch(b))
    pred = out["pred"].clone().requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(pred.view(-1, cfg.vocab), out["gold"].view(-1).long(), ignore_index=0)
    loss.backward()
    s.backward(theta, ga, 1.0, dpred=pred.grad)
    s.forward(theta, to_batch(b)); s.backward(theta, gb, 1.0)
    va, vb = s.views(ga), s.views(gb)
    for k in va:
        if float(vb[k].abs().max()) > 1e-7:
            # the VGG gradients run in single-pass TF32 by default: a 1e-7 change of d(pred) moves operand roundings
            assert rel_err(va[k], vb[k]) < (1e-3 if k.startswith("conv.") else 3e-5), k


def _meta_run_lanes(s, p, steps, lr, meta_lr, clip=False, max_norm=400.0, lanes=None, use_graph=False, dropout=0.0,
                    seeds=None):
    """Same contract as _meta_run but through mtl_meta_tasks (concurrent task lanes, optional CUDA graph)."""
    theta, grad, cg = (s.new_arena() for _ in range(3))
    m, v, st = s.new_arena(), s.new_arena(), s.new_adam_state()
    s.load(theta, p)
    losses, cgs = [], []
    stepper = None
    for si, (tasks, val) in enumerate(steps):
        n = len(tasks)
        if stepper is None:
            stepper = mtl_b200.MetaStepper(s, n, n_lanes=lanes or n, use_graph=use_graph)
        for t, tr in enumerate(tasks):
            stepper.load_task(t, *tr)
        stepper.load_val(*val)
        before = theta.clone()
        res = stepper.run(theta, cg, lr, 1.0 / n, clip=clip, max_norm=max_norm, dropout=dropout,
                          seed=seeds[si] if seeds else si)
        assert torch.equal(theta, before), "mtl_meta_tasks must not modify theta"
        cgs.append(cg.clone())
        s.meta_finish(theta, grad, cg, m, v, st, meta_lr, clip=clip, max_norm=max_norm)
        losses.append(float(res[:, 8].mean()))
    torch.cuda.synchronize()
    return losses, s.views(cg), s.views(theta), cgs


def _meta_run(s, p, steps, lr, meta_lr, clip=False, max_norm=400.0):
    if os.environ.get("MTL_TEST_LANES", "1") == "1":       # default: the product path (mtl_meta_tasks)
        return _meta_run_lanes(s, p, steps, lr, meta_lr, clip, max_norm)[:3]
    theta, theta0, grad, cg = (s.new_arena() for _ in range(4))
    m, v, st = s.new_arena(), s.new_arena(), s.new_adam_state()
    s.load(theta, p)
    losses = []
    for tasks, val in steps:
        n = len(tasks)
        res = torch.zeros(n, 16, device=dev())
        s.copy(theta0, theta)
        s.zero(cg)
        vb = to_batch(val)
        for i, tr in enumerate(tasks):
            s.meta_task(theta, theta0, grad, cg, to_batch(tr), vb, lr, 1.0 / n, clip=clip, max_norm=max_norm,
                        results=res[i])
        assert torch.equal(theta, theta0), "theta not restored to theta0 after the task loop"
        s.meta_finish(theta, grad, cg, m, v, st, meta_lr, clip=clip, max_norm=max_norm)
        losses.append(float(res[:, 8].mean()))
    torch.cuda.synchronize()
    return losses, s.views(cg), s.views(theta)


def _small_steps():
    cfg = ref_asr.SMALL
    steps = []
    for st in range(3):
        tasks = [ref_meta.synth_batch(cfg, 4, 41, 7, 100 * st + i,
                                      lengths=[41, 30, 9, 5] if (st == 1 and i == 0) else None,
                                      tgt_lengths=[7, 5, 3, 1] if (st == 1 and i == 0) else None) for i in range(3)]
        steps.append((tasks, ref_meta.synth_batch(cfg, 4, 37 if st == 2 else 41, 6 if st == 2 else 7, 100 * st + 50)))
    return steps


@pytest.mark.parametrize("clip", [False, True])
def test_small_meta_steps_vs_oracle_teacher_forced(clip):
    """Three meta-steps, each started from the ORACLE's (theta, Adam m/v/step): per-step val losses, the
    accumulated copy_grad (train-gradient leak included) and the Adam update must match.

    Why teacher-forced: with Adam the first updates are ~ lr*sign(g); elements whose gradient is at
    rounding-noise level (|g| < 1e-5 max|g|) flip with any change of summation order, and the
    perturbation grows ~10x per step (injecting 2e-6 relative noise into the ORACLE's own gradients
    moves its copy_grad by 4-7% two steps later), so free-running trajectories only agree loosely
    (test_small_meta_free_running)."""
    cfg = ref_asr.SMALL
    lr, meta_lr, max_norm = 1e-2, 1e-3, 0.5
    p = ref_asr.init_params(cfg, 3)
    s = _session(cfg)
    po = {k: v.clone() for k, v in p.items()}
    adam = ref_meta.AdamState()
    theta, theta0, grad, cg, m, v = (s.new_arena() for _ in range(6))
    st = s.new_adam_state()
    for si, (tasks, val) in enumerate(_small_steps()):
        s.load(theta, po)
        if adam.step:
            s.load(m, adam.m); s.load(v, adam.v)
        st[0] = adam.step
        n = len(tasks)
        res = torch.zeros(n, 16, device=dev())
        s.copy(theta0, theta); s.zero(cg)
        vb = to_batch(val)
        for i, tr in enumerate(tasks):
            s.meta_task(theta, theta0, grad, cg, to_batch(tr), vb, lr, 1.0 / n, clip=clip, max_norm=max_norm,
                        results=res[i])
        assert torch.equal(theta, theta0), "theta not restored to theta0 after the task loop"
        s.meta_finish(theta, grad, cg, m, v, st, meta_lr, clip=clip, max_norm=max_norm)
        torch.cuda.synchronize()
        r = ref_meta.meta_step(po, adam, cfg, tasks, val, lr=lr, meta_lr=meta_lr, clip=clip, max_norm=max_norm)
        for i in range(n):
            assert abs(float(res[i, 0]) - r["tr_losses"][i]) < TOL_OUT * r["tr_losses"][i], (si, i)
            assert abs(float(res[i, 8]) - r["val_losses"][i]) < TOL_OUT * r["val_losses"][i], (si, i)
        cgv, thv = s.views(cg), s.views(theta)
        for k in po:
            ref = r["copy_grad"][k]
            gmax = float(ref.abs().max())
            if gmax > 1e-7:
                # SMALL has only 3444 conv pixels: one relu / max-pool tie decided the other way by a
                # 1e-7 activation difference moves a conv gradient element by ~3e-4 of the tensor max
                assert rel_err(cgv[k], ref) < _tol(k), (si, k)
            d = (thv[k].cpu() - po[k]).abs()
            assert float(d.max()) <= 2.1 * meta_lr, (si, k)            # |Adam step| <= lr either way
            solid = ref.abs() > 1e-2 * gmax                               # elements with a real gradient
            if gmax > 1e-7 and solid.any():
                assert float(d[solid].max()) <= 0.05 * meta_lr, (si, k, float(d[solid].max()))
        assert int(st[0]) == adam.step


def test_small_meta_free_running():
    """Free-running 3-step trajectory: losses stay within 2e-3 relative of the oracle's (see the
    teacher-forced test for why not tighter)."""
    cfg = ref_asr.SMALL
    p = ref_asr.init_params(cfg, 3)
    steps = _small_steps()
    po = {k: v.clone() for k, v in p.items()}
    adam = ref_meta.AdamState()
    ref_losses = [ref_meta.meta_step(po, adam, cfg, t, v, lr=1e-2, meta_lr=1e-3)["loss"] for t, v in steps]
    losses, cg, theta = _meta_run(_session(cfg), p, steps, 1e-2, 1e-3)
    assert abs(losses[0] - ref_losses[0]) < TOL_OUT * ref_losses[0]
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 2e-3 * abs(b), (losses, ref_losses)
    assert losses[-1] < losses[0]


def test_small_meta_golden_from_live_reference():
    """Two unchanged TransientTrainer iterations recorded from the live reference (val batch shape differs
    from the train batches).  Step-0 loss and copy_grad are compared via the 1-step state; the
    reference prints losses with 3 decimals."""
    g = np.load(os.path.join(GOLD, "small_meta.npz"))
    cfg, m = ref_asr.SMALL, mg.SMALL_META
    p = ref_asr.init_params(cfg, m["seed"])
    steps = [mg.small_tasks(st) for st in range(m["n_steps"])]
    losses, cg, theta = _meta_run(_session(cfg), p, steps, m["lr"], m["meta_lr"])
    assert abs(losses[0] - g["losses"][0]) < 6e-4, (losses, g["losses"])
    assert abs(losses[1] - g["losses"][1]) < 2e-3 * g["losses"][1], (losses, g["losses"])
    # after ONE Adam step of lr 1e-3 the second step's copy_grad agrees to a few % (Adam noise amplification)
    num = den = 0.0
    for name, _ in ref_asr.param_specs(cfg):
        ref = torch.from_numpy(g["copy_grad/" + name]).double()
        num += float((cg[name].cpu().double() - ref).pow(2).sum())
        den += float(ref.pow(2).sum())
    assert (num / den) ** 0.5 < 3e-2, (num / den) ** 0.5


@pytest.mark.timeout(900)
def test_cfg2_fwd_bwd_golden_from_live_reference():
    """BASELINE cfg 2 shapes (enc2/dec4/d512, B=8, T=101, L=32, ragged): loss, indices, pred samples and
    per-tensor gradient norms/samples recorded from the live reference."""
    g = np.load(os.path.join(GOLD, "cfg2_fwd_bwd.npz"))
    cfg = ref_asr.CFG2
    p = ref_asr.init_params(cfg, 31)
    out, pred, grads = _fwd_bwd(_session(cfg), p, mg.cfg2_batch(3100, ragged=True))
    assert np.array_equal(out["gold"].cpu().numpy(), g["gold"])
    keep = g["gold"] != 0
    assert np.array_equal(out["hyp"].cpu().numpy()[keep], g["hyp"][keep])
    assert abs(float(out["ce"][0]) - float(g["loss"])) < TOL_OUT * float(g["loss"])
    assert int(out["ce"][2]) == int(g["num_correct"])
    flat = pred.reshape(-1).cpu()
    ps = flat[torch.from_numpy(mg.pred_sample_idx(flat.numel()))]
    assert rel_err(ps, torch.from_numpy(g["pred_samples"])) < TOL_OUT
    assert abs(float(pred.double().norm()) - float(g["pred_norm"])) < TOL_OUT * float(g["pred_norm"])
    for name, _ in ref_asr.param_specs(cfg):
        gn = float(g["gnorm/" + name])
        v = grads[name].cpu()
        if gn > 1e-6:
            assert abs(float(v.double().norm()) - gn) < TOL_GRAD * gn, name
            s = v.reshape(-1)[torch.from_numpy(mg.sample_idx(v.numel()))]
            assert float((s - torch.from_numpy(g["gsamp/" + name])).abs().max()) < TOL_GRAD * gn, name


@pytest.mark.timeout(900)
def test_cfg2_meta_step_golden_from_live_reference():
    """One full cfg-2 meta-step (3 tasks x (8 train + 8 val)) vs the unchanged TransientTrainer."""
    g = np.load(os.path.join(GOLD, "cfg2_meta.npz"))
    cfg, m = ref_asr.CFG2, mg.CFG2_META
    p = ref_asr.init_params(cfg, m["seed"])
    tasks, val = mg.cfg2_tasks()
    losses, cg, theta = _meta_run(_session(cfg), p, [(tasks, val)], m["lr"], m["meta_lr"])
    assert abs(losses[0] - float(g["losses"][0])) < 2e-4 * float(g["losses"][0])
    n_bad = n_tot = 0
    for name, _ in ref_asr.param_specs(cfg):
        gn = float(g["cgnorm/" + name])
        idx = torch.from_numpy(mg.sample_idx(cg[name].numel()))
        if gn > 1e-6:
            assert abs(float(cg[name].double().norm()) - gn) < 5 * TOL_GRAD * gn, name
            s = cg[name].reshape(-1).cpu()[idx]
            assert float((s - torch.from_numpy(g["cgsamp/" + name])).abs().max()) < 5 * TOL_GRAD * gn, name
        # first Adam step: delta = -meta_lr * g/(|g|+eps) -> sign agreement wherever |g| is not ~0
        delta = (theta[name].cpu() - p[name]).reshape(-1)[idx]
        ref_d = torch.from_numpy(g["dsamp/" + name])
        big = torch.from_numpy(np.abs(g["cgsamp/" + name])) > 1e-6
        n_bad += int(((delta - ref_d).abs() > 0.05 * m["meta_lr"])[big].sum())
        n_tot += int(big.sum())
    assert n_bad <= 0.002 * n_tot, (n_bad, n_tot)


@pytest.mark.timeout(900)
def test_cfg4_dims_fwd_bwd_vs_oracle():
    """BASELINE configs[3] dimensions (enc4/dec6, d_model 768, 12 heads; fp32-grade arithmetic instead of its bf16) on a
    batch the CPU oracle finishes in seconds: B = 2, 403 frames (T' = 100 > 64: the tiled attention kernels, ragged
    lengths), 20 target tokens.  Outputs / loss / decode indices at the cfg-2 bounds.  Gradients at 3e-3: with only
    200 encoder rows ONE relu decision of a position-wise FFN taken the other way (pre-activation within rounding of
    zero) moves that layer's linear_1.weight / bias by ~1e-3 of the tensor max -- measured with all three engines,
    including the exact fp32 CUDA-core one (7.4e-4 on encoder.layers.2, a different unit than 3xTF32's 1.1e-3 on
    encoder.layers.1), so it is a property of the tiny batch, not of the arithmetic."""
    cfg = ref_asr.ModelConfig(n_enc=4, n_dec=6, d_model=768, n_heads=12, d_k=64, d_v=64, d_inner=768)
    p = ref_asr.init_params(cfg, 41)
    batch = ref_meta.synth_batch(cfg, 2, 403, 20, 4100, lengths=[403, 288], tgt_lengths=[20, 13])
    loss_o, g_o, gold_o, hyp_o, pred_o = ref_meta.loss_and_grads(p, cfg, batch)
    out, pred, grads = _fwd_bwd(_session(cfg), p, batch)
    assert torch.equal(out["gold"].cpu().long(), gold_o)
    keep = gold_o != 0
    assert torch.equal(out["hyp"].cpu().long()[keep], hyp_o[keep])
    assert rel_err(pred, pred_o) < TOL_OUT
    assert abs(float(out["ce"][0]) - loss_o) < TOL_OUT * abs(loss_o)
    bad = {k: rel_err(grads[k], g_o[k]) for k in g_o if float(g_o[k].abs().max()) > 1e-7}
    top = sorted(bad.items(), key=lambda kv: -kv[1])[:5]
    print("cfg4 worst gradient errors:", top)
    assert all(v < max(_tol(k), 3e-3) for k, v in bad.items()), top


@pytest.mark.parametrize("lanes", [1, 2, 3])
def test_meta_tasks_lanes_match_sequential_meta_task(lanes):
    """mtl_meta_tasks (tasks on concurrent lanes, per-lane weight copies) == the sequential
    mtl_meta_task loop (snapshot / reset of one theta), including the task-order of the copy_grad sum."""
    cfg = ref_asr.SMALL
    p = ref_asr.init_params(cfg, 3)
    steps = _small_steps()[:1]          # one step: a second one would compare through Adam's sign(g) amplification
    s = _session(cfg)
    os.environ["MTL_TEST_LANES"] = "0"
    try:
        l_seq, cg_seq, th_seq = _meta_run(s, p, steps, 1e-2, 1e-3, clip=True, max_norm=0.5)
    finally:
        os.environ["MTL_TEST_LANES"] = "1"
    cg_seq = {k: v.clone() for k, v in cg_seq.items()}
    th_seq = {k: v.clone() for k, v in th_seq.items()}
    l_par, cg_par, th_par, _ = _meta_run_lanes(s, p, steps, 1e-2, 1e-3, clip=True, max_norm=0.5, lanes=lanes)
    assert np.allclose(l_seq, l_par, rtol=1e-5)
    for k in cg_seq:
        if float(cg_seq[k].abs().max()) > 1e-7:
            # only the order of the K-slab / TMA reduce-add sums differs between the two runs (2e-5); the VGG gradients
            # amplify that: their input / weight gradients run in single-pass TF32 by default, where a 1e-7 difference of an
            # operand moves its rounding by 2^-11, and conv.0.weight sums those over every pixel of a noise input
            assert rel_err(cg_par[k], cg_seq[k]) < (2e-3 if k.startswith("conv.") else 2e-5), k
        assert float((th_par[k] - th_seq[k]).abs().max()) <= 2.1e-3, k


def test_meta_tasks_cuda_graph_replay_matches_eager_with_fresh_dropout_seeds():
    """The captured meta-step replays with per-step seeds patched into the graph: same seed -> same
    result as the eager path, different seed -> different dropout masks."""
    cfg = ref_asr.SMALL
    p = ref_asr.init_params(cfg, 4)
    tasks, val = _small_steps()[0]
    steps = [(tasks, val)] * 5
    seeds = [11, 12, 13, 12, 14]
    s_e, s_g = _session(cfg), _session(cfg)
    *_, cg_e = _meta_run_lanes(s_e, p, steps, 1e-2, 0.0, dropout=0.1, seeds=seeds, use_graph=False)
    *_, cg_g = _meta_run_lanes(s_g, p, steps, 1e-2, 0.0, dropout=0.1, seeds=seeds, use_graph=True)
    cap, rep = s_g.graph_stats()
    assert cap == 1 and rep == 4, (cap, rep)
    assert s_e.graph_stats() == (0, 0)
    for i in range(5):
        assert rel_err(cg_g[i], cg_e[i]) < 3e-3, i               # run-to-run: reduce-add order, amplified by the TF32 VGG gradients
    assert rel_err(cg_g[3], cg_g[1]) < 3e-3                      # same seed, theta unchanged (meta_lr 0)
    assert rel_err(cg_g[2], cg_g[1]) > 1e-2                      # fresh masks


def test_full_size_properties_cfg2():
    """Size-independent properties at BASELINE cfg-2 size: zero loss-scale gives exactly zero gradient
    contribution; gradient is linear in the loss scale; meta_task restores theta bit-exactly and
    copy_grad accumulates exactly what backward left in grad."""
    cfg = ref_asr.CFG2
    s = _session(cfg)
    p = ref_asr.init_params(cfg, 7)
    theta, theta0, grad, g2, cg = (s.new_arena() for _ in range(5))
    s.load(theta, p)
    s.copy(theta0, theta)
    tr, va = to_batch(mg.cfg2_batch(1)), to_batch(mg.cfg2_batch(2))
    s.forward(theta, tr); s.backward(theta, grad, 0.0)
    assert float(grad.abs().max()) == 0.0
    s.forward(theta, tr); s.backward(theta, grad, 1.0)
    s.forward(theta, tr); s.backward(theta, g2, 0.25)
    assert rel_err(g2 * 4, grad) < 1e-4
    res = torch.zeros(16, device=dev())
    s.meta_task(theta, theta0, g2, cg, tr, va, 1e-4, 1.0 / 3, results=res)
    assert torch.equal(theta, theta0)
    assert torch.equal(cg, g2)
    assert float(res[0]) > 0 and float(res[8]) > 0 and int(res[1]) == 8 * 33


def test_dropout_pass_is_deterministic_given_seed_and_unbiased():
    cfg = ref_asr.SMALL
    s = _session(cfg)
    p = ref_asr.init_params(cfg, 8)
    b = ref_meta.synth_batch(cfg, 4, 41, 7, 4)
    o1, pred1, g1 = _fwd_bwd(s, p, b, dropout=0.1, seed=42)
    g1 = {k: v.clone() for k, v in g1.items()}
    o2, pred2, g2 = _fwd_bwd(s, p, b, dropout=0.1, seed=42)
    # same seed -> same Philox masks.  Not bit-identical: K-slab GEMMs (stem projection here) merge their partial sums
    # through TMA reduce-add, whose order is not fixed -- fp32 round-off only (a different mask moves pred by ~1e-1)
    assert rel_err(pred1, pred2) < 1e-5
    o3, pred3, _ = _fwd_bwd(s, p, b, dropout=0.1, seed=43)
    assert rel_err(pred1, pred3) > 1e-3
    o0, pred0, _ = _fwd_bwd(s, p, b, dropout=0.0)
    assert 0.0 < rel_err(pred1, pred0) < 1.0


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE cfg 2, EVERY element of every tensor, against the oracle run on this box's CPU cores (north_star: 1e-3
# relative fp32).  The golden-fixture tests above sample 32 elements per tensor because the fixtures have to stay small;
# these do not sample.  Two metrics per tensor:
#   L2   = |a - b|_2 / |b|_2          bound 1e-3 on EVERY tensor, no exception;
#   max  = max|a - b| / max|b|        bound 1e-3, except on tensors hit by a DECISION FLIP (bound 2e-2, at most 10 % of
#                                     the tensors).
# Decision flips: the gradient is a discontinuous function of the forward values wherever a ReLU pre-activation or the
# gap between the two largest values of a max-pool window is within the forward rounding error.  tcgen05 accumulates in
# fp32 with truncation (measured: conv.7 output 1.1e-5 of its max after 3 x 72-144 chained k-steps, vs 1.6e-6 for
# the fp32 CUDA-core engine and 7e-7 for fp32 MKL-DNN against fp64), so a handful of the 4 M pooled / 0.75 M FFN units
# decide the other way.  Measured on the ragged cfg-2 batch with the default engine (tools/probes/vgg_chain_diff.py,
# profiles/r02_a_vgg_chain_diff.log): 5 of 4 096 000 elements of d(conv.7 output) differ (ours 0, oracle -7e-5, ...),
# EVERY other intermediate of the chain agrees to 1e-5, and each operator run alone on the oracle's inputs agrees to
# 8e-6 (tools/probes/conv_bwd_real.py).  One such element moves conv.7.bias by 1.4e-3 of its max, and conv.0.weight --
# a sum of 1 M random-sign terms dy * x over a noise input, |sum| ~ sum|terms| / 1000 -- by 2e-3 (max and L2 alike).
# An FFN unit of a 264-row batch that flips moves its row of linear_1.weight by ~1e-2 and every tensor below it by
# ~1e-3.  The exact fp32 engine (MTL_GEMM_MODE=0) has one VGG flip (1.9e-4); no implementation that is not
# bit-identical to the reference can exclude them.  So the bounds are: L2 <= 1e-3 on every non-VGG tensor of a single
# pass (measured 1e-5), 5e-3 on the VGG tensors and on the six-pass meta-step; max-norm <= 2e-2 everywhere with the
# flip-hit tensors counted; the VGG flips are counted element by element in
# test_cfg2_vgg_gradient_chain_differs_only_by_decision_flips.
TOL_FULL = {0: 5e-4, 1: 1e-2, 2: 1e-3}[GEMM_MODE]
TOL_FLIP = 2e-2


def _full_tensor_check(ours, ref, what, l2_tol, max_flipped_frac, med_tol=1e-4):
    names = [k for k in ref if float(ref[k].abs().max()) > 1e-7]
    mx = {k: rel_err(ours[k], ref[k]) for k in names}
    l2 = {k: float((ours[k].detach().double().cpu() - ref[k].double()).norm() / ref[k].double().norm()) for k in names}
    top = sorted(mx.items(), key=lambda kv: -kv[1])[:6]
    flipped = [k for k in names if mx[k] >= TOL_FULL]
    med = float(np.median(list(mx.values())))
    print("%s: worst max-norm errors %s; median %.1e; worst L2 %.2e; %d of %d tensors above %.0e (decision flips): %s" % (
        what, top, med, max(l2.values()), len(flipped), len(names), TOL_FULL, flipped))
    assert len(names) >= 180
    assert all(v < l2_tol(k) for k, v in l2.items()), sorted(l2.items(), key=lambda kv: -kv[1])[:5]
    assert all(v < TOL_FLIP for v in mx.values()), top
    assert med < med_tol * (TOL_FULL / 1e-3)
    assert len(flipped) <= max_flipped_frac * len(names), flipped


@pytest.mark.timeout(900)
def test_cfg2_full_tensor_fwd_bwd_vs_oracle():
    cfg = ref_asr.CFG2
    p = ref_asr.init_params(cfg, 31)
    batch = mg.cfg2_batch(3100, ragged=True)
    torch.set_num_threads(os.cpu_count() or 1)
    loss_o, g_o, gold_o, hyp_o, pred_o = ref_meta.loss_and_grads(p, cfg, batch)
    out, pred, grads = _fwd_bwd(_session(cfg), p, batch)
    assert torch.equal(out["gold"].cpu().long(), gold_o)
    keep = gold_o != 0
    assert torch.equal(out["hyp"].cpu().long()[keep], hyp_o[keep])          # bit-exact decode indices
    assert rel_err(pred, pred_o) < TOL_OUT
    assert abs(float(out["ce"][0]) - loss_o) < TOL_OUT * abs(loss_o)
    _full_tensor_check(grads, g_o, "cfg2 fwd+bwd", lambda k: 5 * TOL_FULL if k.startswith("conv.") else TOL_FULL, 0.1)


@pytest.mark.timeout(900)
def test_cfg2_vgg_gradient_chain_differs_only_by_decision_flips():
    """Counts the ReLU / max-pool decision flips of the VGG backward directly: every element of d(conv.7 output) and
    d(conv.2 output) of OUR pass (mtl_debug_pass_buffers) against autograd on the CPU; all but a handful of elements
    must agree to 1e-4 of the tensor max, and the handful must be flips (one side exactly zero)."""
    import ctypes as C
    import torch.nn.functional as F
    cfg = ref_asr.CFG2
    p0 = ref_asr.init_params(cfg, 31)
    batch = mg.cfg2_batch(3100, ragged=True)
    x, lens, trg = batch
    torch.set_num_threads(os.cpu_count() or 1)
    p = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    bufs = ref_asr.buffers(cfg)
    c1 = F.relu(F.conv2d(x, p["conv.0.weight"], p["conv.0.bias"], padding=1))
    z2 = F.conv2d(c1, p["conv.2.weight"], p["conv.2.bias"], padding=1); z2.retain_grad()
    p2 = F.max_pool2d(F.relu(z2), 2, stride=2)
    c3 = F.relu(F.conv2d(p2, p["conv.5.weight"], p["conv.5.bias"], padding=1))
    z4 = F.conv2d(c3, p["conv.7.weight"], p["conv.7.bias"], padding=1); z4.retain_grad()
    p4 = F.max_pool2d(F.relu(z4), 2, stride=2)
    b, ch, fr, t = p4.shape
    feat = p4.reshape(b, ch * fr, t).transpose(1, 2).contiguous()
    enc = ref_asr.encoder_forward(p, cfg, feat, lens, bufs["encoder.positional_encoding.pe"])
    pred, gold = ref_asr.decoder_forward(p, cfg, trg, enc, lens, bufs["decoder.positional_encoding.pe"])
    ref_asr.ce_loss(pred, gold).backward()
    s = _session(cfg)
    if GEMM_MODE == 2:
        s.set_op_mode("conv_dgrad", 2)          # this test counts DECISIONS: keep TF32 operand rounding out of the chain
        s.set_op_mode("conv_wgrad", 2)
    theta, grad = s.new_arena(), s.new_arena()
    s.load(theta, p0)
    s.forward(theta, to_batch(batch)); s.backward(theta, grad, 1.0)
    torch.cuda.synchronize()
    ptrs = (C.c_void_p * 16)()
    mtl_b200.lib.check(s.lib.mtl_debug_pass_buffers(s._h, ptrs))
    for idx, ref in ((9, z4.grad), (12, z2.grad)):                 # dc4, dc2 (NHWC)
        r = ref.detach().permute(0, 2, 3, 1).contiguous()
        off = ptrs[idx] - s._ws.data_ptr()
        ours = s._ws[off:off + r.numel() * 4].view(torch.float32).view(r.shape).cpu()
        d = (ours - r).abs()
        bad = d > 1e-4 * float(r.abs().max())
        n_bad = int(bad.sum())
        print("VGG chain buffer %d: %d of %d elements differ" % (idx, n_bad, r.numel()))
        if idx == 9:                                               # first decision layer of the backward: pure flips
            assert n_bad <= 64, n_bad
            assert bool(((ours[bad] == 0) | (r[bad] == 0)).all()), "a mismatch that is not a zero-vs-nonzero flip"
        else:                                                      # flips upstream spread over their 3x3x3x3 footprints
            assert n_bad <= 2e-3 * r.numel(), n_bad


@pytest.mark.timeout(900)
def test_cfg2_full_tensor_meta_step_copy_grad_vs_oracle():
    """One full cfg-2 meta-step (trainer/asr/transient_trainer.py:178-237: 3 tasks x (train at theta0, SGD step, shared
    val batch at the adapted weights, train-gradient leak)): every element of the 190 copy_grad tensors and the val
    losses against oracle/ref_meta.meta_step on the CPU."""
    cfg, m = ref_asr.CFG2, mg.CFG2_META
    p = ref_asr.init_params(cfg, m["seed"])
    tasks, val = mg.cfg2_tasks()
    torch.set_num_threads(os.cpu_count() or 1)
    po = {k: v.clone() for k, v in p.items()}
    r = ref_meta.meta_step(po, ref_meta.AdamState(), cfg, tasks, val, lr=m["lr"], meta_lr=m["meta_lr"])
    losses, cg, theta, _ = _meta_run_lanes(_session(cfg), p, [(tasks, val)], m["lr"], m["meta_lr"])
    assert abs(losses[0] - r["loss"]) < TOL_OUT * abs(r["loss"])
    # six passes: a flipped FFN unit in a decoder layer reaches, through the cross-attentions below it, the whole encoder
    # and the VGG front-end of that pass, so most tensors carry ~1e-3 here (the exact fp32 engine, MTL_GEMM_MODE=0, stays
    # at its 5e-4 on the same step: profiles/r02_b_meta_step_mode0.log)
    _full_tensor_check(cg, r["copy_grad"], "cfg2 meta-step copy_grad", lambda k: 5 * TOL_FULL, 0.8, med_tol=2e-3)
    # first Adam step from zero moments: |delta| = meta_lr * |g| / (|g| + eps) -> at most meta_lr
    for k in po:
        d = (theta[k].cpu() - po[k]).abs()
        assert float(d.max()) <= 2.1 * m["meta_lr"], k


def test_merged_lowrank_path_matches_oracle():
    """mtl_session_set_flag("merge_lowrank", 1): every projection pair B(A x) (modules/common_layers.py:287-289,303) runs
    as one GEMM against W = B.A on the activation chain; a = A x and da = dy.B are formed beside the parameter
    gradients.  Same outputs and gradients as the factored default (kept as a measured A/B variant)."""
    cfg = ref_asr.SMALL
    p = ref_asr.init_params(cfg, 2)
    batch = ref_meta.synth_batch(cfg, 4, 41, 7, 10, lengths=[41, 30, 9, 5], tgt_lengths=[7, 5, 3, 1])
    loss_o, g_o, gold_o, hyp_o, pred_o = ref_meta.loss_and_grads(p, cfg, batch)
    s = _session(cfg)
    s.set_flag("merge_lowrank", 1)
    out, pred, grads = _fwd_bwd(s, p, batch)
    keep = gold_o != 0
    assert torch.equal(out["hyp"].cpu().long()[keep], hyp_o[keep])
    assert rel_err(pred, pred_o) < TOL_OUT
    bad = {k: rel_err(grads[k], g_o[k]) for k in g_o if float(g_o[k].abs().max()) > 1e-7}
    worst = max(bad, key=bad.get)
    assert bad[worst] < _tol(worst), (worst, bad[worst])


def test_fused_lowrank_path_matches_oracle():
    """mtl_session_set_flag("fuse_lowrank", 1): every projection pair linear_b(linear_a(x)) (modules/common_layers.py:
    287-289,303) and its input gradient run as ONE kernel (q | k | v grouped into one launch; the rank-r tile never leaves
    the SM).  Same outputs and gradients as the two-launch default (kept as a measured A/B variant)."""
    cfg = ref_asr.SMALL
    batch = ref_meta.synth_batch(cfg, 4, 41, 7, 10, lengths=[41, 30, 9, 5], tgt_lengths=[7, 5, 3, 1])
    p = ref_asr.init_params(cfg, 2)
    loss_o, g_o, gold_o, hyp_o, pred_o = ref_meta.loss_and_grads(p, cfg, batch)
    s = _session(cfg)
    s.set_flag("fuse_lowrank", 1)
    out, pred, grads = _fwd_bwd(s, p, batch)
    keep = gold_o != 0
    assert torch.equal(out["hyp"].cpu().long()[keep], hyp_o[keep])
    assert rel_err(pred, pred_o) < TOL_OUT
    bad = {k: rel_err(grads[k], g_o[k]) for k in g_o if float(g_o[k].abs().max()) > 1e-7}
    worst = max(bad, key=bad.get)
    assert bad[worst] < _tol(worst), (worst, bad[worst])


def test_fused_lowrank_path_equals_default_path_at_cfg2():
    """cfg-2 shapes (M = 200 / 264 rows, d = 512, rank 100), ragged batch: the fused pair path against the two-launch
    default on the same device -- the VGG forward is shared, so every ReLU / max-pool decision is too, and all 190
    gradients have to agree to accumulation-order noise."""
    cfg = ref_asr.CFG2
    batch = mg.cfg2_batch(3100, ragged=True)
    p = ref_asr.init_params(cfg, 31)
    res = []
    for fuse in (0, 1):
        s = _session(cfg)
        s.set_flag("fuse_lowrank", fuse)
        out, pred, grads = _fwd_bwd(s, p, batch)
        res.append((out["hyp"].cpu().clone(), pred.cpu().clone(), {k: v.cpu().clone() for k, v in grads.items()}))
    assert torch.equal(res[0][0], res[1][0])
    tol = {0: 1e-5, 1: 5e-3, 2: 2e-5}[GEMM_MODE]
    assert rel_err(res[1][1], res[0][1]) < tol
    bad = {k: rel_err(res[1][2][k], res[0][2][k]) for k in res[0][2] if float(res[0][2][k].abs().max()) > 1e-7}
    worst = max(bad, key=bad.get)
    assert bad[worst] < 20 * tol, (worst, bad[worst])


def test_precision_policy_classes():
    """mtl_session_set_op_mode: the default policy runs the VGG input / weight gradients in single-pass TF32; forcing them
    back to 3xTF32 or everything to TF32 changes the arithmetic (different bits) but stays inside the per-mode bounds."""
    cfg = ref_asr.SMALL
    p = ref_asr.init_params(cfg, 2)
    batch = ref_meta.synth_batch(cfg, 4, 41, 7, 10)
    _, g_o, *_ = ref_meta.loss_and_grads(p, cfg, batch)
    res = {}
    for name, over in (("default", {}), ("all3x", {"conv_dgrad": 2, "conv_wgrad": 2}), ("convfwd_tf32", {"conv_fwd": 1})):
        s = _session(cfg)
        for k, v in over.items():
            s.set_op_mode(k, v)
        _, pred, grads = _fwd_bwd(s, p, batch)
        res[name] = (pred.clone(), {k: v.clone() for k, v in grads.items()})
    if GEMM_MODE == 2:
        assert not torch.equal(res["default"][1]["conv.2.weight"], res["all3x"][1]["conv.2.weight"])
        assert rel_err(res["default"][0], res["all3x"][0]) < 1e-5                 # the forward is untouched by the policy
        assert rel_err(res["default"][0], res["convfwd_tf32"][0]) > 1e-6          # ... and a forward class changes it
    for name, (pred, grads) in res.items():
        for k in g_o:
            if float(g_o[k].abs().max()) > 1e-7:
                tol = 0.3 if name == "convfwd_tf32" else (5e-2 if k.startswith("conv.") else TOL_GRAD)   # TF32 forward: ReLU flips
                assert rel_err(grads[k], g_o[k]) < tol, (name, k)
    with pytest.raises(mtl_b200.MtlError):
        _session(cfg).set_op_mode(3, 7)
    with pytest.raises(mtl_b200.MtlError):
        _session(cfg).set_flag("no_such_flag", 1)


def test_fused_pooling_epilogue_in_the_engine():
    """MTL_CONV_POOL_FUSE=1: conv.2 / conv.7 write their max-pooled outputs from their own epilogues (no maxpool2_fwd launch);
    the switch is read once per process, so the SMALL forward / backward parity test runs in a fresh one."""
    import subprocess
    import sys
    env = dict(os.environ, MTL_CONV_POOL_FUSE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-k",
                        "small_fwd_bwd_vs_oracle or cfg2_fwd_bwd_golden"], capture_output=True, text=True, env=env, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-1500:] + r.stderr[-500:]
