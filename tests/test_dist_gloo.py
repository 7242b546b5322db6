"""CPU, world_size 2, gloo: the host logic of the multi-GPU meta-step (SURVEY 8e / DESIGN 7).

Every rank owns the tasks ``mtl_b200.shard.task_shard`` gives it, computes their copy_grad contribution
(here with the CPU oracle standing in for the CUDA engine -- the sharding / exchange code under test is the
product's), packs it into ONE flat fp32 arena, runs ``exchange_copy_grad`` and applies the Adam step.  Both
ranks must end with bit-identical parameters that match the single-process oracle meta-step."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flat(d, names):
    return torch.cat([d[k].reshape(-1) for k in names]).contiguous()


def _worker(rank, world, port, n_tasks, out_dir):
    for p in (os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    torch.set_num_threads(2)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mtl_b200.shard import dist_env, exchange_copy_grad, reduce_stats, task_shard
        from oracle import ref_asr, ref_meta
        cfg = ref_asr.SMALL
        params = ref_asr.init_params(cfg, 3)
        names = list(params)
        tasks = [ref_meta.synth_batch(cfg, 2, 41, 7, 100 + i) for i in range(n_tasks)]
        val = ref_meta.synth_batch(cfg, 2, 37, 6, 150)
        lr, meta_lr = 1e-2, 1e-3
        d, r, w = dist_env()
        assert (r, w) == (rank, world) and d is not None
        mine = task_shard(n_tasks, r, w)
        # local copy_grad: transient_trainer.py:178-237 for the owned tasks, theta reset after each
        cg = {k: torch.zeros_like(v) for k, v in params.items()}
        theta0 = {k: v.clone() for k, v in params.items()}
        loss_sum = 0.0
        for t in mine:
            _, g, *_ = ref_meta.loss_and_grads(params, cfg, tasks[t])
            ref_meta.sgd_step_(params, g, lr)
            lv, gv, *_ = ref_meta.loss_and_grads(params, cfg, val, 1.0 / n_tasks)
            loss_sum += lv
            for k in cg:
                cg[k] += g[k] + gv[k]
            for k in params:
                params[k].copy_(theta0[k])
        flat = _flat(cg, names)
        exchange_copy_grad(flat, d)                                   # the ONE collective of the step
        (loss_all,) = reduce_stats((loss_sum,), torch.device("cpu"), d)
        off = 0
        grads = {}
        for k in names:
            n = params[k].numel()
            grads[k] = flat[off:off + n].view_as(params[k]).clone()
            off += n
        ref_meta.adam_step_(params, grads, ref_meta.AdamState(), meta_lr)
        torch.save(dict(theta=_flat(params, names), cg=flat, loss=loss_all / n_tasks, mine=mine),
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_tasks", [3, 2])
def test_two_rank_meta_step_matches_single_process_oracle(tmp_path, n_tasks):
    sys.path[:0] = [p for p in (os.path.join(ROOT, "meta-transfer-learning_b200"), ROOT) if p not in sys.path]
    from oracle import ref_asr, ref_meta
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_tasks, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    assert sorted(r0["mine"] + r1["mine"]) == list(range(n_tasks))   # a partition of the tasks
    assert torch.equal(r0["theta"], r1["theta"]) and torch.equal(r0["cg"], r1["cg"])   # replicas stay bit-identical

    cfg = ref_asr.SMALL
    params = ref_asr.init_params(cfg, 3)
    names = list(params)
    tasks = [ref_meta.synth_batch(cfg, 2, 41, 7, 100 + i) for i in range(n_tasks)]
    val = ref_meta.synth_batch(cfg, 2, 37, 6, 150)
    ref = ref_meta.meta_step(params, ref_meta.AdamState(), cfg, tasks, val, lr=1e-2, meta_lr=1e-3)
    cg_ref = _flat(ref["copy_grad"], names)
    # the fp32 sum over tasks is re-associated by the all-reduce: ~1e-6 relative
    assert float((r0["cg"] - cg_ref).abs().max()) <= 2e-6 * float(cg_ref.abs().max())
    assert abs(r0["loss"] - ref["loss"]) <= 1e-6 * abs(ref["loss"])
    # Adam moves every entry by ~lr; entries with a solid gradient must land where the oracle lands
    solid = cg_ref.abs() > 1e-3 * float(cg_ref.abs().max())
    assert float((r0["theta"] - _flat(params, names)).abs()[solid].max()) <= 0.02 * 1e-3


def test_shard_helpers_reject_bad_input():
    sys.path[:0] = [p for p in (os.path.join(ROOT, "meta-transfer-learning_b200"),) if p not in sys.path]
    from mtl_b200.shard import exchange_copy_grad, task_shard
    assert task_shard(3, 0, 4) == [0] and task_shard(3, 3, 4) == [] and task_shard(8, 1, 4) == [1, 5]
    with pytest.raises(ValueError):
        task_shard(3, 2, 2)
    with pytest.raises(ValueError):
        exchange_copy_grad(torch.zeros(2, 2))
    t = torch.ones(4)
    assert exchange_copy_grad(t, None) is t
