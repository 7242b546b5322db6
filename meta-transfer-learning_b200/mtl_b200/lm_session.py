"""Torch-tensor front of the LSTM language-model entry points of libmtl_b200 (mtl_lm_*): parameter arena layout,
workspace, one forward(+backward) pass and the whole first-order meta-step of lm/main_meta_transfer.py:277-372."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib as _l


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@dataclass(frozen=True)
class LmSpec:
    """RNNModel('LSTM', vocab, ninp, nhid, nlayers) -- lm/model/rnn_model.py:15-46."""
    vocab: int
    ninp: int = 200
    nhid: int = 200
    nlayers: int = 2


def lm_param_specs(spec: LmSpec) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) in ``RNNModel.parameters()`` order (encoder, rnn layer by layer, decoder)."""
    out = [("encoder.weight", (spec.vocab, spec.ninp))]
    for l in range(spec.nlayers):
        n_in = spec.ninp if l == 0 else spec.nhid
        out += [(f"rnn.weight_ih_l{l}", (4 * spec.nhid, n_in)), (f"rnn.weight_hh_l{l}", (4 * spec.nhid, spec.nhid)),
                (f"rnn.bias_ih_l{l}", (4 * spec.nhid,)), (f"rnn.bias_hh_l{l}", (4 * spec.nhid,))]
    out += [("decoder.weight", (spec.vocab, spec.nhid)), ("decoder.bias", (spec.vocab,))]
    return out


class LmSession:
    """Layout + workspace of one LSTM LM on one CUDA device.  There is no CPU path."""

    def __init__(self, spec: LmSpec, device="cuda", gemm_mode: int = 2):
        if not torch.cuda.is_available():
            raise _l.MtlError("libmtl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _l.get_lib()
        self.spec = spec
        self.device = torch.device(device)
        self.gemm_mode = int(gemm_mode)
        self.cfg = _l.LmCfg(spec.vocab, spec.ninp, spec.nhid, spec.nlayers)
        self.n_floats = int(self.lib.mtl_lm_param_floats(C.byref(self.cfg)))
        if self.n_floats <= 0:
            raise _l.MtlError(f"unsupported LM configuration {spec}: {self.lib.mtl_last_error().decode()}")
        self.specs = lm_param_specs(spec)
        assert self.lib.mtl_lm_param_count(C.byref(self.cfg)) == len(self.specs)
        self.table = []
        off, num = C.c_longlong(), C.c_longlong()
        for i, (name, shape) in enumerate(self.specs):
            _l.check(self.lib.mtl_lm_param_info(C.byref(self.cfg), i, C.byref(off), C.byref(num)))
            n = 1
            for s in shape:
                n *= s
            assert n == num.value, (name, shape, num.value)
            self.table.append((name, shape, off.value, n))
        self._ws = None
        self._ws_key = None
        self._scratch = torch.zeros(1032 + 64, device=self.device)

    # ------------------------------------------------------------------ arenas
    def new_arena(self) -> torch.Tensor:
        return torch.zeros(self.n_floats, device=self.device, dtype=torch.float32)

    def views(self, arena: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {name: arena[off:off + n].view(shape) for name, shape, off, n in self.table}

    def load(self, arena: torch.Tensor, params: Dict[str, torch.Tensor]) -> None:
        v = self.views(arena)
        for name, *_ in self.table:
            v[name].copy_(params[name].to(self.device, torch.float32))

    def new_hidden(self, bsz: int):
        z = torch.zeros(self.spec.nlayers, bsz, self.spec.nhid, device=self.device)
        return z, z.clone()

    def _workspace(self, T: int, B: int) -> torch.Tensor:
        if self._ws_key != (T, B):
            need = int(self.lib.mtl_lm_workspace_bytes(C.byref(self.cfg), T, B))
            if need <= 0:
                raise _l.MtlError("mtl_lm_workspace_bytes failed")
            if self._ws is None or self._ws.numel() < need + 256:
                self._ws = torch.empty(need + 256, device=self.device, dtype=torch.uint8)
            self._ws_key = (T, B)
        return self._ws

    def _ws_ptr(self, ws):
        p = ws.data_ptr()
        a = (p + 255) & ~255
        return C.c_void_p(a), ws.numel() - (a - p)

    # ------------------------------------------------------------------ one pass
    def run(self, theta, tokens, targets=None, hidden=None, grad=None, dropout=0.0, seed=0, scale=1.0,
            want_logits=False, hidden_out=True):
        """RNNModel.forward (+ CE + backward into ``grad`` when given).  tokens (T, B) int64, targets (T*B,) int64.
        Returns dict(loss block (8,), logits (T, B, V) or None, hidden (h, c) or None)."""
        tokens = tokens.to(self.device, torch.int64).contiguous()
        T, B = tokens.shape
        if targets is not None:
            targets = targets.to(self.device, torch.int64).contiguous().view(-1)
            assert targets.numel() == T * B
        ws = self._workspace(T, B)
        wp, wbytes = self._ws_ptr(ws)
        h0 = c0 = None
        if hidden is not None:
            h0, c0 = (t.to(self.device, torch.float32).contiguous() for t in hidden)
        hT = cT = None
        if hidden_out:
            hT, cT = self.new_hidden(B)
        loss = torch.zeros(8, device=self.device) if targets is not None else None
        logits = torch.empty(T, B, self.spec.vocab, device=self.device) if want_logits else None
        _l.check(self.lib.mtl_lm_pass(C.byref(self.cfg), self.gemm_mode, _ptr(theta), _ptr(grad), _ptr(tokens), _ptr(targets),
                                      T, B, _ptr(h0), _ptr(c0), _ptr(hT), _ptr(cT), float(dropout), int(seed), float(scale),
                                      wp, wbytes, _ptr(loss), _ptr(logits), _stream()))
        return {"loss": loss, "logits": logits, "hidden": (hT, cT) if hidden_out else None}

    # ------------------------------------------------------------------ meta-step
    def _enqueue_meta(self, theta, theta_work, grad, meta_grad, hidden, toks, trgs, vt, vy, weights, lr, meta_lr_factor, clip,
                      dropout, seed, seed_slot, results):
        n = len(toks)
        T, B = vt.shape
        ws = self._workspace(T, B)
        wp, wbytes = self._ws_ptr(ws)
        tok_arr = (C.c_void_p * n)(*[t.data_ptr() for t in toks])
        trg_arr = (C.c_void_p * n)(*[t.data_ptr() for t in trgs])
        w_arr = (C.c_float * n)(*[float(w) for w in weights])
        _l.check(self.lib.mtl_lm_meta_step(C.byref(self.cfg), self.gemm_mode, _ptr(theta), _ptr(theta_work), _ptr(grad),
                                           _ptr(meta_grad), _ptr(hidden[0]), _ptr(hidden[1]), n, tok_arr, trg_arr, _ptr(vt),
                                           _ptr(vy), T, B, w_arr, float(lr), float(meta_lr_factor), float(clip),
                                           float(dropout), int(seed), _ptr(seed_slot), wp, wbytes, _ptr(results),
                                           _ptr(self._scratch), _stream()))

    def meta_step(self, theta, theta_work, grad, meta_grad, hidden, train: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                  val: Tuple[torch.Tensor, torch.Tensor], weights: Sequence[float], lr: float, meta_lr_factor: float,
                  clip: float, dropout: float, seed: int, results: Optional[torch.Tensor] = None, graph: bool = False):
        """One iteration of lm/main_meta_transfer.py:293-372 on the device, without a host sync.  ``hidden`` (h, c) is
        updated in place; ``results`` (n_tasks, 16) receives the train / val loss blocks.  graph=True: the token blocks are
        copied into static device buffers and the step (about a thousand small kernels) is replayed from a CUDA graph
        captured at the second call with the same shapes and hyper-parameters; the dropout seed lives in a device word."""
        n = len(train)
        T, B = val[0].shape
        if not all(tuple(t.shape) == (T, B) for t, _ in train):
            # blocks at the end of a corpus are shorter than bptt (lm/util/data.py:37): compose the same iteration from
            # single passes and arena kernels (still no host sync)
            return self._meta_step_ragged(theta, theta_work, grad, meta_grad, hidden, train, val, weights, lr, meta_lr_factor,
                                          clip, dropout, seed, results)
        if not graph:
            toks = [t.to(self.device, torch.int64).contiguous() for t, _ in train]
            trgs = [y.to(self.device, torch.int64).contiguous().view(-1) for _, y in train]
            vt = val[0].to(self.device, torch.int64).contiguous()
            vy = val[1].to(self.device, torch.int64).contiguous().view(-1)
            self._enqueue_meta(theta, theta_work, grad, meta_grad, hidden, toks, trgs, vt, vy, weights, lr, meta_lr_factor, clip,
                               dropout, seed, None, results)
            self._keep = (toks, trgs, vt, vy)      # alive until the stream has consumed them
            return
        key = (n, T, B, theta.data_ptr(), theta_work.data_ptr(), grad.data_ptr(), meta_grad.data_ptr(), hidden[0].data_ptr(),
               hidden[1].data_ptr(), tuple(float(w) for w in weights), float(lr), float(meta_lr_factor), float(clip),
               float(dropout), None if results is None else results.data_ptr())
        g = getattr(self, "_graph", None)
        if g is None or g["key"] != key:
            mk = lambda *shape: torch.zeros(*shape, device=self.device, dtype=torch.int64)
            g = {"key": key, "seen": 0, "graph": None, "toks": [mk(T, B) for _ in range(n)], "trgs": [mk(T * B) for _ in range(n)],
                 "vt": mk(T, B), "vy": mk(T * B), "seed": torch.zeros(1, device=self.device, dtype=torch.int64)}
            self._graph = g
        for i, (t, y) in enumerate(train):
            g["toks"][i].copy_(t, non_blocking=True)
            g["trgs"][i].copy_(y.reshape(-1), non_blocking=True)
        g["vt"].copy_(val[0], non_blocking=True)
        g["vy"].copy_(val[1].reshape(-1), non_blocking=True)
        g["seed"].fill_(int(seed))
        run = lambda: self._enqueue_meta(theta, theta_work, grad, meta_grad, hidden, g["toks"], g["trgs"], g["vt"], g["vy"], weights,
                                         lr, meta_lr_factor, clip, dropout, 0, g["seed"], results)
        if g["graph"] is not None:
            g["graph"].replay()
        elif g["seen"] == 0:
            g["seen"] = 1
            run()                                  # first sighting: eager (warms every kernel / attribute / tensor map up)
        else:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(cur)
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(cg, stream=side):
                    run()
            cur.wait_stream(side)
            g["graph"] = cg
            cg.replay()                            # the capture itself executed nothing

    def _meta_step_ragged(self, theta, theta_work, grad, meta_grad, hidden, train, val, weights, lr, meta_lr_factor, clip,
                          dropout, seed, results):
        """mtl_lm_meta_step restated over mtl_lm_pass + mtl_arena_* for task blocks of different lengths."""
        n_f, st = self.n_floats, _stream
        lib = self.lib
        _l.check(lib.mtl_arena_zero(_ptr(meta_grad), n_f, st()))
        for i, (tok, trg) in enumerate(train):
            _l.check(lib.mtl_arena_copy(_ptr(theta_work), _ptr(theta), n_f, st()))
            _l.check(lib.mtl_arena_zero(_ptr(grad), n_f, st()))
            out = self.run(theta_work, tok, trg, hidden=hidden, grad=grad, dropout=dropout, seed=int(seed) * 128 + 2 * i)
            if results is not None:
                results[i, :8].copy_(out["loss"])
            if clip and clip > 0:
                _l.check(lib.mtl_arena_clip(_ptr(grad), n_f, float(clip), _ptr(self._scratch), st()))
            _l.check(lib.mtl_arena_sgd(_ptr(theta_work), _ptr(grad), float(lr) / float(meta_lr_factor), n_f, st()))
            hidden[0].copy_(out["hidden"][0])
            hidden[1].copy_(out["hidden"][1])
            # a val block of another batch width cannot take this hidden state; the reference would fail there as well
            out = self.run(theta_work, val[0], val[1], hidden=hidden, grad=meta_grad, dropout=dropout,
                           seed=int(seed) * 128 + 2 * i + 1, scale=float(weights[i]), hidden_out=False)
            if results is not None:
                results[i, 8:].copy_(out["loss"])
        if clip and clip > 0:
            _l.check(lib.mtl_arena_clip(_ptr(meta_grad), n_f, float(clip), _ptr(self._scratch), st()))
        _l.check(lib.mtl_arena_sgd(_ptr(theta), _ptr(meta_grad), float(lr), n_f, st()))
