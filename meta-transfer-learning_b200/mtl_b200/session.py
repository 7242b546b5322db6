"""Torch-tensor front of the C engine: owns the C session, the workspace and pointer plumbing."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import lib as _l
from .spec import ModelSpec, param_specs


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def positional_table(max_len: int, d: int) -> torch.Tensor:
    """modules/common_layers.py:93-99 -- the ``pe`` buffer, shape (1, max_len, d)."""
    import math
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len).unsqueeze(1).float()
    div = torch.exp(torch.arange(0, d, 2).float() * -(math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


@dataclass
class Batch:
    """One (inputs, input_sizes, targets) batch on the device (utils/data_loader.py:245-321)."""
    x: torch.Tensor        # (B,1,F,T) fp32
    lens: torch.Tensor     # (B,) int32 raw frame counts
    trg: torch.Tensor      # (B,L) int64, PAD=0
    n: int                 # 1 + max non-PAD target length (Decoder.preprocess, decoder.py:55-69)

    @staticmethod
    def from_host(x, lens, trg, device, non_blocking=True) -> "Batch":
        """Host tensors -> device batch; n is derived on the host so the step needs no sync."""
        n = int((trg != 0).sum(dim=1).max().item()) + 1
        return Batch(x.to(device=device, dtype=torch.float32, non_blocking=non_blocking).contiguous(),
                     lens.to(device=device, dtype=torch.int32, non_blocking=non_blocking).contiguous(),
                     trg.to(device=device, dtype=torch.int64, non_blocking=non_blocking).contiguous(), n)

    @property
    def B(self):
        return self.x.shape[0]


class Session:
    """A model layout + workspace bound to one CUDA device."""

    def __init__(self, spec: ModelSpec, device="cuda", gemm_mode: int = 2):
        if not torch.cuda.is_available():
            raise _l.MtlError("libmtl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _l.get_lib()
        self.spec = spec
        self.device = torch.device(device)
        cfg = _l.ModelCfg(spec.n_enc, spec.n_dec, spec.d_model, spec.n_heads, spec.d_k, spec.d_v,
                          spec.d_inner, spec.rank, spec.vocab, spec.n_freq)
        h = C.c_void_p()
        _l.check(self.lib.mtl_session_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.set_gemm_mode(gemm_mode)
        self.n_floats = int(self.lib.mtl_param_arena_floats(h))
        self.specs = param_specs(spec)
        assert self.lib.mtl_param_count(h) == len(self.specs)
        self.table = []
        off, num = C.c_longlong(), C.c_longlong()
        for i, (name, shape) in enumerate(self.specs):
            _l.check(self.lib.mtl_param_info(h, i, C.byref(off), C.byref(num)))
            n = 1
            for s in shape:
                n *= s
            assert n == num.value, (name, shape, num.value)
            self.table.append((name, shape, off.value, n))
        self.pe_enc = positional_table(spec.src_max_len, spec.d_model)[0].contiguous().to(self.device)
        self.pe_dec = positional_table(spec.tgt_max_len, spec.d_model)[0].contiguous().to(self.device)
        self._ws = None
        self._scratch = torch.zeros(1040, dtype=torch.float32, device=self.device)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.mtl_session_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ arenas
    def set_gemm_mode(self, mode: int):
        _l.check(self.lib.mtl_session_set_gemm_mode(self._h, int(mode)))
        self.gemm_mode = int(mode)

    OP_CLASSES = ("conv_fwd", "conv_dgrad", "conv_wgrad", "lin_fwd", "lin_dgrad", "lin_wgrad", "stem", "vocab", "attn")

    def set_op_mode(self, op_class, mode: int):
        """Per-operation-class engine (include/mtl_b200.h: mtl_session_set_op_mode); mode -1 = follow gemm_mode."""
        idx = self.OP_CLASSES.index(op_class) if isinstance(op_class, str) else int(op_class)
        _l.check(self.lib.mtl_session_set_op_mode(self._h, idx, int(mode)))

    def set_flag(self, name: str, value: int):
        _l.check(self.lib.mtl_session_set_flag(self._h, name.encode(), int(value)))

    def new_arena(self) -> torch.Tensor:
        return torch.zeros(self.n_floats, dtype=torch.float32, device=self.device)

    def views(self, arena: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {name: arena[off:off + n].view(shape) for name, shape, off, n in self.table}

    def load(self, arena: torch.Tensor, params: Dict[str, torch.Tensor]):
        v = self.views(arena)
        for name, t in params.items():
            v[name].copy_(t.to(self.device))

    def workspace(self, B: int, T: int, n: int) -> torch.Tensor:
        need = int(self.lib.mtl_workspace_bytes(self._h, B, T, n))
        if need < 0:
            _l.check(-2)
        need += (1024 + 8) * 4 + 512          # clip scratch tail used by mtl_meta_task
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need + (64 << 20), dtype=torch.uint8, device=self.device)
        return self._ws

    def _cbatch(self, b: Batch, hyp=None, gold=None, ce=None) -> _l.CBatch:
        assert b.x.is_cuda and b.x.is_contiguous() and b.x.dtype == torch.float32
        assert b.lens.dtype == torch.int32 and b.trg.dtype == torch.int64 and b.trg.is_contiguous()
        B, _, F, T = b.x.shape
        assert F == self.spec.n_freq, "input frequency bins != model n_freq"
        cb = _l.CBatch()
        cb.x, cb.lens, cb.trg = b.x.data_ptr(), b.lens.data_ptr(), (b.trg.data_ptr() if b.trg.numel() else None)
        cb.B, cb.T, cb.L, cb.n = B, T, b.trg.shape[1], b.n
        cb.hyp_out = hyp.data_ptr() if hyp is not None else None
        cb.gold_out = gold.data_ptr() if gold is not None else None
        cb.ce_out = ce.data_ptr() if ce is not None else None
        return cb

    # ------------------------------------------------------------------ passes
    def forward(self, theta: torch.Tensor, b: Batch, dropout: float = 0.0, seed: int = 0,
                smoothing: float = 0.0):
        """Transformer.forward + CE.  Returns dict(pred (B,n,V) view into the workspace -- valid until the
        next pass --, gold, hyp (B,n) int32, ce (8,) fp32 = [loss, n_valid, n_correct, ...])."""
        B, n = b.B, b.n
        ws = self.workspace(B, b.x.shape[3], n)
        hyp = torch.empty(B * n, dtype=torch.int32, device=self.device)
        gold = torch.empty(B * n, dtype=torch.int32, device=self.device)
        ce = torch.empty(8, dtype=torch.float32, device=self.device)
        cb = self._cbatch(b, hyp, gold, ce)
        pred_p, ldp = C.c_void_p(), C.c_int()
        _l.check(self.lib.mtl_asr_forward(self._h, _ptr(theta), _ptr(self.pe_enc), _ptr(self.pe_dec), _ptr(ws),
                                          ws.numel(), C.byref(cb), float(dropout), int(seed), float(smoothing),
                                          _stream(), C.byref(pred_p), C.byref(ldp)))
        self._live = (b, theta)               # keep the inputs alive until backward() has been enqueued
        off = pred_p.value - ws.data_ptr()
        flat = ws[off:off + B * n * ldp.value * 4].view(torch.float32)
        pred = flat.view(B, n, ldp.value)[:, :, :self.spec.vocab]
        return dict(pred=pred, gold=gold.view(B, n), hyp=hyp.view(B, n), ce=ce)

    # ------------------------------------------------------------------ inference
    def encode(self, theta: torch.Tensor, x: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
        """Transformer.encode (models/asr/transformer.py:78-98): (B,1,F,T) spectrograms + raw frame counts ->
        encoder output (B, T', d_model) with T' = (T // 2) // 2."""
        B, _, F, T = x.shape
        need = int(self.lib.mtl_encode_workspace_bytes(self._h, B, T))
        if need < 0:
            _l.check(-2)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need + (64 << 20), dtype=torch.uint8, device=self.device)
        b = Batch(x, lens, torch.zeros(B, 0, dtype=torch.int64, device=self.device), 0)
        cb = self._cbatch(b)
        out = torch.empty(B, (T // 2) // 2, self.spec.d_model, dtype=torch.float32, device=self.device)
        _l.check(self.lib.mtl_asr_encode(self._h, _ptr(theta), _ptr(self.pe_enc), _ptr(self._ws), self._ws.numel(),
                                         C.byref(cb), _ptr(out), _stream()))
        self._live = (b, theta)
        return out

    def greedy(self, theta: torch.Tensor, enc_out: torch.Tensor, start_token: int, max_steps: int = 300) -> torch.Tensor:
        """Decoder.greedy_search (modules/decoder.py:131-184) -> (B, max_steps) int32 arg-max tokens (not yet cut at EOS)."""
        enc_out = enc_out.to(device=self.device, dtype=torch.float32).contiguous()
        B, Tp, d = enc_out.shape
        assert d == self.spec.d_model
        if max_steps + 1 > self.pe_dec.shape[0]:
            raise ValueError(f"max_steps {max_steps} exceeds the decoder positional table ({self.pe_dec.shape[0]} rows)")
        need = int(self.lib.mtl_greedy_workspace_bytes(self._h, B, Tp, int(max_steps)))
        if need < 0:
            _l.check(-2)
        ws = torch.empty(need + 4096, dtype=torch.uint8, device=self.device)
        out = torch.empty(B, max_steps, dtype=torch.int32, device=self.device)
        _l.check(self.lib.mtl_asr_greedy(self._h, _ptr(theta), _ptr(self.pe_dec), _ptr(ws), ws.numel(), _ptr(enc_out),
                                         B, Tp, int(start_token), int(max_steps), _ptr(out), _stream()))
        self._live_greedy = (ws, enc_out, theta)      # keep the buffers alive until the stream has consumed them
        return out

    def backward(self, theta: torch.Tensor, grad: torch.Tensor, scale: float = 1.0,
                 dpred: Optional[torch.Tensor] = None):
        """Accumulates d(scale*CE)/dtheta (or the vjp of ``dpred``) of the last forward into ``grad``."""
        ld = 0
        if dpred is not None:
            dpred = dpred.contiguous().float()
            ld = dpred.shape[-1]
        _l.check(self.lib.mtl_asr_backward(self._h, _ptr(theta), _ptr(grad), float(scale), _ptr(dpred), ld,
                                           _stream()))

    def meta_task(self, theta, theta0, grad, copy_grad, train: Batch, val: Batch, lr: float, val_scale: float,
                  clip: bool = False, max_norm: float = 400.0, dropout: float = 0.0, smoothing: float = 0.0,
                  seed: int = 0, results: Optional[torch.Tensor] = None, out=None):
        """transient_trainer.py:178-237 for one task (see include/mtl_b200.h: mtl_meta_task)."""
        T = max(train.x.shape[3], val.x.shape[3])
        ws = self.workspace(max(train.B, val.B), T, max(train.n, val.n))
        hp = _l.MetaHParams(float(lr), float(val_scale), int(bool(clip)), float(max_norm), float(dropout),
                            float(smoothing), int(seed))
        o = out or {}
        ctr = self._cbatch(train, o.get("tr_hyp"), o.get("tr_gold"))
        cva = self._cbatch(val, o.get("val_hyp"), o.get("val_gold"))
        _l.check(self.lib.mtl_meta_task(self._h, _ptr(theta), _ptr(theta0), _ptr(grad), _ptr(copy_grad),
                                        _ptr(self.pe_enc), _ptr(self.pe_dec), _ptr(ws), ws.numel(),
                                        C.byref(ctr), C.byref(cva), C.byref(hp), _ptr(results), _stream()))

    def region_a_floats(self) -> int:
        """Leading floats of the arena (everything but the VGG front-end) whose copy_grad is final early (mtl_b200.h)."""
        return int(self.lib.mtl_region_a_floats(self._h))

    def wait_region_a(self, stream: torch.cuda.Stream):
        """`stream` waits until region A of copy_grad of the latest MetaStepper.run holds every task's contribution."""
        _l.check(self.lib.mtl_stream_wait_region_a(self._h, C.c_void_p(stream.cuda_stream)))

    def graph_stats(self):
        cap, rep = C.c_ulonglong(), C.c_ulonglong()
        _l.check(self.lib.mtl_graph_stats(self._h, C.byref(cap), C.byref(rep)))
        return int(cap.value), int(rep.value)

    def meta_finish(self, theta, grad, copy_grad, adam_m, adam_v, adam_state, meta_lr: float, clip: bool = False,
                    max_norm: float = 400.0, betas=(0.9, 0.999), eps: float = 1e-8):
        _l.check(self.lib.mtl_meta_finish(_ptr(theta), _ptr(grad), _ptr(copy_grad), _ptr(adam_m), _ptr(adam_v),
                                          _ptr(adam_state), float(meta_lr), float(betas[0]), float(betas[1]),
                                          float(eps), int(bool(clip)), float(max_norm),
                                          _ptr(self._scratch), self.n_floats, _stream()))

    def new_adam_state(self) -> torch.Tensor:
        """16-byte device block {int step; float step_size; float bc2_sqrt; pad}."""
        return torch.zeros(4, dtype=torch.int32, device=self.device)

    # arena helpers
    def zero(self, a):
        _l.check(self.lib.mtl_arena_zero(_ptr(a), a.numel(), _stream()))

    def copy(self, dst, src):
        _l.check(self.lib.mtl_arena_copy(_ptr(dst), _ptr(src), dst.numel(), _stream()))

    def axpy(self, y, x, a: float):
        _l.check(self.lib.mtl_arena_axpy(_ptr(y), _ptr(x), float(a), y.numel(), _stream()))

    def sgd(self, p, g, lr: float):
        _l.check(self.lib.mtl_arena_sgd(_ptr(p), _ptr(g), float(lr), p.numel(), _stream()))

    def clip(self, g, max_norm: float):
        _l.check(self.lib.mtl_arena_clip(_ptr(g), g.numel(), float(max_norm), _ptr(self._scratch), _stream()))
        return self._scratch[1024:1026]

    def adam(self, p, g, m, v, state, lr: float, b1=0.9, b2=0.999, eps=1e-8):
        _l.check(self.lib.mtl_arena_adam(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(state), float(lr), b1, b2, eps,
                                         p.numel(), _stream()))


class MetaStepper:
    """All tasks of one meta-step through ``mtl_meta_tasks``: tasks run concurrently on per-lane streams
    and, once the same shapes have been seen twice, the whole step replays from a CUDA graph.

    Owns what a graph bakes in: the per-lane adapted-weight / gradient arenas and workspaces, the device seed
    slot, the results block and STATIC input slots (one set per distinct batch shape) into which every step's
    batches are copied, so pointers never change between steps."""

    def __init__(self, session: Session, n_tasks: int, n_lanes: Optional[int] = None, use_graph: bool = True):
        self.s = session
        self.n_tasks = int(n_tasks)
        self.n_lanes = int(n_lanes or min(self.n_tasks, 8))
        self.use_graph = bool(use_graph)
        dev = session.device
        self.lane_theta = [session.new_arena() for _ in range(self.n_lanes)]
        self.lane_grad = [session.new_arena() for _ in range(self.n_lanes)]
        self.lane_ws = [None] * self.n_lanes
        self.seed_slot = torch.zeros(1, dtype=torch.int64, device=dev)
        self.results = torch.zeros(self.n_tasks, 16, dtype=torch.float32, device=dev)
        self._slots = {}
        self.train = [None] * self.n_tasks
        self.val = None

    # ------------------------------------------------------------------ static input slots
    def _slot(self, who, B, F, T, L):
        """ONE set of static device buffers per input role (`who`), sized by capacity: SpectrogramDataset.sample pads to
        the batch's own longest utterance / transcript, so (T, L) change almost every iteration of a real run.  The
        batch of this step is a contiguous PREFIX view of the capacity buffers (the engine takes the logical (B, T, L)
        from mtl_batch), so the base pointers stay put while shapes vary; capacities grow geometrically (rounded up to
        64 frames / 16 tokens) and the outgrown buffers are released."""
        need_x, need_t = B * F * T, max(1, B * L)
        cur = self._slots.get(who)
        if cur is None or cur[0].numel() < need_x or cur[2].numel() < need_t or cur[1].numel() < B:
            dev = self.s.device
            cap_x = B * F * ((T + 63) // 64 * 64)
            cap_t = B * ((L + 15) // 16 * 16 + 1)
            if cur is not None:
                cap_x = max(cap_x, min(2 * cur[0].numel(), 2 * need_x))
                cap_t = max(cap_t, min(2 * cur[2].numel(), 2 * need_t))
                self._slots[who] = cur = None
            self._slots[who] = cur = (torch.zeros(cap_x, dtype=torch.float32, device=dev),
                                      torch.zeros(B, dtype=torch.int32, device=dev),
                                      torch.zeros(cap_t, dtype=torch.int64, device=dev),
                                      torch.zeros(cap_t + B, dtype=torch.int32, device=dev),     # hyp
                                      torch.zeros(cap_t + B, dtype=torch.int32, device=dev))     # gold
        bx, bl, bt, hyp, gold = cur
        return bx[:need_x].view(B, 1, F, T), bl[:B], bt[:B * L].view(B, L), hyp, gold

    def _stage(self, who, x, lens, trg, n):
        B, _, F, T = x.shape
        L = trg.shape[1]
        sx, sl, st, hyp, gold = self._slot(who, B, F, T, L)
        sx.copy_(x, non_blocking=True)
        sl.copy_(lens, non_blocking=True)
        if L:
            st.copy_(trg, non_blocking=True)
        if n is None:
            n = int((trg != 0).sum(dim=1).max().item()) + 1      # host tensors: no device sync
        return Batch(sx, sl, st, n), hyp, gold

    def load_task(self, t: int, x, lens, trg, n: Optional[int] = None):
        """Copies task t's training batch (host pinned or device tensors) into its static slot."""
        self.train[t] = self._stage(("tr", t), x, lens, trg, n)

    def load_val(self, x, lens, trg, n: Optional[int] = None):
        self.val = self._stage(("val",), x, lens, trg, n)

    # ------------------------------------------------------------------ the step
    def run(self, theta, copy_grad, lr: float, val_scale: float, clip: bool = False, max_norm: float = 400.0,
            dropout: float = 0.0, smoothing: float = 0.0, seed: int = 0):
        """copy_grad <- sum over tasks of [grad(train; theta) + val_scale * grad(val; theta - lr*grad_train)];
        self.results[t] = [train CE block (8), val CE block (8)].  theta is not modified."""
        s = self.s
        B = max(max(b.B for b, _, _ in self.train), self.val[0].B)
        T = max(max(b.x.shape[3] for b, _, _ in self.train), self.val[0].x.shape[3])
        n = max(max(b.n for b, _, _ in self.train), self.val[0].n)
        need = int(s.lib.mtl_workspace_bytes(s._h, B, T, n))
        if need < 0:
            _l.check(-2)
        need += (1024 + 8) * 4 + 512
        for l in range(self.n_lanes):
            if self.lane_ws[l] is None or self.lane_ws[l].numel() < need:
                self.lane_ws[l] = None
                self.lane_ws[l] = torch.empty(need + (16 << 20), dtype=torch.uint8, device=s.device)
        ctr = (_l.CBatch * self.n_tasks)()
        for t, (b, hyp, gold) in enumerate(self.train):
            ctr[t] = s._cbatch(b, hyp, gold)
        cva = s._cbatch(self.val[0], self.val[1], self.val[2])
        lanes = (_l.CLane * self.n_lanes)()
        for l in range(self.n_lanes):
            lanes[l].theta = self.lane_theta[l].data_ptr()
            lanes[l].grad = self.lane_grad[l].data_ptr()
            lanes[l].workspace = self.lane_ws[l].data_ptr()
            lanes[l].workspace_bytes = self.lane_ws[l].numel()
        a = _l.MetaStepArgs()
        a.theta, a.copy_grad = theta.data_ptr(), copy_grad.data_ptr()
        a.pe_enc, a.pe_dec = s.pe_enc.data_ptr(), s.pe_dec.data_ptr()
        a.n_tasks, a.train, a.val = self.n_tasks, ctr, C.pointer(cva)
        a.n_lanes, a.lanes = self.n_lanes, lanes
        a.hp = _l.MetaHParams(float(lr), float(val_scale), int(bool(clip)), float(max_norm), float(dropout),
                              float(smoothing), int(seed))
        a.results = self.results.data_ptr()
        a.seed_slot = self.seed_slot.data_ptr()
        a.use_graph = int(self.use_graph)
        _l.check(s.lib.mtl_meta_tasks(s._h, C.byref(a), _stream()))
        return self.results

    def train_outputs(self, t: int):
        """(hyp, gold) int32 (B, n) of task t's TRAINING pass (what the trainer's CER is computed from)."""
        b, hyp, gold = self.train[t]
        return hyp[:b.B * b.n].view(b.B, b.n), gold[:b.B * b.n].view(b.B, b.n)
