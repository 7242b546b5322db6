"""``torch.optim``-compatible optimizers whose state lives in flat arenas next to the model's parameter arena.

``ArenaSGD`` / ``ArenaAdam`` subclass ``torch.optim.SGD`` / ``Adam`` so that everything the reference does with
its optimizers keeps working -- ``param_groups[0]['lr']`` (trainer ``get_lr``), ``state_dict()`` /
``load_state_dict()`` and pickling the optimizer object into a checkpoint (utils/functions.py:117-126,183-186)
-- while ``step()`` is one fused kernel over the whole arena (csrc/arena.cu) instead of a 190-tensor loop.
The per-parameter ``exp_avg`` / ``exp_avg_sq`` entries of ``state`` are views into the moment arenas."""
from __future__ import annotations

import torch


def _arena_model(params_owner):
    theta, grad = params_owner.arenas()
    return params_owner.session, theta, grad


class ArenaSGD(torch.optim.SGD):
    """Plain SGD (no momentum / weight decay: trainer/asr/transient_trainer.py:106) on the parameter arena."""

    def __init__(self, model, lr):
        super().__init__(model.parameters(), lr=lr)
        self._model = model

    @torch.no_grad()
    def step(self, closure=None):
        s, theta, grad = _arena_model(self._model)
        if not self._model.grads_pending():          # torch semantics: parameters without a gradient are skipped
            return
        self._model._adopt_foreign_grads()
        s.sgd(theta, grad, float(self.param_groups[0]['lr']))


class ArenaAdam(torch.optim.Adam):
    """Adam with torch defaults (betas (0.9, 0.999), eps 1e-8: transient_trainer.py:109, joint_trainer.py:124)."""

    def __init__(self, model, lr):
        super().__init__(model.parameters(), lr=lr)
        self._model = model
        s, theta, _ = _arena_model(model)
        self.m, self.v = s.new_arena(), s.new_arena()
        self.dev_state = s.new_adam_state()          # {int step, float step_size, float bc2_sqrt, pad} on the device
        self._bind_state()

    def _bind_state(self):
        s = self._model.session
        mv, vv = s.views(self.m), s.views(self.v)
        step = self.dev_state[0:1]
        for (name, *_), p in zip(s.table, self._model._params):
            self.state[p] = {"step": step.float().cpu().reshape(()), "exp_avg": mv[name], "exp_avg_sq": vv[name]}

    @property
    def step_count(self) -> int:
        return int(self.dev_state[0].item())

    @torch.no_grad()
    def step(self, closure=None):
        s, theta, grad = _arena_model(self._model)
        if not self._model.grads_pending():
            return
        self._model._adopt_foreign_grads()
        g = self.param_groups[0]
        s.adam(theta, grad, self.m, self.v, self.dev_state, float(g['lr']), g['betas'][0], g['betas'][1], g['eps'])

    def _refresh_steps(self):
        if hasattr(self, "dev_state"):               # (an unpickled copy only carries the base-class state)
            n = float(self.step_count)
            for st in self.state.values():
                st["step"] = torch.tensor(n)

    def state_dict(self):
        self._refresh_steps()
        return super().state_dict()

    def __getstate__(self):                          # checkpoints pickle the optimizer OBJECT (utils/functions.py:117-126)
        self._refresh_steps()
        return super().__getstate__()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)          # casts / copies the loaded tensors
        loaded = [self.state[p] for p in self._model._params if p in self.state]
        if loaded:
            s = self._model.session
            mv, vv = s.views(self.m), s.views(self.v)
            for (name, *_), st in zip(s.table, loaded):
                mv[name].copy_(st["exp_avg"])
                vv[name].copy_(st["exp_avg_sq"])
            self.dev_state[0] = int(float(loaded[0]["step"]))
        self._bind_state()


def _cpu(obj):
    if torch.is_tensor(obj):
        return obj.detach().cpu().clone()
    if isinstance(obj, dict):
        return {k: _cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_cpu(v) for v in obj)
    return obj


def to_torch(opt, host_params=None):
    """A plain ``torch.optim.SGD`` / ``Adam`` on the host carrying ``opt``'s hyper-parameters and state: what goes
    into a checkpoint, so that the reference's ``load_meta_model`` (utils/functions.py:183-186: it only calls
    ``.state_dict()`` on the pickled object) can read files written here without this package being importable.
    ``host_params``: host tensors in ``model.parameters()`` order to build the optimizer over (shared storage with
    the checkpoint's model_state_dict, so the pickle stores them once)."""
    if not isinstance(opt, (ArenaSGD, ArenaAdam)):
        return opt
    if host_params is None:
        host_params = [p.detach().cpu().clone() for p in opt._model._params]
    params = [torch.nn.Parameter(t, requires_grad=True) for t in host_params]
    g = opt.param_groups[0]
    if isinstance(opt, ArenaSGD):
        new = torch.optim.SGD(params, lr=g['lr'])
    else:
        new = torch.optim.Adam(params, lr=g['lr'], betas=tuple(g['betas']), eps=g['eps'])
    sd = _cpu(opt.state_dict())
    if len(sd["state"]):
        new.load_state_dict(sd)
    else:
        new.param_groups[0]['lr'] = g['lr']
    return new


def adopt(opt, model, kind):
    """Returns an arena optimizer equivalent to ``opt`` (a torch.optim.SGD / Adam over model.parameters(), e.g.
    one restored by a reference-style ``load_meta_model``); arena optimizers pass through.  Options the arena
    kernels do not implement are refused instead of being dropped silently."""
    if isinstance(opt, (ArenaSGD, ArenaAdam)):
        return opt
    if len(opt.param_groups) != 1:
        raise NotImplementedError("arena optimizers take one parameter group")
    g = opt.param_groups[0]
    if g.get('weight_decay', 0) != 0:
        raise NotImplementedError("weight_decay is not implemented by the arena optimizers")
    if kind == "sgd":
        if g.get('momentum', 0) != 0 or g.get('nesterov', False) or g.get('dampening', 0) != 0:
            raise NotImplementedError("ArenaSGD is plain SGD (transient_trainer.py:106): no momentum / nesterov")
        return ArenaSGD(model, g['lr'])
    if g.get('amsgrad', False) or g.get('maximize', False):
        raise NotImplementedError("ArenaAdam implements torch.optim.Adam without amsgrad / maximize")
    new = ArenaAdam(model, g['lr'])
    new.param_groups[0]['betas'] = tuple(g['betas'])
    new.param_groups[0]['eps'] = g['eps']
    if len(opt.state):
        new.load_state_dict(opt.state_dict())
    return new
