"""Learning-rate wrappers with the reference's names (utils/optimizer.py:3-50).  The meta / joint trainers use
plain SGD / Adam (trainer/asr/transient_trainer.py:105-109, joint_trainer.py:123-130); these wrappers only
have to exist for ``utils.functions.init_optimizer`` callers."""


class NoamOpt:
    """lr = max(min_lr, factor * d^-0.5 * min(step^-0.5, step * warmup^-1.5)), applied before every step."""

    def __init__(self, model_size, factor, warmup, optimizer, min_lr=1e-5):
        self.optimizer, self.model_size, self.factor, self.warmup, self.min_lr = optimizer, model_size, factor, warmup, min_lr
        self._step, self._rate = 0, 0

    def rate(self, step=None):
        step = self._step if step is None else step
        return max(self.min_lr, self.factor * self.model_size ** -0.5 * min(step ** -0.5, step * self.warmup ** -1.5))

    def step(self):
        self._step += 1
        self._rate = self.rate()
        for group in self.optimizer.param_groups:
            group['lr'] = self._rate
        self.optimizer.step()

    def zero_grad(self):
        self.optimizer.zero_grad()


class AnnealingOpt:
    """Divides the learning rate of the first param group by lr_anneal on every ``step`` call."""

    def __init__(self, lr, lr_anneal, optimizer):
        self.optimizer, self.lr, self.lr_anneal = optimizer, lr, lr_anneal

    def step(self):
        state = self.optimizer.state_dict()
        state['param_groups'][0]['lr'] = state['param_groups'][0]['lr'] / self.lr_anneal
        self.optimizer.load_state_dict(state)
