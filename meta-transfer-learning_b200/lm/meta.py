"""The meta-loop of lm/main_meta_transfer.py:277-411 on the B200 engine.

One iteration = snapshot of the weights; for every task: train forward / backward from the running hidden state, clip,
inner SGD(lr / meta_lr_factor), meta-validation forward on the LAST task's next block with the train pass's hidden
state, loss weighted (1 - ratio) / 2 for the first two tasks and ratio for the third; weights reset; one backward of the
weighted sum, clip, outer SGD(lr).  The whole iteration is ONE call into libmtl_b200 (mtl_lm_meta_step) on flat
parameter arenas, without a host sync; the losses are read back only when a log line needs them.

Semantics note: the reference's ``batch_loss.backward()`` after ``load_state_dict`` is rejected
by autograd on torch >= 1.5 (in-place modified saved tensors); this trainer implements the first-order meta-gradient --
each task's validation gradient at its adapted weights -- which is what the ASR trainer of the same repository does."""
from __future__ import annotations

import math
import time

import torch


def task_weights(n_tasks: int, ratio: float):
    """lm/main_meta_transfer.py:346-349: (1 - ratio) / 2 for tasks 0 and 1, ratio from task 2 on."""
    return [(1.0 - ratio) / 2.0 if i < 2 else ratio for i in range(n_tasks)]


class LMMetaTrainer(object):
    def __init__(self, model, args):
        self.model = model
        self.args = args
        s = model.session
        self.theta = model.arena()
        self.theta_work, self.grad, self.meta_grad = s.new_arena(), s.new_arena(), s.new_arena()
        self.hidden = s.new_hidden(args.batch_size)            # model.init_hidden(batch_size), carried across iterations
        self.lr = args.lr
        self._results = None                                     # static loss-block buffer: part of the CUDA graph's identity
        self.use_graph = bool(getattr(args, "cuda_graph", True))   # replay the iteration from a CUDA graph (same shapes)

    def step(self, dataset, it: int, results: torch.Tensor = None):
        """One meta-iteration on ``dataset`` (an ``LMDataset``); returns the (n_tasks, 16) device loss blocks (a static
        buffer unless ``results`` is given: clone it to keep it past the next iteration)."""
        a, s = self.args, self.model.session
        n = len(dataset.task_list)
        _, _, val_x, val_y = dataset.sample(-1, it)
        train = []
        for i in range(n):
            tr_x, tr_y, _, _ = dataset.sample(i, it)
            train.append((tr_x, tr_y))
        if results is None:
            if self._results is None or self._results.shape[0] != n:
                self._results = torch.zeros(n, 16, device=s.device)
            results = self._results                               # same pointer every iteration, so the captured graph replays
        seed = (int(getattr(a, "seed", 0)) * 1000003 + it) & 0x7FFFFFFFFFFF
        s.meta_step(self.theta, self.theta_work, self.grad, self.meta_grad, self.hidden, train, (val_x, val_y),
                    task_weights(n, a.ratio), self.lr, a.meta_lr_factor, a.clip if a.clip else 0.0,
                    a.dropout if self.model.training else 0.0, seed, results, graph=self.use_graph)
        return results

    def evaluate(self, data_source, eval_batch_size=10):
        """lm/main_meta_transfer.py:217-266 without the prediction dump: mean token CE of a batchified stream."""
        a, s = self.args, self.model.session
        hidden = s.new_hidden(eval_batch_size)
        total = torch.zeros((), device=s.device, dtype=torch.float64)
        for i in range(0, data_source.size(0) - 1, a.bptt):
            seq_len = min(a.bptt, len(data_source) - 1 - i)
            data, targets = data_source[i:i + seq_len], data_source[i + 1:i + 1 + seq_len].reshape(-1)
            out = s.run(self.theta, data, targets, hidden=hidden)
            hidden = out["hidden"]
            total += seq_len * out["loss"][0].double()
        return float(total) / len(data_source)

    def train(self, dataset, val_data_source, start_it, num_it, log_interval, valid_interval, test_data_source=None,
              save_path=None):
        self.model.train()
        it, total, best_val, counter = start_it, 0.0, 0.0, 0
        n = len(dataset.task_list)
        w = torch.tensor(task_weights(n, self.args.ratio), device=self.model.session.device)
        pending = []
        start = time.time()
        while it < num_it:
            pending.append(self.step(dataset, it).clone())        # the static block is overwritten by the next iteration
            if it % log_interval == 0 and it > 0:
                total += float(sum((r[:, 8] * w).sum() for r in pending))         # batch_loss = sum_i w_i val_loss_i
                pending = []
                cur = total / (valid_interval if it % valid_interval == 0 else it % valid_interval)
                print('| it {:3d} | lr {:02.2f} | ms/batch {:5.2f} | word_loss {:5.2f} | avg ppl {:8.2f}'.format(
                    it, self.lr, (time.time() - start) * 1000 / log_interval, cur, math.exp(min(cur, 50.0))))
                start = time.time()
            if it % valid_interval == 0 and it > 0:
                val_loss = self.evaluate(val_data_source)
                print("it {} | val loss {:5f} | ppl {:5f}".format(it, val_loss, math.exp(val_loss)))
                if test_data_source is not None:
                    test_loss = self.evaluate(test_data_source)
                    print("it {} | test loss {:5f} | ppl {:5f}".format(it, test_loss, math.exp(test_loss)))
                self.model.train()
                if not best_val or val_loss < best_val:
                    if save_path:
                        torch.save(self.model.state_dict(), save_path)
                    best_val, counter = val_loss, 0
                else:
                    self.lr /= 4.0
                    counter += 1
                if counter == 5:
                    break
                total = 0.0
            it += 1
        return it
