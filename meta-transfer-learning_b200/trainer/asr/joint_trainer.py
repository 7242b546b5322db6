"""``JointTrainer``: the multi-task joint-training baseline with the reference's signature and logging
(trainer/asr/joint_trainer.py:18-404) on the B200 engine.

One iteration (joint_trainer.py:178-271, ``discriminator=None`` path):

    grad <- 0 ; for task i: grad += grad [CE(theta; train_i) / N] ; [clip] ; theta <- Adam(theta, grad)

Each task is one fused forward + one fused backward that accumulates straight into the flat gradient arena;
the Adam step is one kernel over the arena.  As in the reference the sampler is asked for one validation shot
per task (``k_valid = 1``, :152) which is then ignored.  The adversarial / multi-task discriminator branches
(:209-247) are outside the hot path and raise.  With ``torch.distributed`` initialised, tasks are sharded
rank = task % world and the gradient arena is all-reduced before the Adam step."""
import logging
import threading
import time
from collections import deque

import torch

from mtl_b200.session import Batch
from trainer.asr.transient_trainer import MAX_RETRIES, TransientTrainer, _cer_counts
from utils.functions import save_joint_model


class JointTrainer(TransientTrainer):
    """
    Trainer class
    """

    def __init__(self):
        logging.info("Joint Trainer is initialized")

    def forward_one_batch(self, model, vocab, src, trg, src_percentages, src_lengths, trg_lengths, smoothing, loss_type,
                          verbose=False, discriminator=None, accent_id=None, multi_task=False):
        if discriminator is not None:
            raise NotImplementedError("adversarial / multi-task training is out of scope of the B200 hot path")
        return super().forward_one_batch(model, vocab, src, trg, src_percentages, src_lengths, trg_lengths, smoothing,
                                         loss_type, verbose=verbose)

    def train(self, model, vocab, train_data_list, valid_loader_list, loss_type, start_it, num_it, args,
              evaluate_every=1000, window_size=100, last_summary_every=1000, last_metrics=None, early_stop=10,
              cpu_state_dict=False, is_copy_grad=False, opt_name="adam", discriminator=None):
        from mtl_b200.optim import ArenaAdam, ArenaSGD
        if discriminator is not None:
            raise NotImplementedError("adversarial / multi-task training is out of scope of the B200 hot path")
        if loss_type != "ce":
            raise NotImplementedError("only the cross-entropy loss is on the B200 hot path")
        history = []
        best_valid_val = 1000000000
        smoothing = args.label_smoothing
        early_stop_criteria, early_stop_val = early_stop.split(",")[0], int(early_stop.split(",")[1])
        count_stop = 0
        logging.info("name " + args.name)
        total_time = 0
        logging.info("TRAIN")
        print("TRAIN")
        if not next(model.parameters()).is_cuda:
            model = model.cuda()               # the reference trains on the CPU without --cuda; the engine cannot
        model.train()
        model.label_smoothing = float(smoothing)
        session = model.session
        theta, grad = model.arenas()
        model._direct_grads = True                   # this trainer writes the gradient arena itself (no autograd)
        if opt_name == "adam":
            opt = ArenaAdam(model, args.lr)
        elif opt_name == "sgd":
            opt = ArenaSGD(model, args.lr)
        else:
            raise ValueError("opt_name must be 'adam' or 'sgd'")

        last_sum_loss = deque(maxlen=window_size)
        last_sum_cer = deque(maxlen=window_size)
        last_sum_char = deque(maxlen=window_size)
        k_train = args.k_train
        n_tasks = len(train_data_list)
        from mtl_b200.shard import dist_env, exchange_copy_grad, reduce_stats, step_seed, task_shard
        dist, rank, world = dist_env()
        my_tasks = task_shard(n_tasks, rank, world)
        buffers = [[] for _ in range(n_tasks)]

        def fetch(buf):
            for manifest_id in range(n_tasks):
                buf[manifest_id].insert(0, train_data_list[manifest_id].sample(k_train, 1, manifest_id))

        prefetch = threading.Thread(target=fetch, args=(buffers,))
        prefetch.start()
        drop = float(model.encoder.dropout_rate)
        it, retries = start_it, 0
        while it < num_it:
            prefetch.join()
            prefetch = threading.Thread(target=fetch, args=(buffers,))
            prefetch.start()
            try:
                start_time = time.time()
                batches = [buffers[m].pop() for m in range(n_tasks)]
                session.zero(grad)                                            # opt.zero_grad()
                outs = []
                for m in my_tasks:
                    (tr_inputs, tr_input_sizes, _, tr_targets, _), _ = batches[m]
                    b = Batch.from_host(tr_inputs, tr_input_sizes, tr_targets, session.device)
                    out = session.forward(theta, b, dropout=drop, seed=step_seed(it * 64 + m), smoothing=float(smoothing))
                    session.backward(theta, grad, 1.0 / n_tasks)              # (tr_loss / N).backward()
                    outs.append(out)
                exchange_copy_grad(grad, dist)                                # one all-reduce of the flat gradient arena
                if args.clip:
                    session.clip(grad, args.max_norm)
                opt.step()

                total_loss, total_cer, total_char = 0.0, 0, 0
                for out in outs:                                              # host read-back after the step is queued
                    total_loss += float(out["ce"][0])
                    c, n = _cer_counts(vocab, out["hyp"].cpu().tolist(), out["gold"].cpu().tolist())
                    total_cer, total_char = total_cer + c, total_char + n
                total_loss, total_cer, total_char = reduce_stats((total_loss, total_cer, total_char), session.device, dist)
                last_sum_cer.append(total_cer)
                last_sum_char.append(total_char)
                last_sum_loss.append(total_loss)
                total_time += time.time() - start_time
                retries = 0
                msg = "(Iteration {}) TRAIN LOSS:{:.4f} CER:{:.2f}% LR:{:.7f} TOTAL TIME:{:.7f}".format(
                    (it + 1), total_loss / n_tasks, total_cer * 100 / max(1, total_char), self.get_lr(opt), total_time)
                print(msg)
                logging.info(msg)
                if (it + 1) % last_summary_every == 0:
                    msg = "(Summary Iteration {} | MA {}) TRAIN LOSS:{:.4f} CER:{:.2f}%".format(
                        (it + 1), window_size, sum(last_sum_loss) / len(last_sum_loss),
                        sum(last_sum_cer) * 100 / max(1, sum(last_sum_char)))
                    print(msg, flush=True)
                    logging.info(msg)

                if (it + 1) % evaluate_every == 0:
                    metrics = self._validate(model, vocab, valid_loader_list, it, args, smoothing, loss_type, history)
                    if (it + 1) % args.save_every == 0:
                        save_joint_model(model, vocab, (it + 1), opt, metrics, args, best_model=False)
                    key = "avg_valid_cer" if early_stop_criteria == "cer" else "avg_valid_loss"
                    print("CRITERIA: CER" if early_stop_criteria == "cer" else "CRITERIA: LOSS")
                    if best_valid_val > metrics[key]:
                        count_stop = 0
                        best_valid_val = metrics[key]
                        save_joint_model(model, vocab, (it + 1), opt, metrics, args, best_model=True)
                    else:
                        count_stop += 1
                        print("count_stop:", count_stop)
                    if count_stop >= early_stop_val:
                        logging.info("EARLY STOP")
                        print("EARLY STOP\n")
                        break
                    model.train()
                it += 1
            except KeyboardInterrupt:
                raise
            except torch.cuda.OutOfMemoryError as e:
                retries += 1
                print('Error: {}, fetching new data...'.format(e), flush=True)
                logging.info('Error: {}, fetching new data...'.format(e))
                torch.cuda.empty_cache()
                if retries > MAX_RETRIES:
                    raise
        prefetch.join()
        return opt
