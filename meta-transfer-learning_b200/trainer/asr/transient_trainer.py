"""``TransientTrainer``: the first-order-MAML ("meta-transfer") training loop with the reference's signature,
logging and checkpoint behaviour (trainer/asr/transient_trainer.py:18-377), on the B200 engine.

One iteration = one meta-step over N = len(train_data_list) tasks:

    cg <- 0
    for task i:   g_i  <- grad CE(theta; train_i)                     inner forward/backward   (:188-199)
                  th_i <- theta - lr * [clip] g_i                      inner SGD step           (:205-207)
                  g_i  += grad [CE(th_i; val) / N]                     outer pass at th_i; NO zero_grad in
                                                                       between (:198-199,226-227), so the
                                                                       train gradient stays in the sum
                  cg   += g_i                                          add_copy_grad            (:229)
    grad <- cg ; [clip] ; theta <- Adam(theta, grad)                   (:248-255)

with the validation batch of the LAST manifest shared by every task (:168-169).  The reference runs this as 2N
autograd passes with N+1 full state_dict copies and ~2*B*L host syncs per pass; here the whole task loop is
ONE call (``mtl_meta_tasks``): tasks run concurrently on per-lane streams with their own adapted-weight arenas,
replayed from a CUDA graph once shapes repeat, and the host reads one small result block per iteration
(losses + the top-1 / gold indices the CER strings are built from).

Deliberate deviations from the reference, all on paths it cannot survive itself:
  * ``is_copy_grad=False`` raises: the reference branch (:231,250) dies with an autograd in-place error on
    torch >= 1.5 and then retries forever;
  * the blanket ``except Exception`` retry loop (:364-377) is narrowed to CUDA out-of-memory, at most
    ``MAX_RETRIES`` times in a row -- any other error propagates;
  * ``args.cuda`` must be set: the reference trainer itself only works with it (:210-215) and the engine has no
    CPU path.
When ``torch.distributed`` is initialised the tasks are sharded rank = task % world and the flat copy_grad
arena is all-reduced once per iteration (every rank then takes the identical Adam step)."""
import logging
import threading
import time
from collections import deque

import torch
from tqdm import tqdm

from utils.functions import post_process, save_meta_model
from utils.metrics import calculate_cer, calculate_metrics, calculate_wer

MAX_RETRIES = 3


def _strings(vocab, ids):
    return ["".join(vocab.id2label[int(x)] for x in row) for row in ids]


def _cer_counts(vocab, hyp_ids, gold_ids):
    """Sum of character edit distances and gold characters over a batch (transient_trainer.py:29-35,54-64)."""
    total_cer = total_char = 0
    for hyp, gold in zip(_strings(vocab, hyp_ids), _strings(vocab, gold_ids)):
        hyp = post_process(hyp, vocab.special_token_list)
        gold = post_process(gold, vocab.special_token_list)
        total_cer += calculate_cer(hyp.replace(' ', ''), gold.replace(' ', ''))
        total_char += len(gold.replace(' ', ''))
    return total_cer, total_char


class TransientTrainer():
    """
    Trainer class
    """

    def __init__(self):
        logging.info("Transient Trainer is initialized")

    def forward_one_batch(self, model, vocab, src, trg, src_percentages, src_lengths, trg_lengths, smoothing, loss_type,
                          verbose=False):
        """One forward pass -> (loss, summed CER, gold characters).  ``loss`` is an autograd scalar wired to the
        fused backward, so ``loss.backward()`` works as in the reference (used by validation and by callers that
        drive the model through autograd)."""
        model.label_smoothing = float(smoothing)
        pred, gold, hyp = model(src, src_lengths, trg, verbose=False)
        src_percentages.mul_(int(pred.size(1)))          # side effect kept (:39); only CTC consumed the result
        loss, _ = calculate_metrics(pred, gold, vocab.PAD_ID, input_lengths=None, target_lengths=trg_lengths,
                                    smoothing=smoothing, loss_type=loss_type)
        if loss.item() == float('Inf'):
            logging.info("Found infinity loss, masking")
            print("Found infinity loss, masking")
            loss = torch.where(loss != loss, torch.zeros_like(loss), loss)
        total_cer, total_char = _cer_counts(vocab, hyp.cpu().tolist(), gold.cpu().tolist())
        if verbose:
            print('Total CER', total_cer)
            print('Total char', total_char)
        return loss, total_cer, total_char

    def get_lr(self, optimizer):
        for param_group in optimizer.param_groups:
            return param_group['lr']

    def train(self, model, vocab, train_data_list, valid_loader_list, loss_type, start_it, num_it, args, inner_opt=None,
              outer_opt=None, evaluate_every=1000, window_size=100, last_summary_every=1000, last_metrics=None,
              early_stop=10, cpu_state_dict=False, is_copy_grad=False):
        """
        args:
            model: models.asr.transformer.Transformer on a CUDA device
            train_data_list: one K-shot sampler per task (``.sample(k_train, k_valid, manifest_id)``)
            valid_loader_list: validation loaders (utils.data_loader.AudioDataLoader)
            start_it / num_it: first / last iteration; last_metrics: (if resume)
        """
        from mtl_b200 import MetaStepper
        from mtl_b200.optim import ArenaAdam, ArenaSGD, adopt
        from mtl_b200.shard import MetaExchange, dist_env, reduce_stats, step_seed, task_shard
        if loss_type != "ce":
            raise NotImplementedError("only the cross-entropy loss is on the B200 hot path")
        if not is_copy_grad:
            raise NotImplementedError("run with --copy-grad: the reference's non-copy-grad branch "
                                      "(transient_trainer.py:231,250) fails on torch >= 1.5")
        if not args.cuda:
            raise RuntimeError("run with --cuda: TransientTrainer needs it in the reference too "
                               "(transient_trainer.py:210-215) and the engine has no CPU path")
        history = []
        best_valid_val = 1000000000
        smoothing = args.label_smoothing
        early_stop_criteria, early_stop_val = early_stop.split(",")[0], int(early_stop.split(",")[1])
        count_stop = 0
        logging.info("name " + args.name)
        total_time = 0
        logging.info("TRAIN")
        print("TRAIN")
        model.train()
        model.label_smoothing = float(smoothing)
        session = model.session
        theta, grad = model.arenas()
        model._direct_grads = True                   # this trainer writes the gradient arena itself (no autograd)

        inner_opt = ArenaSGD(model, args.lr) if inner_opt is None else adopt(inner_opt, model, "sgd")
        outer_opt = ArenaAdam(model, args.meta_lr) if outer_opt is None else adopt(outer_opt, model, "adam")

        last_sum_loss = deque(maxlen=window_size)
        last_sum_cer = deque(maxlen=window_size)
        last_sum_char = deque(maxlen=window_size)

        k_train, k_valid = args.k_train, args.k_valid
        n_tasks = len(train_data_list)
        dist, rank, world = dist_env()
        my_tasks = task_shard(n_tasks, rank, world)
        stepper = MetaStepper(session, max(1, len(my_tasks))) if my_tasks else None
        exchange = MetaExchange(session, dist)                        # the one exchange step of the meta-step + Adam
        model.zero_copy_grad()
        copy_grad = model._cg

        # one sampling thread ahead of the step, as in the reference (:120-139)
        buffers = [[] for _ in range(n_tasks)]

        def _pin(batch):
            # host batches go up asynchronously from pinned memory (the copy into the static slots never blocks the host);
            # device-computed features (utils/data_loader.py: feature_device) are already there
            return tuple(t.pin_memory() if (torch.is_tensor(t) and not t.is_cuda) else t for t in batch)

        def fetch(buf):
            for manifest_id in range(n_tasks):
                tr, va = train_data_list[manifest_id].sample(k_train, k_valid, manifest_id)
                buf[manifest_id].insert(0, (_pin(tr), _pin(va)))

        prefetch = threading.Thread(target=fetch, args=(buffers,))
        prefetch.start()

        it, retries = start_it, 0
        host_res = torch.zeros(max(1, len(my_tasks)), 16).pin_memory()
        while it < num_it:
            prefetch.join()
            prefetch = threading.Thread(target=fetch, args=(buffers,))
            prefetch.start()
            try:
                start_time = time.time()
                # the validation shots of the LAST manifest serve every task (:168-169)
                _, val_data = buffers[-1][-1]
                batches = [buffers[m].pop() for m in range(n_tasks)]
                val_inputs, val_input_sizes, _, val_targets, _ = val_data

                if stepper is not None:
                    for slot, m in enumerate(my_tasks):
                        (tr_inputs, tr_input_sizes, _, tr_targets, _), _ = batches[m]
                        stepper.load_task(slot, tr_inputs, tr_input_sizes, tr_targets)
                    stepper.load_val(val_inputs, val_input_sizes, val_targets)
                    stepper.run(theta, copy_grad, self.get_lr(inner_opt), 1.0 / n_tasks, clip=args.clip,
                                max_norm=args.max_norm, dropout=float(model.encoder.dropout_rate),
                                smoothing=float(smoothing), seed=step_seed(it, rank, world))
                    host_res.copy_(stepper.results, non_blocking=True)
                else:
                    session.zero(copy_grad)
                g = outer_opt.param_groups[0]
                exchange.ran_tasks = stepper is not None
                exchange.finish(theta, grad, copy_grad, outer_opt.m, outer_opt.v, outer_opt.dev_state, g['lr'],
                                clip=args.clip, max_norm=args.max_norm, betas=g['betas'], eps=g['eps'])

                # the only host read-back of the iteration: losses + train indices for the CER strings
                total_loss, total_cer, total_char = 0.0, 0, 0
                if stepper is not None:
                    ids = [tuple(t.cpu() for t in stepper.train_outputs(slot)) for slot in range(len(my_tasks))]
                    torch.cuda.current_stream().synchronize()
                    total_loss = float(host_res[:len(my_tasks), 8].sum())
                    for hyp, gold in ids:
                        c, n = _cer_counts(vocab, hyp.tolist(), gold.tolist())
                        total_cer, total_char = total_cer + c, total_char + n
                total_loss, total_cer, total_char = reduce_stats((total_loss, total_cer, total_char), session.device, dist)

                last_sum_cer.append(total_cer)
                last_sum_char.append(total_char)
                last_sum_loss.append(total_loss / n_tasks)
                total_time += time.time() - start_time
                retries = 0

                msg = "(Iteration {}) TRAIN LOSS:{:.4f} CER:{:.2f}% LR:{:.7f} TOTAL TIME:{:.7f}".format(
                    (it + 1), total_loss / n_tasks, total_cer * 100 / max(1, total_char), self.get_lr(outer_opt), total_time)
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
                        save_meta_model(model, vocab, (it + 1), inner_opt, outer_opt, metrics, args, best_model=False)
                    key = "avg_valid_cer" if early_stop_criteria == "cer" else "avg_valid_loss"
                    print("CRITERIA: CER" if early_stop_criteria == "cer" else "CRITERIA: LOSS")
                    if best_valid_val > metrics[key]:
                        count_stop = 0
                        best_valid_val = metrics[key]
                        save_meta_model(model, vocab, (it + 1), inner_opt, outer_opt, metrics, args, best_model=True)
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
        return inner_opt, outer_opt

    def _validate(self, model, vocab, valid_loader_list, it, args, smoothing, loss_type, history):
        """Periodic validation (transient_trainer.py:280-331): eval mode, no dropout, per-loader loss and CER."""
        print("")
        logging.info("VALID")
        model.eval()
        losses, cers = [], []
        with torch.no_grad():
            for ind, loader in enumerate(valid_loader_list):
                tot_loss, tot_cer, tot_char = 0.0, 0, 0
                bar = tqdm(iter(loader), leave=True, total=len(loader))
                for i, (src, trg, src_percentages, src_lengths, trg_lengths) in enumerate(bar):
                    loss, cer, n_char = self.forward_one_batch(model, vocab, src, trg, src_percentages, src_lengths,
                                                               trg_lengths, smoothing, loss_type)
                    tot_cer, tot_char, tot_loss = tot_cer + cer, tot_char + n_char, tot_loss + loss.item()
                    bar.set_description("(Iteration {}) VALID SET {} LOSS:{:.4f} CER:{:.2f}%".format(
                        (it + 1), ind, tot_loss / (i + 1), tot_cer * 100 / max(1, tot_char)))
                losses.append(tot_loss / len(loader))
                cers.append(tot_cer * 100 / max(1, tot_char))
                msg = "(Iteration {}) VALID SET {} LOSS:{:.4f} CER:{:.2f}%".format((it + 1), ind, losses[-1], cers[-1])
                print(msg)
                logging.info(msg)
        metrics = {"avg_valid_loss": sum(losses) / len(losses), "avg_valid_cer": sum(cers) / len(cers),
                   "valid_loss": losses, "valid_cer": cers, "history": history}
        history.append(metrics)
        msg = "(Iteration {}) AVG VALID LOSS:{:.4f} AVG CER:{:.2f}%".format(
            (it + 1), metrics["avg_valid_loss"], metrics["avg_valid_cer"])
        print(msg)
        logging.info(msg)
        return metrics
