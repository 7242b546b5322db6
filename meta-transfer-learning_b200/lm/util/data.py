"""Token-block data of the LM meta-loop (lm/util/data.py:12-150): ``Dictionary``, ``Corpus`` (one sentence per line,
``<eos>`` appended, ids in reading order) and ``LMDataset`` -- per task a (len // batch_size, batch_size) column-major
block matrix (``batchify``) from which iteration ``i`` takes the bptt-block ``i`` for training and block ``i + 1`` for
meta-validation (``sample``).  Host-side, vectorised; nothing here touches the device except ``batchify(.., cuda=True)``."""
from __future__ import annotations

import os
import random

import torch


class Dictionary(object):
    def __init__(self):
        self.word2idx = {}
        self.idx2word = {}

    def add_word(self, word):
        if word not in self.word2idx:
            self.idx2word[len(self.idx2word)] = word
            self.word2idx[word] = len(self.idx2word) - 1
        return self.word2idx[word]

    def __len__(self):
        return len(self.idx2word)


class Corpus(object):
    """lm/util/data.py:83-150: train / valid / test files -> LongTensors of word ids over one shared Dictionary."""

    def __init__(self, train_path, valid_path=None, test_path=None, dictionary=None, seed=1000):
        random.seed(seed)
        self.dictionary = Dictionary() if dictionary is None else dictionary
        self.train = self.tokenize(train_path)
        self.valid = self.tokenize(valid_path) if valid_path is not None else None
        self.test = self.tokenize(test_path) if test_path is not None else None

    def tokenize(self, path):
        assert os.path.exists(path), path
        ids = []
        with open(path, 'r', encoding='utf-8') as f:
            for line in f:
                words = line.strip().replace("  ", " ").split() + ['<eos>']
                ids.extend(self.dictionary.add_word(w) for w in words)
        return torch.tensor(ids, dtype=torch.long)


class LMDataset(object):
    def __init__(self, task_list, args):
        self.bptt = args.bptt
        self.batch_size = args.batch_size
        self.args = args
        self.task_list = [self.batchify(t, self.batch_size) for t in task_list]

    def batchify(self, data, bsz):
        nbatch = data.size(0) // bsz
        data = data.narrow(0, 0, nbatch * bsz)
        data = data.view(bsz, -1).t().contiguous()
        if getattr(self.args, "cuda", False):
            data = data.cuda()
        return data

    def get_batch(self, source, i, evaluation=False):
        seq_len = min(self.bptt, len(source) - 1 - i)
        data = source[i:i + seq_len]
        target = source[i + 1:i + 1 + seq_len].reshape(-1)
        return data, target

    def sample(self, manifest_id, i):
        """(train inputs, train targets, val inputs, val targets) of iteration i (lm/util/data.py:46-67): the bptt-aligned
        block i (mod corpus length) and the one after it."""
        ids = self.task_list[manifest_id]
        tr_pos, val_pos = (i * self.bptt) % len(ids), ((i + 1) * self.bptt) % len(ids)
        tr_ids, val_ids = tr_pos - tr_pos % self.bptt, val_pos - val_pos % self.bptt
        tr_src, tr_target = self.get_batch(ids, tr_ids)
        val_src, val_target = self.get_batch(ids, val_ids)
        return (tr_src, tr_target, val_src, val_target)
