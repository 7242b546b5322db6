"""Token-block data of the LM meta-loop (API of lm/util/data.py:12-150): ``Dictionary``, ``Corpus`` (one sentence per
line, ``<eos>`` appended, ids in reading order) and ``LMDataset`` -- per task a (len // batch_size, batch_size)
column-major block matrix from which iteration ``i`` takes bptt-block ``i`` for training and block ``i + 1`` for
meta-validation.  Host-side and vectorised; nothing here touches the device except ``batchify`` with ``args.cuda``."""
from __future__ import annotations

import os
import random

import torch


class Dictionary(object):
    """word <-> id tables in first-seen order (lm/util/data.py:69-81); ``idx2word`` is indexable by id."""

    def __init__(self):
        self.word2idx, self.idx2word = {}, []

    def add_word(self, word):
        idx = self.word2idx.get(word)
        if idx is None:
            idx = self.word2idx[word] = len(self.idx2word)
            self.idx2word.append(word)
        return idx

    def __len__(self):
        return len(self.idx2word)


class Corpus(object):
    """train / valid / test text files -> 1-D LongTensors of word ids over one shared Dictionary (lm/util/data.py:83-150)."""

    def __init__(self, train_path, valid_path=None, test_path=None, dictionary=None, seed=1000):
        random.seed(seed)
        self.dictionary = dictionary if dictionary is not None else Dictionary()
        self.train = self.tokenize(train_path)
        self.valid = None if valid_path is None else self.tokenize(valid_path)
        self.test = None if test_path is None else self.tokenize(test_path)

    def tokenize(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        add = self.dictionary.add_word
        with open(path, encoding="utf-8") as f:
            ids = [add(w) for line in f for w in line.strip().replace("  ", " ").split() + ["<eos>"]]
        return torch.tensor(ids, dtype=torch.long)


class LMDataset(object):
    """``task_list[k]``: task k's token stream folded into ``batch_size`` parallel columns (lm/util/data.py:12-34)."""

    def __init__(self, task_list, args):
        self.args = args
        self.bptt, self.batch_size = args.bptt, args.batch_size
        self.task_list = [self.batchify(stream, self.batch_size) for stream in task_list]

    def batchify(self, data, bsz):
        rows = data.size(0) // bsz                                 # the remainder that does not fill a row is dropped
        cols = data[:rows * bsz].view(bsz, rows).t().contiguous()  # column c = the c-th contiguous slice of the stream
        return cols.cuda() if getattr(self.args, "cuda", False) else cols

    def get_batch(self, source, i, evaluation=False):
        """(inputs (n, B), targets (n * B,)) starting at row i; n = bptt, or what is left before the last row."""
        n = min(self.bptt, len(source) - 1 - i)
        return source[i:i + n], source[i + 1:i + 1 + n].reshape(-1)

    def _block_start(self, block, n_rows):
        pos = (block * self.bptt) % n_rows
        return pos - pos % self.bptt                               # aligned down to a bptt boundary (lm/util/data.py:60-61)

    def sample(self, manifest_id, i):
        """(train inputs, train targets, val inputs, val targets) of iteration i: blocks i and i + 1 (lm/util/data.py:46-67)."""
        ids = self.task_list[manifest_id]
        train = self.get_batch(ids, self._block_start(i, len(ids)))
        val = self.get_batch(ids, self._block_start(i + 1, len(ids)))
        return train + val
