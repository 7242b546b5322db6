"""CLI of the LM meta-loop with the reference script's flags (lm/main_meta_transfer.py:20-54) on the B200 engine.

The reference reads ./data/{seame,cv,hkust}_{train,valid,test,dev}.txt (lm/main_meta_transfer.py:119-128), files its
repository does not ship; ``--data-dir`` points at a directory that holds them.  Tasks, in the script's order:
CV, HKUST, SEAME (the last one also provides the shared meta-validation block and the validation / test streams)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lm.meta import LMMetaTrainer  # noqa: E402
from lm.model.rnn_model import RNNModel  # noqa: E402
from lm.util import data  # noqa: E402


def build_parser():
    p = argparse.ArgumentParser(description='SEAME RNN/LSTM language model, meta-transfer training (B200 engine)')
    p.add_argument('--name', type=str, default='')
    p.add_argument('--model', type=str, default='LSTM')
    p.add_argument('--emsize', type=int, default=200)
    p.add_argument('--nhid', type=int, default=200)
    p.add_argument('--nlayers', type=int, default=2)
    p.add_argument('--lr', type=float, default=20)
    p.add_argument('--meta_lr_factor', type=float, default=3)
    p.add_argument('--clip', type=float, default=0.25)
    p.add_argument('--epochs', type=int, default=40)
    p.add_argument('--batch_size', type=int, default=20, metavar='N')
    p.add_argument('--bptt', type=int, default=35)
    p.add_argument('--dropout', type=float, default=0.2)
    p.add_argument('--ratio', type=float, default=0.8)
    p.add_argument('--tied', action='store_true')
    p.add_argument('--pad', action='store_true')
    p.add_argument('--seed', type=int, default=1111)
    p.add_argument('--cuda', action='store_true')
    p.add_argument('--log_path', type=str, default='./log')
    p.add_argument('--log-interval', type=int, default=200, metavar='N')
    p.add_argument('--save', type=str, default='./model')
    p.add_argument('--data-dir', type=str, default='./data', help='directory with the nine corpus text files')
    p.add_argument('--iterations', type=int, default=1000000, help='upper iteration limit (the reference uses 1000000)')
    p.add_argument('--valid-interval', type=int, default=600)
    return p


def batchify(t, bsz, cuda):
    nbatch = t.size(0) // bsz
    t = t.narrow(0, 0, nbatch * bsz).view(bsz, -1).t().contiguous()
    return t.cuda() if cuda else t


def main(argv=None):
    args = build_parser().parse_args(argv)
    torch.manual_seed(args.seed)
    if not args.cuda:
        raise SystemExit("this implementation runs on the GPU only: pass --cuda")
    d = args.data_dir
    seame = data.Corpus(os.path.join(d, "seame_train.txt"), os.path.join(d, "seame_valid.txt"),
                        os.path.join(d, "seame_test.txt"), None, args.seed)
    cv = data.Corpus(os.path.join(d, "cv_train.txt"), os.path.join(d, "cv_valid.txt"), os.path.join(d, "cv_test.txt"),
                     seame.dictionary, args.seed)
    hkust = data.Corpus(os.path.join(d, "hkust_train.txt"), None, os.path.join(d, "hkust_dev.txt"), cv.dictionary, args.seed)
    dictionary = hkust.dictionary
    print("vocab:", len(dictionary))
    lm_dataset = data.LMDataset([cv.train, hkust.train, seame.train], args)
    seame_val = batchify(seame.valid, 10, True)
    seame_test = batchify(seame.test, 10, True)
    model = RNNModel(args.model, len(dictionary), args.emsize, args.nhid, args.nlayers, args.dropout, args.tied).cuda()
    print(model)
    trainer = LMMetaTrainer(model, args)
    os.makedirs(args.save, exist_ok=True)
    save_path = os.path.join(args.save, (args.name or "lm_meta") + ".pt")
    print("############# TRAIN data #############")
    trainer.train(lm_dataset, seame_val, 0, args.iterations, args.log_interval, args.valid_interval, seame_test, save_path)
    print('| End of training | SEAME test loss {:5.2f}'.format(trainer.evaluate(seame_test)))


if __name__ == "__main__":
    main()
