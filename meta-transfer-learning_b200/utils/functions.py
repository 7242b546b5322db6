"""Model factory, parameter counting and the checkpoint wire format, with the reference's names and
signatures (utils/functions.py:16-364).

Checkpoints keep the reference's on-disk layout -- a pickled dict {vocab, args, epoch, model_state_dict,
inner_opt, outer_opt, metrics} with the optimizer OBJECTS inside (functions.py:117-126) -- and the
reference's ``state_dict`` keys, so files interchange in both directions.  ``torch.load`` needs
``weights_only=False`` for such pickles on torch >= 2.6."""
import logging
import math
import os

import numpy as np
import torch

from models.asr.transformer import Transformer
from modules import Decoder, Discriminator, Encoder
from utils.optimizer import AnnealingOpt, NoamOpt


def generate_labels(labels, special_token_list):
    """label <-> id maps: special tokens first, then unseen labels in order (functions.py:16-34)."""
    label2id, id2label = {}, {}
    for lab in list(special_token_list) + list(labels):
        if lab in label2id:
            if lab not in special_token_list:
                print("multiple label: ", lab)
            continue
        label2id[lab] = len(label2id)
        id2label[label2id[lab]] = lab
    return label2id, id2label


def compute_num_params(model):
    """(trainable, non-trainable) parameter counts (functions.py:36-41)."""
    train = sum(int(np.prod(p.shape)) for p in model.parameters() if p.requires_grad)
    frozen = sum(int(np.prod(p.shape)) for p in model.parameters() if not p.requires_grad)
    return train, frozen


def post_process(string, special_token_list):
    """Strips the special tokens and maps the sentencepiece space marker to ' ' (functions.py:360-364)."""
    for tok in special_token_list:
        string = string.replace(tok, "")
    return string.replace("▁", " ")


# ------------------------------------------------------------------ model factory (functions.py:307-351)
def init_transformer_model(args, vocab, train=True, is_factorized=False, r=100):
    """Builds Encoder / Decoder / Transformer from ``args``.  As in the reference, ``args.dim_input`` is
    OVERWRITTEN from the sample rate and window size: bins = floor(sr*win/2)+1, two floor-poolings, x128."""
    if args.feat_extractor == 'vgg_cnn':
        bins = int(math.floor((args.sample_rate * args.window_size) / 2) + 1)
        args.dim_input = int(math.floor(int(math.floor(bins) / 2) / 2)) * 128
        if getattr(args, "feat", "spectrogram") == "logfbank":
            args.dim_input = 2560
    else:
        raise NotImplementedError(f"feat_extractor={args.feat_extractor!r} is outside the B200 hot path (vgg_cnn only)")
    encoder = Encoder(args.num_enc_layers, num_heads=args.num_heads, dim_model=args.dim_model, dim_key=args.dim_key,
                      dim_value=args.dim_value, dim_input=args.dim_input, dim_inner=args.dim_inner,
                      src_max_length=args.src_max_len, dropout=args.dropout, is_factorized=is_factorized, r=r)
    decoder = Decoder(vocab, num_layers=args.num_dec_layers, num_heads=args.num_heads, dim_emb=args.dim_emb,
                      dim_model=args.dim_model, dim_inner=args.dim_inner, dim_key=args.dim_key,
                      dim_value=args.dim_value, trg_max_length=args.tgt_max_len, dropout=args.dropout,
                      emb_trg_sharing=args.emb_trg_sharing, is_factorized=is_factorized, r=r)
    return Transformer(encoder, decoder, vocab, feat_extractor=args.feat_extractor, train=train)


def init_discriminator_model(args):
    return Discriminator(args.dim_model, args.num_class)


def init_optimizer(args, model, opt_type="noam"):
    """Noam-scheduled Adam or annealed Nesterov SGD (functions.py:289-305); unused by the meta / joint trainers."""
    if opt_type == "noam":
        return NoamOpt(args.dim_input, args.k_lr, args.warmup,
                       torch.optim.Adam(model.parameters(), betas=(0.9, 0.98), eps=1e-9), min_lr=args.min_lr)
    if opt_type == "sgd":
        return AnnealingOpt(args.lr, args.lr_anneal,
                            torch.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum, nesterov=True))
    print("Optimizer is not defined")
    return None


# ------------------------------------------------------------------ checkpoints
def _ckpt_path(args, epoch, best_model):
    folder = "{}/{}".format(args.save_folder, args.name)
    os.makedirs(folder, exist_ok=True)
    return "{}/best_model.th".format(folder) if best_model else "{}/epoch_{}.th".format(folder, epoch)


def _host_state_dict(model):
    """state_dict with every tensor materialised on its own storage (arena views would drag the whole arena)."""
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def save_meta_model(model, vocab, epoch, inner_opt, outer_opt, metrics, args, best_model=False):
    """{vocab, args, epoch, model_state_dict, inner_opt, outer_opt, metrics} -> <save_folder>/<name>/...
    (functions.py:101-126)."""
    path = _ckpt_path(args, epoch, best_model)
    print("SAVE MODEL to", path)
    logging.info("SAVE MODEL to " + path)
    from mtl_b200.optim import to_torch
    sd = _host_state_dict(model)
    host_params = [sd[name] for name, _ in model.named_parameters()]
    # the optimizers go in as plain torch.optim objects over the same host tensors: the reference's
    # load_meta_model (functions.py:183-186) reads such a file without this package
    torch.save({'vocab': vocab, 'args': args, 'epoch': epoch, 'model_state_dict': sd,
                'inner_opt': to_torch(inner_opt, host_params), 'outer_opt': to_torch(outer_opt, host_params),
                'metrics': metrics}, path)


def save_joint_model(model, vocab, epoch, opt, metrics, args, best_model=False):
    """{vocab, args, epoch, model_state_dict, opt, metrics}.  (The reference's save_joint_model writes an EMPTY
    dict for loss 'ce' -- functions.py:43-66 -- which its own load_joint_model cannot read; this writes the
    fields load_joint_model expects.)"""
    path = _ckpt_path(args, epoch, best_model)
    print("SAVE MODEL to", path)
    logging.info("SAVE MODEL to " + path)
    from mtl_b200.optim import to_torch
    sd = _host_state_dict(model)
    torch.save({'vocab': vocab, 'args': args, 'epoch': epoch, 'model_state_dict': sd,
                'opt': to_torch(opt, [sd[name] for name, _ in model.named_parameters()]), 'metrics': metrics}, path)


save_model = save_joint_model


def _restore_model(checkpoint, train):
    args, vocab = checkpoint['args'], checkpoint['vocab']
    model = init_transformer_model(args, vocab, train=train, is_factorized=args.is_factorized, r=args.r)
    model.load_state_dict(checkpoint['model_state_dict'])
    if args.cuda:
        print("CUDA")
        model = model.cuda()
    return model, vocab, args


def load_meta_model(load_path, train=True):
    """-> (model, vocab, inner_opt, outer_opt, epoch, metrics, args)  (functions.py:158-188).  The optimizers are
    arena optimizers carrying the checkpoint's state (SGD has none; Adam: step, exp_avg, exp_avg_sq)."""
    from mtl_b200.optim import ArenaAdam, ArenaSGD
    checkpoint = torch.load(load_path, map_location=torch.device('cpu'), weights_only=False)
    model, vocab, args = _restore_model(checkpoint, train)
    if not args.cuda:
        raise RuntimeError("checkpoint was trained with cuda=False; the B200 engine has no CPU path")
    inner_opt, outer_opt = ArenaSGD(model, args.lr), ArenaAdam(model, args.meta_lr)
    inner_opt.load_state_dict(checkpoint['inner_opt'].state_dict())
    outer_opt.load_state_dict(checkpoint['outer_opt'].state_dict())
    return model, vocab, inner_opt, outer_opt, checkpoint['epoch'], checkpoint['metrics'], args


def load_joint_model(load_path, train=True):
    """-> (model, vocab, opt, epoch, metrics, args)  (functions.py:190-218)."""
    from mtl_b200.optim import ArenaAdam
    checkpoint = torch.load(load_path, map_location=torch.device('cpu'), weights_only=False)
    model, vocab, args = _restore_model(checkpoint, train)
    opt = ArenaAdam(model, args.lr)
    opt.load_state_dict(checkpoint['opt'].state_dict())
    return model, vocab, opt, checkpoint['epoch'], checkpoint['metrics'], args


load_model = load_joint_model


def load_discriminator(load_path, train=True):
    raise NotImplementedError("adversarial / multi-task training is out of scope of the B200 hot path")
