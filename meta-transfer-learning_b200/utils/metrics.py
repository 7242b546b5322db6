"""Loss and error-rate metrics with the reference's names (utils/metrics.py:38-126).

``calculate_metrics`` returns the CE of the model's logits.  When ``pred`` is the latest output of the B200
``Transformer`` the loss tensor is the one the fused CE kernel already produced, wired (via autograd) to the
fused backward; for any other tensor it is plain ``F.cross_entropy`` on that tensor.  CER / WER use a native
edit distance (python-Levenshtein is a third-party C extension the reference imports, utils/metrics.py:3)."""
import torch
import torch.nn.functional as F


def _edit_distance(a, b):
    """Levenshtein distance between two sequences (two-row DP)."""
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i] + [0] * len(b)
        for j, cb in enumerate(b, 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb))
        prev = cur
    return prev[-1]


def calculate_cer(s1, s2):
    """Character edit distance between hypothesis s1 and gold s2 (utils/metrics.py:38-46)."""
    return _edit_distance(s1, s2)


def calculate_wer(s1, s2):
    """Word edit distance between two space-separated sentences (utils/metrics.py:48-66)."""
    return _edit_distance(s1.split(), s2.split())


def calculate_loss(pred, gold, pad_id, input_lengths=None, target_lengths=None, non_pad_mask=None, smoothing=0.0,
                   loss_type="ce"):
    """Mean CE over non-pad positions, optional label smoothing eps: target (1-eps) on gold, eps/V elsewhere
    (utils/metrics.py:96-126)."""
    if loss_type != "ce":
        raise NotImplementedError("only the cross-entropy loss is on the B200 hot path (ctc is a non-default flag)")
    owner = getattr(pred, "_mtl_owner", None)
    if owner is not None:
        fused = owner[0].fused_loss(pred, smoothing)
        if fused is not None:
            return fused
    flat, g = pred.reshape(-1, pred.size(-1)), gold.contiguous().view(-1).long()
    if smoothing > 0.0:
        keep = g.ne(pad_id)
        logp = F.log_softmax(flat, dim=1)
        v = flat.size(1)
        target = torch.full_like(logp, smoothing / v).scatter_(1, (g * keep).view(-1, 1), 1.0 - smoothing)
        return -(target * logp).sum(dim=1)[keep].sum() / keep.sum()
    return F.cross_entropy(flat, g, ignore_index=pad_id, reduction="mean")


def calculate_metrics(pred, gold, pad_id, input_lengths=None, target_lengths=None, non_pad_mask=None, smoothing=0.0,
                      loss_type="ce"):
    """(loss, number of correct non-pad top-1 predictions)  (utils/metrics.py:68-94)."""
    if non_pad_mask is None:
        non_pad_mask = gold.ne(pad_id)
    else:
        gold.masked_fill_(torch.logical_not(non_pad_mask), pad_id)
    loss = calculate_loss(pred, gold, pad_id, input_lengths, target_lengths, non_pad_mask, smoothing, loss_type)
    owner = getattr(pred, "_mtl_owner", None)
    if owner is not None and owner[0]._last is not None and owner[0]._last["pred"].shape == pred.shape:
        return loss, int(owner[0]._last["ce"][2].item())          # counted by the fused CE kernel
    hyp = pred.detach().reshape(-1, pred.size(-1)).argmax(dim=1)
    return loss, int((hyp.eq(gold.view(-1)) & non_pad_mask.view(-1)).sum().item())
