"""Accent discriminator (modules/discriminator.py): only used by the adversarial / multi-task flags of
joint_train.py, which are outside the B200 hot path.  Kept importable so ``from modules import Discriminator``
(utils/functions.py:12) resolves."""
import torch.nn as nn


class Discriminator(nn.Module):
    def __init__(self, dim_model, num_class):
        super().__init__()
        raise NotImplementedError("adversarial / multi-task training (joint_train.py --adversarial/--multitask) is "
                                  "out of scope of the B200 hot path")
