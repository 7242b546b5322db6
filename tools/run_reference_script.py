#!/usr/bin/env python
"""Runs one of the reference's CLI scripts UNCHANGED against the B200 packages.

    python tools/run_reference_script.py /path/to/meta-transfer-learning/meta_transfer_train.py --cuda --copy-grad ...

The script file is executed byte-for-byte with ``runpy`` as ``__main__``; only ``sys.path`` differs: this
repository's ``meta-transfer-learning_b200/`` directory comes first, so the script's
``from trainer.asr.transient_trainer import TransientTrainer``, ``from utils.data_loader import ...``,
``from utils.functions import ...`` and ``from torchsummary import summary`` resolve to the B200 packages
instead of the reference's own modules.  Everything after the script path is passed on as its argv.
(With ``torchrun`` in front, the trainers shard the tasks over the ranks: set ``MTL_DIST=1`` to have this
launcher initialise ``torch.distributed`` (NCCL) before the script starts.)"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "meta-transfer-learning_b200")


def main(argv):
    if len(argv) < 2:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[1])
    script_dir = os.path.dirname(script)
    sys.path[:] = [PKG] + [p for p in sys.path if os.path.abspath(p or ".") not in (script_dir, PKG)]
    sys.argv = [script] + argv[2:]
    if os.environ.get("MTL_DIST") == "1" and "RANK" in os.environ:
        import torch
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
