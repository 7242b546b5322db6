"""Import shim: the reference scripts do ``from torchsummary import summary`` (meta_transfer_train.py:13,
joint_train.py:13) but never call it; the third-party package is not part of this image."""


def summary(model, *args, **kwargs):
    print(model)
    n = sum(p.numel() for p in model.parameters())
    print(f"Total params: {n:,}")
    return n
