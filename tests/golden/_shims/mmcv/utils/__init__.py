from torch.nn.modules.batchnorm import _BatchNorm  # noqa: F401


def print_log(*a, **k):
    pass
