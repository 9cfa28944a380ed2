"""Import shim used ONLY by tests/golden/make_golden.py to import the reference's modules in the
build container (pytorch_lightning is not installed).  The reference only subclasses
pl.LightningModule; no Lightning behaviour is exercised on the inference path."""
import torch.nn as nn


class LightningModule(nn.Module):
    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            import torch
            return torch.device("cpu")

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass
