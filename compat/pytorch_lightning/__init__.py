"""Import shim for code that only needs `pl.LightningModule` as an nn.Module base
(reference compute_pose.py:1,6).  Put <repo>/compat on sys.path to use it."""
import torch


class LightningModule(torch.nn.Module):
    pass
