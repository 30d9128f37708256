from torch import nn


class Residual(nn.Module):
    """Identity marker used for residual branches (reference: modules/skip_connections.py)."""

    def forward(self, x):
        return x
