"""Drop-path (stochastic depth per sample).  reference: training/utils/regularization_utils.py:4-49.

The reference multiplies the residual branch by a per-sample Bernoulli mask in a separate elementwise pass.  Here the module only
DRAWS the mask (`sample_scale`): the multiply is folded into the fused BatchNorm + residual + activation kernel of the branch's
last convolution (`SgbBnDesc.sample_scale`, csrc/bn_kernels.cu) and into its backward passes, so drop-path costs no extra pass
over the activation."""
import torch
from torch import nn


class DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def sample_scale(self, x: torch.Tensor):
        """fp32 [N]: 0 or 1 / keep_prob per image (None when drop-path is inactive), the `random_tensor` of drop_path()."""
        if self.drop_prob == 0.0 or not self.training:
            return None
        keep_prob = 1 - self.drop_prob
        m = torch.empty(x.shape[0], dtype=torch.float32, device=x.device).bernoulli_(keep_prob)
        if keep_prob > 0.0 and self.scale_by_keep:
            m.div_(keep_prob)
        return m

    def forward(self, x):
        """Stand-alone use (not on the fused path): the reference's elementwise form."""
        m = self.sample_scale(x)
        return x if m is None else x * m.to(x.dtype).view(-1, *([1] * (x.dim() - 1)))

    def extra_repr(self):
        return f"drop_prob={round(self.drop_prob, 3):0.3f}"
