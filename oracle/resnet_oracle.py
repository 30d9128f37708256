"""Whole-graph CPU oracle of the ImageNet ResNets (TEST INFRASTRUCTURE -- see oracle/sg_oracle.py for the rules).

A functional fp32 restatement of ResNet.forward (training/models/classification_models/resnet.py:194-210: 7x7 stride-2 conv, BN,
ReLU, 3x3 stride-2 max-pool, four stages of Bottleneck / BasicResNetBlock, global average pool, Linear) over a reference-format
state dict, with the drop-path scale of every block as an explicit input (training/utils/regularization_utils.py:4-15).
Pinned by tests/test_oracle_golden.py::test_resnet50_oracle_matches_reference against tests/golden/other_configs.pt (logits,
loss and gradient norms produced by the unmodified reference for its seeded initialisation).
Used by bench.py's CPU arm for config 4 (`cpu_baseline` / `--impl reference`), nowhere on the product path.
"""
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import sg_oracle as O

STRUCTURE = {"resnet18": ("basic", (2, 2, 2, 2)), "resnet34": ("basic", (3, 4, 6, 3)), "resnet50": ("bottleneck", (3, 4, 6, 3)), "resnet101": ("bottleneck", (3, 4, 23, 3))}


def block_list(name: str) -> List[str]:
    kind, layers = STRUCTURE[name]
    return [f"layer{li + 1}.{bi}." for li, n in enumerate(layers) for bi in range(n)]


def resnet_forward(name: str, p: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = True, sample_scales: Optional[Dict[str, torch.Tensor]] = None, eps=1e-5, momentum=0.1):
    """logits [N, num_classes].  `sample_scales`: {block prefix: [N] drop-path scale} (missing / None: no drop-path)."""
    kind, layers = STRUCTURE[name]
    out = O.q(F.relu(O.batch_norm(O.q(F.conv2d(O.q(x), O.qw(p["conv1.weight"]), stride=2, padding=3)), p, "bn1.", training, eps, momentum)))
    out = F.max_pool2d(out, kernel_size=3, stride=2, padding=1)
    for li, n in enumerate(layers):
        for bi in range(n):
            pre = f"layer{li + 1}.{bi}."
            stride = 2 if (bi == 0 and li > 0) else 1
            has_sc = (pre + "shortcut.0.weight") in p
            ss = None if sample_scales is None else sample_scales.get(pre)
            fn = O.resnet_bottleneck if kind == "bottleneck" else O.resnet_basic_block
            out = fn(out, p, pre, stride, has_sc, training, eps, momentum, sample_scale=ss)
    out = O.q(F.adaptive_avg_pool2d(out, 1)).flatten(1)
    return F.linear(out, O.qw(p["linear.weight"]), p["linear.bias"])


def train_step(name: str, state: Dict[str, torch.Tensor], x: torch.Tensor, y: torch.Tensor, live: Sequence[str], droppath_prob: float = 0.0, generator=None):
    """fp32 CPU forward + cross entropy + backward; returns (loss, {param: grad})."""
    p = dict(state)
    for k in live:
        p[k] = p[k].detach().clone().requires_grad_(True)
    scales = None
    if droppath_prob > 0:
        keep = 1.0 - droppath_prob
        scales = {pre: torch.empty(x.shape[0]).bernoulli_(keep, generator=generator).div_(keep) for pre in block_list(name)}
    logits = resnet_forward(name, p, x, True, scales)
    loss = F.cross_entropy(logits, y)
    loss.backward()
    return loss.detach(), {k: p[k].grad for k in live if p[k].grad is not None}
