"""models.get() (reference: training/models/model_factory.py:192-256) over the ARCHITECTURES registry."""
from typing import Optional

import torch

from ...common.registry import ARCHITECTURES
from ..utils import HpmStruct


def get(model_name: str, arch_params: Optional[dict] = None, num_classes: int = None, strict_load: bool = True, checkpoint_path: str = None, pretrained_weights: str = None, load_backbone: bool = False, download_required_code: bool = True, checkpoint_num_classes: int = None, num_input_channels: int = None):
    """Builds a registered architecture.  `pretrained_weights` (a download in the reference) is not available offline;
    `checkpoint_path` loads an SG checkpoint (the state-dict keys are identical to the reference's)."""
    if pretrained_weights is not None:
        raise NotImplementedError("pretrained weight download is out of scope (no network); pass checkpoint_path instead")
    if model_name not in ARCHITECTURES:
        raise KeyError(f"unknown model `{model_name}`; registered: {sorted(ARCHITECTURES)}")
    arch_params = dict(arch_params or {})
    # the checkpoint's head first, the requested one afterwards (model_factory.py:227-254)
    checkpoint_num_classes = checkpoint_num_classes or num_classes
    if checkpoint_num_classes is not None:
        arch_params["num_classes"] = checkpoint_num_classes
    net = ARCHITECTURES[model_name](HpmStruct(**arch_params))
    if checkpoint_path is not None:
        ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        sd = ckpt.get("ema_net", ckpt.get("net", ckpt)) if isinstance(ckpt, dict) else ckpt
        sd = {k[len("module.") :] if k.startswith("module.") else k: v for k, v in sd.items()}
        net.load_state_dict(sd, strict=bool(strict_load))
    if checkpoint_num_classes != num_classes:
        net.replace_head(new_num_classes=num_classes)
    if num_input_channels is not None and num_input_channels != net.get_input_channels():
        net.replace_input_channels(in_channels=num_input_channels)
    return net
