from torch import nn


class SgModule(nn.Module):
    """Subset of the reference's SgModule surface (training/models/sg_module.py:9-79) that the hot path needs."""

    def initialize_param_groups(self, lr: float, training_params) -> list:
        return [{"named_params": self.named_parameters()}]

    def update_param_groups(self, param_groups: list, lr: float, epoch: int, iter: int, training_params, total_batch: int) -> list:
        for param_group in param_groups:
            param_group["lr"] = lr
        return param_groups

    def get_include_attributes(self) -> list:
        return []

    def get_exclude_attributes(self) -> list:
        return []

    def prep_model_for_conversion(self, input_size=None, **kwargs):
        pass

    def replace_head(self, **kwargs):
        raise NotImplementedError

    def get_finetune_lr_dict(self, lr: float):
        raise NotImplementedError
