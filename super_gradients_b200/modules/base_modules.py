from abc import ABC, abstractmethod
from typing import List, Union

from torch import nn


class BaseDetectionModule(nn.Module, ABC):
    """Detection-module contract (reference: modules/base_modules.py:9-27): built with `in_channels`, exposes
    `out_channels`."""

    def __init__(self, in_channels: Union[List[int], int], **kwargs):
        super().__init__()
        self.in_channels = in_channels

    @property
    @abstractmethod
    def out_channels(self) -> Union[List[int], int]:
        raise NotImplementedError()
