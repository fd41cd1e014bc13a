"""Drop-in for reference ``src/models/sed_model.py`` (the `SEDModel` ABC the trainers program against)."""
from abc import ABC, abstractmethod

import torch.nn as nn


class SEDModel(nn.Module, ABC):

    def __init__(self) -> None:
        super().__init__()

    def _replicate_for_data_parallel(self):
        """`nn.DataParallel` over SEVERAL devices replicates the module every forward (src/utils/__init__.py:11-21 wraps every net).
        The B200 path is one process per GPU (torchrun + `training.ParamArena`): replicas would read parameters out of torch's
        coalesced broadcast buffers, which are not 16-byte aligned as the kernels require, and would bypass the arena's bf16 operand
        copies.  With one visible device DataParallel calls the module directly and works unchanged; anything else fails loudly."""
        from .. import _lib
        raise _lib.T4sError("nn.DataParallel over more than one device is not supported by the B200 mirrors: run one process per GPU "
                            "(torchrun) or pass device_ids=[local_rank]; see INTEGRATION.md §2")

    @abstractmethod
    def get_feature_extractor(self):
        pass

    @abstractmethod
    def get_model_name(self) -> str:
        pass

    @abstractmethod
    def get_backbone_upsample_ratio(self):
        pass
