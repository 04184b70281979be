"""``fireflies/utils/torch_grads.py``."""
from typing import List

import torch


def retain_grads(non_leaf_tensor: List[torch.Tensor]) -> None:
    for tensor in non_leaf_tensor:
        tensor.retain_grad()
