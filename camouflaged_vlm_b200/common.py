"""Parameter containers mirroring models/mmseg/models/sam/common.py (MLPBlock :13-26, LayerNorm2d
:31-43) so that state_dict keys (``mlp.lin1.weight``, ``neck.1.weight`` ...) match the reference.
The encoder's compute for both lives in the CUDA library (lin1/lin2 GEMM epilogues, rowops.cu)."""
from typing import Type

import torch
import torch.nn as nn


class MLPBlock(nn.Module):
    def __init__(self, embedding_dim: int, mlp_dim: int, act: Type[nn.Module] = nn.GELU) -> None:
        super().__init__()
        self.lin1 = nn.Linear(embedding_dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, embedding_dim)
        self.act = act()


class LayerNorm2d(nn.Module):
    def __init__(self, num_channels: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps
