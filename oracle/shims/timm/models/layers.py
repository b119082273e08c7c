import torch.nn as nn
from torch.nn.init import trunc_normal_
class DropPath(nn.Identity):
    def __init__(self, p=0.0, *a, **k):
        super().__init__()
def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)
