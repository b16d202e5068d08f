"""Parameter initialisers matching the reference's schemes (modules/initialization.py:37-98)."""
import math

import torch.nn as nn


def init_with_xavier_uniform(n, p):
    if p.dim() == 1:
        nn.init.constant_(p, 0.)
    elif p.dim() in (2, 3, 4):
        nn.init.xavier_uniform_(p)
    else:
        raise ValueError(n)


def init_with_lecun_normal(n, p, param_init):
    if p.dim() == 1:
        nn.init.constant_(p, 0.)
    elif p.dim() in (2, 3, 4):
        fan_in = p.size(1) * (p[0][0].numel() if p.dim() > 2 else 1)
        nn.init.normal_(p, mean=0., std=1. / math.sqrt(fan_in))
    else:
        raise ValueError(n)


def init_with_uniform(n, p, param_init):
    if p.dim() == 1:
        nn.init.constant_(p, 0.)
    elif p.dim() in (2, 3, 4):
        nn.init.uniform_(p, a=-param_init, b=param_init)
    else:
        raise ValueError(n)
