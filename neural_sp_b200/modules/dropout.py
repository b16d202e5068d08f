"""nn.Dropout on the library's kernel (csrc/dropout.cu): identity in eval mode, Philox mask regenerated in the backward
(neural_sp_b200/autograd.py _DropoutFn) in training mode.  Subclasses nn.Dropout so that code inspecting ``.p`` or
``isinstance(m, nn.Dropout)`` keeps working; there are no parameters, so state_dict keys are unaffected."""
import torch.nn as nn


class Dropout(nn.Dropout):
    def forward(self, x):
        if not self.training or self.p <= 0:
            return x
        from .. import autograd as ag
        return ag.dropout(x, self.p)
