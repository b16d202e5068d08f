"""Intermediate time subsamplers (reference encoders/subsampling.py).  MaxPoolSubsampler (:175-209) is the
one the LibriSpeech Conformer recipes use; the other five variants are not on the B200 path yet."""
import torch
import torch.nn as nn

from .. import ops


def update_lens_pool1d(xlens, factor):
    """ceil-mode MaxPool1d(kernel=stride=factor): n -> (n + 1 - factor) // factor + 1  (conv.py:443-445)."""
    return torch.IntTensor([(int(n) + 1 - factor) // factor + 1 for n in xlens])


class MaxPoolSubsampler(nn.Module):
    def __init__(self, subsampling_factor):
        super().__init__()
        self.factor = subsampling_factor

    def forward(self, xs, xlens, batch_first=True):
        if self.factor == 1:
            return xs, xlens
        assert batch_first
        return ops.maxpool_time(xs, self.factor), update_lens_pool1d(xlens, self.factor)
