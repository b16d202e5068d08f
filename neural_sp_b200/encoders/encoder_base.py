"""Base class for encoders (reference encoders/encoder_base.py:19-73; plotting helpers omitted)."""
import torch.nn as nn

from ..modules._prep import VALID_PRECISIONS


class EncoderBase(nn.Module):
    @property
    def output_dim(self):
        return self._odim

    @property
    def output_dim_sub1(self):
        return getattr(self, '_odim_sub1', self._odim)

    @property
    def output_dim_sub2(self):
        return getattr(self, '_odim_sub2', self._odim)

    @property
    def subsampling_factor(self):
        return self._factor

    @property
    def subsampling_factor_sub1(self):
        return self._factor_sub1

    @property
    def subsampling_factor_sub2(self):
        return self._factor_sub2

    def reset_cache(self):
        raise NotImplementedError

    def set_precision(self, precision):
        """'bf16' (performance), 'tf32' (one tf32 pass), 'fp32' (3xTF32 GEMMs + fp32 attention: parity mode)."""
        assert precision in VALID_PRECISIONS, precision
        for m in self.modules():
            m.precision = precision
        return self

    def turn_on_ceil_mode(self, encoder):
        pass   # pooling here is always ceil-mode (conv.py:333)

    def turn_off_ceil_mode(self, encoder):
        raise NotImplementedError("floor-mode pooling is not on the B200 path")

    def _plot_attention(self, save_path=None, n_cols=2):
        """Reference encoder_base.py:75: plots `self.aws_dict`.  The flash-style attention kernels never materialise the
        `[B, H, T', T']` weights, so `aws_dict` stays empty and there is nothing to draw (kept for the trainer's call)."""
        return None
