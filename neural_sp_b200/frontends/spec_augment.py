"""SpecAugment (reference frontends/spec_augment.py:12-140), B200-native.

Same constructor, presets, `freq_mask` / `time_mask` properties and `__call__(xs)` on the padded device batch `[B, T, F]`
(in place, one mask set for the whole batch, exactly like the reference).  The mask parameters are drawn on the host with
the SAME `np.random.uniform` call sequence as the reference -- seeded runs pick identical rectangles -- and ALL of them are
applied by ONE kernel launch (nsp_mask_rects) instead of one slice-assignment kernel per mask; no mask tensor, no H2D copy
(the rectangles travel as kernel arguments)."""
import numpy as np

from .. import ops


class SpecAugment(object):
    def __init__(self, F, T, n_freq_masks, n_time_masks, p=1.0, W=40, adaptive_number_ratio=0, adaptive_size_ratio=0,
                 max_n_time_masks=20):
        self.W, self.F, self.T = W, F, T
        self.n_freq_masks, self.n_time_masks, self.p = n_freq_masks, n_time_masks, p
        self.adaptive_number_ratio = adaptive_number_ratio
        self.adaptive_size_ratio = adaptive_size_ratio
        self.max_n_time_masks = max_n_time_masks
        if adaptive_number_ratio > 0:
            self.n_time_masks = 0
        if adaptive_size_ratio > 0:
            self.T = 0
        self._freq_mask = None
        self._time_mask = None

    def librispeech_basic(self):
        self.W, self.F, self.T, self.n_freq_masks, self.n_time_masks, self.p = 80, 27, 100, 1, 1, 1.0

    def librispeech_double(self):
        self.W, self.F, self.T, self.n_freq_masks, self.n_time_masks, self.p = 80, 27, 100, 2, 2, 1.0

    def switchboard_mild(self):
        self.W, self.F, self.T, self.n_freq_masks, self.n_time_masks, self.p = 40, 15, 70, 2, 2, 0.2

    def switchboard_strong(self):
        self.W, self.F, self.T, self.n_freq_masks, self.n_time_masks, self.p = 40, 27, 70, 2, 2, 0.2

    @property
    def freq_mask(self):
        return self._freq_mask

    @property
    def time_mask(self):
        return self._time_mask

    def draw(self, n_frames, n_bins):
        """The reference's random draws in its order (mask_freq :112-120 first, then mask_time :122-140)
        -> (freq rectangles [(f0, f1)], time rectangles [(t0, t1)])."""
        fm, tm = [], []
        for _ in range(self.n_freq_masks):
            f = int(np.random.uniform(low=0, high=self.F))
            f_0 = int(np.random.uniform(low=0, high=n_bins - f))
            fm.append((f_0, f_0 + f))
            self._freq_mask = (f_0, f_0 + f)
        if self.adaptive_number_ratio > 0:
            n_masks = min(int(n_frames * self.adaptive_number_ratio), self.max_n_time_masks)
        else:
            n_masks = self.n_time_masks
        T = self.adaptive_size_ratio * n_frames if self.adaptive_size_ratio > 0 else self.T
        for _ in range(n_masks):
            t = int(np.random.uniform(low=0, high=T))
            t = min(t, int(n_frames * self.p))
            t_0 = int(np.random.uniform(low=0, high=n_frames - t))
            tm.append((t_0, t_0 + t))
            self._time_mask = (t_0, t_0 + t)
        return fm, tm

    def __call__(self, xs):
        """xs `[B, T, F]` fp32 CUDA tensor, masked in place (returned for the reference's `xs = self.specaug(xs)` idiom)."""
        fm, tm = self.draw(xs.size(1), xs.size(-1))
        if fm or tm:
            ops.mask_rects_(xs, fm, tm)
        return xs
