"""Loss weights shared by tests/golden/gen_golden_conv.py (main_grads) and the tests that replay its gradient fixtures; kept
apart from the generator so that the tests need not import the reference."""
import numpy as np


def loss_weights(shape, ylens):
    w = np.random.default_rng(11).standard_normal(shape).astype(np.float32)
    for b, n in enumerate(ylens):
        w[b, int(n):] = 0
    return w
