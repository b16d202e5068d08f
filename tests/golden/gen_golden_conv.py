#!/usr/bin/env python3
"""CNN front-end fixtures from the UNMODIFIED reference `ConvEncoder` (eval mode) for the general block shapes: BatchNorm2d with
non-trivial running statistics, LayerNorm2D, residual, strided convolutions, (2,1) pooling, bottleneck.
    python tests/golden/gen_golden_conv.py      ->  tests/golden/zz_conv_*.npz (sd.*, xs, xlens, ys, ys_lens, cfg)"""
import importlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_import import import_reference  # noqa: E402

import_reference()

BASE = dict(input_dim=80, in_channel=1, channels="32_32_32", kernel_sizes="(3,3)_(3,3)_(3,3)", strides="(1,1)_(1,1)_(1,1)",
            poolings="(2,2)_(2,2)_(2,2)", dropout=0.0, normalization='', residual=False, bottleneck_dim=0, param_init=0.1)
CASES = {
    "bn_res": dict(normalization='batch_norm', residual=True, bottleneck_dim=16),
    "ln2d": dict(normalization='layer_norm', poolings="(2,2)_(2,1)_(1,1)"),
    "strided": dict(channels="32_32", kernel_sizes="(3,3)_(3,3)", strides="(2,2)_(2,2)", poolings="(1,1)_(1,1)", bottleneck_dim=24),
    "stride_ln": dict(channels="32_32", kernel_sizes="(3,3)_(3,3)", strides="(1,1)_(2,2)", poolings="(1,1)_(1,1)",
                      normalization='layer_norm'),
}


def main():
    mod = importlib.import_module('neural_sp.models.seq2seq.encoders.conv')
    for name, ov in CASES.items():
        torch.manual_seed(0)
        args = dict(BASE)
        args.update(ov)
        enc = mod.ConvEncoder(**args).eval()
        g = torch.Generator().manual_seed(1)
        for m in enc.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        rng = np.random.default_rng(7)
        xlens = [61, 50, 38]
        xs = np.zeros((3, 61, 80), np.float32)
        for b, n in enumerate(xlens):
            xs[b, :n] = rng.standard_normal((n, 80)).astype(np.float32)
        with torch.no_grad():
            ys, ylens = enc(torch.from_numpy(xs), torch.IntTensor(xlens))
        save = {"sd." + k: v.numpy() for k, v in enc.state_dict().items()}
        save.update(xs=xs, xlens=np.array(xlens, np.int32), ys=ys.numpy(), ys_lens=ylens.numpy().astype(np.int32),
                    cfg=np.array(json.dumps(args)))
        np.savez_compressed(os.path.join(HERE, "zz_conv_%s.npz" % name), **save)
        print(name, tuple(ys.shape), ylens.tolist())


sys.path.insert(0, HERE)
from gen_golden_conv_weights import loss_weights  # noqa: E402


def main_grads():
    """Training-mode gradients of the LayerNorm2D cases (train(): same function, dropout 0): d sum(ys * w) / d parameter, with
    non-trivial LayerNorm affine parameters.  -> zz_convgrad_*.npz (sd.*, xs, xlens, ys, g.*); w = loss_weights(...)."""
    mod = importlib.import_module('neural_sp.models.seq2seq.encoders.conv')
    for name in ("ln2d", "stride_ln"):
        torch.manual_seed(0)
        args = dict(BASE)
        args.update(CASES[name])
        enc = mod.ConvEncoder(**args).train()
        g = torch.Generator().manual_seed(2)
        with torch.no_grad():
            for k, p in enc.named_parameters():
                if ".norm" in k:
                    p.add_(0.3 * torch.randn(p.shape, generator=g))
        rng = np.random.default_rng(7)
        xlens = [61, 50, 38]
        xs = np.zeros((3, 61, 80), np.float32)
        for b, n in enumerate(xlens):
            xs[b, :n] = rng.standard_normal((n, 80)).astype(np.float32)
        ys, ylens = enc(torch.from_numpy(xs), torch.IntTensor(xlens))
        w = loss_weights(tuple(ys.shape), ylens.tolist())
        (ys * torch.from_numpy(w)).sum().backward()
        save = {"sd." + k: v.detach().numpy() for k, v in enc.state_dict().items()}
        save.update({"g." + k: p.grad.numpy() for k, p in enc.named_parameters()})
        save.update(xs=xs, xlens=np.array(xlens, np.int32), ys=ys.detach().numpy(), ys_lens=ylens.numpy().astype(np.int32),
                    cfg=np.array(json.dumps(args)))
        np.savez_compressed(os.path.join(HERE, "zz_convgrad_%s.npz" % name), **save)
        print("grad", name, tuple(ys.shape), len([k for k in save if k.startswith("g.")]))


if __name__ == "__main__":
    main()
    main_grads()
