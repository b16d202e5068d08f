#!/usr/bin/env python3
"""Length arithmetic of the reference (conv.py:423-477 update_lens_1d/2d, subsampling.py) for every input length 1..260
and the front-end / subsampler configurations of the recipes  ->  tests/golden/lens_contract.json (build container only)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CONVS = {"(1,1)_(2,2)": "(1,1)_(2,2)", "(2,2)_(2,2)": "(2,2)_(2,2)", "(1,1)_(1,1)": "(1,1)_(1,1)", "(3,2)_(2,1)": "(3,2)_(2,1)"}
SUBS = [("max_pool", 2), ("mean_pool", 3), ("drop", 2), ("add", 2), ("concat", 2), ("conv1d", 2), ("max_pool", 4)]
LENS = list(range(1, 261))

if __name__ == "__main__":
    import torch
    from oracle.ref_import import import_reference
    import_reference()
    import importlib
    conv_mod = importlib.import_module('neural_sp.models.seq2seq.encoders.conv')
    sub_mod = importlib.import_module('neural_sp.models.seq2seq.encoders.subsampling')
    out = {"conv": {}, "sub": {}}
    for name, pool in CONVS.items():
        enc = conv_mod.ConvEncoder(80, in_channel=1, channels="32_32", kernel_sizes="(3,3)_(3,3)", strides="(1,1)_(1,1)",
                                   poolings=pool, dropout=0., normalization='', residual=False, bottleneck_dim=0, param_init=0.1)
        res = []
        for n in LENS:
            xl = torch.IntTensor([n])
            for blk in enc.layers:
                xl = conv_mod.update_lens_2d(xl, blk.conv1, dim=0)
                xl = conv_mod.update_lens_2d(xl, blk.conv2, dim=0)
                if blk.pool is not None:
                    xl = conv_mod.update_lens_2d(xl, blk.pool, dim=0)
            res.append(int(xl[0]))
        out["conv"][name] = res
    cls = {"max_pool": "MaxPoolSubsampler", "mean_pool": "MeanPoolSubsampler", "drop": "DropSubsampler", "add": "AddSubsampler",
           "concat": "ConcatSubsampler", "conv1d": "Conv1dSubsampler"}
    for typ, f in SUBS:
        C = getattr(sub_mod, cls[typ])
        m = C(f, 8) if typ in ("concat", "conv1d") else C(f)
        res = []
        for n in LENS:
            xs = torch.zeros(1, n, 8)
            try:
                _, xl = m(xs, torch.IntTensor([n]))
                res.append(int(xl[0]))
            except Exception:          # noqa: BLE001
                res.append(None)
        out["sub"]["%s/%d" % (typ, f)] = res
    json.dump(out, open(os.path.join(HERE, "lens_contract.json"), "w"))
    print("wrote lens_contract.json")
