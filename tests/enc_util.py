"""Shared helpers: rebuild encoders (ours and the oracle's cfg) from a golden fixture."""
import json

import numpy as np
import torch


def golden_cfg(g):
    c = json.loads(str(g["cfg"]))
    return c["args"], c["conv"], c["kind"]


def oracle_cfg(g):
    a, conv, kind = golden_cfg(g)
    nl = a["n_layers"]
    sub, la = [1] * nl, [0] * nl
    for i, s in enumerate(a["subsample"].split("_")[:nl]):
        sub[i] = int(s)
    for i, s in enumerate(a["lookahead"].split("_")[:nl]):
        la[i] = int(s)
    cc = None
    if conv:
        pools = [tuple(int(v) for v in t.strip("()").split(",")) for t in conv["poolings"].split("_")]
        cc = dict(in_channel=conv["in_channel"], poolings=pools)
    n_c = int(str(a["chunk_size_current"]).split("_")[-1])
    lc = n_c > 0 and "uni" not in a["enc_type"]
    st = a["streaming_type"] if lc else ""
    extra = dict(streaming_type=st, N_l=int(str(a["chunk_size_left"]).split("_")[-1]), N_c=n_c,
                 N_r=int(str(a["chunk_size_right"]).split("_")[-1]))
    causal_attn = "uni" in a["enc_type"]
    cfg = dict(kind=kind, n_layers=nl, n_heads=a["n_heads"], d_model=a["d_model"], pe_type=a["pe_type"],
                clamp_len=a["clamp_len"], layer_norm_eps=a["layer_norm_eps"],
                normalization=a.get("normalization", "layer_norm"), causal=causal_attn, lookaheads=la,
                subsample=sub, dropout_layer=a["dropout_layer"], conv=cc, ffn_activation=a["ffn_activation"],
                n_layers_sub1=a["n_layers_sub1"])
    cfg.update(extra)
    cfg["causal_conv"] = causal_attn or st == "mask"
    return cfg


def state_dict_of(g):
    return {k[3:]: torch.from_numpy(np.asarray(g[k])) for k in g.files if k.startswith("sd.")}


def build_ours(g, device, precision):
    """Instantiate neural_sp_b200's encoder with the fixture's constructor args and load the reference weights."""
    from neural_sp_b200.encoders.conformer import ConformerEncoder
    from neural_sp_b200.encoders.conv import ConvEncoder
    from neural_sp_b200.encoders.transformer import TransformerEncoder
    a, conv, kind = golden_cfg(g)
    a = dict(a)
    a["frontend_conv"] = ConvEncoder(**conv) if conv else None
    if kind == "rnn":
        from neural_sp_b200.encoders.rnn import RNNEncoder
        enc = RNNEncoder(**a)
    elif kind == "conformer":
        enc = ConformerEncoder(**a)
    else:
        a.pop("kernel_size"), a.pop("normalization")
        enc = TransformerEncoder(**a)
    missing, unexpected = enc.load_state_dict(state_dict_of(g), strict=True)
    enc = enc.to(device).eval()
    enc.set_precision(precision)
    return enc


def replay_stream(enc, g, device):
    """Feed the fixture's chunk schedule (tests/golden/gen_golden_streaming.py) through `enc` with streaming=True.
    Returns the list of per-chunk outputs (fp32, CPU) and their lengths."""
    xs = torch.from_numpy(g["xs"])
    outs, lens = [], []
    enc.reset_cache()
    for start, end, pl, pr, xlen, lookback, lookahead in g["sched"].tolist():
        chunk = xs[:, start:end]
        if pl or pr:
            chunk = torch.cat([chunk.new_zeros(1, pl, chunk.size(2)), chunk, chunk.new_zeros(1, pr, chunk.size(2))], dim=1)
        o = enc(chunk.contiguous().to(device), torch.IntTensor([xlen]), task='all', streaming=True,
                lookback=bool(lookback), lookahead=bool(lookahead))['ys']
        outs.append(o['xs'].float().cpu())
        lens.append(int(o['xlens'][0]))
    return outs, lens
