import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_golden
from enc_util import build_ours, golden_cfg
name = sys.argv[1] if len(sys.argv) > 1 else "enc_lc_reshape.npz"
g = load_golden(name)
for prec in ("fp32", "bf16"):
    enc = build_ours(g, torch.device("cuda:0"), prec)
    acts = []
    hooks = [enc.conv.register_forward_hook(lambda m, i, o: acts.append(o[0].float().cpu().numpy().copy()))]
    for layer in enc.layers:
        hooks.append(layer.register_forward_hook(lambda m, i, o: acts.append(o[0].float().cpu().numpy().copy())))
    out = enc(torch.from_numpy(g["xs"]).cuda(), torch.IntTensor(g["xlens"].tolist()), task="all")
    a, conv, kind = golden_cfg(g)
    for i, got in enumerate(acts):
        ref = g["act.%d" % i]
        if i == 0: ref = ref * np.sqrt(a["d_model"])
        print(prec, "act", i, got.shape, ref.shape, np.abs(got - ref).max() / np.abs(ref).max(), np.isnan(got).sum())
    ys = out["ys"]["xs"].float().cpu().numpy()
    print(prec, "out", np.abs(ys - g["ys"]).max() / np.abs(g["ys"]).max())
