"""Encoder / RNN-T golden fixtures from the UNMODIFIED reference (see gen_golden.py for how to run).

Each enc_*.npz holds: the reference module's state_dict (``sd.<key>``), the seeded padded input batch,
``xlens``, the encoder output ``ys`` (+ ``xlens_out``), per-layer activations (``act.N``: conv front-end
output, then every block) and the constructor arguments as a JSON string (``cfg``)."""
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

BASE = dict(input_dim=80, enc_type='conv_conformer', n_heads=4, kernel_size=7, normalization='layer_norm',
            n_layers=3, n_layers_sub1=0, n_layers_sub2=0, d_model=64, d_ff=128, ffn_bottleneck_dim=0,
            ffn_activation='swish', pe_type='relative', layer_norm_eps=1e-12, last_proj_dim=0, dropout_in=0.0,
            dropout=0.0, dropout_att=0.0, dropout_layer=0.0, subsample="1_2_1", subsample_type='max_pool',
            n_stacks=1, n_splices=1, frontend_conv=None, task_specific_layer=False, param_init='xavier_uniform',
            clamp_len=10, lookahead="0_0_0", chunk_size_left="0", chunk_size_current="0", chunk_size_right="0",
            streaming_type='mask')
CONV = dict(input_dim=80, in_channel=1, channels="32_32", kernel_sizes="(3,3)_(3,3)", strides="(1,1)_(1,1)",
            poolings="(1,1)_(2,2)", dropout=0.0, normalization='', residual=False, bottleneck_dim=64, param_init=0.1)

CASES = {
    # LibriSpeech-recipe structure in miniature (poolings (1,1)_(2,2), hierarchical max-pool, relative + clamp 10)
    "enc_conformer_small": dict(args={}, conv={}, B=3, T=64, xlens=[64, 57, 40]),
    # relative_xl (u/v bias), unclamped distances, BatchNorm conv module (eval), (2,2)_(2,2) front-end,
    # LayerDrop rescale, sub-task output
    "enc_conformer_xl_bn": dict(args=dict(pe_type='relative_xl', clamp_len=-1, normalization='batch_norm', n_heads=2,
                                          d_model=32, d_ff=64, kernel_size=3, dropout_layer=0.2, n_layers=2,
                                          n_layers_sub1=1, subsample="1_1", lookahead="0_0"),
                                conv=dict(poolings="(2,2)_(2,2)", bottleneck_dim=32), B=2, T=90, xlens=[90, 71]),
    # Transformer blocks, relative_xl, relu FFN, output bridge
    "enc_transformer_xl": dict(args=dict(enc_type='conv_transformer', pe_type='relative_xl', ffn_activation='relu',
                                         n_layers=2, subsample="1_1", lookahead="0_0", last_proj_dim=48, clamp_len=5),
                               conv=dict(poolings="(2,2)_(2,2)"), B=2, T=70, xlens=[70, 33], kind='transformer'),
    # Transformer with pe_type='relative' -> plain MHA through the reference's typo; no CNN (embed Linear)
    "enc_transformer_plain": dict(args=dict(enc_type='transformer', pe_type='relative', ffn_activation='relu',
                                            n_layers=2, subsample="1_2", lookahead="0_0", n_heads=2, d_model=32, d_ff=64),
                                  conv=None, B=2, T=50, xlens=[50, 44], kind='transformer'),
    # latency-controlled, overlapped windows (no mask inside a window), centre extraction
    "enc_lc_reshape": dict(args=dict(n_layers=2, subsample="1_1", lookahead="0_0", chunk_size_left="16",
                                     chunk_size_current="32", chunk_size_right="16", streaming_type='reshape',
                                     kernel_size=3), conv=dict(poolings="(2,2)_(2,2)"), B=2, T=90, xlens=[90, 61]),
    # latency-controlled, chunk-wise attention mask + causal conv module, CNN applied chunk by chunk
    "enc_lc_mask": dict(args=dict(n_layers=2, subsample="1_1", lookahead="0_0", chunk_size_left="32",
                                  chunk_size_current="32", chunk_size_right="0", streaming_type='mask',
                                  kernel_size=3), conv=dict(poolings="(2,2)_(2,2)"), B=2, T=96, xlens=[96, 70]),
    # unidirectional Conformer: causal attention with lookahead, causal depthwise conv
    "enc_uni_conformer": dict(args=dict(enc_type='conv_uni_conformer', lookahead="1_0", n_layers=2, subsample="1_1",
                                        kernel_size=5), conv=dict(poolings="(2,2)_(2,2)"), B=2, T=60, xlens=[60, 48]),
}


def build_reference(name):
    import importlib
    case = CASES[name]
    args = dict(BASE)
    args.update(case["args"])
    torch.manual_seed(0)
    conv_args = None
    if case["conv"] is not None:
        conv_args = dict(CONV)
        conv_args.update(case["conv"])
        conv_args["bottleneck_dim"] = args["d_model"]
        conv_mod = importlib.import_module('neural_sp.models.seq2seq.encoders.conv')
        args["frontend_conv"] = conv_mod.ConvEncoder(**conv_args)
    kind = case.get("kind", "conformer")
    if kind == "conformer":
        mod = importlib.import_module('neural_sp.models.seq2seq.encoders.conformer')
        enc = mod.ConformerEncoder(**args)
    else:
        mod = importlib.import_module('neural_sp.models.seq2seq.encoders.transformer')
        a = dict(args)
        a.pop("kernel_size"); a.pop("normalization")
        enc = mod.TransformerEncoder(**a)
    # make BatchNorm running statistics non-trivial
    g = torch.Generator().manual_seed(1)
    for m in enc.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    enc.eval()
    return enc, args, conv_args, kind


def gen_encoder():
    for name, case in CASES.items():
        enc, args, conv_args, kind = build_reference(name)
        rng = np.random.default_rng(1234)
        B, T = case["B"], case["T"]
        xs = np.zeros((B, T, 80), np.float32)
        for b, n in enumerate(case["xlens"]):
            xs[b, :n] = rng.standard_normal((n, 80)).astype(np.float32)
        xlens = torch.IntTensor(case["xlens"])
        acts = []
        hooks = []
        if enc.conv is not None:
            hooks.append(enc.conv.register_forward_hook(lambda m, i, o: acts.append(o[0].detach().numpy().copy())))
        elif hasattr(enc, "embed"):
            hooks.append(enc.embed.register_forward_hook(lambda m, i, o: acts.append(o.detach().numpy().copy())))
        for layer in enc.layers:
            hooks.append(layer.register_forward_hook(lambda m, i, o: acts.append(o[0].detach().numpy().copy())))
        with torch.no_grad():
            out = enc(torch.from_numpy(xs), xlens.clone(), task='all')
        for h in hooks:
            h.remove()
        save = {"sd." + k: v.numpy() for k, v in enc.state_dict().items()}
        for i, a in enumerate(acts):
            save["act.%d" % i] = a
        save.update(xs=xs, xlens=np.array(case["xlens"], np.int32), ys=out['ys']['xs'].numpy(),
                    xlens_out=out['ys']['xlens'].numpy().astype(np.int32))
        if out['ys_sub1']['xs'] is not None:
            save["ys_sub1"] = out['ys_sub1']['xs'].numpy()
        cfg = {k: v for k, v in args.items() if k != "frontend_conv"}
        save["cfg"] = np.array(json.dumps(dict(args=cfg, conv=conv_args, kind=kind)))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
        print("encoder", name, out['ys']['xs'].shape, out['ys']['xlens'].tolist(), len(acts), "acts")


def gen_rnnt():
    """RNN-T loss goldens: the reference's warp_rnnt / warprnnt_pytorch are not installable offline
    (rnn_transducer.py:248-256); torchaudio.functional.rnnt_loss is the stand-in oracle."""
    import torchaudio
    cases = {"rnnt_small": (3, 20, 6, 12, [20, 17, 9], [6, 4, 2]), "rnnt_mid": (2, 50, 15, 40, [50, 41], [15, 11])}
    for name, (B, T, U, V, flens, ylens) in cases.items():
        torch.manual_seed(3)
        rng = np.random.default_rng(11)
        logits = (torch.randn(B, T, U + 1, V) * 1.5).requires_grad_(True)
        ys = np.zeros((B, U), np.int32)
        for b in range(B):
            ys[b, :ylens[b]] = rng.integers(1, V, size=ylens[b])
        lp = logits.log_softmax(-1)
        nll = torchaudio.functional.rnnt_loss(lp, torch.from_numpy(ys), torch.IntTensor(flens), torch.IntTensor(ylens),
                                              blank=0, reduction='none', fused_log_softmax=False)
        loss = nll.mean()                     # warp_rnnt reduction='mean' (rnn_transducer.py:249-252)
        (glp,) = torch.autograd.grad(loss, lp, retain_graph=True)
        loss.backward()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), logits=logits.detach().numpy(), ys=ys,
                            flens=np.array(flens, np.int32), ylens=np.array(ylens, np.int32), nll=nll.detach().numpy(),
                            loss=loss.detach().numpy(), grad_logits=logits.grad.numpy(), grad_log_probs=glp.numpy())
        print("rnnt", name, float(loss))


if __name__ == "__main__":
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle.ref_import import import_reference
    import_reference()
    gen_encoder()
    gen_rnnt()


def gen_rnn_encoder():
    """RNNEncoder goldens: BASELINE configs[0] structure (BLSTM 2x256, no CNN) and a small conv_lstm with projections."""
    import importlib
    mod = importlib.import_module('neural_sp.models.seq2seq.encoders.rnn')
    conv_mod = importlib.import_module('neural_sp.models.seq2seq.encoders.conv')
    base = dict(input_dim=80, enc_type='blstm', n_units=256, n_projs=0, last_proj_dim=0, n_layers=2, n_layers_sub1=0,
                n_layers_sub2=0, dropout_in=0.0, dropout=0.0, subsample="1_1", subsample_type='drop', n_stacks=1,
                n_splices=1, frontend_conv=None, bidir_sum_fwd_bwd=False, task_specific_layer=False, param_init=0.1,
                chunk_size_current="0", chunk_size_right="0", cnn_lookahead=True, rsp_prob=0.)
    cases = {
        "rnn_c1_blstm": dict(args={}, conv=None, B=2, T=200, xlens=[200, 173]),
        "rnn_conv_lstm_proj": dict(args=dict(enc_type='conv_lstm', n_units=64, n_projs=32, n_layers=3, subsample="1_2_1",
                                             subsample_type='max_pool', last_proj_dim=40, n_layers_sub1=2),
                                   conv=dict(CONV, poolings="(2,2)_(2,2)", bottleneck_dim=0), B=3, T=90, xlens=[90, 77, 41]),
        "rnn_blstm_sum": dict(args=dict(n_units=32, bidir_sum_fwd_bwd=True, subsample="2_1", subsample_type='concat'),
                              conv=None, B=3, T=37, xlens=[37, 30, 5]),
    }
    for name, case in cases.items():
        torch.manual_seed(0)
        args = dict(base)
        args.update(case["args"])
        if case["conv"] is not None:
            args["frontend_conv"] = conv_mod.ConvEncoder(**case["conv"])
        enc = mod.RNNEncoder(**args)
        with torch.no_grad():                      # weights rounded to fp16-representable values: fixture stored as fp16
            for p_ in enc.parameters():
                p_.copy_(p_.half().float())
        enc.eval()
        rng = np.random.default_rng(4321)
        B, T = case["B"], case["T"]
        xs = np.zeros((B, T, 80), np.float32)
        for b, n in enumerate(case["xlens"]):
            xs[b, :n] = rng.standard_normal((n, 80)).astype(np.float32)
        with torch.no_grad():
            out = enc(torch.from_numpy(xs), torch.IntTensor(case["xlens"]), task='all')
        save = {"sd." + k: v.numpy().astype(np.float16) for k, v in enc.state_dict().items()}
        save.update(xs=xs, xlens=np.array(case["xlens"], np.int32), ys=out['ys']['xs'].numpy(),
                    xlens_out=np.asarray(out['ys']['xlens']).astype(np.int32))
        if out['ys_sub1']['xs'] is not None:
            save["ys_sub1"] = out['ys_sub1']['xs'].numpy()
        cfg = {k: v for k, v in args.items() if k != "frontend_conv"}
        save["cfg"] = np.array(json.dumps(dict(args=cfg, conv=case["conv"])))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
        print("rnn", name, out['ys']['xs'].shape, np.asarray(out['ys']['xlens']).tolist())


GRAD_CASES = ["enc_conformer_small", "enc_transformer_xl", "enc_transformer_plain", "enc_uni_conformer"]


def grad_loss_weights(shape, xlens_out, seed=4321):
    """Fixed projection of the encoder output to a scalar: loss = sum(ys * w), w = 0 on padded output frames (as every
    real loss of the path: CTC / attention only read frames < xlens).  Shared by generator and tests."""
    w = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
    for b, n in enumerate(xlens_out):
        w[b, int(n):] = 0.0
    return w


def gen_encoder_grads():
    """Parameter gradients of the UNMODIFIED reference (torch autograd, CPU fp32) for loss = sum(ys * w):
    encgrad_<case>.npz holds ``g.<param>`` for every parameter; inputs / weights are those of enc_<case>.npz."""
    for name in GRAD_CASES:
        case = CASES[name]
        enc, args, conv_args, kind = build_reference(name)      # eval(): dropouts are 0, LayerDrop is 0 in these cases
        base = np.load(os.path.join(HERE, name + ".npz"))
        for k, v in enc.state_dict().items():
            assert np.array_equal(v.numpy(), base["sd." + k]), k
        xs = torch.from_numpy(base["xs"])
        out = enc(xs, torch.IntTensor(case["xlens"]), task='all')
        ys = out['ys']['xs']
        w = torch.from_numpy(grad_loss_weights(tuple(ys.shape), out['ys']['xlens'].tolist()))
        loss = (ys * w).sum()
        loss.backward()
        save = {"g." + k: p.grad.numpy() for k, p in enc.named_parameters() if p.grad is not None}
        missing = [k for k, p in enc.named_parameters() if p.grad is None]
        save["loss"] = np.array(float(loss.detach()), np.float32)
        np.savez_compressed(os.path.join(HERE, "encgrad_" + name[4:] + ".npz"), **save)
        print("encoder grads", name, float(loss), len(save) - 1, "tensors; no grad:", missing)


RNN_GRAD_CASES = ["rnn_conv_lstm_proj", "rnn_blstm_sum"]


def gen_rnn_grads():
    """Parameter gradients of the UNMODIFIED reference RNNEncoder (torch autograd through nn.LSTM on packed sequences, CPU
    fp32) for loss = sum(ys * w) [+ sum(ys_sub1 * w1)]: rnngrad_<case>.npz; inputs / weights are those of rnn_<case>.npz."""
    import importlib
    mod = importlib.import_module('neural_sp.models.seq2seq.encoders.rnn')
    conv_mod = importlib.import_module('neural_sp.models.seq2seq.encoders.conv')
    for name in RNN_GRAD_CASES:
        base = np.load(os.path.join(HERE, name + ".npz"))
        c = json.loads(str(base["cfg"]))
        args = dict(c["args"])
        torch.manual_seed(0)
        args["frontend_conv"] = conv_mod.ConvEncoder(**c["conv"]) if c["conv"] else None
        enc = mod.RNNEncoder(**args)
        enc.load_state_dict({k[3:]: torch.from_numpy(base[k].astype(np.float32)) for k in base.files if k.startswith("sd.")})
        enc.eval()                                   # dropouts are 0 in these cases
        out = enc(torch.from_numpy(base["xs"]), torch.IntTensor(base["xlens"].tolist()), task='all')
        ys = out['ys']['xs']
        assert np.allclose(ys.detach().numpy(), base["ys"], atol=1e-6)
        loss = (ys * torch.from_numpy(grad_loss_weights(tuple(ys.shape), out['ys']['xlens'].tolist()))).sum()
        if out['ys_sub1']['xs'] is not None:
            s1 = out['ys_sub1']['xs']
            loss = loss + (s1 * torch.from_numpy(grad_loss_weights(tuple(s1.shape), out['ys_sub1']['xlens'].tolist(), seed=99))).sum()
        loss.backward()
        save = {"g." + k: p.grad.numpy() for k, p in enc.named_parameters() if p.grad is not None}
        missing = [k for k, p in enc.named_parameters() if p.grad is None]
        save["loss"] = np.array(float(loss.detach()), np.float32)
        np.savez_compressed(os.path.join(HERE, "rnngrad_" + name[4:] + ".npz"), **save)
        print("rnn grads", name, float(loss.detach()), len(save) - 1, "tensors; no grad:", missing)
