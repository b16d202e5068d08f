"""Conformer training with BatchNorm in the convolution module (the reference's DEFAULT `conformer_normalization`), CPU part:
in train() mode the batch statistics are taken over all B*T frames, the running statistics are updated, and the gradient flows
through the statistics.  Ours (ops replaced by their torch restatements) against the UNMODIFIED reference in train() mode with
identical weights: outputs, every parameter gradient, and the BatchNorm running statistics / num_batches_tracked after the step.
Needs /root/reference (build container only): skipped elsewhere."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/neural_sp"), reason="reference tree not available")

BASE = dict(input_dim=80, enc_type='conv_conformer', n_heads=2, kernel_size=7, normalization='batch_norm', n_layers=2,
            n_layers_sub1=0, n_layers_sub2=0, d_model=32, d_ff=64, ffn_bottleneck_dim=0, ffn_activation='swish',
            pe_type='relative', layer_norm_eps=1e-12, last_proj_dim=0, dropout_in=0.0, dropout=0.0, dropout_att=0.0,
            dropout_layer=0.0, subsample="2_1", subsample_type='max_pool', n_stacks=1, n_splices=1, frontend_conv=None,
            task_specific_layer=False, param_init='xavier_uniform', clamp_len=10, lookahead="0_0", chunk_size_left="0",
            chunk_size_current="0", chunk_size_right="0", streaming_type='mask')
CONV = dict(input_dim=80, in_channel=1, channels="32_32", kernel_sizes="(3,3)_(3,3)", strides="(1,1)_(1,1)",
            poolings="(2,2)_(2,2)", dropout=0.0, normalization='', residual=False, bottleneck_dim=32, param_init=0.1)


@pytest.mark.parametrize("ov", [dict(), dict(enc_type='conv_uni_conformer'), dict(enc_type='conv_conformer_v2'),
                                dict(enc_type='conformer', kernel_size=3, pe_type='relative_xl')])
def test_batchnorm_training_matches_reference(ov, monkeypatch):
    import ops_doubles
    from oracle.ref_import import import_reference
    import_reference()
    from neural_sp_b200.encoders.conformer import ConformerEncoder
    from neural_sp_b200.encoders.conv import ConvEncoder
    ops_doubles.install_training(monkeypatch)
    torch.manual_seed(0)
    a_ref = dict(BASE)
    a_ref.update(ov)
    a_our = dict(a_ref)
    if 'conv' in a_ref['enc_type']:
        a_ref['frontend_conv'] = importlib.import_module('neural_sp.models.seq2seq.encoders.conv').ConvEncoder(**CONV)
        a_our['frontend_conv'] = ConvEncoder(**CONV)
    ref = importlib.import_module('neural_sp.models.seq2seq.encoders.conformer').ConformerEncoder(**a_ref)
    ours = ConformerEncoder(**a_our)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.set_precision("fp32")
    ref.train(), ours.train()
    rng = np.random.RandomState(3)
    for step in range(2):                    # two steps: the running statistics accumulate
        xs = torch.from_numpy(rng.randn(3, 60, 80).astype(np.float32))
        xs[1, 50:] = 0
        xs[2, 41:] = 0
        xlens = torch.IntTensor([60, 50, 41])
        for m in (ref, ours):
            m.zero_grad()
        r = ref(xs.clone(), xlens.clone(), task='all')['ys']
        o = ours(xs.clone(), xlens.clone(), task='all')['ys']
        assert torch.equal(r['xlens'], o['xlens'])
        assert float((r['xs'] - o['xs']).abs().max()) <= 1e-4 * float(r['xs'].abs().max())
        w = torch.from_numpy(np.random.RandomState(5).randn(*r['xs'].shape).astype(np.float32))
        for b, n in enumerate(r['xlens'].tolist()):
            w[b, n:] = 0
        (r['xs'] * w).sum().backward()
        (o['xs'] * w).sum().backward()
        rg = dict(ref.named_parameters())
        gmax = max(float(p.grad.abs().max()) for p in rg.values() if p.grad is not None)
        bad = []
        for k, p in ours.named_parameters():
            g = rg[k].grad
            if g is None:
                continue
            assert p.grad is not None, k
            e = float((p.grad - g).abs().max() / max(float(g.abs().max()), 1e-3 * gmax))
            if not e <= 1e-3:
                bad.append((k, e))
        assert not bad, (step, bad[:8], len(bad))
        rb = dict(ref.named_buffers())
        for k, buf in ours.named_buffers():
            if 'running' in k:
                assert torch.allclose(buf, rb[k], atol=1e-5, rtol=1e-4), (step, k)
            if 'num_batches_tracked' in k:
                assert int(buf) == int(rb[k]) == step + 1
