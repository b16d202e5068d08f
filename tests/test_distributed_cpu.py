"""N>1 host logic on CPU: gloo, world_size 2 (rendezvous on 127.0.0.1)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neural_sp_b200.dist import flat_allreduce_grads, shard_batch
    utts = list(range(10))                       # length-sorted utterance ids
    mine = shard_batch(utts, rank, world)
    torch.manual_seed(0)
    w1, w2 = torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))
    w1.grad = torch.full((3, 2), float(rank + 1))
    w2.grad = torch.arange(5.0) * (rank + 1)
    flat = flat_allreduce_grads([w1, w2])
    q.put((rank, mine, w1.grad.tolist(), w2.grad.tolist(), flat.numel()))
    dist.destroy_process_group()


def test_shard_and_single_flat_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6, 8] and res[1][1] == [1, 3, 5, 7, 9]       # indices[rank::world]
    for r in res:
        assert r[2] == [[3.0, 3.0]] * 3                                          # 1 + 2 summed over ranks
        assert r[3] == [0.0, 3.0, 6.0, 9.0, 12.0]
        assert r[4] == 11                                                        # one flat buffer for all grads


def _worker_buckets(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neural_sp_b200 import autograd as ag
    works = []
    ag.set_grad_sync(lambda flat: works.append(dist.all_reduce(flat, async_op=True)))
    wq, wk, wv = (torch.nn.Parameter(torch.zeros(2, 3)) for _ in range(3))
    b = torch.nn.Parameter(torch.zeros(4))
    G = ag._Grads([wq, wk, wv, b])                     # one node = one flat bucket; q, k, v adjacent
    fused = G.fused([wq, wk, wv])
    assert fused is not None and tuple(fused.shape) == (6, 3)
    fused += float(rank + 1)                           # "kernels" accumulate into the views
    G.buf(b).add_(torch.arange(4.0) * (rank + 1))
    G.put(wk, torch.full((3, 2), 10.0 * (rank + 1)).t())
    grads = [G.get(p) for p in (wq, wk, wv, b)]
    G.done()                                           # bucket handed to the (async) all-reduce
    for w in works:
        w.wait()
    ag.set_grad_sync(None)
    q.put((rank, [g.tolist() for g in grads], G.flat.numel(), len(works)))
    dist.destroy_process_group()


def test_bucketed_grad_sync_world2():
    """Per-node flat gradient buckets: the returned grads are views of the bucket, so the in-place all-reduce of the
    bucket IS the gradient exchange (sum over ranks), one collective per autograd node."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_buckets, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        gq, gk, gv, gb = r[1]
        assert gq == [[3.0] * 3] * 2 and gv == [[3.0] * 3] * 2          # (1) + (2)
        assert gk == [[30.0] * 3] * 2                                    # put() overwrote the view: 10 + 20
        assert gb == [0.0, 3.0, 6.0, 9.0]
        assert r[2] == 22 and r[3] == 1


def _worker_train_step(rank, world, port, q, tests_dir, model="conformer"):
    """One data-parallel training step (encoder + CTC) per rank on its own utterances, gradients exchanged through the bucketed
    grad-sync hook; ops replaced by the torch restatements (CPU)."""
    import sys
    sys.path.insert(0, tests_dir)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    import ops_doubles
    from conftest import load_golden
    from enc_util import build_ours
    from neural_sp_b200 import autograd as ag
    from neural_sp_b200.decoders.ctc import CTC

    class MP:
        def setattr(self, o, n, v):
            setattr(o, n, v)
    real_frontend = ag.frontend_forward          # keep the real CNN node (its bucket goes through the grad-sync hook too);
    ops_doubles.install_training(MP())           # its ops have restatements, only the whole-node double is not wanted here
    ag.frontend_forward = real_frontend
    torch.manual_seed(0)
    if model == "conformer":
        g = load_golden("enc_conformer_small.npz")
        enc = build_ours(g, torch.device("cpu"), "fp32").train()
        odim = 64
    else:                                               # CNN + LSTM stack with projections, max-pool subsampling, bridge
        import json
        from neural_sp_b200.encoders.conv import ConvEncoder
        from neural_sp_b200.encoders.rnn import RNNEncoder
        g = load_golden("rnn_conv_lstm_proj.npz")
        c = json.loads(str(g["cfg"]))
        a = dict(c["args"])
        a["frontend_conv"] = ConvEncoder(**c["conv"]) if c["conv"] else None
        enc = RNNEncoder(**a)
        enc.load_state_dict({k[3:]: torch.from_numpy(np.asarray(g[k]).astype(np.float32)) for k in g.files if k.startswith("sd.")})
        enc.set_precision("fp32")
        enc.train()
        odim = enc.output_dim
    ctc = CTC(eos=2, blank=0, enc_n_units=odim, vocab=40, lsm_prob=0.1, fc_list="32").train()
    ctc.set_precision("fp32")
    xs_all, xlens_all = torch.from_numpy(g["xs"]), g["xlens"].tolist()
    ys_all = [[5, 6, 7, 8], [9, 10, 11], [12, 13]]
    shards = [[0, 2], [1]]                              # utterances per rank
    works = []
    if world > 1:
        ag.set_grad_sync(lambda flat: works.append(dist.all_reduce(flat, async_op=True)))

    def step(idx):
        out = enc(xs_all[idx], torch.IntTensor([xlens_all[i] for i in idx]), task="ys")
        loss, _ = ctc(out["ys"]["xs"], out["ys"]["xlens"], [ys_all[i] for i in idx])
        loss.backward()
        return float(loss.detach())

    params = list(enc.named_parameters()) + [("ctc." + k, p) for k, p in ctc.named_parameters()]
    if world > 1:
        loss = step(shards[rank])
        for w in works:
            w.wait()
        ag.set_grad_sync(None)
        q.put((rank, loss, {k: p.grad.clone().numpy() for k, p in params}))
        dist.destroy_process_group()
    else:                                               # single process: gradients of the SUM of the per-shard mean losses
        tot = {k: torch.zeros_like(p) for k, p in params}
        for sh in shards:
            for _, p in params:
                p.grad = None
            step(sh)
            for k, p in params:
                tot[k] += p.grad
        q.put((0, 0.0, {k: v.numpy() for k, v in tot.items()}))


import pytest  # noqa: E402


@pytest.mark.parametrize("model", ["conformer", "conv_lstm"])
def test_data_parallel_training_step_world2(model):
    """every rank back-propagates its own utterances; each autograd node's flat gradient bucket is all-reduced
    (sum) as soon as its backward is enqueued; the result equals one process accumulating the per-shard mean losses."""
    import numpy as np
    tests_dir = os.path.dirname(os.path.abspath(__file__))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p0 = ctx.Process(target=_worker_train_step, args=(0, 1, 0, q, tests_dir, model))
    p0.start()
    ref = q.get(timeout=300)[2]
    p0.join(timeout=60)
    world, port = 2, _free_port()
    procs = [ctx.Process(target=_worker_train_step, args=(r, world, port, q, tests_dir, model)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gmax = max(float(np.abs(v).max()) for v in ref.values())
    for rank, _, grads in res:
        for k, v in ref.items():
            err = float(np.abs(grads[k] - v).max()) / max(float(np.abs(v).max()), 1e-3 * gmax)
            assert err <= 1e-4, (rank, k, err)
