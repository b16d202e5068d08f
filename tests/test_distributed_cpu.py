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
