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
