"""N>1 host logic on CPU: world_size-2 gloo, sample sharding + the single padded gather of logits."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from insmos_b200 import distributed as D


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = D.shard_samples(5, rank, world)
        g = torch.Generator().manual_seed(rank)
        logits = torch.randn((1000 + 137 * rank, 3), generator=g)
        parts = D.gather_logits(logits, world, pad_rows=2048)
        ok = len(parts) == world
        g = D.gather_logits_padded(logits, world, 2048)                       # counts travel in the payload, no size exchange
        ok = ok and g.counts() == [1000 + 137 * r for r in range(world)] and tuple(g.buf.shape) == (world, 2049, 3)
        ok = ok and bool((g.buf[rank, 1000 + 137 * rank:2048] == 0).all())
        for r in range(world):
            exp = torch.randn((1000 + 137 * r, 3), generator=torch.Generator().manual_seed(r))
            ok = ok and torch.equal(parts[r], exp)
        ret[rank] = (ok, mine)
    finally:
        dist.destroy_process_group()


def test_shard_and_gather_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29611, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] == [0, 2, 4] and ret[1][1] == [1, 3]


def test_single_rank_is_identity():
    x = torch.randn(10, 3)
    assert D.gather_logits(x, 1)[0] is x
    assert D.shard_samples(8, 3, 8) == [3]


# ---- training step (N3): the flat gradient buffer and its single all-reduce, world 2 over gloo -------------------------
def _train_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from insmos_b200.train import TrainStep
        torch.manual_seed(0)                                                   # replicated weights
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
        ts = TrainStep(net, lr=1e-3)
        ok = ts.world == world and ts.flat.numel == sum(p.numel() for p in net.parameters())
        ok = ok and all(p.data_ptr() >= ts.flat.data.data_ptr() for p in net.parameters())
        x = torch.randn((8, 6), generator=torch.Generator().manual_seed(100 + rank))      # each rank its own samples
        ts.flat.zero_grad()
        net(x).pow(2).sum().backward()
        local = ts.flat.grad.clone()
        ts.all_reduce_gradients()                                               # ONE collective on the flat buffer
        ret[rank] = (ok, local, ts.flat.grad.clone())
        try:
            ts.optimizer_step()                                                 # the Adam kernel is CUDA-only: must refuse, not fall back
            ret["adam_cpu_%d" % rank] = "no error"
        except (ValueError, RuntimeError) as e:
            ret["adam_cpu_%d" % rank] = type(e).__name__
    finally:
        dist.destroy_process_group()


def test_flat_gradient_all_reduce_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_train_worker, args=(world, 29613, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0]
    total = ret[0][1] + ret[1][1]
    assert torch.allclose(ret[0][2], total) and torch.equal(ret[0][2], ret[1][2])      # sum on every rank; 1/world is folded into Adam
    assert ret["adam_cpu_0"] in ("ValueError", "RuntimeError") and ret["adam_cpu_1"] in ("ValueError", "RuntimeError")
