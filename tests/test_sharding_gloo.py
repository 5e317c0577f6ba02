"""World-size-2 gloo test (CPU) of the multi-GPU host logic: shards partition the clips, no item is lost or
duplicated, and the reported throughput is total items / max-over-ranks time."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dmm_net_b200.sharding import FlatGradBucket, aggregate_throughput, max_over_ranks, shard_indices


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_items, rank, world)
    # every rank "processes" its clips: here just a checksum over the indices
    local = torch.tensor([float(sum(mine)), float(len(mine))], dtype=torch.float64)
    dist.all_reduce(local)
    elapsed = 10.0 * (rank + 1)                                   # rank 1 is the slow one
    worst = max_over_ranks(elapsed)
    rate = aggregate_throughput(len(mine), elapsed)
    torch.save({"mine": mine, "sum": local.tolist(), "worst": worst, "rate": rate}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    n_items, world = 37, 2
    mp.spawn(_worker, args=(world, _free_port(), n_items, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    allidx = sorted(res[0]["mine"] + res[1]["mine"])
    assert allidx == list(range(n_items))                          # a partition: nothing lost, nothing duplicated
    assert abs(len(res[0]["mine"]) - len(res[1]["mine"])) <= 1
    for r in res:
        assert r["sum"] == [float(sum(range(n_items))), float(n_items)]
        assert r["worst"] == 20.0
        assert abs(r["rate"] - n_items / 20e-3) < 1e-6


def test_single_process_identity():
    assert shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert max_over_ranks(3.5) == 3.5
    assert aggregate_throughput(10, 100.0) == 100.0


def _grad_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                           # same weights on both ranks
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    bucket = FlatGradBucket(net.parameters())
    views = [p.grad.data_ptr() for p in net.parameters()]
    x = torch.full((4, 6), float(rank + 1))                       # different data per rank
    for _ in range(2):                                             # second step: grads still live in the bucket
        bucket.zero()
        net(x).pow(2).sum().backward()
        local = bucket.flat.clone()
        bucket.rendezvous()
        bucket.all_reduce_mean()
    assert [p.grad.data_ptr() for p in net.parameters()] == views  # autograd accumulated in place: still the views
    torch.save({"local": local, "mean": bucket.flat.clone(), "nbytes": bucket.nbytes}, os.path.join(out_dir, f"g{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_bucket_all_reduce(tmp_path):
    """The training step's only collective (examples/synthetic_train_step.py, bench.py train leg): every gradient is a
    view into one flat buffer, one all-reduce averages them over the ranks."""
    world = 2
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"g{r}.pt")) for r in range(world)]
    want = (res[0]["local"] + res[1]["local"]) / 2
    assert not torch.equal(res[0]["local"], res[1]["local"])
    for r in res:
        assert torch.allclose(r["mean"], want, rtol=0, atol=1e-7)
        assert r["nbytes"] == 4 * (6 * 5 + 5 + 5 * 3 + 3)
