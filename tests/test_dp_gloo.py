"""CPU, world_size 2, gloo: the data-parallel harness (graph sharding + flat-buffer gradient all-reduce)
that bench.py uses with NCCL on the GPU box.  The conv kernels themselves need CUDA, so the model
here is a plain torch module: what is under test is the N>1 host logic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alignn_b200 import dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(6, 8)
        self.b = torch.nn.Linear(8, 1)
        self.unused = torch.nn.Linear(3, 3)  # never receives a gradient (like the dead bn_edges pairs, App. D-11)

    def forward(self, x):
        return self.b(torch.nn.functional.silu(self.a(x)))


def _make_model():
    torch.manual_seed(0)
    return _Net()


def _worker(rank, world, port, ret):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, _, w = dp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    model = _make_model()
    if rank == 1:                               # ranks start different; broadcast must fix that
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    dp.broadcast_parameters(model)
    torch.manual_seed(1)
    X, Y = torch.randn(10, 6), torch.randn(10, 1)                 # 10 "graphs" in the global batch
    mine = dp.shard_range(10, rank, world)
    reducer = dp.FlatGradAllReducer(model.parameters())
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    for step in range(3):
        reducer.zero_grad()
        idx = list(mine)
        # per-rank mean loss weighted so that the average over ranks equals the global mean
        loss = ((model(X[idx]) - Y[idx]) ** 2).sum() / 10 * world
        loss.backward()
        reducer.all_reduce()
        opt.step()
    assert model.unused.weight.grad is None                        # optimizers skip it, as under the reference's DDP
    assert reducer.nbytes() == sum(p.numel() for n, p in model.named_parameters() if not n.startswith("unused")) * 4
    ret[rank] = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    dist.destroy_process_group()


def test_flat_allreduce_matches_single_process_training():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert torch.equal(ret[0], ret[1])                             # replicas stay in lock-step
    # single-process reference on the full batch
    model = _make_model()
    torch.manual_seed(1)
    X, Y = torch.randn(10, 6), torch.randn(10, 1)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    for step in range(3):
        opt.zero_grad(set_to_none=True)
        (((model(X) - Y) ** 2).sum() / 10).backward()
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert torch.allclose(ret[0], ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n,world", [(64, 8), (10, 4), (3, 8), (0, 2)])
def test_shard_range_partitions_exactly(n, world):
    parts = [dp.shard_range(n, r, world) for r in range(world)]
    flat = [i for p in parts for i in p]
    assert flat == list(range(n))
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1


def test_single_process_is_a_noop_group():
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    assert dp.init_from_env("gloo") == (0, 0, 1)
    m = _make_model()
    red = dp.FlatGradAllReducer(m.parameters())
    red.zero_grad()
    m(torch.randn(4, 6)).sum().backward()
    assert red.all_reduce() is None and red.flat is not None
    g0 = m.a.weight.grad
    assert g0.data_ptr() >= red.flat.data_ptr()                   # grads are views into the flat buffer
    red.zero_grad()
    assert m.a.weight.grad is None                                 # autograd writes fresh gradients next step
    m(torch.randn(4, 6)).sum().backward()
    red.all_reduce()
    assert m.a.weight.grad.data_ptr() == g0.data_ptr()             # ... which land in the same flat slice
