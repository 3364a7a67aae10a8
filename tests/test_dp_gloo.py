"""CPU, world_size 2, gloo: the host-side logic of the batch-sharded data parallel path
(shard_batch, broadcast, bucketed overlapped gradient all-reduce) -- no GPU involved."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mamba_tts_project_b200.dp import GradAllReducer, broadcast_parameters, shard_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16),
                               torch.nn.Tanh(), torch.nn.Linear(16, 4))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _make_model(seed=100 + rank)          # replicas start different ...
        broadcast_parameters(model)                    # ... and are made identical
        reducer = GradAllReducer(model, bucket_bytes=600)   # tiny buckets -> several all-reduces
        assert len(reducer.buckets) >= 3
        g = torch.Generator().manual_seed(0)
        x, y = torch.randn(8, 8, generator=g), torch.randn(8, 4, generator=g)
        xs, ys = shard_batch([x, y], rank, world)
        for _ in range(2):                             # two steps: the reducer re-arms itself
            model.zero_grad(set_to_none=True)
            # sum of local squared errors, scaled so that the averaged gradient equals the
            # gradient of the GLOBAL mean loss (scale sums, not local means)
            loss = ((model(xs) - ys) ** 2).sum() / (x.shape[0] * 4) * world
            loss.backward()
            reducer.finish()
        if rank == 0:
            torch.save([p.grad.clone() for p in model.parameters()], out)
    finally:
        dist.destroy_process_group()


def test_sharded_gradients_equal_full_batch(tmp_path):
    out = str(tmp_path / "grads.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    # single-process reference: rank 0's initial weights, the whole batch
    model = _make_model(seed=100)
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(8, 8, generator=g), torch.randn(8, 4, generator=g)
    ((model(x) - y) ** 2).mean().backward()
    for a, p in zip(got, model.parameters()):
        assert torch.allclose(a, p.grad, rtol=1e-5, atol=1e-6)


def test_shard_batch_rejects_ragged():
    import pytest
    with pytest.raises(ValueError):
        shard_batch([torch.zeros(5, 2)], 0, 2)
    a, b = shard_batch([torch.arange(6), None], 1, 3)
    assert a.tolist() == [2, 3] and b is None
