"""world_size-2 gloo tests of the data-parallel plumbing (cnn_graph_b200/dist.py): batch shard
bounds and the single flat gradient all-reduce.  CPU only."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    from cnn_graph_b200 import dist as cgdist
    r, w, _ = cgdist.init_from_env('gloo')
    assert (r, w) == (rank, world)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2))]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(5, dtype=torch.float32) * (rank + 1)
    # params[2] has no gradient on purpose (e.g. an unused variable)
    cgdist.GradAllReducer(average=True)(params)
    ok = torch.allclose(params[0].grad, torch.full((3, 4), 1.5)) and \
        torch.allclose(params[1].grad, torch.arange(5, dtype=torch.float32) * 1.5) and params[2].grad is None
    mx = cgdist.max_over_ranks(10.0 + rank, torch.device('cpu'))
    cgdist.barrier()
    out[rank] = bool(ok) and mx == 11.0
    torch.distributed.destroy_process_group()


def test_grad_allreduce_world2_gloo():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        assert dict(out) == {0: True, 1: True}


@pytest.mark.parametrize('n,world', [(100, 8), (7, 2), (3, 4), (0, 2), (1024, 3)])
def test_shard_bounds_partition(n, world):
    from cnn_graph_b200 import dist as cgdist
    spans = [cgdist.shard_bounds(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [e - b for b, e in spans]
    assert max(sizes) - min(sizes) <= 1
