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


def _overlap_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    from cnn_graph_b200 import dist as cgdist
    cgdist.init_from_env('gloo')
    torch.manual_seed(0)
    big = torch.nn.Parameter(torch.randn(40, 50))          # >= min_numel: all-reduced from the autograd hook
    small = torch.nn.Parameter(torch.randn(7))             # flat bucket after the backward pass
    unused = torch.nn.Parameter(torch.randn(3))
    red = cgdist.OverlappedGradAllReducer([big, small, unused], average=True, min_numel=1000)
    ok = True
    for step in range(2):                                   # twice: the hooks and the bucket are reusable
        for q in (big, small, unused):
            q.grad = None
        x = torch.full((50,), float(rank + 1 + step))
        loss = (big @ x).sum() * (rank + 1) + (small * (rank + 2)).sum()
        loss.backward()
        red()
        # d/d big = (rank+1) * x broadcast over rows; averaged over ranks 0, 1
        want_big = sum((r + 1) * float(r + 1 + step) for r in range(world)) / world
        want_small = sum(r + 2 for r in range(world)) / world
        ok = ok and torch.allclose(big.grad, torch.full((40, 50), want_big)) and \
            torch.allclose(small.grad, torch.full((7,), want_small)) and unused.grad is None
    red.remove()
    out[rank] = bool(ok)
    torch.distributed.destroy_process_group()


def test_overlapped_grad_allreduce_world2_gloo():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_overlap_worker, args=(2, port, out), nprocs=2, join=True)
        assert dict(out) == {0: True, 1: True}


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


# ---------------------------------------------------------------- row partition (config C5) -- host logic
def _partition_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    import numpy as np
    import scipy.sparse
    from cnn_graph_b200 import dist as cgdist, partition
    from oracle import graph_ref
    cgdist.init_from_env('gloo')
    rng = np.random.RandomState(3)
    M, C, K = 203, 6, 7
    A = scipy.sparse.random(M, M, density=0.04, random_state=rng, format='csr', dtype=np.float32)
    L = scipy.sparse.csr_matrix(0.1 * (A + A.T), dtype=np.float32)
    X = rng.standard_normal((M, C)).astype(np.float32)
    pb_holder = {}

    def step(x1_ext, x0, alpha):          # host stand-in for cg_cheb_step on the padded local operator
        part = pb_holder['pb'].part
        y = alpha * (part.local @ x1_ext.numpy())[:part.nloc]
        if x0 is not None:
            y = y - x0.numpy()
        return torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))

    pb = partition.PartitionedBasis(L, device=torch.device('cpu'), step_fn=step)
    pb_holder['pb'] = pb
    r0, r1 = pb.part.r0, pb.part.r1
    got = pb.basis(torch.from_numpy(X[r0:r1].copy()), K).numpy()
    ref = graph_ref.chebyshev(L, X, K)[:, r0:r1]
    err = float(np.abs(got - ref).max()) / float(np.abs(ref).max())
    out[rank] = (err < 1e-5, pb.part.nhalo > 0, sum(pb.part.recv_counts) == pb.part.nhalo)
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_row_partitioned_basis_matches_global(world):
    """Halo lists, column remapping and the all-to-all of packed rows reproduce graph.chebyshev on every block."""
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_partition_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {r: (True, True, True) for r in range(world)}


def _filter_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    import numpy as np
    import scipy.sparse
    from cnn_graph_b200 import dist as cgdist, partition
    from oracle import tf_ref
    cgdist.init_from_env('gloo')
    rng = np.random.RandomState(5)
    M, Fin, Fout, K = 157, 5, 7, 6
    A = scipy.sparse.random(M, M, density=0.05, random_state=rng, format='csr', dtype=np.float32)
    Lr = scipy.sparse.csr_matrix(0.1 * A, dtype=np.float32)              # directed on purpose: the backward needs L~^T
    x = rng.standard_normal((M, Fin)).astype(np.float32)
    W = (0.3 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    gy = rng.standard_normal((M, Fout)).astype(np.float32)
    holder = {}

    def make_step(which):
        def step(x1_ext, x0, alpha):      # host stand-in for cg_cheb_step on the padded local operator
            part = getattr(holder['pf'], which).part
            y = alpha * (part.local @ x1_ext.numpy())[:part.nloc]
            if x0 is not None:
                y = y - x0.numpy()
            return torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))
        return step

    def contract(stack, Wt, transposed):  # host stand-ins for cg_cheb_contract / cg_cheb_contract_dw
        Wk = Wt.reshape(Fin, K, Fout)                                     # row f*K + k
        return torch.einsum('krf,fko->ro', stack, Wk) if not transposed else torch.einsum('kro,fko->rf', stack, Wk)

    def dw(stack, g):
        return torch.einsum('krf,ro->fko', stack, g).reshape(Fin * K, Fout)

    pf = partition.PartitionedFilter(Lr, K, device=torch.device('cpu'), step_fn=make_step('fwd'),
                                     step_fn_t=make_step('bwd'), contract_fn=contract, dw_fn=dw)
    holder['pf'] = pf
    r0, r1 = pf.part.r0, pf.part.r1
    y = pf.forward(torch.from_numpy(x[r0:r1].copy()), torch.from_numpy(W)).numpy()
    dx, dW = pf.backward(torch.from_numpy(gy[r0:r1].copy()))
    # oracle on the whole graph (N = 1): lmax = 2 with an already rescaled operator means L~ = Lr + I - I ... so feed
    # the un-rescaled twin Lr + I, whose rescaling (L - I) is Lr
    Lfull = scipy.sparse.csr_matrix(Lr + scipy.sparse.identity(M, dtype=np.float32, format='csr'))
    ref_y = tf_ref.chebyshev5(x[None], Lfull, W, K)[0]
    ref_dx, ref_dW = tf_ref.chebyshev5_backward(x[None], Lfull, W, K, gy[None])

    def rel(a, b):
        return float(np.abs(a - b).max()) / max(float(np.abs(b).max()), 1e-30)

    out[rank] = (rel(y, ref_y[r0:r1]) < 1e-5, rel(dx.numpy(), ref_dx[0][r0:r1]) < 1e-5, rel(dW.numpy(), ref_dW) < 1e-5)
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_row_partitioned_filter_forward_backward(world):
    """Row-partitioned filter (config C5) against the oracle on the whole graph: y and dx block by block, dW after
    the all-reduce; a directed operator checks that the backward exchanges the halo of L~^T."""
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_filter_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {r: (True, True, True) for r in range(world)}


def _subgroup_worker(rank, world, port, out):
    """The filter partitioned over a NON-default subgroup {1, 2} of a 3-rank world: halo exchange and the dW
    all-reduce must both run on that group with the group's own rank / size."""
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    import numpy as np
    import scipy.sparse
    from cnn_graph_b200 import dist as cgdist, partition
    from oracle import tf_ref
    cgdist.init_from_env('gloo')
    group = torch.distributed.new_group([1, 2])
    if rank == 0:
        out[rank] = (True, True, True)
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
        return
    rng = np.random.RandomState(9)
    M, Fin, Fout, K = 91, 3, 4, 5
    A = scipy.sparse.random(M, M, density=0.06, random_state=rng, format='csr', dtype=np.float32)
    Lr = scipy.sparse.csr_matrix(0.1 * A, dtype=np.float32)
    x = rng.standard_normal((M, Fin)).astype(np.float32)
    W = (0.3 * rng.standard_normal((Fin * K, Fout))).astype(np.float32)
    gy = rng.standard_normal((M, Fout)).astype(np.float32)
    holder = {}

    def make_step(which):
        def step(x1_ext, x0, alpha):
            part = getattr(holder['pf'], which).part
            y = alpha * (part.local @ x1_ext.numpy())[:part.nloc]
            if x0 is not None:
                y = y - x0.numpy()
            return torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))
        return step

    def contract(stack, Wt, transposed):
        Wk = Wt.reshape(Fin, K, Fout)
        return torch.einsum('krf,fko->ro', stack, Wk) if not transposed else torch.einsum('kro,fko->rf', stack, Wk)

    def dw(stack, g):
        return torch.einsum('krf,ro->fko', stack, g).reshape(Fin * K, Fout)

    pf = partition.PartitionedFilter(Lr, K, device=torch.device('cpu'), step_fn=make_step('fwd'), step_fn_t=make_step('bwd'),
                                     contract_fn=contract, dw_fn=dw, group=group)
    holder['pf'] = pf
    assert (pf.part.rank, pf.part.world) == (rank - 1, 2)
    r0, r1 = pf.part.r0, pf.part.r1
    y = pf.forward(torch.from_numpy(x[r0:r1].copy()), torch.from_numpy(W)).numpy()
    dx, dW = pf.backward(torch.from_numpy(gy[r0:r1].copy()))
    Lfull = scipy.sparse.csr_matrix(Lr + scipy.sparse.identity(M, dtype=np.float32, format='csr'))
    ref_y = tf_ref.chebyshev5(x[None], Lfull, W, K)[0]
    ref_dx, ref_dW = tf_ref.chebyshev5_backward(x[None], Lfull, W, K, gy[None])

    def rel(a, b):
        return float(np.abs(a - b).max()) / max(float(np.abs(b).max()), 1e-30)

    out[rank] = (rel(y, ref_y[r0:r1]) < 1e-5, rel(dx.numpy(), ref_dx[0][r0:r1]) < 1e-5, rel(dW.numpy(), ref_dW) < 1e-5)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_row_partitioned_filter_on_subgroup():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_subgroup_worker, args=(3, port, out), nprocs=3, join=True)
        assert dict(out) == {r: (True, True, True) for r in range(3)}


def _deferred_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    from cnn_graph_b200 import dist as cgdist
    cgdist.init_from_env('gloo')

    class Store:
        def __init__(self, ps):
            self.ps = ps

        def parameters(self):
            return self.ps

    class Model:
        joins_deferred_update = True

    def make():
        g = torch.Generator().manual_seed(0)
        ps = [torch.nn.Parameter(torch.randn(40, 50, generator=g)), torch.nn.Parameter(torch.randn(7, generator=g)),
              torch.nn.Parameter(torch.randn(3, 5, generator=g))]
        m = Model()
        m.store = Store(ps)
        m.optimizer = torch.optim.SGD(ps, lr=0.1, momentum=0.9)
        return m

    def grads(step):
        g = torch.Generator().manual_seed(100 * step + rank)
        return [torch.randn(40, 50, generator=g), torch.randn(7, generator=g), torch.randn(3, 5, generator=g)]

    lrs = [0.1, 0.1, 0.1, 0.05, 0.05, 0.025]
    plain, deferred = make(), make()
    flat = cgdist.GradAllReducer(average=True)
    hook = cgdist.DeferredGradAllReducer(deferred, min_numel=1000)
    assert hook.active() and len(hook.big) == 1 and len(hook.small) == 2
    for step, lr in enumerate(lrs):
        for m, h in ((plain, flat), (deferred, hook)):
            for grp in m.optimizer.param_groups:
                grp['lr'] = lr
            m.optimizer.zero_grad(set_to_none=True)
            if h is hook:
                h.begin_step()
                h.join()
            for p_, g_ in zip(m.store.parameters(), grads(step)):
                p_.grad = g_.clone()
            h(m.store.parameters())
            m.optimizer.step()
            if h is hook:
                h.set_lr(lr)
        # the small variables agree after every step, the large one is exactly one update behind until flushed
        ok_small = all(torch.allclose(a, b, rtol=0, atol=1e-6) for a, b in zip(plain.store.parameters()[1:], deferred.store.parameters()[1:]))
        assert ok_small, step
    assert hook.valid
    hook.flush()
    out[rank] = all(torch.allclose(a, b, rtol=0, atol=1e-6) for a, b in zip(plain.store.parameters(), deferred.store.parameters()))
    cgdist.barrier()
    torch.distributed.destroy_process_group()


def test_deferred_reducer_same_trajectory_world2():
    """DeferredGradAllReducer: large gradients exchanged and applied at the start of the next step (with that gradient's own
    learning rate, across staircase boundaries) -- same weights as the flat all-reduce after every step + flush."""
    world, port = 2, _free_port()
    with mp.Manager() as manager:
        out = manager.dict()
        mp.spawn(_deferred_worker, args=(world, port, out), nprocs=world, join=True)
        assert all(out.get(r) for r in range(world)), dict(out)
