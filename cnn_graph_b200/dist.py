"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over
NVLink/NVSwitch on the B200 box, gloo in CPU tests).

The Chebyshev hot path shards over the batch with no data-path collective: L~ and the
weights are replicated, every sample (every column of the signal slab) is independent
through filter, activation and pooling.  The only exchange per step is the
sum of the weight gradients, done here as ONE all-reduce over a single flat fp32 bucket
(launch-latency bound at ~8 MB for the MNIST-shaped model, so one bucket, not many).
"""
import os
import sys

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment; returns
    (rank, world_size, local_rank).  Single-process runs return (0, 1, 0) untouched."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
        if backend == 'nccl':
            # NCCL printf()s its version banner to stdout when the communicator is created (NCCL_DEBUG=WARN / VERSION).
            # Create it now, with file descriptor 1 pointed at stderr, so that stdout stays the caller's (bench.py
            # prints exactly one JSON line there).
            sys.stdout.flush()
            saved = os.dup(1)
            try:
                os.dup2(2, 1)
                t = torch.zeros(1, device=torch.device('cuda', local_rank))
                dist.all_reduce(t)
                torch.cuda.synchronize()
            finally:
                os.dup2(saved, 1)
                os.close(saved)
    return rank, world, local_rank


def shard_bounds(n_items, rank, world):
    """Contiguous [begin, end) shard of ``n_items`` for ``rank`` (remainder spread over the
    first ranks) -- the batch partition of the data-parallel path."""
    base, rem = divmod(int(n_items), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class GradAllReducer:
    """Averages gradients across ranks with one flat all-reduce per step."""

    def __init__(self, average=True):
        self.average = average
        self._flat = None

    def __call__(self, params):
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return
        total = sum(g.numel() for g in grads)
        if self._flat is None or self._flat.numel() != total or self._flat.device != grads[0].device:
            self._flat = torch.empty(total, dtype=torch.float32, device=grads[0].device)
        flat = self._flat
        views, offset = [], 0
        for g in grads:
            views.append(flat[offset:offset + g.numel()].view_as(g))
            offset += g.numel()
        torch._foreach_copy_(views, grads)                       # one multi-tensor kernel each way
        if self.average and dist.get_backend() == 'nccl':
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            if self.average:
                flat.div_(dist.get_world_size())
        torch._foreach_copy_(grads, views)


class OverlappedGradAllReducer:
    """Gradient averaging that overlaps with the backward pass.

    Parameters with at least `min_numel` elements (for cgcnn: the 3968 x 512 fc weight, 97 % of the bytes, whose
    gradient is ready first) are all-reduced IN PLACE from a post-accumulate-grad hook, asynchronously, while autograd
    keeps running the graph-conv backward kernels; the remaining small gradients go through one flat bucket after
    the backward pass, and the step waits for the outstanding work before the optimiser runs.  NCCL averages in
    the collective (ReduceOp.AVG); other backends sum and divide."""

    def __init__(self, params, average=True, min_numel=1 << 16):
        self.average = average
        self.params = list(params)
        self.big = [p for p in self.params if p.numel() >= min_numel]
        big_ids = {id(p) for p in self.big}
        self.small = [p for p in self.params if id(p) not in big_ids]
        self._works = []
        self._flat = GradAllReducer(average=average)
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.big]

    def _active(self):
        return dist.is_initialized() and dist.get_world_size() > 1

    def _avg_in_collective(self):
        return self.average and dist.get_backend() == 'nccl'

    def _on_grad(self, p):
        if not self._active() or p.grad is None:
            return
        op = dist.ReduceOp.AVG if self._avg_in_collective() else dist.ReduceOp.SUM
        self._works.append((dist.all_reduce(p.grad, op=op, async_op=True), p))

    def __call__(self, params=None):
        if not self._active():
            return
        self._flat(self.small)
        for work, p in self._works:
            work.wait()
            if self.average and not self._avg_in_collective():
                p.grad.div_(dist.get_world_size())
        self._works = []

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


class DeferredGradAllReducer:
    """Data-parallel averaging that takes the LARGE gradients off the critical path without changing the trajectory.

    After the backward pass of step t the small gradients are averaged (one flat all-reduce) and applied as usual; the
    gradients of the variables with at least `min_numel` elements (cgcnn: the first dense weight, 97 % of the bytes) are
    only copied aside.  Their all-reduce and momentum-SGD update run at the START of step t + 1 on a side stream, under
    the graph-convolution forward kernels -- whose first kernel leaves SMs free, unlike the persistent backward kernels
    that made an all-reduce overlapped with the backward pass slower -- and the model joins the side stream right before
    the dense head first reads those variables (`join`).  Every variable therefore has received the update of step t
    before step t + 1 uses it: same numbers as the plain reducer, step for step.

    The update uses the learning rate of the step that produced the gradient, read from a device scalar so that a replayed
    CUDA graph follows the staircase schedule.  `flush()` applies a pending update at once (evaluation, checkpoints).
    `force=True` runs the deferral without a process group (single-process tests of the capture logic)."""

    def __init__(self, model, min_numel=1 << 16, average=True, force=False):
        self.model, self.average, self.force = model, average, force
        params = list(model.store.parameters())
        self.big = [p for p in params if p.numel() >= min_numel]
        big_ids = {id(p) for p in self.big}
        self.small = [p for p in params if id(p) not in big_ids]
        self.bufs = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.big]
        self.valid = False                     # a gradient is waiting in bufs
        self.lr = None                         # learning rate of the step that produced it
        dev = self.big[0].device if self.big else torch.device('cpu')
        self.lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.side = torch.cuda.Stream(device=dev) if dev.type == 'cuda' else None
        self._flat = GradAllReducer(average=average)
        self._launched = False

    def active(self):
        return bool(self.big) and (self.force or (dist.is_initialized() and dist.get_world_size() > 1))

    def set_lr(self, lr):
        """Host side, outside any captured graph: the rate the next deferred update will read."""
        if self.lr != lr:
            self.lr = lr
            self.lr_dev.fill_(float(lr))

    # ---- pieces called by GraphModel.train_step
    def begin_step(self):
        if not self.active() or not self.valid:
            return
        if self.side is None:
            self._apply()
            return
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            self._apply()
        self._launched = True
        if not getattr(self.model, 'joins_deferred_update', False):
            self.join()                        # the model does not say where it first reads the large variables

    def join(self):
        if self._launched:
            torch.cuda.current_stream().wait_stream(self.side)
            self._launched = False

    def __call__(self, params=None):
        """After the backward pass: average + keep the small gradients, set the large ones aside."""
        if not self.active():
            return
        self.join()
        if dist.is_initialized() and dist.get_world_size() > 1:
            self._flat(self.small)
        for p, b in zip(self.big, self.bufs):
            if p.grad is not None:
                b.copy_(p.grad)
                p.grad = None                  # the optimiser step that follows skips it
        self.valid = True

    def flush(self):
        if self.active() and self.valid:
            self.join()
            self._apply()

    # ---- internals
    @torch.no_grad()
    def _apply(self):
        if dist.is_initialized() and dist.get_world_size() > 1:
            for b in self.bufs:
                if self.average and dist.get_backend() == 'nccl':
                    dist.all_reduce(b, op=dist.ReduceOp.AVG)
                else:
                    dist.all_reduce(b, op=dist.ReduceOp.SUM)
                    if self.average:
                        b.div_(dist.get_world_size())
        opt = self.model.optimizer
        momentum = opt.param_groups[0].get('momentum', 0.0)
        if hasattr(opt, 'apply'):              # ops.NativeMomentumSGD: one launch, rate from the device scalar
            opt.apply(self.big, self.bufs, self.lr, momentum, lr_dev=self.lr_dev)
        else:                                  # torch.optim.SGD (CPU / gloo tests): the same update by hand
            for p, b in zip(self.big, self.bufs):
                st = opt.state[p]
                mb = st.get('momentum_buffer')
                if momentum != 0:
                    if mb is None:
                        mb = st['momentum_buffer'] = torch.clone(b).detach()
                    else:
                        mb.mul_(momentum).add_(b)
                    p.add_(mb, alpha=-self.lr)
                else:
                    p.add_(b, alpha=-self.lr)
        self.valid = False

    def snapshot(self):
        return {'bufs': [b.clone() for b in self.bufs], 'valid': self.valid, 'lr': self.lr}

    def restore(self, state):
        with torch.no_grad():
            for b, old in zip(self.bufs, state['bufs']):
                b.copy_(old)
        self.valid = state['valid']
        self.lr = None
        if state['lr'] is not None:
            self.set_lr(state['lr'])


def max_over_ranks(value, device):
    """Max of a python float over all ranks (device-side all-reduce)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
