"""bench.py arm for BASELINE.json configs[4] (c5): one Chebyshev filter (Fin = Fout = 64, K = 20) forward + backward on
a large synthetic kNN graph (2^20 vertices, exact 16-NN symmetrised => ~18 entries / row, Morton order), the rows of L~
and of every X_k partitioned over the GPUs with a halo exchange per recurrence step (cnn_graph_b200/partition.py,
SURVEY.md 8(e)(2)).  The operator does not fit shared memory, so the recurrence is one CSR step per launch from HBM
(k_spmm_step) -- the genuinely HBM-bound SpMM of the metric.  One "step" = forward (K-1 recurrence steps + contraction)
and backward (dW + all-reduce; K-1 steps on L~^T + contraction for dx) of one signal; total work is fixed as GPUs are added
("scaling": "strong")."""
import json
import os
import time

import numpy as np

from . import common, workloads
from .common import METRIC


def workload_name(args):
    return ('C5 large synthetic kNN graph: 2^%d vertices, 16-NN (~18 nnz/row), %s order, Fin=Fout=64, K=%d, one filter forward + '
            'backward (dx, dW), row-partitioned L~ with halo exchange' % (args.log2m, args.order, args.K or 20))


def cpu_filter_steps(log2m, order, K, F, steps, warmup):
    """oracle/tf_ref.py on the same kind of graph: scipy CSR x dense recurrence (single thread), numpy BLAS contraction,
    backward from the kept stack (dW) and the adjoint recurrence (dx)."""
    from oracle import tf_ref
    L = workloads.knn_graph_laplacian(log2m, 16, order)
    M = L.shape[0]
    rng = np.random.RandomState(0)
    x = rng.standard_normal((1, M, F)).astype(np.float32)
    W = (0.05 * rng.standard_normal((F * K, F))).astype(np.float32)
    gy = rng.standard_normal((1, M, F)).astype(np.float32)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        y, kept = tf_ref.chebyshev5(x, L, W, K, return_stack=True)
        dx, dW = tf_ref.chebyshev5_backward(x, L, W, K, gy, a=kept)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times, M


def run_reference(args, config):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    K = args.K or 20
    log2m = min(args.log2m, 17)              # bounded sample: a 2^17-vertex graph of the same construction, scaled by rows
    times, M = cpu_filter_steps(log2m, args.order, K, 64, args.steps, args.warmup)
    scale = float(1 << log2m) / float(1 << args.log2m)
    total = float(np.sum(times))
    value = len(times) / total * scale
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times) / scale, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'note': 'CPU sample: the same filter on a 2^%d-vertex graph of the same construction, '
                   'time scaled by the vertex ratio (work is linear in M); oracle/tf_ref.py (scipy SpMM single-threaded, numpy BLAS all cores)' % log2m},
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': os.cpu_count(), 'kind': 'port',
                         'sample': '%d filter forward+backward passes on 2^%d vertices, scaled x%g' % (len(times), log2m, scale)},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}), flush=True)


def run_ours(args, config):
    import torch
    from cnn_graph_b200 import _native, dist as cgdist, ops, partition

    rank, world, local_rank = cgdist.init_from_env('nccl')
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the hot path has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    lib = _native.lib()
    K, F = args.K or 20, 64
    M = 1 << args.log2m
    t0 = time.time()
    L = workloads.knn_graph_laplacian(args.log2m, 16, args.order)
    Lr = ops.rescale_csr(L, 2)
    t_build = time.time() - t0
    pf = partition.PartitionedFilter(Lr, K, exchange=os.environ.get('CG_C5_EXCHANGE', 'peer'))
    part = pf.part
    gen = torch.Generator().manual_seed(7)
    x_host = torch.randn(M, F, generator=gen)[part.r0:part.r1].contiguous().pin_memory()
    gy_host = torch.randn(M, F, generator=gen)[part.r0:part.r1].contiguous().pin_memory()
    W = (0.05 * torch.randn(F * K, F, generator=gen)).to(device)
    x_loc, gy_loc = x_host.to(device), gy_host.to(device)
    timer = common.Timer(device, local_rank)
    result = {}

    def step():
        y = pf.forward(x_loc, W)
        dx, dW = pf.backward(gy_loc)
        result['y'], result['dx'], result['dW'] = y, dx, dW

    n0 = lib.cg_launch_count()
    step()
    native_per_step = int(lib.cg_launch_count() - n0)
    W_ = max(args.warmup, 3)
    ms_total, clocks, _ = timer.run(step, args.steps, W_, sample_clocks=True)
    value = args.steps / (ms_total * 1e-3)

    # adjoint identity across the partition: <y, gy> = <W, dW> (the filter is linear in W), summed over ranks
    lhs = (result['y'].double() * gy_loc.double()).sum().reshape(1)
    if world > 1:
        torch.distributed.all_reduce(lhs)
    rhs = float((W.double() * result['dW'].double()).sum())
    adj_err = abs(float(lhs) - rhs) / max(abs(rhs), 1e-30)

    # end to end: this rank's rows of x and gy from pinned host memory every step, a checksum of y and dx read back
    copy_stream = torch.cuda.Stream(device=device)
    bufs = [(torch.empty_like(x_loc), torch.empty_like(gy_loc)) for _ in range(2)]
    out_host = torch.zeros(2, 2, dtype=torch.float32).pin_memory()

    def run_e2e(steps):
        cur = torch.cuda.current_stream()
        done = [None, None]
        for i in range(steps):
            b = i % 2
            with torch.cuda.stream(copy_stream):
                if done[b] is not None:
                    copy_stream.wait_event(done[b])
                bufs[b][0].copy_(x_host, non_blocking=True)
                bufs[b][1].copy_(gy_host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            cur.wait_event(ev)
            y = pf.forward(bufs[b][0], W)
            dx, dW = pf.backward(bufs[b][1])
            out_host[b].copy_(torch.stack([y.sum(), dx.sum()]), non_blocking=True)
            done[b] = torch.cuda.Event()
            done[b].record(cur)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(out_host).all())

    run_e2e(2)
    cgdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(args.steps)
    e1.record()
    torch.cuda.synchronize()
    cgdist.barrier()
    ms_e2e = cgdist.max_over_ranks(e0.elapsed_time(e1), device)

    kernel_ms = common.profile_kernels(step, min(args.steps, 3), timer) if rank == 0 else {}
    if rank != 0:
        for _ in range(min(args.steps, 3)):
            timer.flush()
            step()
        torch.cuda.synchronize()
    cgdist.barrier()
    if rank != 0:
        return
    nnz_loc = int(part.local.nnz)
    b_step_loc = common.b_step(part.nloc, nnz_loc, F)
    work = {'spmm_step': {'bound': 'hbm', 'bytes': b_step_loc, 'flops': 2.0 * nnz_loc * F + 2.0 * part.nloc * F}}
    note = 'fp32-equivalent flops (three bf16 MMAs each) against the dense bf16 peak'
    gem = common.f_gemm(1, part.nloc, F, K, F)
    work['gemm_pipe'] = {'bound': 'tensor', 'bytes': 0, 'flops': gem, 'note': note}
    work['gemm_umma'] = {'bound': 'tensor', 'bytes': 0, 'flops': gem, 'note': note}
    lines = common.roofline_entries(work, kernel_ms, common.measured_traffic(args.traffic_tag + '_c5') if world == 1 else {})
    roof = None
    for e in lines:
        if e['kernel'] == 'spmm_step':
            roof = dict(e)
            roof['all'] = lines
            roof['frac_of_8TBps_nominal'] = e['achieved'] / 8000.0
            roof['note'] = ('one recurrence step on this rank\'s rows: B_step = 8 nnz + 4 (M+1) + 12 M C with the local nnz / rows; '
                            'average over the 2 (K-1) launches of a step (forward on L~, backward on L~^T)')
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        t0 = time.perf_counter()
        log2m = min(args.log2m, 16)
        times, _ = cpu_filter_steps(log2m, args.order, K, F, 1, 0)
        scale = float(1 << log2m) / float(M)
        cpu = {'value': len(times) / float(np.sum(times)) * scale, 'unit': 'samples/s', 'cores': os.cpu_count(), 'kind': 'port',
               'sample': '1 filter forward+backward on a 2^%d-vertex graph of the same construction, scaled by the vertex ratio x%g '
                         '(work is linear in M), %.1f s of CPU work incl. graph build; oracle/tf_ref.py' % (log2m, scale, time.perf_counter() - t0)}
    print(json.dumps({
        'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': W_,
        'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': ('bf16' if getattr(args, 'precision', 'fp32') == 'bf16' else 'f32'),
        'data': 'synthetic',
        'config': {'workload': workload_name(args), 'name': 'c5', 'M': M, 'nnz': int(Lr.nnz), 'F': F, 'K': K, 'rows_rank0': part.nloc,
                   'halo_rows_rank0': part.nhalo, 'parallelism': 'rows%d' % world, 'exchange': pf.exchange_kind,
                   'l2': 'flushed between timed iterations (256 MB fill); operands are 268 MB per slab',
                   'timing': 'CUDA events per step on the launch stream, summed; max over ranks', 'graph_build_s': round(t_build, 1),
                   'adjoint_rel_err': adj_err},
        'clocks': clocks,
        'e2e': {'value': args.steps / (ms_e2e * 1e-3), 'unit': 'samples/s',
                'h2d_bytes_per_step': int(2 * x_host.numel() * 4), 'd2h_bytes_per_step': 8, 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': native_per_step * args.steps, 'native_launches_per_step': native_per_step,
        'roofline': roof, 'cpu_baseline': cpu, 'kernels_ms_per_step': kernel_ms}), flush=True)
