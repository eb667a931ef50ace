#!/usr/bin/env python3
"""BASELINE config C5 on one GPU: large synthetic kNN graph (default 2^20 points in [0,1]^2, exact 16-NN,
Morton order), Fin = Fout = 64, K = 20.  The operator does not fit shared memory, so the recurrence runs one
CSR step per launch from HBM (k_spmm_step) -- the genuinely HBM-bound SpMM of the metric.  Prints one JSON line:
achieved algorithmic GB/s per step (B_step = 8 nnz + 4 (M+1) + 12 M C) against the measured HBM peak.

    python scripts/bench_c5.py [--log2m 20] [--k 16] [--F 64] [--K 20] [--order morton|random]
"""
import argparse, ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cnn_graph_b200 import _native, ops
from cnn_graph_b200.lib import graph


def morton_order(z, bits=16):
    q = np.minimum((z * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    def spread(v):
        v = v & 0xFFFF
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    return np.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << 1), kind='stable')


ap = argparse.ArgumentParser()
ap.add_argument('--log2m', type=int, default=20)
ap.add_argument('--k', type=int, default=16)
ap.add_argument('--F', type=int, default=64)
ap.add_argument('--K', type=int, default=20)
ap.add_argument('--order', default='morton')
ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--partition', type=int, default=0, help='row-partition over WORLD_SIZE ranks (torchrun) with halo exchange')
ap.add_argument('--filter', type=int, default=0, help='with --partition: time the whole filter (forward + backward: dx, dW) instead of the recurrence')
a = ap.parse_args()
from cnn_graph_b200 import dist as cgdist
rank, world, local_rank = cgdist.init_from_env('nccl') if a.partition else (0, 1, 0)
torch.cuda.set_device(local_rank)
M = 1 << a.log2m
t0 = time.time()
rng = np.random.RandomState(2017)
z = rng.uniform(0, 1, (M, 2)).astype(np.float32)
z = z[morton_order(z)] if a.order == 'morton' else z[rng.permutation(M)]
dist, idx = graph.knn_kdtree(z, k=a.k)
A = graph.adjacency(dist, idx)
L = graph.laplacian(A, normalized=True)
t_build = time.time() - t0
if a.partition and a.filter:
    from cnn_graph_b200 import partition
    Lr = ops.rescale_csr(L, 2)
    pf = partition.PartitionedFilter(Lr, a.K)
    part = pf.part
    gen = torch.Generator(device='cuda').manual_seed(7)
    x_full = torch.randn(M, a.F, device='cuda', generator=gen)
    gy_full = torch.randn(M, a.F, device='cuda', generator=gen)
    W = 0.05 * torch.randn(a.F * a.K, a.F, device='cuda', generator=gen)
    x_loc, gy_loc = x_full[part.r0:part.r1].contiguous(), gy_full[part.r0:part.r1].contiguous()
    tf, tb = [], []
    for it in range(a.iters + 1):
        torch.cuda.synchronize(); cgdist.barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        y = pf.forward(x_loc, W)
        e[1].record()
        dx, dW = pf.backward(gy_loc)
        e[2].record()
        torch.cuda.synchronize(); cgdist.barrier()
        tf.append(cgdist.max_over_ranks(e[0].elapsed_time(e[1]), x_loc.device))
        tb.append(cgdist.max_over_ranks(e[1].elapsed_time(e[2]), x_loc.device))
    # adjoint identity across the partition: <y, gy> = <W, dW> (the filter is linear in W), summed over ranks
    lhs = (y.double() * gy_loc.double()).sum().reshape(1)
    if world > 1:
        torch.distributed.all_reduce(lhs)
    rhs = float((W.double() * dW.double()).sum())
    if rank == 0:
        print(json.dumps({'workload': 'C5 row-partitioned filter forward + backward (dx, dW) with halo exchange', 'M': M,
                          'nnz': int(Lr.nnz), 'F': a.F, 'K': a.K, 'n_gpus': world, 'fwd_ms': min(tf[1:]), 'bwd_ms': min(tb[1:]),
                          'adjoint_rel_err': abs(float(lhs) - rhs) / max(abs(rhs), 1e-30), 'halo_rows_rank0': part.nhalo,
                          'rows_rank0': part.nloc, 'timing': 'CUDA events, max over ranks'}))
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)          # (no blocking NCCL teardown)
elif a.partition:
    from cnn_graph_b200 import partition
    Lr = ops.rescale_csr(L, 2)
    pb = partition.PartitionedBasis(Lr)
    part = pb.part
    gen = torch.Generator(device='cuda').manual_seed(7)
    x_full = torch.randn(M, a.F, device='cuda', generator=gen)          # same on every rank
    x_loc = x_full[part.r0:part.r1].contiguous()
    times = []
    for it in range(a.iters + 1):
        torch.cuda.synchronize(); cgdist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        Xt = pb.basis(x_loc, a.K)
        e1.record()
        torch.cuda.synchronize(); cgdist.barrier()
        times.append(cgdist.max_over_ranks(e0.elapsed_time(e1), x_loc.device))
    best = min(times[1:]) / (a.K - 1)
    # correctness of this rank's block against the single-GPU recurrence (first 4 slabs)
    h = ops.GraphHandle(Lr)
    ref = ops.cheb_basis(h, x_full, 4)[:, part.r0:part.r1]
    err = float((Xt[:4] - ref).abs().max()) / float(ref.abs().max())
    err = cgdist.max_over_ranks(err, x_loc.device)
    if rank == 0:
        nnz = Lr.nnz
        b_step = 8 * nnz + 4 * (M + 1) + 12 * M * a.F
        print(json.dumps({'workload': 'C5 row-partitioned recurrence with halo exchange', 'M': M, 'nnz': int(nnz), 'C': a.F,
                          'K': a.K, 'n_gpus': world, 'ms_per_step': best, 'achieved_GBps_aggregate': b_step / (best * 1e-3) / 1e9,
                          'halo_rows_rank0': part.nhalo, 'rows_rank0': part.nloc, 'max_rel_err_vs_single_gpu': err,
                          'timing': 'CUDA events around the K-1 steps incl. halo all-to-all, max over ranks'}))
    torch.distributed.destroy_process_group() if world > 1 else None
    sys.exit(0)
h = ops.GraphHandle.from_laplacian(L, 2)
info = h.info()
C = a.F
x = torch.randn(M, C, device='cuda')
lib = _native.lib()
best = None
for it in range(a.iters):
    lib.cg_profile_reset(); lib.cg_profile_enable(1)
    Xt = ops.cheb_basis(h, x, a.K)
    torch.cuda.synchronize()
    lib.cg_profile_enable(0)
    name = ctypes.create_string_buffer(64); ms = ctypes.c_double(); cnt = ctypes.c_int64()
    n = lib.cg_profile_query(-1, name, 64, ctypes.byref(ms), ctypes.byref(cnt))
    for i in range(n):
        lib.cg_profile_query(i, name, 64, ctypes.byref(ms), ctypes.byref(cnt))
        if name.value == b'spmm_step':
            per = ms.value / cnt.value
            best = per if best is None else min(best, per)
    del Xt
nnz = info['nnz']
b_step = 8 * nnz + 4 * (M + 1) + 12 * M * C
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
gbs = b_step / (best * 1e-3) / 1e9
print(json.dumps({'workload': 'C5 large kNN graph, one recurrence step (k_spmm_step)', 'M': M, 'nnz': int(nnz),
                  'max_row': int(info['width']), 'C': C, 'K': a.K, 'order': a.order, 'ms_per_step': best,
                  'algorithmic_bytes_per_step': b_step, 'achieved_GBps': gbs, 'hbm_peak_GBps': peaks['hbm_gbs'],
                  'frac': gbs / peaks['hbm_gbs'], 'onchip': info['onchip'], 'graph_build_s': round(t_build, 1)}))
