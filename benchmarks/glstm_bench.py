"""bench.py arm for BASELINE.json configs[3] (c4): the human-flow gconv-LSTM model on a Beijing-taxi-shaped 32x32 8-NN
grid graph (M = 1024, nnz 8332; nips2016/gconvTest.py:79-164, humanflow-bjtaxi.ipynb cells 3-4): GconvModel with
infer_func='inference_glstm' -- T unrolled GConvLSTMCell steps (Chebyshev K = 3 cells, Fin = 2, H = 128) and the output
filter H -> 2; one "step" = one full training step on one batch of 50 (forward, MSE, backward, Adam)."""
import json
import os
import time

import numpy as np

from . import common, workloads
from .common import METRIC


def workload_name(args):
    return ('C4 humanflow gconv-LSTM: 32x32 8-NN grid (M=1024), GconvModel inference_glstm, T=%d unrolled cell steps, '
            'Fin=2, H=%d, Chebyshev K=%d, 1 LSTM layer + output filter, full training step' % (args.T, args.H, args.K or 3))


def synthetic_batch(batch, T, seed):
    rng = np.random.RandomState(seed)
    x = rng.uniform(0, 1, (batch, 1024, 2 * T)).astype(np.float32)      # humanflow-bjtaxi.ipynb cell 4: (.., 1024, 2T)
    y = rng.uniform(0, 1, (batch, 1024, 2)).astype(np.float32)
    return x, y


# ---------------------------------------------------------------------------------------------------------------
def cpu_training_steps(args, batch, steps, warmup):
    """oracle/torch_ref.py: op-for-op torch-CPU mirror of the reference's TF graph (eight filters per cell step,
    K-1 concats, restack transposes), autograd backward, Adam."""
    import torch
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count())
    T, H, K = args.T, args.H, args.K or 3
    L = workloads.grid32_laplacian(workloads.host_lib('oracle'))
    Ls = torch_ref.sparse_operator(L, 2)
    g = torch.Generator().manual_seed(0)
    uni = lambda *s: ((torch.rand(*s, generator=g) * 0.2 - 0.1)).requires_grad_(True)
    Wx = {k: uni(K * 2, H) for k in 'zifo'}
    Wh = {k: uni(K * H, H) for k in 'zifo'}
    b = {k: torch.zeros(H, requires_grad=True) for k in 'zifo'}
    Wout = (0.1 * torch.randn(K * H, 2, generator=g)).requires_grad_(True)
    params = list(Wx.values()) + list(Wh.values()) + list(b.values()) + [Wout]
    opt = torch.optim.Adam(params, lr=1e-3)
    x_np, y_np = synthetic_batch(batch, T, 0)
    x, y = torch.from_numpy(x_np), torch.from_numpy(y_np)
    frames = list(torch.unbind(x.reshape(batch, 1024, 2, T), dim=3))
    times = []
    for it in range(warmup + steps):
        mask = (torch.rand(T, batch, 1024, H, generator=g) < 0.8).float() / 0.8
        t0 = time.perf_counter()
        opt.zero_grad()
        c = torch.zeros(batch, 1024, H)
        h = torch.zeros(batch, 1024, H)
        out = None
        for t in range(T):
            h, c = torch_ref.lstm_cell(frames[t], c, h, Ls, K, Wx, Wh, b, 'fork')
            out = h * mask[t]                      # DropoutWrapper(output_keep_prob=0.8): output only, state untouched
        pred = torch_ref.cheby_conv(out, Ls, K, Wout)
        loss = ((y - pred) ** 2).mean()
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times


def run_reference(args, config):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    batch = args.batch or 50
    times = cpu_training_steps(args, batch, args.steps, args.warmup)
    total = float(np.sum(times))
    value = batch * len(times) / total
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'batch_per_gpu': batch, 'global_batch': batch,
                   'note': 'op-for-op torch-CPU mirror of the reference TF graph (oracle/torch_ref.py, pinned to outputs of the '
                           'reference sources), all host threads'},
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': os.cpu_count(), 'kind': 'port',
                         'sample': '%d steps of batch %d' % (len(times), batch)},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def step_work(args, L, N):
    from cnn_graph_b200 import ops
    T, H, K = args.T, args.H, args.K or 3
    Lr = ops.rescale_csr(L, 2)
    M, nnz = Lr.shape[0], Lr.nnz
    # filters of one training step: per cell step an x-path (2 -> 4H) and an h-path (H -> 4H) filter, then H -> 2
    fx, fh, fo = common.f_gemm(N, M, 2, K, 4 * H), common.f_gemm(N, M, H, K, 4 * H), common.f_gemm(N, M, H, K, 2)
    total_gemm = T * (2 * fx + 3 * fh) + 3 * fo          # forward + dW everywhere, dx except for the network input
    bs_h = common.b_stream(M, nnz, N * H, K)
    return {'total_gemm_flops': total_gemm, 'b_stream_h': bs_h, 'b_step_h': common.b_step(M, nnz, N * H), 'M': M, 'nnz': nnz}


def run_ours(args, config):
    import torch
    from cnn_graph_b200 import _native, dist as cgdist
    from cnn_graph_b200.lib import gconv_lstm

    rank, world, local_rank = cgdist.init_from_env('nccl')
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the hot path has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    lib = _native.lib()
    T, H, K = args.T, args.H, args.K or 3
    B = args.batch or 50
    L = workloads.grid32_laplacian(workloads.host_lib('product'))
    torch.manual_seed(1234)
    model = gconv_lstm.GconvModel(L, seq_num_closeness=T, seq_num_period=0, seq_num_trend=0, filter_num=H, conv_layer_num=0,
                                  filter='cheby_conv', batch_size=B, kernel_num=K, in_feature_num=2, out_feature_num=2,
                                  feature_num=2 * T, infer_func='inference_glstm', lstm_layer_count=1, learning_rate=1e-3,
                                  decay_rate=1)
    if world > 1:
        for p_ in model.store.parameters():
            torch.distributed.broadcast(p_.data, src=0)
        model.grad_hook = cgdist.GradAllReducer(average=True)
    x_np, y_np = synthetic_batch(B, T, 99 + rank)
    x_host, y_host = torch.from_numpy(x_np).pin_memory(), torch.from_numpy(y_np).pin_memory()
    x_dev, y_dev = x_host.to(device), y_host.to(device)
    timer = common.Timer(device, local_rank)
    mode = 'eager' if args.eager else 'cuda_graph'
    if mode == 'cuda_graph':
        try:
            model.train_step_graphed(x_dev, y_dev)
            torch.cuda.synchronize()
        except Exception as exc:      # noqa: BLE001
            import sys
            sys.stderr.write('bench: CUDA-graph capture failed (%s); eager launches\n' % (exc,))
            mode = 'eager'
    step = (lambda: model.train_step_graphed(x_dev, y_dev)) if mode == 'cuda_graph' else (lambda: model.train_step(x_dev, y_dev))
    n0 = lib.cg_launch_count()
    model.train_step(x_dev, y_dev)
    native_per_step = int(lib.cg_launch_count() - n0)
    W = max(args.warmup, 3)
    ms_total, clocks, _ = timer.run(step, args.steps, W, sample_clocks=True)
    value = world * B * args.steps / (ms_total * 1e-3)

    trainer = model.pipelined_trainer(perm=None, depth=2, use_graph=(mode == 'cuda_graph'))

    def run_e2e(steps):
        for _ in range(steps):
            trainer.submit(x_host, y_host)
        losses = trainer.drain()
        assert len(losses) == steps and all(np.isfinite(v) for v in losses), losses
    run_e2e(3)
    torch.cuda.synchronize()
    cgdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(args.steps)
    e1.record()
    torch.cuda.synchronize()
    cgdist.barrier()
    ms_e2e = cgdist.max_over_ranks(e0.elapsed_time(e1), device)

    # forward / backward split and per-kernel times (eager)
    kernel_ms = common.profile_kernels(lambda: model.train_step(x_dev, y_dev), min(args.steps, 5), timer) if rank == 0 else {}
    if rank != 0:
        for _ in range(min(args.steps, 5)):
            timer.flush()
            model.train_step(x_dev, y_dev)
        torch.cuda.synchronize()
    cgdist.barrier()
    trainer = None
    model._captured = None
    if rank != 0:
        return
    w = step_work(args, L, B)
    pk = common.peaks()
    roof = None
    gemm = [kernel_ms[k] for k in ('gemm_pipe', 'gemm_umma') if k in kernel_ms]
    if gemm:
        ms = sum(g['ms_per_step'] for g in gemm)
        n_l = sum(g['launches_per_step'] for g in gemm)
        ach = w['total_gemm_flops'] / (ms * 1e-3) / 1e12
        roof = {'kernel': 'gemm_pipe', 'bound': 'tensor', 'achieved': ach, 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s',
                'frac': ach / pk['bf16_tflops'], 'traffic': None, 'ms_per_step': ms, 'launches_per_step': n_l,
                'algorithmic_flops_per_launch': w['total_gemm_flops'] / max(n_l, 1), 'peak_source': pk['source'],
                'note': 'all contraction GEMMs of the step (x-path, h-path, output filter; forward, dX, dW): fp32-equivalent '
                        'algorithmic flops 2 N M Fin K Fout each (three bf16 MMAs per product are issued) against the dense bf16 peak'}
        rec = [(k, kernel_ms[k]) for k in ('basis_onchip', 'spmm_step', 'clenshaw_step') if k in kernel_ms]
        roof['spmm'] = {k: dict(v) for k, v in rec}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        t0 = time.perf_counter()
        times = cpu_training_steps(args, B, 2, 1)
        cpu = {'value': B * len(times) / float(np.sum(times)), 'unit': 'samples/s', 'cores': os.cpu_count(), 'kind': 'port',
               'sample': '2 timed steps of batch %d after 1 warm-up, %.1f s of CPU work; oracle/torch_ref.py (torch-CPU mirror of the '
                         'reference TF graph, all host threads)' % (B, time.perf_counter() - t0)}
    print(json.dumps({
        'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': W,
        'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': ('bf16' if getattr(args, 'precision', 'fp32') == 'bf16' else 'f32'),
        'data': 'synthetic',
        'config': {'workload': workload_name(args), 'name': 'c4', 'batch_per_gpu': B, 'global_batch': B * world, 'T': T, 'H': H, 'K': K,
                   'gate_variant': 'fork (lib/gconv_lstm.py:185-215)', 'parallelism': 'dp%d' % world, 'launch': mode,
                   'l2': 'flushed between timed iterations (256 MB fill)',
                   'timing': 'CUDA events per step on the launch stream, summed; max over ranks'},
        'clocks': clocks,
        'e2e': {'value': world * B * args.steps / (ms_e2e * 1e-3), 'unit': 'samples/s',
                'h2d_bytes_per_step': int(trainer_bytes(x_host, y_host)), 'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': native_per_step * args.steps, 'native_launches_per_step': native_per_step,
        'roofline': roof, 'cpu_baseline': cpu, 'kernels_ms_per_step': kernel_ms}), flush=True)


def trainer_bytes(x_host, y_host):
    return x_host.numel() * x_host.element_size() + y_host.numel() * y_host.element_size()
