#!/usr/bin/env python3
"""Benchmark of the Chebyshev graph-conv hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                   # this framework, default config c2
    python bench.py --impl reference --steps K --warmup W           # the reference's CPU path (oracle port), same config + batch
    python bench.py --config c1|c2|c3|c4|c5 [--batch B] ...         # the other BASELINE.json configs

Default workload (config.workload): BASELINE config C2 -- MNIST-shaped synthetic data on the 28x28 8-NN grid graph,
4-level Graclus coarsening (M = 992 after fake-node padding), cgcnn GC32-P4-GC64-P4-FC512-FC10 with K = 25 Chebyshev
terms (nips2016/mnist.ipynb cells 1,3,14,17), batch 1024 per GPU in BOTH arms (--batch; the reference notebooks' batch 100
is in `batch_sweep`).  One "step" = one full training step of that model on one batch: forward, softmax cross-entropy +
L2, backward, momentum-SGD update (and, for N > 1 GPUs, the all-reduce of the weight gradients).  Every graph-conv
kernel and the dense head are native (cnn_graph_b200/csrc).  The filter arithmetic is fp32 throughout; the tensor-core
products split every fp32 operand into bf16 hi + mid (three MMAs, fp32 accumulate in TMEM), which stays inside the
reference's fp32 tolerance (rtol 1e-4, tests/test_gpu_parity.py, tests/test_reference_fixtures.py).

Prints ONE JSON line (rank 0).  `value`: samples/s with the batch resident in HBM; `e2e`: the same step driven from
pinned HOST buffers (raw signals -> H2D -> device perm_data -> train step -> loss read back) every step.
The harness lives in benchmarks/ (one module per config family); oracle/ is only used by the CPU legs.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from benchmarks import workloads  # noqa: E402  (numpy / scipy only)

WORKLOAD = workloads.CGCNN['c2']['workload']


def build_graphs(seed=0):
    """C2 graphs through the product's host library (kept for tests / scripts)."""
    return workloads.cgcnn_graphs('c2', workloads.host_lib('product'), seed)


def step_work(L, N):
    from benchmarks import cgcnn_bench
    return cgcnn_bench.step_work('c2', L, N)


def main():
    if os.environ.get('CG_BENCH_WATCHDOG'):      # debugging aid: dump every thread's stack after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['CG_BENCH_WATCHDOG']), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c2', choices=['c1', 'c2', 'c3', 'c4', 'c5'])
    ap.add_argument('--batch', type=int, default=0, help='samples per GPU per step in BOTH arms (0: the config default; c2: 1024)')
    ap.add_argument('--ref-batch', type=int, default=0, help='deprecated alias of --batch for the reference arm')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: --batch per GPU; strong: --batch is the GLOBAL batch, split over the GPUs')
    ap.add_argument('--sweep', type=int, nargs='*', default=None,
                    help='device-timed batch sweep added to the line as `batch_sweep` (default for c2 at 1 GPU: 100 256 1024 4096)')
    ap.add_argument('--no-sweep', action='store_true')
    ap.add_argument('--sustain', type=float, default=0.0, help='also run the resident step for this many seconds (rate under power cap)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--overlap-allreduce', action='store_true', help='all-reduce the large gradients from autograd hooks, '
                    'overlapped with the backward pass (default: one flat all-reduce after it, which measured faster)')
    ap.add_argument('--allreduce', default='flat', choices=['flat', 'deferred'],
                    help='N > 1: flat = one bucket after the backward pass (default, measured fastest); deferred = small gradients averaged '
                         'after the backward pass, the large dense weight\'s exchange and update moved under the next forward pass '
                         '(same trajectory; 1.529 vs 1.517 ms per step at 2 GPUs)')
    ap.add_argument('--eager', action='store_true', help='kernel-by-kernel launches instead of CUDA-graph replay')
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'bf16'],
                    help='tensor-core products: fp32-equivalent bf16 hi+mid split x3 (default, rtol 1e-4) or single-pass bf16 (rtol 2e-2)')
    ap.add_argument('--traffic-tag', default='r2', help='profiles/roofline_traffic_<tag>.json: ncu dram bytes per launch')
    # c4 / c5 knobs
    ap.add_argument('--T', type=int, default=3, help='c4: unrolled cell steps')
    ap.add_argument('--H', type=int, default=128, help='c4: hidden features')
    ap.add_argument('--K', type=int, default=0, help='c4 / c5: Chebyshev terms (0: config default 3 / 20)')
    ap.add_argument('--log2m', type=int, default=20, help='c5: 2^log2m vertices')
    ap.add_argument('--order', default='morton', choices=['morton', 'random'], help='c5: vertex order')
    args = ap.parse_args()
    if args.ref_batch and not args.batch:
        args.batch = args.ref_batch
    if args.sweep is None:
        args.sweep = [100, 256, 1024, 4096] if (args.config == 'c2' and args.impl == 'ours' and not args.no_sweep
                                                and args.scaling == 'weak') else []
    if args.no_sweep:
        args.sweep = []

    if args.config in ('c1', 'c2', 'c3'):
        from benchmarks import cgcnn_bench as mod
    elif args.config == 'c4':
        from benchmarks import glstm_bench as mod
    else:
        from benchmarks import rowpart_bench as mod
    if args.impl == 'reference':
        mod.run_reference(args, args.config)
    else:
        if args.precision != 'fp32':
            from cnn_graph_b200 import ops
            ops.set_precision(args.precision)
        mod.run_ours(args, args.config)
        from benchmarks import common
        common.shutdown()


if __name__ == '__main__':
    main()
