"""Shared pieces of the bench harness: measured peaks, clock sampling, the timed loop (CUDA events on the launch
stream, L2 flush between iterations, max over ranks), per-kernel event timing through the library's own profiling
hooks, and SURVEY.md 8(d)'s algorithmic work formulas."""
import ctypes
import json
import os
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRIC = 'cheb_graphconv_train_samples_per_sec'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'],
                'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured'}
    # B200_PROFILING.md fallback figures
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1590.0, 'source': 'fallback'}


# ---------------------------------------------------------------------------------------------------------------
# algorithmic bytes / flops
# ---------------------------------------------------------------------------------------------------------------
def b_step(M, nnz, C):
    """One recurrence step X_k = 2 L~ X_{k-1} - X_{k-2} on an [M x C] fp32 operand with int32 CSR."""
    return 8 * nnz + 4 * (M + 1) + 12 * M * C


def b_stream(M, nnz, C, K):
    """sum_{k=1}^{K-1} B_step; the first step reads one operand fewer."""
    return (K - 1) * b_step(M, nnz, C) - 4 * M * C if K > 1 else 0


def b_floor(M, nnz, C, K):
    """Cache-infinite floor of the same op with a materialised stack."""
    return 8 * nnz + 4 * (M + 1) + 4 * M * C + 4 * K * M * C


def spmm_flops(M, nnz, C, K):
    return (K - 1) * 2 * nnz * C + max(K - 2, 0) * 2 * M * C


def f_gemm(N, M, Fin, K, Fout):
    return 2.0 * N * M * Fin * K * Fout


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            # wait for the first sample: nvidia-smi's start-up (NVML initialisation over every GPU of the box) stalls the
            # GPUs for tens of milliseconds -- it must be over before the timed region begins, the 100 ms polls that follow
            # are what the timed region should see
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:
                time.sleep(0.02)
        except OSError:
            self.proc = None

    def mark(self):
        """Samples from here on count (start of the timed region)."""
        self.first = len(self.lines)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, mx, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines[self.first:]:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[4:8]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None, 'reasons': sorted(reasons), 'samples': len(sm)}


class Timer:
    """The timed loop of the contract: W untimed warm-up calls, then K calls each bracketed by CUDA events on the
    current (launch) stream, an L2 flush (256 MB fill, > 126 MB L2) before every timed call and outside its events,
    barrier + synchronize on both sides, total = max over ranks of the summed event times."""

    def __init__(self, device, local_rank, flush=True):
        import torch
        self.torch = torch
        self.device, self.local_rank = device, local_rank
        if os.environ.get('CG_BENCH_NOFLUSH'):            # debugging aid only: a published number needs the flush
            flush = False
        self.flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device) if flush else None

    def flush(self):
        if self.flush_buf is not None:
            self.flush_buf.fill_(0.0)

    def run(self, fn, steps, warmup, sample_clocks=False, min_seconds=0.0):
        """Returns (total_ms over `steps` calls, clocks, extra) -- when `min_seconds` > 0 the same loop is continued
        untimed-in-`total_ms` but timed separately until that much wall time has passed (sustained-rate check)."""
        from cnn_graph_b200 import dist as cgdist
        torch = self.torch
        # the clock sampler (ONE nvidia-smi subprocess, on local rank 0) starts before the warm-up and is given time to
        # finish initialising: started after the barrier, its spawn skew landed inside the first timed step of every rank
        # and its NVML start-up stalled all GPUs for ~28 ms in the middle of the timed region (8 GPUs: +1 ms per step
        # over 20 steps; per-step times printed with CG_BENCH_STEPTIMES=1 showed 1.53 ms steps and one 28 ms outlier)
        sampler = ClockSampler(self.local_rank) if (sample_clocks and self.local_rank == 0) else None
        if sampler:
            sampler.start()
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        cgdist.barrier()
        if sampler:
            sampler.mark()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            self.flush()
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        cgdist.barrier()
        if sample_clocks:
            # the timed region is a few tens of milliseconds, nvidia-smi polls every 100 ms: keep the same step running
            # (untimed) for 0.3 s so that the clocks / throttle reasons are sampled under this very load
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < 0.3:
                for _ in range(10):
                    fn()
                torch.cuda.synchronize()
            cgdist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        if os.environ.get('CG_BENCH_STEPTIMES'):          # debugging aid: per-step device times of this rank
            import sys
            sys.stderr.write('rank %s step ms: %s\n' % (os.environ.get('RANK', '0'), ' '.join('%.3f' % a.elapsed_time(b) for a, b in ev)))
        sustained = None
        if min_seconds > 0:
            t0 = time.perf_counter()
            n, tot = 0, 0.0
            while time.perf_counter() - t0 < min_seconds:
                chunk = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
                for a, b in chunk:
                    self.flush()
                    a.record()
                    fn()
                    b.record()
                torch.cuda.synchronize()
                tot += sum(a.elapsed_time(b) for a, b in chunk)
                n += len(chunk)
            sustained = {'steps': n, 'ms_per_step': tot / max(n, 1), 'wall_s': time.perf_counter() - t0}
        clocks = sampler.stop() if sampler else None
        return cgdist.max_over_ranks(ms, self.device), clocks, sustained


def profile_kernels(fn, steps, timer=None):
    """Per-kernel device time of `fn` (eager launches): the library brackets each of its launches with CUDA events on
    the launch stream (cg_profile_*).  Returns name -> {'ms_per_step', 'launches_per_step'}."""
    import torch
    from cnn_graph_b200 import _native
    lib = _native.lib()
    lib.cg_profile_reset()
    lib.cg_profile_enable(1)
    for _ in range(steps):
        if timer is not None:
            timer.flush()
        fn()
    torch.cuda.synchronize()
    lib.cg_profile_enable(0)
    out = {}
    name = ctypes.create_string_buffer(64)
    tot, cnt = ctypes.c_double(), ctypes.c_int64()
    n = lib.cg_profile_query(-1, None, 0, None, None)
    for i in range(n):
        lib.cg_profile_query(i, name, 64, ctypes.byref(tot), ctypes.byref(cnt))
        out[name.value.decode()] = {'ms_per_step': tot.value / steps, 'launches_per_step': cnt.value / steps}
    return out


def measured_traffic(tag):
    """dram__bytes (read + write) per launch from the round's `ncu --set full` capture, keyed by bench kernel name:
    written by scripts/summarize_ncu.py into profiles/roofline_traffic_<tag>.json; None when absent."""
    path = os.path.join(ROOT, 'profiles', 'roofline_traffic_%s.json' % tag)
    if os.path.exists(path):
        return json.load(open(path))
    return {}


def roofline_entries(work, kernel_ms, traffic=None):
    """work: name -> {'bound', 'bytes', 'flops', 'floor_bytes'?, 'note'?} PER LAUNCH (average); kernel_ms from
    profile_kernels.  Returns the list sorted by time, each with achieved / peak / frac per SURVEY.md 8(d)."""
    pk = peaks()
    traffic = traffic or {}
    out = []
    for name, w in work.items():
        if name not in kernel_ms:
            continue
        n_l = max(kernel_ms[name]['launches_per_step'], 1)
        avg_s = kernel_ms[name]['ms_per_step'] * 1e-3 / n_l
        e = {'kernel': name, 'bound': w['bound'], 'ms_per_step': kernel_ms[name]['ms_per_step'], 'launches_per_step': n_l,
             'avg_launch_ms': avg_s * 1e3, 'peak_source': pk['source'], 'traffic': traffic.get(name)}
        if w['bound'] == 'hbm':
            e.update(achieved=w['bytes'] / avg_s / 1e9, peak=pk['hbm_gbs'], unit='GB/s',
                     algorithmic_bytes_per_launch=w['bytes'])
            if w.get('flops'):
                e['tensor_TFLOPs'] = w['flops'] / avg_s / 1e12
            if w.get('floor_bytes'):
                e['floor_GBps'] = w['floor_bytes'] / avg_s / 1e9          # B_floor / t
                e['floor_frac'] = e['floor_GBps'] / pk['hbm_gbs']
            if e['traffic']:
                e['dram_GBps'] = e['traffic'] / avg_s / 1e9
                e['dram_frac'] = e['dram_GBps'] / pk['hbm_gbs']
        else:
            e.update(achieved=w['flops'] / avg_s / 1e12, peak=pk['bf16_tflops'], unit='TFLOP/s',
                     algorithmic_flops_per_launch=w['flops'])
        if w.get('note'):
            e['note'] = w['note']
        e['frac'] = e['achieved'] / e['peak']
        out.append(e)
    out.sort(key=lambda e: -e['ms_per_step'])
    return out


def shutdown():
    """Tear the process group down.  destroy_process_group() blocks when CUDA graphs that captured NCCL work are
    still alive, so callers drop those graphs first; a watchdog turns a hang into a clean exit (every rank has passed
    the final barrier and printed by then)."""
    import sys
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            sys.stdout.flush()
            sys.stderr.flush()
            t = threading.Timer(20.0, lambda: os._exit(0))
            t.daemon = True
            t.start()
            dist.destroy_process_group()
            t.cancel()
    except Exception:
        pass
