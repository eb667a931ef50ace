"""bench.py contract on CPU: the reference arm prints exactly one JSON line with the keys the driver reads, and the
per-kernel work table the roofline line is computed from is consistent with SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS='2')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                          '--batch', '8'], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'cheb_graphconv_train_samples_per_sec' and d['unit'] == 'samples/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['steps'] == 1
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['value'] == d['value'] and d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config'] and 'model' not in d['config'] and d['config']['batch_per_gpu'] == 8
    assert 'libcnn_graph_b200' not in out.stderr      # the reference arm is built from oracle/ only


def test_step_work_matches_survey_figures():
    sys.path.insert(0, ROOT)
    import bench
    L, _ = bench.build_graphs()
    work = bench.step_work(L, 100)
    # SURVEY.md 8(d), N = 100: C2 layer 2 B_stream 225.7 MB, contraction 2539.5 MFLOP (graph sizes are seed-dependent: 3 %)
    b, f = work['fused_fwd']['bytes'], work['fused_fwd']['flops']
    assert abs(b / 225.7e6 - 1) < 0.03
    assert abs(work['fused_fwd']['floor_bytes'] / 82.6e6 - 1) < 0.03              # B_floor of the same layer
    assert abs((2.0 * 100 * L[2].shape[0] * 32 * 25 * 64) / 2539.5e6 - 1) < 0.03 and f > 2.0 * 100 * L[2].shape[0] * 32 * 25 * 64
    assert abs(work['basis_onchip']['bytes'] / 29.5e6 - 1) < 0.03                  # C2 layer 1 B_stream
    for name, w in work.items():
        assert w['bound'] in ('hbm', 'tensor') and (w['bytes'] or w['flops']), name


def test_reference_arm_does_not_load_the_product_library():
    """The CPU arm builds its graphs with oracle/ and never imports cnn_graph_b200 (a clean reference arm)."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--batch', '4'];"
            "runpy.run_path(%r, run_name='__main__');"
            "bad = [m for m in sys.modules if m.startswith('cnn_graph_b200')];"
            "import ctypes; maps = open('/proc/self/maps').read();"
            "assert not bad and 'libcnn_graph_b200' not in maps, (bad,)") % os.path.join(ROOT, 'bench.py')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=dict(os.environ, OMP_NUM_THREADS='2'))
    assert out.returncode == 0, out.stderr[-2000:]


def test_reference_arm_other_configs():
    for cfg, extra in (('c1', ['--batch', '8']), ('c4', ['--batch', '2', '--T', '2', '--H', '8']), ('c5', ['--log2m', '10', '--K', '4'])):
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', cfg, '--steps', '1',
                              '--warmup', '0'] + extra, capture_output=True, text=True, timeout=600, cwd=ROOT,
                             env=dict(os.environ, OMP_NUM_THREADS='2'))
        assert out.returncode == 0, (cfg, out.stderr[-2000:])
        d = json.loads([l for l in out.stdout.splitlines() if l.strip()][-1])
        assert d['impl'] == 'reference' and d['value'] > 0 and d['cpu_baseline']['kind'] == 'port', cfg
