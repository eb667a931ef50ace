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
                          '--ref-batch', '8'], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'cheb_graphconv_train_samples_per_sec' and d['unit'] == 'samples/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['steps'] == 1
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['value'] == d['value'] and d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_step_work_matches_survey_figures():
    sys.path.insert(0, ROOT)
    import bench
    L, _ = bench.build_graphs()
    work = bench.step_work(L, 100)
    # SURVEY.md 8(d), N = 100: C2 layer 2 B_stream 225.7 MB, contraction 2539.5 MFLOP (graph sizes are seed-dependent: 3 %)
    b, f = work['fused_fwd']['launches'][0]
    assert abs(b / 225.7e6 - 1) < 0.03
    assert abs((2.0 * 100 * L[2].shape[0] * 32 * 25 * 64) / 2539.5e6 - 1) < 0.03 and f > 2.0 * 100 * L[2].shape[0] * 32 * 25 * 64
    for name, w in work.items():
        assert w['bound'] in ('hbm', 'tensor') and w['launches'], name
