"""Turn ncu output into the small tracked summaries under profiles/.

  python scripts/summarize_ncu.py launches <launches.csv> <out_summary.txt> "<header line>"
  python scripts/summarize_ncu.py full <report.ncu-rep> <out_summary.csv> [<traffic.json>]
"""
import csv, io, json, subprocess, sys
from collections import OrderedDict

COLS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__shared_mem_per_block_dynamic']
# profile-scope names of bench.py (cg_profile_*) for the kernels whose DRAM traffic the roofline line quotes
SCOPE = {'k_cheb_fused': 'fused_fwd', 'k_cheb_clenshaw': 'clenshaw_dx', 'k_dw_planes': 'dw_planes', 'k_dw_umma': 'dw_umma', 'k_dw_thin': 'dw_thin',
         'k_contract_umma': 'contract_umma', 'k_gemm_pipe': 'gemm_pipe', 'k_gemm_stream': 'gemm_pipe', 'k_pack_b': 'gemm_pack_b', 'k_thin_contract': 'thin_contract', 'k_thin_dw': 'thin_dw', 'k_gemm_umma': 'gemm_umma', 'k_basis_onchip': 'basis_onchip',
         'k_spmm_tile_p': 'spmm_step', 'k_spmm_step': 'spmm_step'}


def launches(path, out, header):
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 5]
    hdr = next(r for r in rows if 'Kernel Name' in r)
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    tot = OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        n, t = tot.get(r[ki], (0, 0.0))
        tot[r[ki]] = (n + 1, t + v)
    unit_ns = any('nsecond' in c or c == 'ns' for r in rows[:20] for c in r)
    scale = 1e-6 if unit_ns else 1e-3
    total = sum(t for _, t in tot.values())
    with open(out, 'w') as f:
        f.write(header + '\n')
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write('%-66s n=%4d total=%9.3f ms share=%5.1f%%\n' % (k[:66], n, t * scale, 100.0 * t / total))


def full(rep, out, traffic_out=None):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    ki = hdr.index('Kernel Name')
    traffic = {}
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['Kernel Name'] + [hdr[i] for i in idx])
        w.writerow([''] + [units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[ki]] + [r[i] for i in idx])
            for key, scope in SCOPE.items():
                if key in r[ki]:
                    rd, wr = float(r[hdr.index('dram__bytes_read.sum')]), float(r[hdr.index('dram__bytes_write.sum')])
                    mul = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.0}[units[hdr.index('dram__bytes_read.sum')]]
                    mulw = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.0}[units[hdr.index('dram__bytes_write.sum')]]
                    traffic.setdefault(scope, []).append(rd * mul + wr * mulw)
    if traffic_out:
        t = {k: sum(v) / len(v) for k, v in traffic.items()}
        t['_note'] = ('dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches of the '
                      'scope), ncu --set full --clock-control none of the same bench.py command with --eager (%s)' % out)
        json.dump(t, open(traffic_out, 'w'), indent=1)


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
