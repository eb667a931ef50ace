"""Per-kernel SASS opcode counts of the built library (evidence of tcgen05 / TMA use): profiles/sass_opcodes_<tag>.txt.

  python scripts/sass_opcodes.py profiles/sass_opcodes_r2.txt
"""
import re, subprocess, sys
from collections import OrderedDict

OPS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'UTCBAR', 'SYNCS', 'LDGSTS', 'HMMA', 'FFMA', 'LDS', 'STS', 'SHFL', 'UTCCP']
LIB = 'cnn_graph_b200/libcnn_graph_b200.so'


def main(out):
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
    counts, cur, i = OrderedDict(), None, 0
    for line in sass.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = names[i]
            i += 1
            counts[cur] = dict.fromkeys(OPS, 0)
            continue
        if cur is None:
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + '.'):
                    counts[cur][o] += 1
    with open(out, 'w') as f:
        f.write('cuobjdump -sass %s: opcode counts per kernel (tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM, cp.async.bulk = UBLKCP, '
                'cp.async.bulk.tensor = UTMALDG, mbarrier = SYNCS)\n' % LIB)
        f.write('%-90s' % 'kernel' + ''.join('%9s' % o for o in OPS) + '\n')
        for k, c in counts.items():
            f.write('%-90s' % k[:90] + ''.join('%9d' % c[o] for o in OPS) + '\n')
        tot = {o: sum(c[o] for c in counts.values()) for o in OPS}
        f.write('%-90s' % 'TOTAL' + ''.join('%9d' % tot[o] for o in OPS) + '\n')


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'profiles/sass_opcodes.txt')
