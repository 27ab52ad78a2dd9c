"""Top stalled SASS instructions of one kernel in an .ncu-rep: python scripts/ncu_hot.py rep kernel_regex [n]"""
import csv, subprocess, sys, io
from collections import Counter
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = None, []
for r in rows:
    if r and r[0] == 'Address':
        if hdr is not None and data:
            break
        hdr, data = r, []
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
i_src, i_s, i_ie = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
tot_i = sum(int(r[i_ie]) for r in data)
tot_s = sum(int(r[i_s]) for r in data)
print('total instr', tot_i, 'samples', tot_s, 'n', len(data))
c, cs = Counter(), Counter()
for r in data:
    t = r[i_src].strip().split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += int(r[i_ie])
    cs[op] += int(r[i_s])
print('  '.join(f'{op} {v / tot_i * 100:.1f}%/{cs[op] / max(tot_s, 1) * 100:.1f}%' for op, v in c.most_common(14)))
for r in sorted(data, key=lambda r: -int(r[i_s]))[:n]:
    print(f'{int(r[i_s]) / max(tot_s, 1) * 100:5.1f}% {int(r[i_ie]):9d}  {r[i_src].strip()[:100]}')
