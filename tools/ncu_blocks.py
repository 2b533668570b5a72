"""Group the SASS of a one-kernel ncu report (--set full --import-source on) into runs of
instructions with the same execution count, with their share of warp instructions, stall
samples and active lanes.  usage: python tools/ncu_blocks.py report.ncu-rep [min_share_pct]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = []
for r in rows[2:]:
    try:
        inst = int(r[ix["Instructions Executed"]]); th = int(r[ix["Thread Instructions Executed"]]); smp = int(r[ix["# Samples"]])
    except Exception:
        continue
    data.append((r[1].strip(), inst, th, smp))
tot = sum(d[1] for d in data); tots = sum(d[3] for d in data) or 1
first = data[0][1] or 1
blocks = []; cur = None
for i, (s, inst, th, smp) in enumerate(data):
    op = s.split()[1] if s.startswith('@') and len(s.split()) > 1 else s.split()[0]
    if cur and abs(inst - cur['inst']) <= max(1, 0.001 * cur['inst']) and not cur['last'].startswith(('BRA', 'EXIT', 'BSYNC', 'BSSY', 'CALL', 'RET')):
        cur['n'] += 1; cur['tot'] += inst; cur['th'] += th; cur['smp'] += smp; cur['last'] = op; cur['ops'].append(op)
    else:
        cur = {'start': i, 'inst': inst, 'n': 1, 'tot': inst, 'th': th, 'smp': smp, 'last': op, 'ops': [op]}
        blocks.append(cur)
print(f"{rows[0][1][:100]}\ntotal warp-instr {tot}, thread-instr {sum(d[2] for d in data)}, avg lanes {sum(d[2] for d in data)/tot:.2f}, {len(data)} SASS instructions")
for b in blocks:
    if 100 * b['tot'] / tot < min_share:
        continue
    c = Counter(o.split('.')[0] for o in b['ops'])
    print(f"@{b['start']:4d} n={b['n']:3d} execs/warp {b['inst']/first:7.2f}  {100*b['tot']/tot:5.2f}% inst {100*b['smp']/tots:5.2f}% smp  lanes {b['th']/max(b['tot'],1):5.1f}  {dict(c.most_common(7))}")
