"""Aggregate an ncu source page (--page source --csv --print-source cuda,sass) by source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_hot_lines.py [top]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
hdr = None
agg = {}
for r in rows:
    if r and r[0] in ("Line No", "Address", "#"):
        hdr = r
        ix = {n: i for i, n in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    try:
        inst = int(r[ix["Instructions Executed"]]); th = int(r[ix["Thread Instructions Executed"]]); smp = int(r[ix["# Samples"]])
    except Exception:
        continue
    src = r[ix["Source"]].strip() if "Source" in ix else ""
    key = r[0] if "Line No" in ix else src
    if key in agg:
        a = agg[key]; a[0] += inst; a[1] += th; a[2] += smp
    else:
        agg[key] = [inst, th, smp, key, src[:120]]
out = [tuple(v) for v in agg.values()]
tot = sum(o[0] for o in out) or 1
tots = sum(o[2] for o in out) or 1
print(f"total warp-instr {tot}  samples {tots}")
for inst, th, smp, k, src in sorted(out, reverse=True)[:top]:
    print(f"{100*inst/tot:5.1f}% inst {100*smp/tots:5.1f}% smp  lanes {th/max(inst,1):5.1f}  {k}  {src}")
