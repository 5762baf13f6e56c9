"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` export per CUDA source line:
stall samples, executed warp instructions and the dominant stall reasons.   python tools/ncu_source_lines.py export.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
fname, hdr, data = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0] in ("Function Name",) or hdr is None or r[0] == "":
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    def num(v):
        try:
            return int(float(v))
        except ValueError:
            return 0
    d = dict(zip(hdr[4:], r[4:]))
    stalls = {k[6:]: num(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k}
    data.append((fname, line, num(d.get("# Samples", "")), num(d.get("Instructions Executed", "")), stalls, r[1]))
tot, toti = sum(d[2] for d in data), sum(d[3] for d in data)
print(f"total samples {tot}, warp instructions {toti}")
for f, line, s, i, st, src in sorted(data, key=lambda x: -x[2])[:top]:
    why = ",".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{f}:{line:4d} {100 * s / tot:5.1f}% smp {100 * i / max(toti, 1):5.1f}% inst  [{why}]  {src.strip()[:100]}")
