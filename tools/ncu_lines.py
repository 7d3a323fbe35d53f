"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line: samples, instructions, top stalls.
usage: ncu_lines.py <export.csv> [top N]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
cur_file = ""
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci = {}
        for i, h in enumerate(hdr):
            ci.setdefault(h, i)
        stall = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":      # per-line summary rows have "-" in the Address column
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    key = (cur_file, line)
    a = agg.setdefault(key, {"src": r[1].strip(), "samples": 0, "inst": 0, "st": {}})
    a["samples"] += int(r[ci["# Samples"]] or 0)
    a["inst"] += int(r[ci["Instructions Executed"]] or 0)
    for h, i in stall:
        v = int(r[i] or 0)
        if v:
            a["st"][h[6:]] = a["st"].get(h[6:], 0) + v
ts = sum(a["samples"] for a in agg.values()) or 1
ti = sum(a["inst"] for a in agg.values()) or 1
print(f"total samples {ts}, warp instructions {ti}")
for (f, line), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(a["st"].items(), key=lambda kv: -kv[1])[:3]
    print(f"{f}:{line:<5d} samp {100 * a['samples'] / ts:5.1f}%  inst {100 * a['inst'] / ti:5.1f}%  {a['src'][:100]:100s} | " + " ".join(f"{k}:{v}" for k, v in st))
