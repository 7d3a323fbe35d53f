"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  launches <csv> <out.md> <title>     : per-kernel totals of an `ncu --metrics gpu__time_duration.sum` launch list
  full <ncu-rep> <out.md> <title>     : key metrics + top stall reasons of an `ncu --set full` capture
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path, out, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit.startswith("ns") or unit == "nsecond" else (v * 1000 if unit.startswith("ms") else v)
        a = agg.setdefault(name, [0, 0.0, 1e18, 0.0])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none --csv` (cold-cache, serialised launches: compare SHARES, not absolutes).\n\n")
        f.write("| kernel | launches | total ms | avg us | min us | max us | share |\n|---|---:|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1] / 1000:.3f} | {a[1] / a[0]:.2f} | {a[2]:.2f} | {a[3]:.2f} | {100 * a[1] / tot:.1f}% |\n")
        f.write(f"\nTotal kernel time in the captured window: {tot / 1000:.2f} ms over {sum(a[0] for a in agg.values())} launches.\n")


WANT = ["Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__cycles_active.avg"]


def full(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `ncu --set full --clock-control none --import-source on` ({rep.split('/')[-1]}); ncu flushes caches between replays, so DRAM bytes are cold-cache.\n")
        for d in data:
            name = re.sub(r"\(.*", "", d[idx["Kernel Name"]])
            f.write(f"\n## `{name}`  (launch id {d[idx['ID']]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in WANT:
                if w in idx:
                    f.write(f"| {w} | {d[idx[w]]} | {units[idx[w]]} |\n")
            try:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd = float(d[idx["dram__bytes_read.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_read.sum"]]]
                wr = float(d[idx["dram__bytes_write.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_write.sum"]]]
                f.write(f"| traffic = dram read + write | {(rd + wr) / 1e6:.3f} | Mbyte |\n")
            except Exception:
                pass
            st = []
            for h in stall_cols:
                try:
                    st.append((float(d[idx[h]].replace(",", "")), h))
                except ValueError:
                    pass
            f.write("\nTop stall reasons (warps per issue-active cycle): " +
                    ", ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0]} {v:.1f}" for v, h in sorted(st, reverse=True)[:5]) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
