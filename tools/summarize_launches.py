"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.
    python tools/summarize_launches.py gpurun_out/launches.csv [steps] > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("mk::", "")
    agg.setdefault((name, r[7], r[8]), []).append(float(r[-1]))
tot = sum(sum(v) for v in agg.values())
print(f"# ncu launch list: {len(rows)} launches, {tot / 1e3:.1f} us total ({steps} step(s) captured; cold-cache, serialised: compare SHARES)\n")
print("| share | total us | launches | avg us | kernel | block | grid |")
print("|---:|---:|---:|---:|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"| {100 * sum(v) / tot:.1f}% | {sum(v) / 1e3:.1f} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | `{k[0][:70]}` | {k[1]} | {k[2]} |")
