"""Hot spots of one kernel of an .ncu-rep captured with `--set full --import-source on`: the SASS instructions with the
most warp-stall samples, and the same aggregated per source line (needs -lineinfo).  Runs on the GPU box right after the
capture so that only the text travels back.     usage: ncu_hot.py file.ncu-rep [kernel-name-substring] [top-n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep, want, top = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else ""), int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"] + (["-k", "regex:" + want] if want else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next((i for i, r in enumerate(rows) if any("Sampling" in c for c in r)), None)
if hdr_i is None:
    print("no source page:", out[:2000])
    sys.exit(0)
hdr = rows[hdr_i]


def col(name):
    for i, h in enumerate(hdr):
        if name.lower() in h.lower():
            return i
    return None


c_src, c_samp, c_exec, c_line = col("Source"), col("Warp Stall Sampling (All"), col("Instructions Executed"), col("# Line") or col("Address")
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = sum(float(r[c_samp] or 0) for r in body) or 1.0
print(f"kernel filter '{want}': {len(body)} SASS instructions, {tot:.0f} stall samples")
print("columns:", hdr)
# cumulative share by SASS opcode
from collections import Counter
by_op = Counter()
for r in body:
    toks = r[c_src].split()
    op = (toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")).split(".")[0]
    by_op[op] += float(r[c_samp] or 0)
print("stall samples by opcode:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in by_op.most_common(25)))
body.sort(key=lambda r: -float(r[c_samp] or 0))
for r in body[:top]:
    print(f"{100 * float(r[c_samp] or 0) / tot:6.2f}%  exec={r[c_exec]:>10s}  {r[c_src][:110]}")
