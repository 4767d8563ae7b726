"""Top CUDA source lines from `ncu -i rep --page source --print-source cuda,sass --csv`, by warp instructions executed."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
cur_file = ""
hdr = None
agg = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 3 and r[0] == "Line No":
        hdr = r
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and len(r) > ie and r[0].strip().isdigit():
        try:
            agg.append((int(r[ie] or 0), int(r[isamp] or 0), cur_file, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot = sum(a[0] for a in agg) or 1
tots = sum(a[1] for a in agg) or 1
print(f"-- CUDA source lines: {tot} warp instructions, {tots} samples")
for a in sorted(agg, key=lambda a: -a[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 60]:
    print(f"{a[0] / tot * 100:5.1f}% inst {a[1] / tots * 100:5.1f}% samp  {a[2]}:{a[3]:<4d} {a[4][:120]}")
if len(sys.argv) > 3:  # line ranges of one file: name:lo-hi,...
    for spec in sys.argv[3].split(","):
        name, rng = spec.split(":")
        lo, hi = (int(v) for v in rng.split("-"))
        sel = [a for a in agg if a[2] == "render_span.cu" and lo <= a[3] <= hi]
        print(f"  {name:12s} L{lo}-{hi}: {sum(a[0] for a in sel) / tot * 100:5.1f}% inst {sum(a[1] for a in sel) / tots * 100:5.1f}% samp")
    other = [a for a in agg if a[2] != "render_span.cu"]
    print(f"  other files: {sum(a[0] for a in other) / tot * 100:5.1f}% inst")
