"""Top stalled SASS instructions of one kernel from `ncu -i rep --page source --csv --kernel-name regex:X` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
num = lambda v: int(float(v)) if v not in ("", None) else 0  # noqa: E731
tot = sum(num(r[isamp]) for r in data)
print("total samples", tot, "SASS lines", len(data), "warp-instr executed", sum(num(r[iex]) for r in data))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + num(r[i])
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(data, key=lambda r: -num(r[isamp]))[:N]:
    st = sorted(((num(r[i]), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(r[isamp].rjust(6), r[iex].rjust(9), r[isrc][:100].ljust(100), st)
