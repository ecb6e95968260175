"""Top stalled instructions of one kernel from `ncu --page source --csv` output (test/profiling helper)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) == len(hdr)]
base = int(data[0][ia], 16)
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "instructions", len(data))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -int(r[isamp]))[:top]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print("%5x %6s %7s  %-60s %s" % (int(r[ia], 16) - base, r[isamp], r[iex], r[isrc].strip()[:60], st))
