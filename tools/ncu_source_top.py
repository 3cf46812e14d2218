"""stdin: `ncu --page source --csv` of one kernel -> the 40 source lines with the most warp stall samples."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
while rows and (not rows[0] or rows[0][0] != "Address"):       # a "Kernel Name" line precedes the header
    rows.pop(0)
if len(rows) < 2:
    sys.exit(0)
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
samp = next((h for h in hdr if "Warp Stall Sampling (All" in h), None) or next((h for h in hdr if "Sampling" in h), None)
src = next((h for h in hdr if h.strip() in ("Source", "SASS")), hdr[1])
out = []
for r in rows[1:]:
    try:
        out.append((float(r[idx[samp]] or 0), r[idx[src]]))
    except Exception:
        pass
tot = sum(v for v, _ in out) or 1.0
for v, s in sorted(out, reverse=True)[:40]:
    print("%6.2f%%  %s" % (100 * v / tot, s[:150]))
