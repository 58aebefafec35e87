"""Summarise an `ncu --page source --csv --print-source sass` export: top SASS lines by stall samples, with the dominant stall
reason of each.  Dev tool.   usage: python tools/ncu_hot_sass.py src.csv [topN]"""
import csv, sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
body = [r for r in rows[2:] if len(r) >= len(hdr) - 2]
ci = {h: i for i, h in enumerate(hdr)}
samp, nis, exe = ci["# Samples"], ci["Warp Stall Sampling (Not-issued Samples)"], ci["Instructions Executed"]
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
if not stall_cols:
    stall_cols = [(i, h) for i, h in enumerate(hdr) if i > ci.get("Divergent Branches", 10) and h and h[0].islower()]
tot = sum(int(r[samp] or 0) for r in body)
print("total samples", tot, " lines", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][samp] or 0))[:top]
for i in sorted(order):
    r = body[i]
    reasons = sorted(((int(r[c] or 0), h) for c, h in stall_cols if r[c] not in ("", "-") and r[c].isdigit()), reverse=True)[:2]
    print("%5d %5.1f%% exec %8s  %-60s %s" % (i, 100.0 * int(r[samp] or 0) / max(tot, 1), r[exe], r[ci["Source"]].strip()[:60],
                                             " ".join("%s=%d" % (h.replace("stall_", ""), v) for v, h in reasons if v)))
