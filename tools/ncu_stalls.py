"""Summarise an `ncu --page source --csv` export: stall reasons over the kernel and the hottest SASS lines."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
lines = []
for r in rows[2:]:
    if len(r) < len(hdr) or not (r[idx["# Samples"]] or "0").isdigit(): continue
    s = int(r[idx["# Samples"]] or 0)
    for h in stall_cols:
        tot[h] += int(r[idx[h]] or 0)
    lines.append((s, r[idx["Source"]].strip(), {h: int(r[idx[h]] or 0) for h in stall_cols if int(r[idx[h]] or 0) > 0}, int(r[idx["Instructions Executed"]] or 0)))
n = sum(tot.values())
print("total samples", n)
for h, v in tot.most_common(8):
    print(f"  {h:28s} {v:8d} {v/n:.3f}")
print("hottest instructions:")
for i, (s, src, d, ie) in sorted(enumerate(lines), key=lambda x: -x[1][0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = sorted(d.items(), key=lambda x: -x[1])[:2]
    print(f"  #{i:5d} {s:7d} exec={ie:9d} {src[:70]:70s} {top}")
