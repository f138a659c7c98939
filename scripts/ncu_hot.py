#!/usr/bin/env python
"""Top stall locations of one kernel from `ncu -i rep --page source --csv --kernel-id :::N` output."""
import csv, sys
f = open(sys.argv[1]); next(f)
rows = [r for r in csv.DictReader(f) if (r.get("# Samples") or "").isdigit()]
# the file holds SASS rows then (maybe) source rows; keep the SASS part (has Address starting with 0x)
rows = [r for r in rows if r["Address"].startswith("0x")]
tot = sum(int(r["# Samples"]) for r in rows)
print("total samples", tot, "n instr", len(rows))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"]))[:n]
stall = [k for k in rows[0].keys() if k and k.startswith("stall_") and "Not Issued" not in k]
for i in sorted(idx):
    r = rows[i]
    top = sorted(((int(r[k] or 0), k) for k in stall), reverse=True)[:2]
    print(i, r["# Samples"], r["Instructions Executed"], r["Source"].strip()[:90], top)
