#!/usr/bin/env python
"""Split a kernel's SASS at BAR.SYNC/EXIT and report samples / executed instructions / op mix per segment.
usage: ncu_segments.py rep kernel-substr units"""
import csv, io, subprocess, sys
rep, filt, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if len(row) >= 2 and row[0] == "Kernel Name": cur = {"name": row[1], "rows": []}; blocks.append(cur)
    elif cur is not None: cur["rows"].append(row)
b = [b for b in blocks if filt in b["name"]][0]
hdr, data = b["rows"][0], b["rows"][1:]
ix = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return int(float(r[ix[k]] or 0))
    except Exception: return 0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
segs, st = [], 0
for i, r in enumerate(data):
    if "BAR.SYNC" in r[ix["Source"]] or "EXIT" in r[ix["Source"]]: segs.append((st, i)); st = i + 1
segs.append((st, len(data) - 1))
tot = sum(num(r, "# Samples") for r in data)
for a, e in segs:
    rows = data[a:e + 1]
    smp = sum(num(r, "# Samples") for r in rows)
    if smp < 0.005 * tot: continue
    exe = sum(num(r, "Instructions Executed") for r in rows)
    mix = {}
    for key in ("HMMA", "LDS", "STS", "LDG", "STG", "SYNCS", "PRMT", "SHF", "MUFU", "BRA"):
        mix[key] = round(sum(num(r, "Instructions Executed") for r in rows if key in r[ix["Source"]]) / units, 1)
    sa = sorted(((h[6:], sum(num(r, h) for r in rows)) for h in stalls), key=lambda kv: -kv[1])[:4]
    print(f"[{a:4d}..{e:4d}] samples {100 * smp / tot:5.1f}%  exe/unit {exe / units:8.1f}  {mix}  stalls {sa}")
