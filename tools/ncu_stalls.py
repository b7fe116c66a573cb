#!/usr/bin/env python
"""Per-kernel stall-reason totals and hottest SASS lines from an .ncu-rep (source page).  usage: ncu_stalls.py rep [kernel-substr] [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; filt = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if len(row) >= 2 and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
seen = set()
for b in blocks:
    if filt not in b["name"] or b["name"] in seen or len(b["rows"]) < 2: continue
    seen.add(b["name"])
    hdr, data = b["rows"][0], b["rows"][1:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    num = lambda r, k: int(float(r[ix[k]] or 0)) if k in ix and len(r) > ix[k] else 0
    tot = sum(num(r, "# Samples") for r in data)
    print("=====", b["name"][:100], "samples", tot, "static instr", len(data), "warp-instr executed", sum(num(r, "Instructions Executed") for r in data))
    agg = sorted(((h, sum(num(r, h) for r in data)) for h in stalls), key=lambda kv: -kv[1])
    print("  stalls:", [(h, v) for h, v in agg if v][:8])
    for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:topn]:
        st = sorted(((h, num(r, h)) for h in stalls), key=lambda kv: -kv[1])[:2]
        print(f"  {num(r, '# Samples'):6d} {100.0 * num(r, '# Samples') / max(tot, 1):5.1f}%  exe={num(r, 'Instructions Executed'):9d}  {r[ix['Source']][:70]:70s} {st}")
