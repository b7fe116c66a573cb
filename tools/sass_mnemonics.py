#!/usr/bin/env python
"""Counts, per kernel of libptpreshape.so, the SASS mnemonics that identify the hardware path (tcgen05 = UTCHMMA / LDTM / STTM,
TMA = UTMALDG / UBLKCP, mma.sync = HMMA, ...) and writes profiles/r2_sass_mnemonics.md.  No GPU needed (cuobjdump)."""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "proxytransformation_b200/csrc/libptpreshape.so")], capture_output=True, text=True).stdout
COLS = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "HMMA", "LDSM", "LDGSTS", "SYNCS")
pat = re.compile(r"\b(" + "|".join(COLS) + r")\b")
kern, counts = None, collections.OrderedDict()
for line in out.split("\n"):
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        kern = m.group(1); counts[kern] = collections.Counter(); continue
    if kern:
        for t in pat.findall(line):
            counts[kern][t] += 1
dem = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.split("\n")
lines = ["# SASS mnemonics of every kernel in libptpreshape.so (round 2)", "",
         "`cuobjdump -sass proxytransformation_b200/csrc/libptpreshape.so`, instruction counts per kernel of the mnemonics that identify the",
         "hardware path (B200_PROFILING.md: `tcgen05.mma` = UTCHMMA, `tcgen05.ld/st` = LDTM/STTM, TMA = UTMALDG / UBLKCP, `mma.sync` = HMMA,",
         "`ldmatrix` = LDSM, `cp.async` = LDGSTS, mbarrier = SYNCS, `tcgen05.commit` = UTCBAR).  Regenerate with `python tools/sass_mnemonics.py`.", "",
         "| kernel | " + " | ".join(COLS) + " |", "|---|" + "---:|" * len(COLS)]
for k, d in zip(counts, dem):
    c = counts[k]
    lines.append(f"| `{re.sub(r'[(].*', '', d)[:70]}` | " + " | ".join(str(c.get(x, 0) or "") for x in COLS) + " |")
open(os.path.join(ROOT, "profiles", "r2_sass_mnemonics.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(l for l in lines if "UTCHMMA" in l or any(ch.isdigit() for ch in l.split("|", 2)[-1])))
