"""Kernel breakdown of one bench step (device-resident inputs): python tools/kb.py [kernel-name-substrings...]"""
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--steps", "20", "--warmup", "40", "--no-e2e", "--no-cpu-baseline", "--no-checks", "--no-extra"],
                     capture_output=True, text=True).stdout.strip().splitlines()[-1]
d = json.loads(out)
kb = d["kernel_breakdown"]
sel = sys.argv[1:]
print(f"ms_per_step {d['ms_per_step']:.4f}  " + "  ".join(f"{k} {v['ms_per_step']:.4f}" for k, v in sorted(kb.items(), key=lambda kv: -kv[1]['ms_per_step']) if not sel or any(s in k for s in sel)))
