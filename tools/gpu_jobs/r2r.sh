#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -6 > gpurun_out/r2r_tests.log
QB=64 timeout 300 python - <<'PY' >> gpurun_out/r2r_tests.log 2>&1
import os, sys, torch
sys.path.insert(0, os.getcwd())
from proxytransformation_b200 import ProxyTransformationNormReverse, ops, _lib, synthetic as syn
B, V = 64, 196
cfg = syn.C2_WIDE
m = ProxyTransformationNormReverse(**cfg.module_kwargs()).eval()
m.load_state_dict(syn.make_state_dict(cfg, 0, bf16_round=True))
m = m.cuda()
w = m._weights(torch.device("cuda"))
for dt in (torch.bfloat16, torch.float16):
    imgs = [(torch.relu(torch.randn(B, V, 512, 15, 15, device="cuda")) * 1.5).to(dt) for _ in range(2)]
    for k in range(3): ops.img_attnpool(imgs[k & 1], w["img"], 8, params=w["img_struct"])
    torch.cuda.synchronize(); _lib.profile_enable(True)
    for k in range(10): ops.img_attnpool(imgs[k & 1], w["img"], 8, params=w["img_struct"])
    torch.cuda.synchronize(); prof = _lib.profile_read(); _lib.profile_enable(False)
    print(dt, {k: round(v[0] / 10, 4) for k, v in prof.items()})
    del imgs
PY
cat gpurun_out/r2r_tests.log
