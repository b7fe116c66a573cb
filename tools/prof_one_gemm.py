import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from proxytransformation_b200 import ops
M, N, K, act = [int(x) for x in sys.argv[1:5]]
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = "cuda"
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5
As, Ws = ops.split_bf16(A), ops.split_bf16(W)
bias = torch.randn(N, device=dev); C = torch.empty(M, N, device=dev)
for _ in range(4):
    ops.gemm_tc(As, Ws, M, N, K, bias=None, act=act, C=C, ldc=N, bn=bn)
torch.cuda.synchronize()
