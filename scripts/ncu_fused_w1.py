"""Tiny driver for ncu: runs the fused SRA kernel (world=1: load -> min/max ->
quantize -> pack -> self-decode) on a 64 MB fp32 buffer a few times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_cgx_b200 as cgx
C = cgx._C
n = 16 * 1024 * 1024
x = torch.randn(n, device="cuda")
g = C.LocalSraGroup(1, int(os.environ.get("LANES", "148")), 80 << 20, 5000, 2048)
for _ in range(4):
    g.allreduce([x], [(0, n, 4, 512)])
torch.cuda.synchronize()
print("ok")
