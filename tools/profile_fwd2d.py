"""A few embedding2affs calls at cfg2 (for ncu).  usage: python tools/profile_fwd2d.py [cfg2|cfg1]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pea_b200 as P
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = bench.WORKLOADS[name]
emb = torch.randn((w["B"], 16) + tuple(w["spatial"]), generator=torch.Generator().manual_seed(0)).cuda()
offsets = P.multi_offset(w["shifts"], 4)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    for _ in range(3):
        flush.zero_()
        a = P.embedding2affs(emb, offsets)
torch.cuda.synchronize()
print("ok", float(a.sum()))
