import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pea_b200 as P
dev = torch.device("cuda", 0)
for name in ("cfg2", "cfg3"):
    r = bench.time_gt(P, bench.WORKLOADS[name], dev, 50)
    print(name, f"{r['ms_per_step']*1e3:.1f} us  {r['hbm_gbs']:.0f} GB/s = {r['hbm_gbs']/6459.6*100:.1f}%", flush=True)
