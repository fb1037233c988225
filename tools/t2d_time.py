"""Timing only of the 2-D bench workloads (cfg2, cfg1, cfg2x) -- for A/B runs of ablation builds via PEA_B200_LIB.
usage: python tools/t2d_time.py [steps] [names...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pea_b200 as P  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
names = sys.argv[2:] or ["cfg2", "cfg1", "cfg2x"]
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in names:
    w = bench.WORKLOADS[name]
    r = bench.time_workload(P, w, steps, 5, dev, flush, False)
    per = sorted(r["per_ms"])
    ab, K = bench.algorithmic_bytes_per_step(w)
    ms = sum(per) / len(per)
    print(f"{name}: mean {ms * 1e3:.1f} us  min {per[0] * 1e3:.1f} us   {ab / (ms * 1e-3) / 1e9:.0f} GB/s = "
          f"{ab / (ms * 1e-3) / 1e9 / 6459.6 * 100:.1f}% of measured HBM peak   loss {r['loss']:.6f}", flush=True)
