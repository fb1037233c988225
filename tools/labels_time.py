"""Time labels -> loss + gradient at cfg2: two-step route (GT kernels + fused loss) vs the label-driven loss kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pea_b200 as P  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
r = bench.time_labels_pipeline(P, bench.WORKLOADS["cfg2"], dev, flush)
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k != "workload"})
