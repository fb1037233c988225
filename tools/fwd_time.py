"""Time the forward-only 3-D kernel (cfg4: 100 x 1024 x 1024, one GPU) -- A/B runs of tile heights via PEA_B200_LIB.
usage: python tools/fwd_time.py [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pea_b200 as P  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
r = bench.time_zshard(P, dev, 0, 1, steps)
print(f"cfg4 forward: {r['ms_per_step']:.3f} ms  {r['hbm_gbs_aggregate']:.0f} GB/s = {r['hbm_gbs_aggregate'] / 6459.6 * 100:.1f}% of measured HBM peak")
r = bench.time_inference_epilogue(P, dev)
print(f"cfg4 forward + border replicate + relu: torch epilogue {r['torch_epilogue_ms']:.3f} ms, fused {r['fused_ms']:.3f} ms")
