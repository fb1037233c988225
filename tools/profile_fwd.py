"""Two forward passes of the 12-channel affinity kernel on the cfg4 volume (100 x 1024 x 1024, one GPU) for ncu.
usage: python tools/profile_fwd.py [D]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pea_b200 as P  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
emb = torch.empty((1, 16, D, 1024, 1024), device=dev).normal_(generator=g)
for _ in range(2):
    out = P.inf_embedding_loss_norm5(emb)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 1, 5, 5]))
