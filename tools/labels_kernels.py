"""Kernel durations (torch.profiler / CUPTI) of the label-driven losses: counting pass + loss kernel, cfg2 and cfg3.
usage: python tools/labels_kernels.py"""
import os, sys, torch
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pea_b200 as P
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
for name in ("cfg2", "cfg3"):
    w = bench.WORKLOADS[name]
    lab = torch.tensor(bench.make_labels(w, 0).astype(np.int32)).to(dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    emb = torch.randn((w["B"], 16) + tuple(w["spatial"]), generator=g).to(dev)
    e = emb.clone().requires_grad_(True)
    crit = P.WeightedMSE()
    if w["nd"] == 2:
        offsets = P.multi_offset(w["shifts"], 4)
        fn = lambda: P.embedding_loss_labels(e, lab, crit, offsets)[0]
    else:
        fn = lambda: P.embedding_loss_norm5_labels(e, lab, crit)[0]
    for _ in range(3):
        fn().backward(); e.grad = None
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            fn().backward(); e.grad = None
        torch.cuda.synchronize()
    print(name)
    for ev in sorted(prof.key_averages(), key=lambda ev: -ev.device_time_total):
        if ev.device_time_total > 0:
            print(f"   {ev.device_time_total / ev.count:8.1f} us x{ev.count:3d}  {ev.key[:120]}")
