"""GT generation timing (labels -> target / mask / weight) at cfg2 and cfg3: eager back-to-back calls (bench.time_gt) and
CUDA-graph replay of the same call with the L2 flushed between replays; kernel names by torch.profiler.
usage: python tools/gt_time.py"""
import os, sys, torch
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pea_b200 as P
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in ("cfg2", "cfg3"):
    w = bench.WORKLOADS[name]
    r = bench.time_gt(P, w, dev, 50)
    print(name, f"eager {r['ms_per_step']*1e3:.1f} us  {r['hbm_gbs']:.0f} GB/s = {r['hbm_gbs']/6459.6*100:.1f}%", flush=True)
    lab = bench.make_labels(w, 0)
    if w["nd"] == 2:
        t = torch.tensor(lab).to(dev)
        offsets = P.multi_offset(w["shifts"], 4)
        fn = lambda: P.make_targets_2d(t, offsets)
    else:
        t = torch.tensor(lab.astype(np.int32)).to(dev)
        fn = lambda: P.make_targets_3d(t)
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        with torch.cuda.graph(g, stream=st):
            out = fn()
    torch.cuda.synchronize()
    per = []
    for i in range(30):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        per.append(a.elapsed_time(b) * 1e3)
    per = sorted(per[3:])
    ab = r["algorithmic_bytes"]
    med = per[len(per) // 2]
    print(name, f"graph replay, L2 flushed: median {med:.1f} us min {per[0]:.1f} us  {ab / med / 1e3:.0f} GB/s = {ab / med / 1e3 / 6459.6 * 100:.1f}%", flush=True)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    for ev in prof.key_averages():
        if ev.device_time_total > 0:
            print("   ", ev.key[:110], f"{ev.device_time_total:.1f} us x{ev.count}")
