"""The six loss calls of one AC3/AC4 training step (scripts_ac3ac4/main.py:218-231): norm5 + two-view norm5 at full
resolution, norm1 at the four deep-supervision scales.  Times the public API eagerly (what a training loop pays,
Python + launches) and as a CUDA-graph replay, and lists the kernels of one step.
usage: python tools/step6_time.py [iters]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pea_b200 as P  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = torch.device("cuda", 0)
B, D = 4, 18
g = torch.Generator(device="cpu").manual_seed(0)


def rnd(*s):
    return torch.randn(s, generator=g).to(dev)


def tw(K, h):
    t = (torch.rand((B, K, D, h, h), generator=g) > 0.5).float().to(dev)
    w = (torch.rand((B, K, D, h, h), generator=g) * 2 + 0.1).to(dev)
    return t, w


emb, ema = rnd(B, 16, D, 160, 160).requires_grad_(True), rnd(B, 16, D, 160, 160)
tgt, wgt = tw(12, 160)
scales = [10, 20, 40, 80]
emds = [rnd(B, 16, D, h, h).requires_grad_(True) for h in scales]
downs = [tw(3, h) for h in scales]
crit = P.WeightedMSE()


MULTI = os.environ.get("PEA_STEP6_MULTI", "1") == "1"


def step():
    l0, pred = P.embedding_loss_norm5(emb, tgt, wgt, crit, affs0_weight=1)
    l1, _ = P.ema_embedding_loss_norm5(emb, ema, tgt, wgt, crit, affs0_weight=1)
    if MULTI:
        ls = [r[0] for r in P.embedding_loss_norm1_multi(emds, [d[0] for d in downs], [d[1] for d in downs], crit, affs0_weight=1)]
    else:
        ls = [P.embedding_loss_norm1(e, t, w, crit, affs0_weight=1)[0] for e, (t, w) in zip(emds, downs)]
    loss = l0 + l1 + ls[0] + ls[1] + ls[2] + ls[3]
    loss.backward()
    emb.grad = None
    for e in emds:
        e.grad = None
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(iters):
    step()
torch.cuda.synchronize()
eager = (time.perf_counter() - t0) / iters * 1e6
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2):
        step()
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    loss = step()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
per = []
for _ in range(iters):
    flush.zero_()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    graph.replay()
    t.record()
    torch.cuda.synchronize()
    per.append(s.elapsed_time(t) * 1e3)
per.sort()
print(f"step6 (3-D, B={B}, multi={MULTI}): eager {eager:.1f} us/step   graph replay median {per[len(per) // 2]:.1f} us  min {per[0]:.1f} us   loss {float(loss):.6f}")
if os.environ.get("PEA_STEP6_PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
