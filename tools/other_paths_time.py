"""Timing of the API paths the headline legs do not cover, at BASELINE's sizes (CUDA-graph replay, L2 flushed): full-resolution
norm1 losses (TRAIN.embedding_mode == 1), validation calls under no_grad, inf_embedding_loss_norm1.
usage: python tools/other_paths_time.py"""
import os, sys, torch
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pea_b200 as P
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
PEAK = 6459.6
crit = P.WeightedMSE()

def report(name, ms, nbytes):
    print(f"{name:58s} {ms * 1e3:8.1f} us   {nbytes / ms / 1e6:7.0f} GB/s = {nbytes / ms / 1e6 / PEAK * 100:5.1f}% of the HBM roofline", flush=True)

# ---- 3-D, cfg3 shape
w = bench.WORKLOADS["cfg3"]
lab, emb, tgt, wgt, msk, offsets = bench.device_inputs(P, w, 0, dev)
vox = w["B"] * int(np.prod(w["spatial"]))
ema = torch.randn_like(emb)
t3, w3 = tgt[:, :3].contiguous(), wgt[:, :3].contiguous()
e = emb.clone().requires_grad_(True)
def step(fn):
    def f():
        loss = fn()
        loss.backward()
        e.grad = None
    return f
report("embedding_loss_norm1 fwd+bwd (cfg3 shape, K=3)", bench.time_graph(step(lambda: P.embedding_loss_norm1(e, t3, w3, crit)[0]), flush), (64 + 12 + 12 + 64 + 12) * vox)
report("ema_embedding_loss_norm1 fwd+bwd", bench.time_graph(step(lambda: P.ema_embedding_loss_norm1(e, ema, t3, w3, crit)[0]), flush), (128 + 12 + 12 + 64 + 12) * vox)
with torch.no_grad():
    report("inf_embedding_loss_norm1 (forward, K=3)", bench.time_graph(lambda: P.inf_embedding_loss_norm1(emb), flush), (64 + 12) * vox)
    report("inf_embedding_loss_norm5 (forward, K=12)", bench.time_graph(lambda: P.inf_embedding_loss_norm5(emb), flush), (64 + 48) * vox)
    report("embedding_loss_norm5 under no_grad (validation)", bench.time_graph(lambda: P.embedding_loss_norm5(emb, tgt, wgt, crit), flush), (64 + 48 + 48 + 48) * vox)
    report("embedding_loss_norm1 under no_grad (validation)", bench.time_graph(lambda: P.embedding_loss_norm1(emb, t3, w3, crit), flush), (64 + 12 + 12 + 12) * vox)
del lab, emb, tgt, wgt, e, ema
# ---- 2-D, cfg2 / cfg1
for name in ("cfg2", "cfg1"):
    w = bench.WORKLOADS[name]
    lab, emb, tgt, wgt, msk, offsets = bench.device_inputs(P, w, 0, dev)
    px = w["B"] * int(np.prod(w["spatial"]))
    K = len(offsets)
    with torch.no_grad():
        report(f"{name} embedding_loss under no_grad (validation)", bench.time_graph(lambda: P.embedding_loss(emb, tgt, wgt, msk, crit, offsets), flush), (64 + 9 * K) * px + 4 * K * px)
        report(f"{name} embedding2affs", bench.time_graph(lambda: P.embedding2affs(emb, offsets), flush), (64 + 4 * K) * px)
# ---- a 12-offset 3-D table no specialised kernel covers (diagonals), cfg3 shape: table-driven kernel vs work-list kernel
from pea_b200 import _ops as ops
w = bench.WORKLOADS["cfg3"]
g = torch.Generator().manual_seed(0)
emb = torch.randn((w["B"], 16) + tuple(w["spatial"]), generator=g).cuda()
K = 12
offs = np.asarray([[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-1, -1, -1], [-1, 1, 1], [-1, -1, 1], [-1, 1, -1], [0, -9, 0], [0, 0, -9],
                   [0, -9, -9], [0, 9, -9], [0, -9, -4]], np.int32)
tgt = (torch.rand((w["B"], K) + tuple(w["spatial"]), generator=g) > 0.5).float().cuda()
wgt = (torch.rand((w["B"], K) + tuple(w["spatial"]), generator=g) + 0.1).cuda()
norm, coef = np.full(K, 1e5, np.float32), np.ones(K, np.float32)
vox = w["B"] * int(np.prod(w["spatial"]))
for name, impl in (("auto (work-list kernel)", ops.IMPL_AUTO), ("table-driven kernel", ops.IMPL_GENERIC)):
    ms = bench.time_graph(lambda: ops.affs_loss_fwd_bwd(emb, None, offs, ops.BORDER_ZERO, tgt, wgt, None, norm, coef, impl=impl), flush)
    report(f"shift_func(15)-like 12 offsets, fwd+bwd, {name}", ms, 272 * vox)
from pea_b200._autograd import Affinities
from pea_b200.data_affinity import NHOOD_NORM5
# ---- K5: gradient through the affinities for criteria the fused kernels do not implement (MSELoss on the affinities here)
for name in ("cfg2", "cfg3"):
    w = bench.WORKLOADS[name]
    lab, emb, tgt, wgt, msk, offsets = bench.device_inputs(P, w, 0, dev)
    vox = w["B"] * int(np.prod(w["spatial"]))
    K = tgt.shape[1]
    e = emb.clone().requires_grad_(True)
    up = torch.randn_like(tgt)
    def k5():
        a = P.embedding2affs(e, offsets) if w["nd"] == 2 else Affinities.apply(e, None, np.asarray(NHOOD_NORM5, np.int32), ops.BORDER_ZERO, ops.IMPL_AUTO)
        a.backward(up)
        e.grad = None
    report(f"{name} affinities forward + backward through them (K5)", bench.time_graph(k5, flush), (64 + 4 * K + 64 + 4 * K + 64) * vox)
