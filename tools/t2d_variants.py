import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
import bench, pea_b200 as P
from pea_b200 import _ops as ops
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["cfg2"]
lab, emb, tgt, wgt, msk, offsets = bench.device_inputs(P, w, 0, dev)
offs3 = ops.offsets3(offsets, 2)
K = len(offsets)
norm = np.full(K, float(emb.shape[0] * emb.shape[-1]), np.float32); coef = np.ones(K, np.float32)
e5 = emb.unsqueeze(2).contiguous(); t5 = tgt.unsqueeze(2).contiguous(); w5 = wgt.unsqueeze(2).contiguous(); m5 = msk.unsqueeze(2).contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(mask, grad=True, partner=None):
    for _ in range(3):
        ops.affs_loss_fwd_bwd(e5, partner, offs3, ops.BORDER_WRAP, t5, w5, mask, norm, coef, want_grad=grad)
    ts = []
    for _ in range(20):
        flush.zero_()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.affs_loss_fwd_bwd(e5, partner, offs3, ops.BORDER_WRAP, t5, w5, mask, norm, coef, want_grad=grad)
        t.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(t))
    return sorted(ts)[len(ts)//2] * 1e3
print("u8 mask   :", timeit(m5))
print("f32 mask  :", timeit(m5.float()))
print("no mask   :", timeit(None))
print("no grad   :", timeit(m5, grad=False))
