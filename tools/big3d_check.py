import os, sys, torch
sys.path.insert(0, "/root/repo")
import pea_b200 as P
from pea_b200 import _ops as ops
dev = torch.device("cuda", 0)
for (B, D, H, W) in [(1, 40, 256, 512), (2, 5, 1000, 36), (1, 100, 1024, 1024)]:
    g = torch.Generator(device="cuda").manual_seed(D)
    emb = torch.randn((B, 16, D, H, W), device=dev, generator=g)
    tgt = (torch.rand((B, 12, D, H, W), device=dev, generator=g) > 0.5).float()
    wgt = torch.rand((B, 12, D, H, W), device=dev, generator=g) + 0.1
    out = {}
    for name, impl in (("generic", ops.IMPL_GENERIC), ("tiled", ops.IMPL_AUTO)):
        ops.set_default_impl(impl)
        e = emb.clone().requires_grad_(True)
        l, a = P.embedding_loss_norm5(e, tgt, wgt, P.WeightedMSE())
        l.backward()
        out[name] = (l.detach(), a, e.grad)
        ops.set_default_impl(ops.IMPL_AUTO)
    errs = [float((out["tiled"][i] - out["generic"][i]).abs().max() / out["generic"][i].abs().max()) for i in range(3)]
    print((B, D, H, W), " ".join(f"{e:.2e}" for e in errs), "OK" if max(errs) < 1e-5 else "FAIL", flush=True)
    del out, emb, tgt, wgt
    torch.cuda.empty_cache()
