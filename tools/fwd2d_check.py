"""Development check of the forward-only 2-D kernel (pea_affs2d_fwd.cu): against the table-driven generic kernel and the
older tiled kernel (PEA_KNOB_NO_CANON) through the C ABI -- both borders, relu on / off, tables with 1 .. 6 shifts up to
27, partial tiles, capped grids -- and against the oracle; then timing at cfg2 / cfg1.
usage: python tools/fwd2d_check.py [check|time|all]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pea_b200 as P  # noqa: E402
from pea_b200 import _ops as ops  # noqa: E402
from pea_b200._lib import lib  # noqa: E402
from oracle import pea_oracle as O  # noqa: E402

KNOB_NO_CANON, KNOB_MAX_CTAS = 0, 1


def knob(k, v):
    assert lib.pea_debug_configure(k, int(v)) == 0


def worst(ours, ref):
    ours, ref = ours.double().cpu(), ref.double().cpu()
    scale = max(float(ref.abs().max()), 1e-30)
    return float(((ours - ref).abs() / (1e-5 * ref.abs() + 1e-6 * scale)).max())


def check():
    bad = 0
    for shifts in ([1, 3, 5, 9, 11], [1, 3, 5, 9, 27], [2], [1, 2, 4, 8, 12, 28], [3, 7, 20]):
        offsets = O.multi_offset(shifts, neighbor=4)
        offs3 = ops.offsets3(offsets, 2)
        for (B, H, W) in ((2, 130, 160), (1, 75, 100), (1, 64, 64), (3, 33, 72)):
            g = torch.Generator().manual_seed(B * 1000 + H)
            emb = torch.randn((B, 16, H, W), generator=g)
            emb[0, :, 3, 5] = 0.0
            emb_d = emb.cuda()
            for border in (ops.BORDER_WRAP, ops.BORDER_ZERO):
                for relu in (False, True):
                    outs = {}
                    for name in ("new", "new_cap2", "old", "generic"):
                        knob(KNOB_NO_CANON, 1 if name == "old" else 0)
                        knob(KNOB_MAX_CTAS, 2 if name == "new_cap2" else 0)
                        impl = ops.IMPL_GENERIC if name == "generic" else ops.IMPL_AUTO
                        outs[name] = ops.affs_fwd(emb_d, None, offs3, border, impl=impl, relu=relu).cpu()
                    knob(KNOB_NO_CANON, 0)
                    knob(KNOB_MAX_CTAS, 0)
                    r = [worst(outs["new"], outs["generic"]), worst(outs["new"], outs["old"])]
                    same = torch.equal(outs["new"], outs["new_cap2"])
                    ok = all(x <= 1.0 for x in r) and same
                    if border == ops.BORDER_WRAP and not relu:
                        ao = O.embedding2affs(emb, offsets)
                        r.append(worst(outs["new"], ao))
                        ok = ok and r[-1] <= 1.0
                    bad += 0 if ok else 1
                    if not ok:
                        print(f"BAD shifts={shifts} B{B} {H}x{W} border={border} relu={relu}: worst/tol {r} capped-identical {same}", flush=True)
    print("fwd2d_check:", "FAILED %d" % bad if bad else "all ok", flush=True)
    return bad


def timeit():
    import bench
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name in ("cfg2", "cfg1"):
        w = bench.WORKLOADS[name]
        g = torch.Generator().manual_seed(0)
        emb = torch.randn((w["B"], 16) + tuple(w["spatial"]), generator=g).cuda()
        offsets = P.multi_offset(w["shifts"], 4)
        K = len(offsets)
        fb = (64 + 4 * K) * w["B"] * int(np.prod(w["spatial"]))
        for kname, nc in (("new", 0), ("old", 1)):
            knob(KNOB_NO_CANON, nc)
            with torch.no_grad():
                ms = bench.time_graph(lambda: P.embedding2affs(emb, offsets), flush)
                ms_r = bench.time_graph(lambda: P.embedding2affs(emb, offsets, relu=True), flush)
            print(f"{name} embedding2affs {kname}: {ms * 1e3:.1f} us = {fb / ms / 1e6 / 6459.6 * 100:.1f}% of the HBM roofline; "
                  f"relu=True {ms_r * 1e3:.1f} us", flush=True)
        knob(KNOB_NO_CANON, 0)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    rc = check() if what in ("check", "all") else 0
    if what in ("time", "all"):
        timeit()
    sys.exit(1 if rc else 0)
