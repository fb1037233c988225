#!/usr/bin/env python
"""Benchmark of the Pixel-Embedded-Affinity hot path on B200 (see BASELINE.json / SURVEY.md 8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg1] [--impl ours|reference]

A "step" is one fused forward+backward of embedding -> multi-offset affinities -> weighted-MSE loss -> gradient
w.r.t. the embedding on one synthetic batch.  Metric: affinity-voxels/s (B*K*prod(spatial) per step), whole job.
Prints ONE JSON line on rank 0.  N > 1: launched under torchrun, one rank per GPU, every rank processes its own
batch (data-parallel, weak scaling, no data-path collective -- SURVEY.md 8e); time = max over ranks.

--impl reference times the reference's CPU implementation of the same path (the oracle port: the reference is
Python/torch-CPU and cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "affinity_voxels_per_s_fwd_bwd"
UNIT = "aff-voxels/s"
C = 16

WORKLOADS = {
    # name: (ndim, B, spatial, shifts / offsets descr)
    "cfg1": dict(nd=2, B=1, spatial=(544, 512), shifts=[1, 3, 5, 9, 27],
                 descr="CVPPP-shaped 2D: B=1, 16ch 544x512, multi_offset([1,3,5,9,27],4), K=10"),
    "cfg2": dict(nd=2, B=8, spatial=(520, 696), shifts=[1, 3, 5, 9, 11],
                 descr="BBBC039V1-shaped 2D batch: B=8, 16ch 520x696, multi_offset([1,3,5,9,11],4), K=10"),
    "cfg3": dict(nd=3, B=4, spatial=(18, 160, 160), shifts=None,
                 descr="AC3/AC4-shaped 3D patch: B=4, 16ch 18x160x160, 12 anisotropic offsets (norm5), K=12"),
    "cfg3x": dict(nd=3, B=4, spatial=(18, 160, 160), shifts=None, cross=True,
                  descr="cfg3 shape, two-view loss (ema_embedding_loss_norm5, second view detached), K=12"),
    "cfg2x": dict(nd=2, B=8, spatial=(520, 696), shifts=[1, 3, 5, 9, 11], cross=True,
                  descr="cfg2 shape, two-view loss (ema_embedding_loss, second view detached), K=10"),
}


def algorithmic_bytes_per_step(w):
    """read E 4C + target 4K + weight 4K (+ mask K, 2D) + write grad 4C + affs 4K, per spatial voxel."""
    K = 12 if w["nd"] == 3 else 2 * len(w["shifts"])
    per_voxel = 4 * C + 4 * K + 4 * K + (K if w["nd"] == 2 else 0) + 4 * C + 4 * K + (4 * C if w.get("cross") else 0)
    return per_voxel * w["B"] * int(np.prod(w["spatial"])), K


# ------------------------------------------------------------------------------------------ synthetic inputs
def make_labels(w, seed):
    from oracle import pea_oracle as O
    out = []
    for b in range(w["B"]):
        if w["nd"] == 2:
            out.append(O.voronoi_labels(w["spatial"], 20, seed * 100 + b, background_frac=0.3, dtype=np.uint8))
        else:
            out.append(O.voronoi_labels(w["spatial"], 60, seed * 100 + b, background_frac=0.1, z_scale=8.0, dtype=np.uint16))
    return np.stack(out)


def config_dict(w, name):
    """The SAME dict in both arms (the driver compares them)."""
    return {"workload": w["descr"], "name": name, "batch": w["B"],
            "l2": "flushed between timed steps (256 MiB memset, untimed); step working set also exceeds L2",
            "timing": "ours: CUDA events around a CUDA-graph replay of the public-API fwd+bwd, max over ranks; "
                      "reference arm: wall clock around the reference's CPU fwd+bwd on the host cores"}


def cpu_impl():
    """(namespace with the reference's loss functions and criterion, kind).  The UNMODIFIED reference files staged under
    baseline/_ref by build() when they travelled with the snapshot (kind "reference"), else the oracle port."""
    from oracle import ref_loader
    if ref_loader.available():
        try:
            import types
            ref = ref_loader.load()
            ns = types.SimpleNamespace(embedding_loss=ref.loss2d.embedding_loss, embedding_loss_norm5=ref.loss3d.embedding_loss_norm5,
                                       WeightedMSE=ref.crit.WeightedMSE, gen_affs_ours=ref.aff2d.gen_affs_ours,
                                       seg_to_aff=ref.aff3d.seg_to_aff, weight_binary_ratio=ref.seg.weight_binary_ratio)
            return ns, "reference"
        except Exception as ex:  # noqa: BLE001
            sys.stderr.write(f"reference files unusable ({type(ex).__name__}: {ex}); timing the oracle port\n")
    from oracle import pea_oracle as O
    return O, "port"


def host_inputs_oracle(w, seed, B=None):
    """Everything on the host through the oracle (reference arm / cpu_baseline)."""
    from oracle import pea_oracle as O
    B = B or w["B"]
    lab = make_labels(dict(w, B=B), seed)
    g = torch.Generator().manual_seed(seed)
    emb = torch.randn((B, C) + tuple(w["spatial"]), generator=g)
    if w["nd"] == 2:
        offsets = O.multi_offset(w["shifts"], 4)
        pairs = [O.gen_affs_ours(l, offsets, ignore=False, padding=True) for l in lab]
        tgt = np.stack([p[0] for p in pairs])
        msk = np.stack([p[1] for p in pairs])
    else:
        offsets = None
        tgt = np.stack([O.seg_to_aff_norm5(l) for l in lab])
        msk = None
    wgt = np.stack([np.stack([O.weight_binary_ratio(c) for c in t]) for t in tgt])
    return emb, torch.tensor(tgt), torch.tensor(wgt), (torch.tensor(msk) if msk is not None else None), offsets


class host_only:
    """While the reference's CPU path is being timed, `Tensor.cuda()` is a no-op: its WeightedMSE moves the normaliser
    with `.cuda()` (scripts_cvppp/loss/loss.py:116), which on a GPU box would put the loss scalar on the device."""

    def __enter__(self):
        self.saved = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t

    def __exit__(self, *exc):
        torch.Tensor.cuda = self.saved


def oracle_step(w, emb, tgt, wgt, msk, offsets, impl=None, full=False):
    """One fwd+bwd of the reference's CPU implementation (reference files or the port, see cpu_impl)."""
    O = impl or cpu_impl()[0]
    e = emb.detach().requires_grad_(True)
    with host_only():
        if w["nd"] == 2:
            loss, affs, _ = O.embedding_loss(e, tgt, wgt, msk, O.WeightedMSE(), offsets)
        else:
            loss, affs = O.embedding_loss_norm5(e, tgt, wgt, O.WeightedMSE())
        loss.backward()
    if full:
        return float(loss.detach()), affs.detach(), e.grad
    return float(loss.detach())


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank):
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    impl, kind = cpu_impl()
    alg_bytes, K = algorithmic_bytes_per_step(w)
    # bounded sample: shrink the batch until K steps fit in the budget
    B = w["B"]
    emb, tgt, wgt, msk, offsets = host_inputs_oracle(w, 0, B)
    t0 = time.perf_counter()
    oracle_step(w, emb, tgt, wgt, msk, offsets, impl)
    t1 = time.perf_counter() - t0
    budget = 240.0
    while B > 1 and t1 * (B / w["B"]) * (args.steps + args.warmup) > budget:
        B //= 2
    if B != w["B"]:
        emb, tgt, wgt, msk, offsets = host_inputs_oracle(w, 0, B)
    for _ in range(max(args.warmup - 1, 0)):
        oracle_step(w, emb, tgt, wgt, msk, offsets, impl)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(w, emb, tgt, wgt, msk, offsets, impl)
    dt = time.perf_counter() - t0
    vox = B * K * int(np.prod(w["spatial"]))
    value = vox * args.steps / dt
    what = "the reference's own files (baseline/_ref, unmodified)" if kind == "reference" else "the oracle port"
    sample = f"{args.steps} fwd+bwd calls of {what} (torch CPU, {cores} threads) at B={B} of the workload's B={w['B']}"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(w, args.workload),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------ our arm
def device_inputs(P, w, seed, device):
    lab = make_labels(w, seed)
    g = torch.Generator(device="cpu").manual_seed(seed)
    emb = torch.randn((w["B"], C) + tuple(w["spatial"]), generator=g).to(device)
    if w["nd"] == 2:
        offsets = P.multi_offset(w["shifts"], 4)
        tgt, msk, wgt = P.make_targets_2d(torch.tensor(lab).to(device), offsets)
    else:
        offsets = None
        tgt, wgt = P.make_targets_3d(torch.tensor(lab.astype(np.int32)).to(device))
        msk = None
    return lab, emb, tgt, wgt, msk, offsets


def our_step(P, w, e, tgt, wgt, msk, offsets, crit, partner=None):
    if w.get("cross"):
        if w["nd"] == 2:
            loss, _ = P.ema_embedding_loss(e, partner, tgt, wgt, msk, crit, offsets)
        else:
            loss, _ = P.ema_embedding_loss_norm5(e, partner, tgt, wgt, crit)
    elif w["nd"] == 2:
        loss, _, _ = P.embedding_loss(e, tgt, wgt, msk, crit, offsets)
    else:
        loss, _ = P.embedding_loss_norm5(e, tgt, wgt, crit)
    loss.backward()
    return loss


# The dominant kernel's OTHER ceilings, as fractions of the HBM roofline ("three ceilings": per-voxel
# need from the SASS counts of the shipped kernel, chip rate = 148 SMs x 1.965 GHz): a gather-form backward with 20 pair
# evaluations per voxel cannot reach the HBM roofline on this chip whatever its memory behaviour is.
ON_CHIP_CEILINGS = {
    "cfg3": {"shared_memory_bandwidth": 0.86, "fp32_issue_port": 0.84, "instruction_issue": 0.96,
             "how": "1815 B of LDS + TMA fill, 58 FP32-port clocks and 51 issue slots per voxel against 37.2 TB/s and "
                    "1163 G scheduler-clocks/s; HBM roofline = 23.75 G voxels/s"},
    "cfg2": {"shared_memory_bandwidth": 0.83, "how": "1792 B of LDS + TMA fill per 258 B pixel"},
}

LAUNCHES_PER_STEP = 3  # fused3d_norm5_tm, finalize_loss, scale_inplace: fallback when the profiler cannot count them


def ncu_traffic(workload):
    """DRAM bytes of one launch of the workload's dominant kernel, from the committed ncu capture (profiles/traffic.json
    names the digest it came from).  A profiler number cannot be taken inside a timed run; None if there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f)[workload]
        return float(e["dram_bytes_read"] + e["dram_bytes_write"]), e["source"], e["kernel"]
    except Exception:  # noqa: BLE001
        return None, None, None


def time_workload(P, w, steps, warmup, device, flush, dist_on):
    """Device-resident throughput: CUDA graph of one public-API fwd+bwd, replayed; L2 flushed between steps."""
    lab, emb, tgt, wgt, msk, offsets = device_inputs(P, w, 0, device)
    crit = P.WeightedMSE()
    e = emb.clone().requires_grad_(True)
    partner = torch.randn_like(emb) if w.get("cross") else None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            our_step(P, w, e, tgt, wgt, msk, offsets, crit, partner)
            e.grad = None
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss = our_step(P, w, e, tgt, wgt, msk, offsets, crit, partner)
    for _ in range(max(warmup, 3)):
        flush.zero_()
        graph.replay()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
        torch.cuda.synchronize()
    for i in range(steps):
        flush.zero_()            # evict the previous step's lines from the 126 MB L2 (untimed)
        starts[i].record()
        graph.replay()
        ends[i].record()
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
        torch.cuda.synchronize()
    per = [s.elapsed_time(t) for s, t in zip(starts, ends)]  # ms
    return dict(total_ms=sum(per), per_ms=per, loss=float(loss.detach()), inputs=(lab, emb, tgt, wgt, msk, offsets))


def time_e2e(P, w, steps, warmup, device, inputs):
    """Public API with HOST buffers: every step copies embedding/target/weight(/mask) from pinned host memory,
    runs fwd+bwd and reads the loss back."""
    lab, emb, tgt, wgt, msk, offsets = inputs
    host = [t.cpu().pin_memory() if t is not None else None for t in (emb, tgt, wgt, msk)]
    dev = [torch.empty_like(t, device=device) if t is not None else None for t in host]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    crit = P.WeightedMSE()
    h2d = sum(t.numel() * t.element_size() for t in host if t is not None)

    def step():
        for h, d in zip(host, dev):
            if h is not None:
                d.copy_(h, non_blocking=True)
        e = dev[0].requires_grad_(True)
        loss = our_step(P, w, e, dev[1], dev[2], dev[3], offsets, crit)
        loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.synchronize()
        dev[0].requires_grad_(False)
        e.grad = None
        return float(loss_host)

    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, h2d, 4


def time_e2e_label_api(P, w, steps, device, inputs):
    """Host buffers through the label-driven API: every step copies the embedding AND the label
    map from pinned host memory, runs embedding_loss_labels / embedding_loss_norm5_labels fwd+bwd and reads the loss
    back -- what a training loop sends when the GT recipe lives in the loss kernel (target / weight / mask never cross
    PCIe or HBM)."""
    lab, emb, _, _, _, offsets = inputs
    lab_t = torch.tensor(lab.astype(np.int32))
    host = [emb.cpu().pin_memory(), lab_t.pin_memory()]
    dev = [torch.empty_like(t, device=device) for t in host]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    crit = P.WeightedMSE()
    h2d = sum(t.numel() * t.element_size() for t in host)

    def step():
        for h, d in zip(host, dev):
            d.copy_(h, non_blocking=True)
        e = dev[0].requires_grad_(True)
        if w["nd"] == 2:
            loss, _, _ = P.embedding_loss_labels(e, dev[1], crit, offsets)
        else:
            loss, _ = P.embedding_loss_norm5_labels(e, dev[1], crit)
        loss.backward()
        loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.synchronize()
        dev[0].requires_grad_(False)
        e.grad = None

    for _ in range(3):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) * 1e3, h2d


def time_e2e_labels_only(P, w, steps, device, inputs):
    """B200-native pipeline: only the label map crosses PCIe; target/mask/weight are generated on the device."""
    lab, emb, _, _, _, offsets = inputs
    lab_t = torch.tensor(lab if w["nd"] == 2 else lab.astype(np.int32))
    host = lab_t.pin_memory()
    dev_lab = torch.empty_like(host, device=device)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    crit = P.WeightedMSE()
    e = emb.clone().requires_grad_(True)

    def step():
        dev_lab.copy_(host, non_blocking=True)
        if w["nd"] == 2:
            tgt, msk, wgt = P.make_targets_2d(dev_lab, offsets)
        else:
            tgt, wgt = P.make_targets_3d(dev_lab)
            msk = None
        loss = our_step(P, w, e, tgt, wgt, msk, offsets, crit)
        loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.synchronize()
        e.grad = None

    for _ in range(3):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) * 1e3, host.numel() * host.element_size()


def time_zshard(P, device, rank, world, steps, shape=(100, 1024, 1024)):
    """cfg4: forward-only 12-channel affinities of a full AC3/AC4 volume, z-sharded over `world` GPUs (112 B/voxel
    algorithmic).  Strong scaling: the volume is fixed.  Halo transport, best available first: "peer" = slabs in
    symmetric memory, the kernel reads the previous rank's last 4 slices in place over NVLink (one launch, no exchange
    step; a device-side barrier orders producer writes before the reads); "nccl" = send/recv in place, overlapped."""
    from pea_b200.zshard import ZShard, PeerZShard, sharded_affinities, fused_halo_affinities, overlapped_halo_affinities
    D, H, W = shape
    dist_on = world > 1
    transport, sh = "none", None
    if dist_on:
        ok = torch.zeros(1, device=device)
        try:
            sh = PeerZShard(1, C, D, H, W, rank, world, device)
            transport, ok[0] = "peer", 1.0
        except Exception as ex:  # noqa: BLE001  (symmetric memory not available on this box)
            sys.stderr.write(f"rank {rank}: peer halo unavailable ({type(ex).__name__}: {ex}); using NCCL\n")
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        if float(ok) < 1.0:
            transport, sh = "nccl", None
    if sh is None:
        sh = ZShard(1, C, D, H, W, rank, world, device)
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    sh.own.normal_(generator=g)
    out = torch.empty((1, 12, sh.depth, H, W), device=device)

    def step():
        if transport == "peer":
            # publish() -> kernel (halo read from the peer in place) -> release(): the contract of a slab that is rewritten
            sh.affinities(out=out, overlapped=os.environ.get("PEA_ZSHARD_OVERLAP", "0") == "1")
        else:
            sharded_affinities(sh, out=out)

    halo_check = zshard_halo_check(P, device, rank, world, transport) if dist_on else "single GPU"

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
        torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        step()
    t.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(t) / steps], device=device, dtype=torch.float64)
    if dist_on:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms = float(ms)
    vox = int(np.prod(shape))
    res = {"workload": f"cfg4 AC3/AC4 full volume 100x1024x1024 forward-only, z-sharded over {world} GPU(s)",
           "ms_per_step": ms, "value": 12 * vox / (ms * 1e-3), "unit": "aff-voxels/s (forward only)",
           "scaling": "strong", "hbm_gbs_aggregate": 112 * vox / (ms * 1e-3) / 1e9,
           "halo_transport": transport, "halo_check": halo_check,
           "halo_bytes_per_boundary": HALO_BYTES(shape) if world > 1 else 0,
           "slab_depths": [b - a for a, b in __import__("pea_b200").zshard.shard_ranges(D, world)]}
    del sh, out
    torch.cuda.empty_cache()
    return res


def zshard_halo_check(P, device, rank, world, transport, shape=(32, 256, 256), iters=3):
    """The REAL multi-GPU path against a single-GPU run, bit for bit, before anything is timed: every rank builds the same
    seeded reduced volume, computes the whole volume locally, rewrites its slab every iteration (so a missing
    release() barrier or a stale halo shows up) and compares its slab of the sharded result."""
    from pea_b200.zshard import ZShard, PeerZShard, sharded_affinities
    D, H, W = shape
    try:
        sh = PeerZShard(1, C, D, H, W, rank, world, device) if transport == "peer" else ZShard(1, C, D, H, W, rank, world, device)
        ok = torch.ones(1, device=device)
        for it in range(iters):
            g = torch.Generator(device="cpu").manual_seed(4321 + it)
            vol = torch.randn((1, C, D, H, W), generator=g).to(device)
            whole = P.inf_embedding_loss_norm5(vol)
            sh.own.copy_(vol[:, :, sh.z0:sh.z1])
            out = sh.affinities() if transport == "peer" else sharded_affinities(sh)
            if not torch.equal(out, whole[:, :, sh.z0:sh.z1]):
                ok[0] = 0.0
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        del sh
        torch.cuda.empty_cache()
        return "ok" if float(ok) == 1.0 else "FAILED: a slab differs from the single-GPU result"
    except Exception as ex:  # noqa: BLE001
        return f"not run ({type(ex).__name__}: {ex})"


def time_step6(P, device, flush, iters=20):
    """The six loss calls of one AC3/AC4 training step (scripts_ac3ac4/main.py:218-231) -- norm5 + two-view norm5 at
    B x 18 x 160 x 160, norm1 at the four deep-supervision scales -- as six separate public-API calls vs the multi-scale
    launch for the four small ones; parity of the summed loss against the CPU reference/port."""
    B, D = 2, 18   # ac3ac4.yaml batch_size
    g = torch.Generator(device="cpu").manual_seed(6)
    emb = torch.randn((B, C, D, 160, 160), generator=g).to(device).requires_grad_(True)
    ema = torch.randn((B, C, D, 160, 160), generator=g).to(device)

    def tw(K, h):
        t = (torch.rand((B, K, D, h, h), generator=g) > 0.5).float().to(device)
        w = (torch.rand((B, K, D, h, h), generator=g) * 2 + 0.1).to(device)
        return t, w
    tgt, wgt = tw(12, 160)
    scales = [10, 20, 40, 80]
    emds = [torch.randn((B, C, D, h, h), generator=g).to(device).requires_grad_(True) for h in scales]
    downs = [tw(3, h) for h in scales]
    crit = P.WeightedMSE()

    def step(multi):
        if multi == "fused":   # one gradient buffer for the two full-resolution losses + the multi-scale launch
            l01, _, _, _ = P.embedding_loss_norm5_self_and_cross(emb, ema, tgt, wgt, crit, affs0_weight=1)
            ls = [r[0] for r in P.embedding_loss_norm1_multi(emds, [d[0] for d in downs], [d[1] for d in downs], crit)]
            loss = l01 + ls[0] + ls[1] + ls[2] + ls[3]
            loss.backward()
            emb.grad = None
            for e in emds:
                e.grad = None
            return loss
        l0, _ = P.embedding_loss_norm5(emb, tgt, wgt, crit, affs0_weight=1)
        l1, _ = P.ema_embedding_loss_norm5(emb, ema, tgt, wgt, crit, affs0_weight=1)
        if multi:
            ls = [r[0] for r in P.embedding_loss_norm1_multi(emds, [d[0] for d in downs], [d[1] for d in downs], crit)]
        else:
            ls = [P.embedding_loss_norm1(e, t, w, crit)[0] for e, (t, w) in zip(emds, downs)]
        loss = l0 + l1 + ls[0] + ls[1] + ls[2] + ls[3]
        loss.backward()
        emb.grad = None
        for e in emds:
            e.grad = None
        return loss

    out = {}
    for name, multi in (("six_calls", False), ("multi", True), ("fused", "fused")):
        n_ours, _, n_all = count_launches(lambda: step(multi))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step(multi)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            step(multi)
        torch.cuda.synchronize()
        out[name + "_eager_ms"] = (time.perf_counter() - t0) / iters * 1e3
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = step(multi)
        per = []
        for _ in range(iters):
            flush.zero_()
            s_, t_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            graph.replay()
            t_.record()
            torch.cuda.synchronize()
            per.append(s_.elapsed_time(t_))
        out[name + "_graph_ms"] = statistics.median(per)
        out[name + "_loss"] = float(loss.detach())
        out[name + "_launches"] = {"pea_kernels": n_ours, "all_kernels": n_all}
        del graph
    # parity of the summed loss against the CPU implementation's six calls
    try:
        impl, kind = cpu_impl()
        from oracle import pea_oracle as O
        with host_only():
            e0 = emb.detach().cpu()
            lo = impl.embedding_loss_norm5(e0, tgt.cpu(), wgt.cpu(), impl.WeightedMSE())[0]
            n5x = O.ema_embedding_loss_norm5 if kind == "port" else __import__("oracle.ref_loader", fromlist=["x"]).load().loss3d.ema_embedding_loss_norm5
            n1 = O.embedding_loss_norm1 if kind == "port" else __import__("oracle.ref_loader", fromlist=["x"]).load().loss3d.embedding_loss_norm1
            lo = lo + n5x(e0, ema.cpu(), tgt.cpu(), wgt.cpu(), impl.WeightedMSE())[0]
            for e, (t, w) in zip(emds, downs):
                lo = lo + n1(e.detach().cpu(), t.cpu(), w.cpu(), impl.WeightedMSE())[0]
        out["loss_cpu_" + kind] = float(lo)
        out["parity_rel"] = max(abs(out[k] - float(lo)) / abs(float(lo)) for k in ("multi_loss", "fused_loss"))
    except Exception as ex:  # noqa: BLE001
        out["parity_rel"] = f"not run ({type(ex).__name__}: {ex})"
    out["workload"] = f"six loss calls of one AC3/AC4 training step, B={B}: norm5 + two-view norm5 at 18x160x160, norm1 at 18x{{10,20,40,80}}^2"
    return out


def time_cfg5(P, device, rank, world, dim=3, steps=5):
    """BASELINE config 5: one data-parallel training step with the REFERENCE's backbone, unmodified --
    dim 3: UNet_PNI_embedding_deep (scripts_ac3ac4/model/model_superhuman.py:443-492, built as main.py:122-133 does with
    ac3ac4.yaml's filters) on (B,1,18,160,160); dim 2: ResidualUNet2D_deep (scripts_cvppp/model/unet2d_residual.py:324-353,
    main.py:111-115) on (B,3,544,544) -- wrapped in DistributedDataParallel (NCCL gradient all-reduce over NVLink), the
    six PEA loss calls of the training loop (main.py:223-230 / :284-293) and the optimizer step.  Reported: step time with
    this package's loss functions and with the reference's own loss functions running as eager torch ops on the GPU
    (context), and the loss functions' share of the step."""
    from oracle import ref_loader
    if not ref_loader.available():
        return {"unavailable": "reference model files not staged (baseline/_ref)"}
    ref = ref_loader.load(models=True)
    B = 2   # ac3ac4.yaml / cvppp.yaml batch_size
    torch.manual_seed(1234)
    g = torch.Generator(device="cpu").manual_seed(77 + rank)
    crit_p, crit_r = P.WeightedMSE(), ref.crit.WeightedMSE()

    def rnd01(*shape):
        return (torch.rand(shape, generator=g) > 0.5).float().to(device)

    def rndw(*shape):
        return (torch.rand(shape, generator=g) * 2 + 0.1).to(device)

    if dim == 3:
        D = 18
        model = ref.unet3d.UNet_PNI_embedding_deep(in_planes=1, out_planes=12, filters=[28, 36, 48, 64, 80], upsample_mode='bilinear',
                                                    decode_ratio=1, merge_mode='add', pad_mode='zero', bn_mode='async',
                                                    relu_mode='elu', init_mode='kaiming_normal', emd=16).to(device)
        name, inp = "UNet_PNI_embedding_deep", (B, 1, D, 160, 160)
        tgt, wgt = rnd01(B, 12, D, 160, 160), rndw(B, 12, D, 160, 160)
        downs = [(rnd01(B, 3, D, h, h), rndw(B, 3, D, h, h)) for h in (10, 20, 40, 80)]
        L = ref.loss3d

        def losses(ours, outs, ema):
            emds, emb = list(outs[:4]), outs[4]
            if ours:
                l0, _ = P.embedding_loss_norm5(emb, tgt, wgt, crit_p)
                l1, _ = P.ema_embedding_loss_norm5(emb, ema, tgt, wgt, crit_p)
                ls = [r[0] for r in P.embedding_loss_norm1_multi(emds, [d[0] for d in downs], [d[1] for d in downs], crit_p)]
            else:
                l0, _ = L.embedding_loss_norm5(emb, tgt, wgt, crit_r)
                l1, _ = L.ema_embedding_loss_norm5(emb, ema, tgt, wgt, crit_r)
                ls = [L.embedding_loss_norm1(e, t, w, crit_r)[0] for e, (t, w) in zip(emds, downs)]
            return l0 + l1 + ls[0] + ls[1] + ls[2] + ls[3]
        opt_of = lambda m: torch.optim.Adam(m.parameters(), lr=1e-4, betas=(0.9, 0.999), eps=0.01, weight_decay=1e-6, amsgrad=True)
    else:
        model = ref.unet2d.ResidualUNet2D_deep(in_channels=3, out_channels=2, nfeatures=[16, 32, 64, 128, 256], emd=16,
                                               if_sigmoid=False).to(device)
        name, inp = "ResidualUNet2D_deep", (B, 3, 544, 544)
        offsets = P.multi_offset([1, 3, 5, 9, 27], 4)
        tgt, wgt, msk = rnd01(B, 10, 544, 544), rndw(B, 10, 544, 544), rnd01(B, 10, 544, 544)
        # outputs come coarsest first (34^2 .. 272^2): emd4, emd3, emd2, emd1 = the first 2, 4, 6, 8 offsets (main.py:266, :284-287)
        ks, hs = (2, 4, 6, 8), (34, 68, 136, 272)
        downs = [(rnd01(B, k, h, h), rndw(B, k, h, h), rnd01(B, k, h, h)) for k, h in zip(ks, hs)]
        L = ref.loss2d

        def losses(ours, outs, ema):
            emds, emb = list(outs[:4]), outs[4]
            if ours:
                l0, _, _ = P.embedding_loss(emb, tgt, wgt, msk, crit_p, offsets)
                l1, _ = P.ema_embedding_loss(emb, ema, tgt, wgt, msk, crit_p, offsets)
                ls = [r[0] for r in P.embedding_loss_multi(emds, [d[0] for d in downs], [d[1] for d in downs], [d[2] for d in downs],
                                                           crit_p, [offsets[:k] for k in ks])]
            else:
                l0, _, _ = L.embedding_loss(emb, tgt, wgt, msk, crit_r, offsets)
                l1, _ = L.ema_embedding_loss(emb, ema, tgt, wgt, msk, crit_r, offsets)
                ls = [L.embedding_loss(e, d[0], d[1], d[2], crit_r, offsets[:k])[0] for e, d, k in zip(emds, downs, ks)]
            return l0 + l1 + ls[0] + ls[1] + ls[2] + ls[3] + torch.nn.functional.mse_loss(emb, ema)   # + criterion_ct (main.py:295)
        opt_of = lambda m: torch.optim.Adam(m.parameters(), lr=1e-4, betas=(0.9, 0.999), eps=1e-8, amsgrad=False)
    nparams = sum(p.numel() for p in model.parameters())
    if world > 1:
        # broadcast_buffers=False: the step runs the model twice (second view), and a buffer broadcast in the second
        # forward would modify BatchNorm buffers the first forward's graph still needs
        # find_unused_parameters for the 2-D net: its mask head (binary_seg) takes no part in the loss the training loop
        # builds (scripts_cvppp/main.py:295-308), which the reference's nn.DataParallel does not mind but DDP does
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index], broadcast_buffers=False,
                                                          find_unused_parameters=(dim == 2))
    opt = opt_of(model)
    x, x2 = torch.rand(inp, generator=g).to(device), torch.rand(inp, generator=g).to(device)

    def train_step(ours):
        opt.zero_grad(set_to_none=True)
        with torch.no_grad():          # second view first: its forward must not touch state the graph below depends on
            ema = model(x2)[4].detach().clone()
        outs = model(x)
        loss = losses(ours, outs, ema)
        loss.backward()
        opt.step()
        return loss

    out = {"model": name + " (reference file, unmodified)", "params": nparams, "per_gpu_batch": B,
           "input": f"{inp} per GPU", "collective": "DDP NCCL all-reduce of the backbone gradients" if world > 1 else "none (1 GPU)"}
    for tag, ours in (("ours", True), ("reference_loss_on_gpu", False)):
        for _ in range(2):
            train_step(ours)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        s_, t_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        for _ in range(steps):
            loss = train_step(ours)
        t_.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s_.elapsed_time(t_) / steps], device=device, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        out[tag + "_step_ms"] = float(ms)
        out[tag + "_loss"] = float(loss.detach())
    # the loss functions alone (fwd + bwd to the embeddings), same tensors, CUDA events
    with torch.no_grad():
        outs = model(x)
        ema = model(x2)[4].detach().clone()
    leaf = [o.detach().requires_grad_(True) for o in outs[:5]]
    for tag, ours in (("ours", True), ("reference_loss_on_gpu", False)):
        for _ in range(2):
            losses(ours, leaf, ema).backward()
        torch.cuda.synchronize()
        s_, t_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        for _ in range(steps):
            losses(ours, leaf, ema).backward()
        t_.record()
        torch.cuda.synchronize()
        out[tag + "_loss_part_ms"] = s_.elapsed_time(t_) / steps
    out["samples_per_s"] = world * B / (out["ours_step_ms"] * 1e-3)
    out["loss_share_of_step"] = out["ours_loss_part_ms"] / out["ours_step_ms"]
    out["loss_share_of_step_with_reference_loss"] = out["reference_loss_on_gpu_loss_part_ms"] / out["reference_loss_on_gpu_step_ms"]
    del model, opt
    torch.cuda.empty_cache()
    return out


def HALO_BYTES(shape):
    return C * 4 * shape[1] * shape[2] * 4


def time_gt(P, w, device, iters=20):
    """Device-side ground truth for one batch of the workload: labels -> target (+mask) + weights (K6-K8)."""
    lab = make_labels(w, 0)
    K = 12 if w["nd"] == 3 else 2 * len(w["shifts"])
    if w["nd"] == 2:
        t = torch.tensor(lab).to(device)
        offsets = P.multi_offset(w["shifts"], 4)
        fn = lambda: P.make_targets_2d(t, offsets)
        per_voxel = 2 * t.element_size() + 4 * K + K + 4 * K
    else:
        t = torch.tensor(lab.astype(np.int32)).to(device)
        fn = lambda: P.make_targets_3d(t)
        per_voxel = 2 * t.element_size() + 4 * K + 4 * K
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    vox = w["B"] * int(np.prod(w["spatial"]))
    return {"workload": "GT generation (labels -> target/mask/weight) for " + w["descr"], "ms_per_step": ms,
            "value": K * vox / (ms * 1e-3), "unit": "aff-voxels/s", "hbm_gbs": per_voxel * vox / (ms * 1e-3) / 1e9,
            "algorithmic_bytes": per_voxel * vox, "timing": "back-to-back calls, working set may sit in L2"}


def time_labels_pipeline(P, w, device, flush, iters=20):
    """2-D, labels resident on the device -> loss + gradient: this package's two-step route (GT kernels, then the fused
    loss) against the label-driven loss kernel (target / mask / weight never materialised).
    Each route is captured in a CUDA graph and replayed; L2 flushed between iterations (untimed)."""
    lab = torch.tensor(make_labels(w, 0)).to(device)
    lab32 = lab.to(torch.int32)
    offsets = P.multi_offset(w["shifts"], 4)
    g = torch.Generator(device="cpu").manual_seed(0)
    emb = torch.randn((w["B"], C) + tuple(w["spatial"]), generator=g).to(device)
    crit = P.WeightedMSE()
    e = emb.clone().requires_grad_(True)

    def two_step():
        tgt, msk, wgt = P.make_targets_2d(lab, offsets)
        loss, _, _ = P.embedding_loss(e, tgt, wgt, msk, crit, offsets)
        loss.backward()
        e.grad = None
        return loss

    def fused():
        loss, _, _ = P.embedding_loss_labels(e, lab32, crit, offsets)
        loss.backward()
        e.grad = None
        return loss

    out = {}
    for name, fn in (("two_step", two_step), ("fused", fused)):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = fn()
        for _ in range(3):
            flush.zero_()
            graph.replay()
        per = []
        for _ in range(iters):
            flush.zero_()
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            graph.replay()
            t.record()
            torch.cuda.synchronize()
            per.append(s.elapsed_time(t))
        out[name + "_ms"] = statistics.median(per)
        out[name + "_loss"] = float(loss.detach())
        del graph
    K = 2 * len(w["shifts"])
    vox = w["B"] * int(np.prod(w["spatial"]))
    per_voxel = 4 * C + 2 * 4 + 4 * C + 4 * K   # read E + labels (twice: count pass, loss) + write grad + affs
    out.update({"workload": "labels -> loss + gradient, " + w["descr"], "value": K * vox / (out["fused_ms"] * 1e-3),
                "unit": UNIT, "algorithmic_bytes": per_voxel * vox, "hbm_gbs": per_voxel * vox / (out["fused_ms"] * 1e-3) / 1e9,
                "timing": "CUDA-graph replay of each route, median of per-iteration CUDA events, L2 flushed"})
    return out


def time_inference_epilogue(P, device, shape=(100, 1024, 1024), iters=5):
    """cfg4 on one GPU with what the reference's inference loop does next (scripts_ac3ac4/inference.py:160-165: three
    border copies + F.relu): torch epilogue on the plain result vs relu / replicate_border folded into the call."""
    D, H, W = shape
    g = torch.Generator(device=device).manual_seed(5)
    emb = torch.empty((1, C, D, H, W), device=device).normal_(generator=g)

    def torch_epilogue():
        pred = P.inf_embedding_loss_norm5(emb)
        pred[:, 1, :, :1, :] = pred[:, 1, :, 1:2, :]
        pred[:, 2, :, :, :1] = pred[:, 2, :, :, 1:2]
        pred[:, 0, :1, :, :] = pred[:, 0, 1:2, :, :]
        return torch.relu(pred)

    def fused():
        return P.inf_embedding_loss_norm5(emb, relu=True, replicate_border=True)

    out = {}
    for name, fn in (("torch_epilogue", torch_epilogue), ("fused", fused)):
        for _ in range(2):
            r = fn()
        del r
        torch.cuda.synchronize()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            r = fn()
            del r
        t.record()
        torch.cuda.synchronize()
        out[name + "_ms"] = s.elapsed_time(t) / iters
    vox = int(np.prod(shape))
    out.update({"workload": "cfg4 forward + border replicate + relu (inference.py:160-165), 1 GPU",
                "value": 12 * vox / (out["fused_ms"] * 1e-3), "unit": "aff-voxels/s (forward only)"})
    del emb
    torch.cuda.empty_cache()
    return out


def time_graph(fn, flush, iters=20, do_flush=True):
    """Median ms of a CUDA-graph replay of fn (CUDA events around each replay; the L2 is flushed between replays unless
    do_flush is False: steady-state number, working set partly resident in the 126 MB L2)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    for _ in range(3):
        graph.replay()
    per = []
    for _ in range(iters):
        if do_flush:
            flush.zero_()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        graph.replay()
        t.record()
        torch.cuda.synchronize()
        per.append(s.elapsed_time(t))
    del graph
    return statistics.median(per)


def time_forward_and_warm(P, device, flush, peak):
    """forward-only affinities (embedding2affs / inf_embedding_loss_norm5: 4C + 4K bytes per voxel) at
    cfg1 - cfg3, and the fused fwd+bwd step of each with a WARM L2 (no flush between replays) next to the cold number the
    other legs report -- a warm number above the HBM roofline is an L2 artefact, not a win."""
    out = {}
    for name in ("cfg1", "cfg2", "cfg3"):
        w = WORKLOADS[name]
        lab, emb, tgt, wgt, msk, offsets = device_inputs(P, w, 0, device)
        K = 12 if w["nd"] == 3 else 2 * len(w["shifts"])
        vox = w["B"] * int(np.prod(w["spatial"]))
        with torch.no_grad():
            fwd = (lambda: P.embedding2affs(emb, offsets)) if w["nd"] == 2 else (lambda: P.inf_embedding_loss_norm5(emb))
            ms_f = time_graph(fwd, flush)
        e = emb.clone().requires_grad_(True)
        crit = P.WeightedMSE()

        def step():
            our_step(P, w, e, tgt, wgt, msk, offsets, crit)
            e.grad = None
        ms_w = time_graph(step, flush, do_flush=False)
        ab, _ = algorithmic_bytes_per_step(w)
        fb = (4 * C + 4 * K) * vox
        out[name] = {"forward_only_ms": ms_f, "forward_only_hbm_gbs": fb / (ms_f * 1e-3) / 1e9,
                     "forward_only_roofline_frac": fb / (ms_f * 1e-3) / 1e9 / peak,
                     "fused_step_warm_l2_ms": ms_w, "fused_step_warm_l2_hbm_gbs_equiv": ab / (ms_w * 1e-3) / 1e9,
                     "fused_step_warm_l2_frac_of_hbm_roofline": ab / (ms_w * 1e-3) / 1e9 / peak}
        del lab, emb, tgt, wgt, msk, e
    out["timing"] = "CUDA-graph replay, median of 20; forward-only with the L2 flushed, warm numbers without a flush"
    return out


def time_reference_on_gpu(P, w, device, flush, iters=10):
    """The reference's OWN loss function (the unmodified files staged under baseline/_ref) running as eager PyTorch ops on
    this GPU, forward + backward on the workload's batch -- the implementation a user
    of the reference has today on the same hardware.  L2 flushed between steps like the headline."""
    O, kind = cpu_impl()
    if kind != "reference":  # the oracle port is the CPU checker only: it is never run as a GPU arm
        return {"unavailable": "the reference's files are not staged under baseline/_ref"}
    lab, emb, tgt, wgt, msk, offsets = device_inputs(P, w, 0, device)
    e = emb.clone().requires_grad_(True)
    crit = O.WeightedMSE()

    def step():
        if w["nd"] == 2:
            loss, _, _ = O.embedding_loss(e, tgt, wgt, msk, crit, offsets)
        else:
            loss, _ = O.embedding_loss_norm5(e, tgt, wgt, crit)
        loss.backward()
        e.grad = None
        return loss
    for _ in range(3):
        loss = step()
    torch.cuda.synchronize()
    per = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = step()
        b.record()
        torch.cuda.synchronize()
        per.append(a.elapsed_time(b))
    n_all = None
    try:
        _, _, n_all = count_launches(step)
    except Exception:  # noqa: BLE001
        pass
    ms = statistics.median(per)
    K = 12 if w["nd"] == 3 else 2 * len(w["shifts"])
    vox = w["B"] * K * int(np.prod(w["spatial"]))
    return {"what": "the reference's loss function as eager torch ops on this GPU, fwd+bwd, same batch", "kind": kind,
            "ms_per_step": ms, "value": vox / (ms * 1e-3), "unit": UNIT, "kernel_launches_per_step": n_all,
            "loss": float(loss.detach())}


def cpu_baseline(w, seconds=12.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    impl, kind = cpu_impl()
    emb, tgt, wgt, msk, offsets = host_inputs_oracle(w, 0)
    oracle_step(w, emb, tgt, wgt, msk, offsets, impl)  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        oracle_step(w, emb, tgt, wgt, msk, offsets, impl)
        n += 1
        dt = time.perf_counter() - t0
        if dt > seconds or n >= 20:
            break
    _, K = algorithmic_bytes_per_step(w)
    vox = w["B"] * K * int(np.prod(w["spatial"]))
    what = "the reference's own files (baseline/_ref, unmodified)" if kind == "reference" else "the oracle port"
    return {"value": vox * n / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n} full-batch fwd+bwd calls of {what} (torch CPU, {cores} threads), {dt:.1f} s"}


def cpu_gt_baseline(w, seconds=6.0):
    """the reference's GT functions (a10-a12) are single-threaded NumPy / SciPy, called per sample inside
    DataLoader workers: time ONE sample on one core, and quote the x num_workers equivalent (cvppp.yaml:49 /
    bbbc039v1.yaml:49 use 2-8 workers)."""
    impl, kind = cpu_impl()
    lab = make_labels(dict(w, B=1), 0)[0]
    K = 12 if w["nd"] == 3 else 2 * len(w["shifts"])
    if w["nd"] == 2:
        from oracle import pea_oracle as O
        offsets = O.multi_offset(w["shifts"], 4)

        def one():
            t, m = impl.gen_affs_ours(lab, offsets, ignore=False, padding=True)
            return [impl.weight_binary_ratio(c) for c in t]
    else:
        from oracle import pea_oracle as O
        nhoods = [O.mknhood3d(1)] + [np.asarray([[-a, 0, 0], [0, -b, 0], [0, 0, -c]]) for a, b, c in ((2, 3, 3), (3, 9, 9), (4, 27, 27))]

        def one():
            t = np.concatenate([impl.seg_to_aff(lab, nh, pad='') for nh in nhoods], axis=0)
            return [impl.weight_binary_ratio(c) for c in t]
    one()
    n, t0 = 0, time.perf_counter()
    while True:
        one()
        n += 1
        dt = time.perf_counter() - t0
        if dt > seconds or n >= 50:
            break
    per = dt / n
    vox = K * int(np.prod(w["spatial"]))
    return {"ms_per_sample": per * 1e3, "value": vox / per, "unit": UNIT + " (GT only, one sample, one core)",
            "value_x8_workers": 8 * vox / per, "cores": 1, "kind": kind,
            "sample": f"{n} samples of gen_affs_ours/seg_to_aff + weight_binary_ratio x {K} channels, {dt:.1f} s"}


def parity_vs_cpu(P, w, device):
    """The GPU path against the reference's CPU path on the SAME seeded batch (the cpu_baseline leg's inputs): loss,
    affinities, gradient.  Tolerance of tests/: |d| <= 1e-5 |ref| + 1e-6 max|ref| element by element."""
    impl, kind = cpu_impl()
    emb, tgt, wgt, msk, offsets = host_inputs_oracle(w, 0)
    lo, ao, go = oracle_step(w, emb, tgt, wgt, msk, offsets, impl, full=True)
    e = emb.to(device).requires_grad_(True)
    crit = P.WeightedMSE()
    if w["nd"] == 2:
        loss, affs, _ = P.embedding_loss(e, tgt.to(device), wgt.to(device), msk.to(device), crit, offsets)
    else:
        loss, affs = P.embedding_loss_norm5(e, tgt.to(device), wgt.to(device), crit)
    loss.backward()
    out = {"against": kind, "loss_gpu": float(loss), "loss_cpu": lo, "loss_rel": abs(float(loss) - lo) / max(abs(lo), 1e-30)}
    worst = out["loss_rel"] / 1e-5
    for name, a, b in (("affs", affs.detach().cpu(), ao), ("grad", e.grad.cpu(), go)):
        scale = float(b.abs().max())
        d = (a.double() - b.double()).abs()
        out[name + "_max_abs_over_scale"] = float(d.max()) / max(scale, 1e-30)
        worst = max(worst, float((d / (1e-5 * b.abs().double() + 1e-6 * scale)).max()))
    out["parity_max_rel"] = max(out["loss_rel"], out["affs_max_abs_over_scale"], out["grad_max_abs_over_scale"])
    out["tolerance_used"] = worst   # <= 1 means inside rtol 1e-5 + atol 1e-6 * scale everywhere
    out["ok"] = bool(worst <= 1.0)
    return out


def count_launches(fn):
    """Kernels launched by one call of fn (torch.profiler / CUPTI), outside any timed region."""
    try:
        from torch.profiler import profile, ProfilerActivity
        fn()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        names = [ev.name for ev in prof.events() if getattr(ev, "device_type", None) is not None
                 and "cuda" in str(ev.device_type).lower() and not ev.name.lower().startswith("memcpy")
                 and not ev.name.lower().startswith("memset")]
        ours = [n for n in names if "pea::" in n]
        return len(ours), sorted(set(n.split("(")[0][-60:] for n in ours)), len(names)
    except Exception as ex:  # noqa: BLE001
        sys.stderr.write(f"launch count via profiler unavailable ({type(ex).__name__}: {ex})\n")
        return None, [], None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_ours(args, rank, world, local_rank):
    import pea_b200 as P
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist_on = world > 1
    if dist_on:
        torch.distributed.init_process_group("nccl", device_id=device)
    w = WORKLOADS[args.workload]
    alg_bytes, K = algorithmic_bytes_per_step(w)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = time_workload(P, w, args.steps, args.warmup, device, flush, dist_on)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = torch.tensor([res["total_ms"]], device=device, dtype=torch.float64)
    if dist_on:
        torch.distributed.all_reduce(total_ms, op=torch.distributed.ReduceOp.MAX)
    total_ms = float(total_ms)
    vox_step = w["B"] * K * int(np.prod(w["spatial"]))
    value = world * vox_step * args.steps / (total_ms * 1e-3)

    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms, h2d, d2h = time_e2e(P, w, e2e_steps, args.warmup, device, res["inputs"])
    e2e_t = torch.tensor([e2e_ms], device=device, dtype=torch.float64)
    if dist_on:
        torch.distributed.all_reduce(e2e_t, op=torch.distributed.ReduceOp.MAX)
    e2e_value = world * vox_step * e2e_steps / (float(e2e_t) * 1e-3)

    # the label-driven end-to-end variants run on EVERY rank (same barrier / max-over-ranks rule as the headline e2e)
    lab_ms, lab_bytes = time_e2e_labels_only(P, w, e2e_steps, device, res["inputs"])
    lapi_ms, lapi_bytes = time_e2e_label_api(P, w, e2e_steps, device, res["inputs"])
    if dist_on:
        both = torch.tensor([lab_ms, lapi_ms], device=device, dtype=torch.float64)
        torch.distributed.all_reduce(both, op=torch.distributed.ReduceOp.MAX)
        lab_ms, lapi_ms = float(both[0]), float(both[1])
    res["inputs"] = None
    zres, c5 = None, {}
    if not args.quick:
        torch.cuda.empty_cache()
        zres = time_zshard(P, device, rank, world, 5)
        for dim in (3, 2):
            try:
                c5[dim] = time_cfg5(P, device, rank, world, dim)
            except Exception as ex:  # noqa: BLE001
                c5[dim] = {"unavailable": f"{type(ex).__name__}: {ex}"}
    if rank == 0:
        peak, peak_src = measured_peak()
        ms_step = total_ms / args.steps
        lab0, emb0, tgt0, wgt0, msk0, offs0 = device_inputs(P, w, 0, device)
        e0 = emb0.clone().requires_grad_(True)

        def one_step():
            our_step(P, w, e0, tgt0, wgt0, msk0, offs0, P.WeightedMSE())
            e0.grad = None
        n_ours, k_names, n_all = count_launches(one_step)
        del lab0, emb0, tgt0, wgt0, msk0, e0
        achieved = alg_bytes / (statistics.mean(res["per_ms"]) * 1e-3) / 1e9
        also = {}
        if world == 1 and not args.quick:
            for name in ("cfg1", "cfg2", "cfg3", "cfg2x", "cfg3x"):
                if name == args.workload:
                    continue
                ww = WORKLOADS[name]
                r = time_workload(P, ww, args.steps, args.warmup, device, flush, False)
                ab, kk = algorithmic_bytes_per_step(ww)
                ms = r["total_ms"] / args.steps
                also[name] = {"workload": ww["descr"], "ms_per_step": ms,
                              "value": ww["B"] * kk * int(np.prod(ww["spatial"])) / (ms * 1e-3), "unit": UNIT,
                              "hbm_gbs": ab / (ms * 1e-3) / 1e9, "roofline_frac": ab / (ms * 1e-3) / 1e9 / peak}
                del r
            for name in ("cfg2", "cfg3"):
                also["gt_" + name] = time_gt(P, WORKLOADS[name], device)
                also["gt_" + name]["roofline_frac"] = also["gt_" + name]["hbm_gbs"] / peak
            also["cfg4_epilogue"] = time_inference_epilogue(P, device)
            also["cfg2_labels"] = time_labels_pipeline(P, WORKLOADS["cfg2"], device, flush)
            also["cfg2_labels"]["roofline_frac"] = also["cfg2_labels"]["hbm_gbs"] / peak
            also["cfg1_labels"] = time_labels_pipeline(P, WORKLOADS["cfg1"], device, flush)   # CVPPP table: axis kernel's label flavour
            also["cfg1_labels"]["roofline_frac"] = also["cfg1_labels"]["hbm_gbs"] / peak
            also["step6"] = time_step6(P, device, flush)
            torch.cuda.empty_cache()
            also["forward_and_warm_l2"] = time_forward_and_warm(P, device, flush, peak)
            torch.cuda.empty_cache()
            try:  # context line: what the reference itself does on this GPU (same batch, eager torch ops)
                also["reference_on_gpu"] = time_reference_on_gpu(P, w, device, flush)
                if "ms_per_step" in also["reference_on_gpu"]:
                    also["reference_on_gpu"]["ours_ms_per_step"] = ms_step
                    also["reference_on_gpu"]["speedup"] = also["reference_on_gpu"]["ms_per_step"] / ms_step
            except Exception as ex:  # noqa: BLE001
                also["reference_on_gpu"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
            torch.cuda.empty_cache()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": config_dict(w, args.workload),
            "parallelism": f"dp{world}: every rank runs the workload's batch (independent batches, no data-path collective)",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": args.traffic if args.traffic is not None else ncu_traffic(args.workload)[0],
                         "traffic_source": "--traffic" if args.traffic is not None else ncu_traffic(args.workload)[1],
                         "peak_source": peak_src,
                         "kernel": "whole fused step (all pea kernels of one fwd+bwd); dominant kernel: " + str(ncu_traffic(args.workload)[2]),
                         "algorithmic_bytes_per_step": alg_bytes,
                         "on_chip_ceilings": ON_CHIP_CEILINGS.get(args.workload)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "what": "embedding+target+weight(+mask) from pinned host each step, loss read back"},
            "e2e_labels_only": {"value": world * vox_step * e2e_steps / (lab_ms * 1e-3), "unit": UNIT, "n_gpus": world,
                                "h2d_bytes_per_step": lab_bytes,
                                "what": "embedding resident (backbone output); labels uploaded, GT generated on device"},
            "e2e_label_api": {"value": world * vox_step * e2e_steps / (lapi_ms * 1e-3), "unit": UNIT, "n_gpus": world,
                              "h2d_bytes_per_step": lapi_bytes, "d2h_bytes_per_step": 4,
                              "what": "embedding + int32 labels from pinned host each step through the label-driven loss "
                                      "(embedding_loss_labels / embedding_loss_norm5_labels), loss read back"},
            "gpu_launches": (n_ours if n_ours is not None else LAUNCHES_PER_STEP) * args.steps,
            "gpu_launches_per_step": {"pea_kernels": n_ours, "all_kernels": n_all, "names": k_names,
                                      "how": "torch.profiler (CUPTI) over one public-API fwd+bwd outside the timed region"
                                             if n_ours is not None else "static table (profiler unavailable)"},
            "clocks": clocks,
            "loss": res["loss"],
        }
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline(w)
            line["cpu_baseline"]["gt"] = cpu_gt_baseline(w)
            line["parity"] = parity_vs_cpu(P, w, device)
            line["parity_max_rel"] = line["parity"]["parity_max_rel"]
            if not line["parity"]["ok"]:
                sys.stderr.write(f"PARITY FAILURE against the CPU {line['parity']['against']}: {line['parity']}\n")
        else:
            line["cpu_baseline"] = None
        if zres is not None:
            zres["roofline_frac_per_gpu"] = zres["hbm_gbs_aggregate"] / world / peak
            also["cfg4"] = zres
        for dim, r5 in c5.items():
            also[f"cfg5_{dim}d"] = r5
        if also:
            line["also"] = also
        emit(line)
    if dist_on:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    # The contract is ONE JSON line on stdout: keep a private handle to the real stdout and point fd 1 at stderr so
    # that library banners (e.g. "NCCL version ...") cannot pollute it.
    global _REAL_STDOUT
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg1", "cfg2", "cfg3"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="skip the other workloads (the `also` block)")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes/launch from an ncu capture (for the roofline object)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
