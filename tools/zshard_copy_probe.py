"""Which engine moves the staged halo?  Profiles pea_zshard_stage_halo (cudaMemcpy3DAsync per band + cuStreamWriteValue32)
with CUPTI and prints the activity names (a "Memcpy ..." activity = copy engine, a kernel name = SMs) and the achieved rate
on an otherwise idle GPU.  One GPU: local buffers; torchrun with 2+ ranks: rank r pulls from rank r-1's symmetric memory."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pea_b200 as P  # noqa: F401,E402
from pea_b200 import zshard as Z  # noqa: E402
from pea_b200._lib import lib, check  # noqa: E402
from pea_b200._ops import _ptr  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
H = W = 1024
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    sh = Z.PeerZShard(1, 16, 13 * world, H, W, rank, world, dev)
    sh.own.normal_()
    tail = sh.prev_tail
    sh.publish()
else:
    buf = torch.randn(1, 16, 13, H, W, device=dev)
    tail = buf[:, :, 9:]
if tail is not None:
    stage = torch.empty(1, 16, 4, H, W, device=dev)
    flags = torch.zeros(64, device=dev, dtype=torch.int32)
    side = torch.cuda.Stream()

    def pull(epoch, row0=96, band=64):
        check("stage", lib.pea_zshard_stage_halo(_ptr(tail), int(tail.stride(1)), _ptr(stage), 1, 16, 4, H, W, row0, band,
                                                 _ptr(flags), epoch, side.cuda_stream))
    pull(1); torch.cuda.synchronize()
    assert torch.equal(stage[:, :, :, 96:], tail[:, :, :, 96:]) and int(flags[:15].min()) == 1
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        pull(2); torch.cuda.synchronize()
    names = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            names[ev.name] = names.get(ev.name, 0) + 1
    for band in (32, 64, 128, 1024):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            a.record()
            for i in range(5):
                pull(10 + i, 0, band)
            b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"rank {rank}: band {band:4d} rows: {ms:.4f} ms per 268 MB pull = {268.4 / ms:.0f} GB/s", flush=True)
    print(f"rank {rank}: device activities of one pull: {names}", flush=True)
if world > 1:
    sh.release(); dist.barrier(); dist.destroy_process_group()
