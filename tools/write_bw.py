"""Write-only HBM bandwidth of this B200 (the roofline of the write-dominated GT kernels): torch fill_ / zero_ (a memset
kernel: 128-bit stores) over buffers far larger than L2, CUDA events, best and median of 10."""
import torch
for mib in (512, 2048):
    x = torch.empty(mib << 18, dtype=torch.float32, device="cuda")
    for name, fn in (("fill_", lambda: x.fill_(1.0)), ("zero_", lambda: x.zero_())):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        by = x.numel() * 4
        print(f"{name} {mib} MiB: best {by / ts[0] / 1e6:.0f} GB/s  median {by / ts[5] / 1e6:.0f} GB/s", flush=True)
    y = torch.empty_like(x)
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); y.copy_(x); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"copy_ {mib} MiB: best {2 * x.numel() * 4 / ts[0] / 1e6:.0f} GB/s (read+write)", flush=True)
    del x, y
