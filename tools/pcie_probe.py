"""Host<->device copy bandwidth of this box with pinned memory (what bounds bench.py's e2e: one 8.3 MB RGBA8 frame per step)."""
import subprocess
import torch

print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pcie.link.width.max",
                      "--format=csv"], capture_output=True, text=True).stdout.strip())
dev = torch.device("cuda:0")
for mb in (8.2944, 64.0, 512.0):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, (dst, src) in (("d2h", (h, d)), ("h2d", (d, h))):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        reps = max(5, int(2e9 / n))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name} {mb:8.1f} MB: {ms:.4f} ms per copy = {n / ms / 1e6:.1f} GB/s")
