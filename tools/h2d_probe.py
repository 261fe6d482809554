"""Host-to-device copy rate of the box, the way the upload thread issues it: one clip of the configs[2] shape
(120 frames x 8 masks x 480 x 640 fp32 = 1.18 GB) from pinned host tensors, per frame and as one block."""
import time
import torch

dev = torch.device("cuda:0")
frames = [torch.rand(8, 480, 640).pin_memory() for _ in range(120)]
big = torch.rand(120 * 8, 480, 640).pin_memory()
stage = torch.empty(120 * 8, 480, 640, device=dev)
nbytes = big.numel() * 4
for name, fn in (("120 per-frame copies", lambda: [stage[8 * i:8 * i + 8].copy_(f, non_blocking=True) for i, f in enumerate(frames)]),
                 ("one 1.18 GB copy", lambda: stage.copy_(big, non_blocking=True))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(f"{name}: {1e3 * min(ts):.1f} ms -> {nbytes / min(ts) / 1e9:.1f} GB/s (median {1e3 * sorted(ts)[2]:.1f} ms)")
pageable = torch.rand(8 * 30, 480, 640)
t0 = time.perf_counter(); stage[:240].copy_(pageable); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"pageable 295 MB: {1e3 * dt:.1f} ms -> {pageable.numel() * 4 / dt / 1e9:.1f} GB/s")
