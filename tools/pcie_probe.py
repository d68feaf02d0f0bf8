"""PCIe floor for the end-to-end SpMV: 64 MiB H2D and 64 MiB D2H, alone and simultaneously (two streams), pinned memory."""
import time
import torch
n = 16777216
hx = torch.empty(n).pin_memory(); hy = torch.empty(n).pin_memory()
dx = torch.empty(n, device="cuda"); dy = torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=20):
    for i in range(reps + 3):
        if i == 3:
            torch.cuda.synchronize(); t = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1):
                dx.copy_(hx, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                hy.copy_(dy, non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


for name, a, b in (("h2d", 1, 0), ("d2h", 0, 1), ("both", 1, 1)):
    ms = run(a, b)
    print(f"PCIE {name} ms={ms:.3f} GBps_each_way={n * 4 / ms / 1e6:.1f}")
