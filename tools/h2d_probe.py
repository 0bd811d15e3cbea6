"""Host-to-device copy rate of this box from pinned memory (what bounds the e2e extraction)."""
import time, torch
n = 102_150_960
pin = torch.empty(n, dtype=torch.uint8).pin_memory()
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3): dev.copy_(pin, non_blocking=True)
torch.cuda.synchronize()
for size in (n, n // 4, n // 31):
    t0 = time.perf_counter()
    for _ in range(10): dev[:size].copy_(pin[:size], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print("H2D %9d bytes: %.3f ms = %.1f GB/s" % (size, dt * 1e3, size / dt / 1e9))
