"""PCIe ceiling for the host-batch path: DMA copies and torch zero-copy-free baselines at the batch size (232 KB) and at 64 MB."""
import time, torch
dev = torch.device("cuda:0")
for nbytes in (231888, 1 << 20, 64 << 20):
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for name, f in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(5): f()
        torch.cuda.synchronize()
        # device time of one copy (events) and wall time of copy + synchronize
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        N = 50
        e0.record()
        for _ in range(N): f()
        e1.record(); torch.cuda.synchronize()
        dev_us = e0.elapsed_time(e1) * 1e3 / N
        t0 = time.perf_counter()
        for _ in range(N):
            f(); torch.cuda.synchronize()
        wall_us = (time.perf_counter() - t0) * 1e6 / N
        print("%s %9d B: %.1f us back-to-back (%.1f GB/s), %.1f us copy+sync" % (name, nbytes, dev_us, nbytes / dev_us / 1e3, wall_us))
