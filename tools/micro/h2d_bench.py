"""H2D bandwidth from pinned host memory: one stream vs several concurrent streams (PCIe Gen5 x16 on the B200 box)."""
import torch, time
n = 800 << 20
host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
dev = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
streams = [torch.cuda.Stream() for _ in range(4)]
for k in (1, 2, 4):
    for chunk in (n, 64 << 20, 16 << 20):
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for off in range(0, n, chunk):
                for i in range(4):
                    with torch.cuda.stream(streams[i % k]):
                        dev[i][off:off + chunk].copy_(host[i][off:off + chunk], non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print(f"streams={k} chunk={chunk >> 20} MiB: {4 * n / best / 1e9:.1f} GB/s")
