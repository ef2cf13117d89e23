import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda:0")
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"H2D 1 GiB pinned: {1e3*(t1-t0):.1f} ms = {n/(t1-t0)/1e9:.1f} GB/s")
# two streams, two halves
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s1): d[:n//2].copy_(h[:n//2], non_blocking=True)
    with torch.cuda.stream(s2): d[n//2:].copy_(h[n//2:], non_blocking=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"H2D 2 streams: {n/(t1-t0)/1e9:.1f} GB/s")
