import torch, time
x = torch.empty(210*1024*1024, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for n in (13, 52, 210):
    s = x[:n*1024*1024]; t = d[:n*1024*1024]
    for _ in range(2): t.copy_(s, non_blocking=True)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(5): t.copy_(s, non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/5
    print("H2D %d MB: %.2f ms  %.1f GB/s" % (n, dt*1e3, n*1.048576/1e3/dt))
y = torch.empty(320000, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t0=time.perf_counter(); y.copy_(d[:320000]); torch.cuda.synchronize(); print("D2H 320KB %.3f ms" % ((time.perf_counter()-t0)*1e3))
