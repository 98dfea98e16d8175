import torch, time
print("peer access 1->0:", torch.cuda.can_device_access_peer(1, 0), "0->1:", torch.cuda.can_device_access_peer(0, 1))
x = torch.randn(4, 4194304, dtype=torch.float64, device="cuda:1")
y = torch.empty_like(x, device="cuda:0")
for dev_ctx in (0, 1):
    with torch.cuda.device(dev_ctx):
        for _ in range(3): y.copy_(x, non_blocking=True)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        t = time.perf_counter()
        for _ in range(10): y.copy_(x, non_blocking=True)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        dt = (time.perf_counter() - t) / 10
        print(f"current device {dev_ctx}: {x.numel()*8/dt/1e9:.1f} GB/s ({dt*1e3:.2f} ms)")
