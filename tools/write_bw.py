import torch, time
n = 741217625
x = torch.empty(n, dtype=torch.float64, device="cuda")
for name, fn in [("fill_", lambda: x.fill_(1.5)), ("zero_", lambda: x.zero_()), ("memset", lambda: torch.cuda.current_stream().synchronize() or x.zero_())]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(name, "ms", ms, "GB/s", n * 8 / ms / 1e6)
y = torch.empty(n // 2, dtype=torch.float64, device="cuda"); z = torch.empty_like(y)
for _ in range(3): z.copy_(y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): z.copy_(y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("copy ms", ms, "GB/s (r+w)", 2 * (n // 2) * 8 / ms / 1e6)
