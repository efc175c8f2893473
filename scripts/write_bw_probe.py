import torch
x = torch.empty(8 << 30, dtype=torch.uint8, device="cuda")
for name, fn in (("memset", lambda: x.zero_()), ("fill_f16", lambda: x.view(torch.float16).fill_(1.5))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(name, round(x.numel() / ms / 1e6, 1), "GB/s")
y = torch.empty(4 << 30, dtype=torch.uint8, device="cuda"); z = torch.empty_like(y)
for _ in range(3): z.copy_(y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): z.copy_(y)
e1.record(); torch.cuda.synchronize()
print("copy r+w", round(2 * y.numel() / (e0.elapsed_time(e1) / 10) / 1e6, 1), "GB/s")
