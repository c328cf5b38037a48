import torch, time
x = torch.empty(3229876224 // 4, dtype=torch.float32, device="cuda")
for fn, name, bytes_ in ((lambda: x.fill_(1.0), "fill_ (write only)", x.numel() * 4), (lambda: x.zero_(), "zero_ (memset)", x.numel() * 4)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.3f} ms  {bytes_ / ms / 1e6:.0f} GB/s")
y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"copy (read+write): {ms:.3f} ms  {2 * x.numel() * 4 / ms / 1e6:.0f} GB/s")
