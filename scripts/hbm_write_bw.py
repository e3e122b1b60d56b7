import torch
x = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
for f, name, bytes_ in ((lambda: x.fill_(1), "fill 2GiB (write only)", 2 << 30), (lambda: y.copy_(x), "copy 2GiB (read+write)", 4 << 30)):
    for _ in range(3): f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(name, round(ms, 3), "ms", round(bytes_ / ms / 1e6, 1), "GB/s")
# read-only stream: a sum over 2 GiB (fp32) -- the ceiling for kernels that only read (the backward's saved-feature load)
z = torch.empty(512 << 20, dtype=torch.float32, device="cuda").normal_()
for _ in range(3): z.sum()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10): z.sum()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print("sum 2GiB (read only)", round(ms, 3), "ms", round((2 << 30) / ms / 1e6, 1), "GB/s")
