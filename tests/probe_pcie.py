"""Probe (not a test): device->host bandwidth into page-locked memory with 1, 2 and 4 concurrent copies (torch streams) - does one
cudaMemcpyAsync saturate the link of this box?"""
import time

import torch

n = 44 * (1 << 20)
src = torch.empty(4 * n, dtype=torch.uint8, device="cuda")
dst = torch.empty(4 * n, dtype=torch.uint8).pin_memory()
streams = [torch.cuda.Stream() for _ in range(4)]


def run(k, total):
    per = total // k
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(k):
        with torch.cuda.stream(streams[i]):
            dst[i * per:(i + 1) * per].copy_(src[i * per:(i + 1) * per], non_blocking=True)
    torch.cuda.synchronize()
    return total / (time.perf_counter() - t0) / 1e9


for total in (n, 4 * n):
    for k in (1, 2, 4):
        run(k, total)
        print(f"{total >> 20} MB D2H in {k} concurrent copies: {max(run(k, total) for _ in range(5)):.1f} GB/s")
h = torch.empty(4 * n, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize()
t0 = time.perf_counter()
src.copy_(h, non_blocking=True)
torch.cuda.synchronize()
print(f"H2D {4 * n / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
