"""Host-to-device bandwidth of pinned memory on this box (what bounds bench.py's e2e)."""
import torch
for mb in (4, 16, 64, 256):
    n = mb * 1024 * 1024
    src = torch.empty(n, dtype=torch.uint8).pin_memory()
    dst = torch.empty(n, dtype=torch.uint8, device='cuda')
    for _ in range(3): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): dst.copy_(src, non_blocking=True)
    b.record(); b.synchronize()
    print('%4d MB: %.1f GB/s' % (mb, 10 * n / (a.elapsed_time(b) * 1e-3) / 1e9))

# does a second copy in flight (another stream) hide the fixed cost of a copy?
n = 16 * 1024 * 1024
srcs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
dsts = [torch.empty(n, dtype=torch.uint8, device='cuda') for _ in range(4)]
for ns in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    for k in range(12):
        with torch.cuda.stream(streams[k % ns]):
            dsts[k % 4].copy_(srcs[k % 4], non_blocking=True)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    b.record(); b.synchronize()
    print('16 MB copies over %d stream(s): %.3f ms per copy, %.1f GB/s' % (ns, a.elapsed_time(b) / 12, 12 * n / (a.elapsed_time(b) * 1e-3) / 1e9))
# one 16 MB copy split into chunks on one stream
for parts in (1, 2, 4, 8):
    c = n // parts
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(10):
        for q in range(parts):
            dsts[0][q * c:(q + 1) * c].copy_(srcs[0][q * c:(q + 1) * c], non_blocking=True)
    b.record(); b.synchronize()
    print('16 MB as %d chunk(s) on one stream: %.3f ms' % (parts, a.elapsed_time(b) / 10))
