"""What can B200 deliver for random row gathers?  torch.index_select (ATen) as an independent yardstick."""
import torch
dev = torch.device("cuda:0")
n, E = 235868, 2358104
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
idx = torch.randint(0, n, (E,), device=dev, generator=g)
idx_sorted = torch.sort(idx).values
def t(fn, it=5):
    fn(); ts = []
    for _ in range(it):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts)//2] * 1e3
for row_bytes in (128, 256, 512, 1024, 2048):
    x = torch.randn(n, row_bytes // 4, device=dev)
    out = torch.empty(E, row_bytes // 4, device=dev)
    for name, ix in (("random", idx), ("sorted", idx_sorted)):
        us = t(lambda: torch.index_select(x, 0, ix, out=out))
        print(f"index_select rows={row_bytes:5d}B {name}: {us:8.1f} us  read {E*row_bytes/us/1e3:7.0f} GB/s  read+write {2*E*row_bytes/us/1e3:7.0f} GB/s  {E/us/1e3:6.2f} Grows/s", flush=True)
a = torch.empty(1 << 28, dtype=torch.float32, device=dev); b = torch.empty_like(a)
us = t(lambda: b.copy_(a)); print(f"copy 1 GiB: {us:.1f} us {2*a.numel()*4/us/1e3:.0f} GB/s")
