"""Pinned host <-> device copy bandwidth of the box (the e2e figure of bench.py is judged against it)."""
import torch, time
x = torch.empty(50331648 // 8, dtype=torch.float64).pin_memory()
d = torch.empty_like(x, device='cuda')
o = torch.empty(1 << 20, dtype=torch.float64, device='cuda'); oh = torch.empty(1 << 20, dtype=torch.float64).pin_memory()
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t = (time.perf_counter() - t0) / 10
print('H2D 50.3 MB: %.3f ms  %.1f GB/s' % (t * 1e3, 50331648 / t / 1e9))
t0 = time.perf_counter()
for _ in range(10): oh.copy_(o, non_blocking=True)
torch.cuda.synchronize()
t = (time.perf_counter() - t0) / 10
print('D2H 8.4 MB: %.3f ms  %.1f GB/s' % (t * 1e3, 8388608 / t / 1e9))
# chunked
n = 8
parts = x.chunk(n); dparts = d.chunk(n)
t0 = time.perf_counter()
for _ in range(10):
    for a, b in zip(parts, dparts): b.copy_(a, non_blocking=True)
torch.cuda.synchronize()
t = (time.perf_counter() - t0) / 10
print('H2D in 8 chunks: %.3f ms' % (t * 1e3))
import subprocess
print(subprocess.run(['nvidia-smi', '--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max', '--format=csv'], capture_output=True, text=True).stdout)
