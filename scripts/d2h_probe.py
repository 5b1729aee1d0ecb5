import torch, time
x = torch.empty(3, 1080, 1920, device="cuda")
h = torch.empty(3, 1080, 1920).pin_memory()
for _ in range(3): h.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): h.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 50
print(f"D2H 24.9 MB pinned: {dt*1e3:.3f} ms = {x.numel()*4/dt/1e9:.1f} GB/s")
x8 = torch.empty(3, 1080, 1920, device="cuda", dtype=torch.uint8); h8 = torch.empty(3,1080,1920,dtype=torch.uint8).pin_memory()
for _ in range(3): h8.copy_(x8, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): h8.copy_(x8, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
print(f"D2H 6.2 MB pinned: {dt*1e3:.3f} ms = {x8.numel()/dt/1e9:.1f} GB/s")
