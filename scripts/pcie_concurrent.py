import os, sys, time, torch
r = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(r)
import torch.distributed as dist
dist.init_process_group("gloo")
h = torch.empty(297209856 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(h, device="cuda")
for _ in range(2):
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
dist.barrier()
t = time.perf_counter()
for _ in range(5):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 5
print(f"rank {r} cpu affinity {sorted(os.sched_getaffinity(0))[:4]}.. n={len(os.sched_getaffinity(0))} concurrent D2H 297 MB: {dt*1e3:.2f} ms = {297.2/dt/1e3:.1f} GB/s", flush=True)
dist.barrier()
if r == 0:
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    print(f"rank 0 alone: {(time.perf_counter()-t)*1e3:.2f} ms", flush=True)
dist.barrier()
