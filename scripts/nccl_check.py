"""Multi-GPU plumbing check (run under torchrun): NCCL init, small and gradient-sized all-reduce, then one ParamArena step."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
t0 = time.time()
dist.init_process_group("nccl", device_id=dev)
x = torch.ones(1024, device=dev) * (rank + 1)
dist.all_reduce(x)
torch.cuda.synchronize()
print(f"[rank {rank}] small all_reduce ok ({x[0].item()}) after {time.time() - t0:.1f}s", flush=True)
big = torch.ones(100_950_000, device=dev)
for i in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    dist.all_reduce(big)
    e.record()
    torch.cuda.synchronize()
    print(f"[rank {rank}] 404 MB all_reduce {s.elapsed_time(e):.2f} ms", flush=True)
from transformer4sed_b200 import functional as F  # noqa: E402
from transformer4sed_b200.training import ParamArena  # noqa: E402
lin = torch.nn.Linear(256, 256).to(dev)
arena = ParamArena(lin, [dict(name="all", params=list(lin.parameters()), lr=1e-3, weight_decay=0.0)])
y = F.linear(torch.randn(64, 256, device=dev).to(torch.bfloat16), lin.weight, lin.bias)
y.float().sum().backward()
arena.step()
torch.cuda.synchronize()
print(f"[rank {rank}] arena step ok", flush=True)
dist.barrier()
torch.cuda.synchronize()
print(f"[rank {rank}] barrier ok", flush=True)
dist.destroy_process_group()
