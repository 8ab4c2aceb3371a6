"""All ranks copy device -> pinned host memory (and, second leg, both directions) at the same time: what the box's PCIe
fabric and host memory deliver in aggregate. Launch with torchrun, one rank per GPU; rank 0 prints one JSON line."""
import json
import os
import time

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
h = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
d = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
s = [torch.cuda.Stream(), torch.cuda.Stream()]
out = {}
for name, both in (("d2h_only", False), ("d2h_and_h2d", True)):
    for rep in range(2):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(12):
            with torch.cuda.stream(s[0]):
                h[0].copy_(d[0], non_blocking=True)
            if both:
                with torch.cuda.stream(s[1]):
                    d[1].copy_(h[1], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[name] = {"per_rank_GBps_each_direction": 12 * n / float(t.item()) / 1e9, "aggregate_GBps_each_direction": world * 12 * n / float(t.item()) / 1e9}
if dist.get_rank() == 0:
    print(json.dumps({"n_gpus": world, "host_cores": os.cpu_count(), **out}))
dist.destroy_process_group()
