"""Bandwidth of PeerGather pushes (torchrun, >= 2 ranks) (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pyneapple_b200 import parallel
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
shape = (4, 4194304)
pg = parallel.PeerGather(shape, torch.float64, dev)
x = torch.full(shape, float(rank + 1), dtype=torch.float64, device=dev)
for _ in range(3): pg.push(x)
pg.wait()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(10): pg.push(x)
pg.flush(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
out = pg.wait()
print(f"rank {rank}: push of {x.numel()*8/1e6:.0f} MB in {ms:.3f} ms -> {x.numel()*8/ms/1e6:.0f} GB/s", flush=True)
if rank == 0:
    print("blocks ok:", [float(out[r].mean()) for r in range(dist.get_world_size())])
dist.destroy_process_group()
