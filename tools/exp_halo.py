"""Dev helper (torchrun, >= 2 GPUs): time the device-resident halo exchange, repeated."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from same_b200 import sharding as S
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rng = np.random.default_rng(rank)
n = 1_000_000
xy = torch.from_numpy(np.column_stack([rng.uniform(0, 100, n), rng.uniform(100 * rank, 100 * (rank + 1), n)])).to(dev)
prob = torch.from_numpy(rng.uniform(0, 1, (n, 3))).to(dev); ty = torch.from_numpy(rng.integers(0, 3, n).astype(np.int32)).to(dev)
S.exchange_halo({"y": np.full((8, 1), -1.0)}, "y", 0.0, device=dev)
for rep in range(5):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    halo, info = S.exchange_halo({"xy": xy, "prob": prob, "type": ty}, ("xy", 1), 100 * rank + 16.0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0: print(f"rep {rep}: {1e3 * (t1 - t0):.2f} ms for {info['bytes'] / 1e6:.1f} MB ({info['rows']} rows)")
dist.barrier(); dist.destroy_process_group()
