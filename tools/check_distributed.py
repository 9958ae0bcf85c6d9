"""2+ GPU functional check (torchrun): distributed_sliding_window_matching == the single-process result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_distributed.py
Each rank runs (on the GPU of its LOCAL_RANK, same_b200.device.default_device) its contiguous block of the window list on its own GPU (IncumbentBackend in place of Gurobi, same seeded rule
everywhere); rank 0 also runs all windows alone and compares the gathered frame with it row for row."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import same_b200
from same_b200 import datagen, sharding
from same_b200.solver import IncumbentBackend
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ref, qry, ct = datagen.make_section_pair(n_tiles=16, n_types=3, seed=9)
optim = dict(window_size=12, overlap=3, min_cells_per_window=10, radius=1.0, knn=6, max_matches=1, min_angle_deg=15, cell_id_col="Cell_Num_Old")

def inc(spec):
    rp = np.asarray(spec.row_ptr, dtype=np.int64)
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    return bench.incumbent(np.column_stack([rows, rows]), seed=len(rows))

with contextlib.redirect_stdout(io.StringIO()):
    got = sharding.distributed_sliding_window_matching(ref, qry, commonCT=list(ct), optim_params=dict(optim), gurobi_params={}, solver=IncumbentBackend(inc))
    full = same_b200.sliding_window_matching(ref, qry, commonCT=list(ct), optim_params=dict(optim), gurobi_params={}, solver=IncumbentBackend(inc)) if rank == 0 else None
if rank == 0:
    cols = [c for c in full.columns if c != "run_time"]
    a = got.sort_values(["window_id", "aligned_idx"]).reset_index(drop=True)[cols]
    b = full.sort_values(["window_id", "aligned_idx"]).reset_index(drop=True)[cols]
    assert len(a) == len(b) and all((a[c].to_numpy() == b[c].to_numpy()).all() for c in cols), "sharded result differs from the single-process result"
    print(f"distributed check ok: world={world}, {b['window_id'].nunique()} windows, {len(b)} matches identical")
dist.barrier()
dist.destroy_process_group()
