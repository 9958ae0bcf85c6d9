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
# neighbour exchange of border cells on the GPUs (NCCL send/recv): rank r receives exactly the band of rank r+1
rng = np.random.default_rng(100 + rank)
n = 4000 + 100 * rank
xy = np.column_stack([rng.uniform(0, 50, n), rng.uniform(10.0 * rank, 10.0 * (rank + 1), n)])
prob = rng.uniform(0, 100, (n, 3))
ty = rng.integers(0, 3, n).astype(np.int32)
halo, info = sharding.exchange_halo({"xy": xy, "prob": prob, "type": ty}, ("xy", 1), 10.0 * rank + 2.5, device=torch.device("cuda", local))
nxt = np.random.default_rng(100 + rank + 1)
n2 = 4000 + 100 * (rank + 1)
xy2 = np.column_stack([nxt.uniform(0, 50, n2), nxt.uniform(10.0 * (rank + 1), 10.0 * (rank + 2), n2)])
prob2 = nxt.uniform(0, 100, (n2, 3)); ty2 = nxt.integers(0, 3, n2).astype(np.int32)
band = xy2[:, 1] < 10.0 * (rank + 1) + 2.5
dev = torch.device("cuda", local)      # the device-resident form: tensors in, tensors out, same rows
halo_d, info_d = sharding.exchange_halo({"xy": torch.from_numpy(xy).to(dev), "prob": torch.from_numpy(prob).to(dev), "type": torch.from_numpy(ty).to(dev)},
                                        ("xy", 1), 10.0 * rank + 2.5)
assert info_d == info and all(np.array_equal(halo_d[k].cpu().numpy(), halo[k]) for k in halo) and halo_d["xy"].is_cuda
if rank < world - 1:
    assert np.array_equal(halo["xy"], xy2[band]) and np.array_equal(halo["prob"], prob2[band]) and np.array_equal(halo["type"][:, 0], ty2[band])
    assert halo["type"].dtype == np.int32 and info["rows"] == int(band.sum()) and info["bytes"] == int(band.sum()) * (16 + 24 + 4)
else:
    assert len(halo["xy"]) == 0 and info["rows"] == 0
# row-sharded upload + NVLink all-gather builds the same section as a plain upload: candidates of a window block agree bit for bit
from same_b200 import _lib as L
from same_b200.device import Section
lut = {c: i for i, c in enumerate(ct)}
fr = (qry[["X", "Y"]].to_numpy(), ref[["X", "Y"]].to_numpy(), qry[list(ct)].to_numpy(), ref[list(ct)].to_numpy(),
      qry["cell_type"].map(lut).to_numpy(np.int32), ref["cell_type"].map(lut).to_numpy(np.int32))
rects = np.array([[x, x + 14.0, y, y + 14.0] for x in (0.0, 10.0, 20.0, 30.0) for y in (0.0, 10.0, 20.0, 30.0)])
with sharding.section_from_row_shards(fr, device_index=local) as s1, Section(*fr, device=local) as s2:
    with s1.batch(rects) as b1, s2.batch(rects) as b2:
        b1.candidates(1.0, 6, False, 1.0); b2.candidates(1.0, 6, False, 1.0)
        for w_ in (L.KEEP_A, L.KEEP_R, L.PAIRS, L.COST, L.ROW_PTR):
            assert np.array_equal(b1.get(w_), b2.get(w_)), "row-sharded section differs from the plain upload"
if rank == 0:
    print("row-sharded section ok")
ok = torch.tensor([1], device=torch.device("cuda", local))
dist.all_reduce(ok)
if rank == 0:
    assert int(ok.item()) == world
    print("halo exchange ok")
dist.barrier()
dist.destroy_process_group()
