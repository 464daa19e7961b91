"""BASELINE.json configs[3] (torchrun, 8 x B200): precipitation-weighted flow accumulation
on a 40 000 x 40 000 synthetic DEM (1.6e9 cells), 5 000 rows per GPU, weights
1650..1700 mm (hash of seed, row, column).  Check at this size: the weighted
accumulation of the cells that do not pass flow on adds up to the total weight."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
os.environ.setdefault("NCCL_DEBUG", "NONE")
import torch, torch.distributed as dist
from hydrotools.cuda import distributed as HD, _native as N, _device as D

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
total = rows * world
eng = HD.BandEngine(rank, world, rows, cols, total, 10.0, 10.0)
eng.generate_synthetic(20261019)
w, wld = D.empty_padded(rows, cols, torch.float32, dev, 4)
N.call("ht_synth_weight_f32", w.data_ptr(), rows, cols, wld, 20261020, rank * rows, D.stream_ptr())
wv = w[:, :cols]
eng.flow_direction_accumulation()
fd, _ = eng.results()
for _ in range(2):
    acc = eng.flow_accumulation(wv)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    acc = eng.flow_accumulation(wv)
e1.record(); torch.cuda.synchronize(); dist.barrier()
t = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
rr = torch.arange(rank * rows, (rank + 1) * rows, device=dev).view(-1, 1)
cc = torch.arange(cols, device=dev).view(1, -1)
edge = (rr == 0) | (cc == 0) | (rr == total - 1) | (cc == cols - 1)
term = ~((fd > 0) & ~edge)
s = torch.stack([acc[term].sum(), wv.double().sum()])
dist.all_reduce(s)
if rank == 0:
    ms = float(t.item())
    print(f"weighted accumulation {total} x {cols} on {world} GPUs: {ms:.2f} ms per pass = "
          f"{total * cols / ms / 1e6:.1f} Gcells/s")
    print(f"weight reaching terminal cells {float(s[0]):.6e} of {float(s[1]):.6e} "
          f"(relative difference {abs(float(s[0]) - float(s[1])) / float(s[1]):.1e})")
    assert abs(float(s[0]) - float(s[1])) / float(s[1]) < 1e-9
    print("config4 ok")
dist.destroy_process_group()
