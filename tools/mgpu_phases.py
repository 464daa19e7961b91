"""Development aid (torchrun): device time of every phase of the row-band
sharded D8 + accumulation step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import torch, torch.distributed as dist
from hydrotools.cuda import distributed as HD

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
rows = cols = 10000
eng = HD.BandEngine(rank, world, rows, cols, rows * world, 10.0, 10.0)
eng.generate_synthetic(20261019)
band = eng.band
for _ in range(3):
    eng.flow_direction_accumulation()
torch.cuda.synchronize(); dist.barrier()

marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))

tot = {}
for it in range(10):
    marks.clear()
    mark("start")
    eng.exchange_halos(2); mark("halo exchange")
    band.direction(); mark("d8")
    band.local(0); mark("stage A")
    band.mark_edges(0); mark("mark band edges")
    band.resolve(True); mark("stage B (roots)")
    both = band.band_tables4(); mark("band tables (1 kernel)")
    P = world
    every = torch.empty((P * 4, cols), dtype=torch.int64, device=both.device)
    dist.all_gather_into_tensor(every, both); mark("all_gather")
    base, nxt = band.band_graph(every, P); mark("band graph (1 kernel)")
    inflow = band.forest_resolve(base, nxt, 0).view(P, 2, cols); mark("band forest resolve")
    band.resolve_delta(inflow[rank, 0], inflow[rank, 1], 0); mark("stage B delta")
    band.final(0); mark("stage C")
    torch.cuda.synchronize()
    for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
        tot[n1] = tot.get(n1, 0.0) + e0.elapsed_time(e1)
dist.barrier()
# tiles whose inflow depends on what the neighbour bands deliver (a marked entry slot)
band.local(0); band.mark_edges(0); band.resolve(True)
tiles = ((rows + 63) // 64) * ((cols + 63) // 64)
agg = band._ws_view(2, torch.int64)[:tiles * 256].view(tiles, 256)
aff = ((agg >> 40) != 0).any(dim=1)
print(f"rank {rank}: {int(aff.sum())} of {tiles} tiles carry a marked entry "
      f"({100.0 * float(aff.float().mean()):.1f} %); per tile row (first 12, last 4): "
      f"{[round(float(x), 2) for x in aff.view((rows + 63) // 64, -1).float().mean(dim=1)[:12]]} ... "
      f"{[round(float(x), 2) for x in aff.view((rows + 63) // 64, -1).float().mean(dim=1)[-4:]]}", flush=True)
names = list(tot.keys())
mine = torch.tensor([tot[k] / 10 for k in names], dtype=torch.float64, device="cuda")
allr = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(allr, mine)
if rank == 0:
    sums = [float(t.sum()) for t in allr]
    slow = max(range(world), key=lambda r: sums[r])
    print("per-rank sums (ms):", " ".join(f"{x:.3f}" for x in sums))
    print(f"{'phase':24s} {'rank 0':>9s} {'rank ' + str(slow) + ' (slowest)':>18s}")
    for i, k in enumerate(names):
        print(f"{k:24s} {float(allr[0][i]):9.4f} {float(allr[slow][i]):18.4f}")
dist.destroy_process_group()
