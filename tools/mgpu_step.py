"""Development aid (torchrun): device time of the row-band sharded D8 + accumulation step
(10 000 rows per GPU, as bench.py times it) and the mass balance of the result."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import torch, torch.distributed as dist
from hydrotools.cuda import distributed as HD

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
rows = cols = 10000
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
eng = HD.BandEngine(rank, world, rows, cols, rows * world, 10.0, 10.0)
eng.generate_synthetic(20261019)
for _ in range(5):
    eng.flow_direction_accumulation()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
t0 = time.perf_counter()
e0.record()
for _ in range(steps):
    eng.flow_direction_accumulation()
e1.record()
host = (time.perf_counter() - t0) / steps * 1e3
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps, host], dtype=torch.float64, device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
host = float(ms[1]); ms = ms[:1]
if rank == 0:
    cells = rows * cols * world
    print(f"world {world}: {float(ms):.4f} ms per step = {cells / float(ms) / 1e6:.1f} Gcells/s; host time to enqueue a step {host:.3f} ms")

# timeline of one step (events on the main and the side stream), slowest rank
band = eng.band
if world > 1:
    acc = {}
    for it in range(10):
        ev = []
        def mark(name, stream=None):
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream) if stream is not None else e.record()
            ev.append((name, e))
        dist.barrier(); torch.cuda.synchronize()
        mark("start")
        work = eng.exchange_halos(2, async_op=True)
        e_ = band.D8_TILE_ROWS
        band.direction_part(e_, rows - e_); mark("d8 interior")
        work(); mark("halo rows placed")
        band.direction_part(0, e_); band.direction_part(rows - e_, rows); mark("d8 edge tile rows")
        band.local(0); mark("stage A")
        band.mark_edges(0); band.resolve(True); mark("stage B (roots)")
        band.tile_phases(0); mine = band.publish_tables(0); mark("tile phases + tables")
        main = torch.cuda.current_stream(); side = band.side_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            every = torch.empty((world * 4, cols), dtype=torch.int64, device=mine.device)
            dist.all_gather_into_tensor(every, mine); mark("side: all_gather", side)
            band.import_flows(every, world, rank, 0, True); mark("side: forest + delta resolve", side)
        band.final_phase(0, 0); mark("stage C phase 0")
        main.wait_stream(side)
        band.final_phase(1, 0); mark("stage C phase 1")
        torch.cuda.synchronize()
        for n, e in ev[1:]:
            acc[n] = acc.get(n, 0.0) + ev[0][1].elapsed_time(e) / 10
    names = list(acc.keys())
    mine_t = torch.tensor([acc[k] for k in names], dtype=torch.float64, device="cuda")
    allr = [torch.empty_like(mine_t) for _ in range(world)]
    dist.all_gather(allr, mine_t)
    if rank == 0:
        print("time since step start (ms) at which each phase ENDS, per rank:")
        for i, k in enumerate(names):
            print(f"  {k:30s} " + " ".join(f"{float(t[i]):7.3f}" for t in allr))
dist.destroy_process_group()
