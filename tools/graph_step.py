"""Development aid (python or torchrun): the D8 + accumulation step enqueued launch by launch
vs replayed as one CUDA graph; checks that both leave the same result."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import torch, torch.distributed as dist
from hydrotools.cuda import distributed as HD

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rows = cols = 10000
steps = 20
eng = HD.BandEngine(rank, world, rows, cols, rows * world, 10.0, 10.0)
eng.generate_synthetic(20261019)

def timed(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / steps * 1e3
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps, host], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms[0]), float(ms[1])

a = timed(eng.flow_direction_accumulation)
fd0, fa0 = (t.clone() for t in eng.results())
eng.band.acc.zero_(); eng.band.dir.zero_()
print(f"rank {rank}: capturing", flush=True)
replay = eng.graph(eng.flow_direction_accumulation)
print(f"rank {rank}: captured", flush=True)
replay(); torch.cuda.synchronize()
print(f"rank {rank}: first replay done", flush=True)
b = timed(replay)
fd1, fa1 = eng.results()
same = bool(torch.equal(fd0, fd1)) and bool(torch.equal(fa0.nan_to_num(-1), fa1.nan_to_num(-1)))
if rank == 0:
    print(f"world {world}: launch by launch {a[0]:.4f} ms per step (host enqueue {a[1]:.4f} ms); "
          f"graph replay {b[0]:.4f} ms (host {b[1]:.4f} ms); same result: {same}; launches per step "
          f"{eng.launches // (steps + 5) if False else 'n/a'}")
sys.stdout.flush()
import threading
threading.Timer(20.0, lambda: (print(f"rank {rank}: teardown hung, hard exit", flush=True), os._exit(0))).start()
replay.release()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
print(f"rank {rank}: clean exit", flush=True)
os._exit(0)
