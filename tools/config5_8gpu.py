"""BASELINE.json configs[4] (torchrun, 8 x B200): D8 + cell-count accumulation on a
100 000 x 100 000 synthetic conditioned DEM (1e10 cells), 12 500 rows per GPU.
Correctness at this size through size-independent properties:
  * acc(c) == 1 + sum of acc over c's donors for EVERY cell, donors in the
    neighbour band included (edge rows of direction and accumulation are exchanged);
  * the accumulation of the cells that do not pass flow on adds up to the cell count.
Smaller sizes: python -m torch.distributed.run ... tools/config5_8gpu.py ROWS_PER_GPU COLS"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
os.environ.setdefault("NCCL_DEBUG", "NONE")
import torch, torch.distributed as dist
from hydrotools.cuda import distributed as HD

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
total = rows * world
eng = HD.BandEngine(rank, world, rows, cols, total, 10.0, 10.0)
eng.generate_synthetic(20261019)
pits, ties = eng.validate()
for _ in range(2):
    eng.flow_direction_accumulation()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.flow_direction_accumulation()
e1.record(); torch.cuda.synchronize(); dist.barrier()
t = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
fd, fa = eng.results()

# edge rows of every band (first, last) for the cross-band donors
edge_fd = torch.stack([fd[0], fd[-1]]).contiguous()
edge_fa = torch.stack([fa[0], fa[-1]]).contiguous()
all_fd = torch.empty((world * 2, cols), dtype=fd.dtype, device=dev)
all_fa = torch.empty((world * 2, cols), dtype=fa.dtype, device=dev)
dist.all_gather_into_tensor(all_fd, edge_fd); dist.all_gather_into_tensor(all_fa, edge_fa)
dr = torch.tensor([0, -1, -1, -1, 0, 1, 1, 1, 0], device=dev)
dc = torch.tensor([0, 1, 0, -1, -1, -1, 0, 1, 1], device=dev)
g0 = rank * rows  # global row of my row 0
inflow = torch.zeros((rows, cols), dtype=torch.float64, device=dev)
cc1 = torch.arange(cols, device=dev).view(1, -1)


def scatter(f, a, grow):
    """Add the accumulation of the donors in rows `grow` (global row numbers, [n, 1]) whose
    receiver lies in my band."""
    f = f.to(torch.int64)
    cc = cc1.expand(f.shape[0], -1)
    rr = grow.expand(-1, cols)
    edge = (rr == 0) | (cc == 0) | (rr == total - 1) | (cc == cols - 1)
    don = (f > 0) & ~edge
    code = f.clamp(min=0)
    tr, tc = rr + dr[code] - g0, cc + dc[code]
    don &= (tr >= 0) & (tr < rows)
    inflow.view(-1).index_add_(0, (tr * cols + tc)[don], a[don])
    return don


CH = 500
term_sum = torch.zeros((), dtype=torch.float64, device=dev)
for r0 in range(0, rows, CH):
    r1 = min(rows, r0 + CH)
    grow = torch.arange(g0 + r0, g0 + r1, device=dev).view(-1, 1)
    f = fd[r0:r1].to(torch.int64)
    rr = grow.expand(-1, cols); cc = cc1.expand(r1 - r0, -1)
    edge = (rr == 0) | (cc == 0) | (rr == total - 1) | (cc == cols - 1)
    passes = (f > 0) & ~edge
    term_sum += fa[r0:r1][~passes].sum()
    scatter(fd[r0:r1], fa[r0:r1], grow)
if rank > 0:          # last row of the band above
    scatter(all_fd[2 * (rank - 1) + 1:2 * (rank - 1) + 2], all_fa[2 * (rank - 1) + 1:2 * (rank - 1) + 2],
            torch.tensor([[g0 - 1]], device=dev))
if rank < world - 1:  # first row of the band below
    scatter(all_fd[2 * (rank + 1):2 * (rank + 1) + 1], all_fa[2 * (rank + 1):2 * (rank + 1) + 1],
            torch.tensor([[g0 + rows]], device=dev))
ok = torch.tensor([1.0 if torch.equal(fa, inflow + 1) else 0.0], dtype=torch.float64, device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
dist.all_reduce(term_sum)
mx = fa.max().reshape(1).clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
mem = torch.tensor([torch.cuda.max_memory_allocated() / 2**30], dtype=torch.float64, device=dev)
dist.all_reduce(mem, op=dist.ReduceOp.MAX)
if rank == 0:
    cells = total * cols
    print(f"{total} x {cols} on {world} GPUs: validate (pits, ties) = ({pits}, {ties}); "
          f"{ms:.2f} ms per step = {cells / ms / 1e6:.1f} Gcells/s; peak {float(mem):.1f} GiB per GPU")
    print(f"acc == 1 + sum(donors) on every rank: {bool(ok.item())}; flow reaching terminal cells "
          f"{float(term_sum):.0f} of {cells} cells; max accumulation {float(mx):.0f}")
    assert bool(ok.item()) and float(term_sum) == cells
    print("config5 ok")
dist.destroy_process_group()
