"""BASELINE.json configs[4] per-GPU shape on ONE GPU: a 12 500 x 100 000 band
(what each of 8 B200s holds of the 100k x 100k DEM), D8 + accumulation, checked
through the size-independent mass balance  acc(c) = 1 + sum of acc of c's donors
(computed on the device with torch) and timed."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import torch
from hydrotools.cuda import distributed as HD

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
eng = HD.BandEngine(0, 1, rows, cols, rows, 10.0, 10.0)
eng.generate_synthetic(20261019)
print("validate (pits, ties):", eng.validate())
for _ in range(2):
    eng.flow_direction_accumulation()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.flow_direction_accumulation()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"{rows} x {cols}: {ms:.3f} ms per step, {rows * cols / ms / 1e6:.2f} Gcells/s, "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
for k in eng.time_kernels(3):
    print(f"  {k['name']:24s} {k['ms']:.3f} ms")
fd, fa = eng.results()
n = rows * cols
# mass balance in row chunks (receivers of a chunk lie within one row of it)
dr = torch.tensor([0, -1, -1, -1, 0, 1, 1, 1, 0], device="cuda")
dc = torch.tensor([0, 1, 0, -1, -1, -1, 0, 1, 1], device="cuda")
inflow = torch.zeros((rows, cols), dtype=torch.float64, device="cuda")
CH = 500
for r0 in range(0, rows, CH):
    r1 = min(rows, r0 + CH)
    f = fd[r0:r1].to(torch.int64)
    rr = torch.arange(r0, r1, device="cuda").view(-1, 1).expand(-1, cols)
    cc = torch.arange(cols, device="cuda").view(1, -1).expand(r1 - r0, -1)
    edge = (rr == 0) | (cc == 0) | (rr == rows - 1) | (cc == cols - 1)
    don = (f > 0) & ~edge
    code = f.clamp(min=0)
    tr, tc = rr + dr[code], cc + dc[code]
    idx = (tr * cols + tc)[don]
    inflow.view(-1).index_add_(0, idx, fa[r0:r1][don])
ok = torch.equal(fa, inflow + 1)
rr = torch.arange(rows, device="cuda").view(-1, 1)
cc = torch.arange(cols, device="cuda").view(1, -1)
edge = (rr == 0) | (cc == 0) | (rr == rows - 1) | (cc == cols - 1)
term = ~((fd > 0) & ~edge)
total = float(fa[term].sum())
print("acc == 1 + sum(donors):", ok, " flow reaching terminal cells:", total, "cells:", n,
      "max acc:", float(fa.max()))
assert ok and total == n
print("config5 band ok")
