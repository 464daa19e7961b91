"""SURVEY.md 8(d) "harder case": natural fBm terrain, conditioned by the oracle's A*
rank on the host (elevation := pop order, mm), then D8 + accumulation on the GPU:
bit-exact against the oracle and timed.  3 600 x 4 000 = 14.4 M cells keeps the rank
inside the supported elevation range (|z| < 16 777 m)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle
from test_watershed_gpu import _fbm
from hydrotools.cuda import distributed as HD

rows, cols = 3600, 4000
t0 = time.time()
dem = oracle.condition_by_rank(_fbm(11, rows, cols), None, 10.0, 10.0)
t1 = time.time()
efd, efa = oracle.rwatershed(dem, None, 10.0, 10.0)
t2 = time.time()
print(f"host: conditioning {t1 - t0:.1f} s, oracle r.watershed {t2 - t1:.1f} s "
      f"({rows * cols / (t2 - t1) / 1e9:.4f} Gcells/s, 1 core)")
eng = HD.BandEngine(0, 1, rows, cols, rows, 10.0, 10.0)
eng.load(dem)
print("validate (pits, ties):", eng.validate())
for _ in range(3):
    eng.flow_direction_accumulation()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.flow_direction_accumulation()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fd, fa = eng.results()
print(f"GPU: {ms:.3f} ms per step = {rows * cols / ms / 1e6:.2f} Gcells/s; stage B rounds {eng.band.resolve_rounds()}")
for k in eng.time_kernels(10):
    print(f"  {k['name']:24s} {k['ms']:.4f} ms")
print("direction mismatches", int((fd.cpu().numpy() != efd).sum()), "accumulation mismatches",
      int((fa.cpu().numpy() != efa).sum()), "max accumulation", float(np.nanmax(efa)))
