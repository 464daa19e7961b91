"""Development aid: device times of the four kernels of the path on the
10k x 10k synthetic DEM + a checksum of the result (compare across builds)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import torch
from hydrotools.cuda import distributed as HD
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
eng = HD.BandEngine(0, 1, n, n, n, 10.0, 10.0)
eng.generate_synthetic(20261019)
for _ in range(3):
    eng.flow_direction_accumulation()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.flow_direction_accumulation()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fd, fa = eng.results()
print(f"step {ms:.4f} ms  {n*n/ms/1e6:.2f} Gcells/s  checksum fa {float(fa.sum()):.0f} max {float(fa.max()):.0f} fd {int(fd.to(torch.int64).sum())}")
for k in eng.time_kernels(10):
    print(f"  {k['name']:28s} {k['ms']:.4f} ms")
