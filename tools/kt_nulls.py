"""Device times of the path on the 10k x 10k synthetic DEM WITH NULL blobs (most D8
tiles take the general path; real DEMs look like this)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import torch
from hydrotools.cuda import distributed as HD
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
eng = HD.BandEngine(0, 1, n, n, n, 10.0, 10.0, nodata=-32767.0)
eng.generate_synthetic(20261019, with_nulls=True)
print("validate:", eng.validate())
for _ in range(3):
    eng.flow_direction_accumulation()
torch.cuda.synchronize()
fd, fa = eng.results()
print("NULL fraction:", float((fd == -128).float().mean()))
for k in eng.time_kernels(10):
    print(f"  {k['name']:28s} {k['ms']:.4f} ms")
