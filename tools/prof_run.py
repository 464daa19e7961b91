"""One D8 + accumulation step on the 10k x 10k synthetic DEM, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import torch
from hydrotools.cuda import distributed as HD
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = HD.BandEngine(0, 1, n, n, n, 10.0, 10.0)
eng.generate_synthetic(20261019)
for _ in range(steps):
    eng.flow_direction_accumulation()
torch.cuda.synchronize()
print("done", float(eng.results()[1].max()))
