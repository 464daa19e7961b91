"""Development aid: FlowPipeline throughput on the 10 000 x 10 000 synthetic DEM for a few depths."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import torch
import hydrotools.cuda as hc
from hydrotools.cuda import synth

n = 10000
dem, ld = synth.synth_dem(20261019, n, n)
host = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
host.copy_(dem[:, :n]); torch.cuda.synchronize()
del dem
for depth in (2, 3, 4):
    pipe = hc.FlowPipeline(n, n, nodata=None, csx=10.0, csy=10.0, depth=depth)
    for _ in pipe.run([host] * 4):
        pass
    torch.cuda.synchronize()
    k = 12
    t0 = time.perf_counter()
    for _ in pipe.run([host] * k):
        pass
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / k
    print(f"depth {depth}: {dt * 1e3:.2f} ms per raster = {n * n / dt / 1e9:.2f} Gcells/s", flush=True)
    del pipe
    torch.cuda.empty_cache()
