"""Fused terrain kernel on an n x n synthetic DEM, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import torch
from hydrotools.cuda import _native as N, _device as D, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nodata = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dem, ld = synth.synth_dem(1, n, n, with_nulls=bool(nodata))
outs = [torch.empty((n, ld), dtype=torch.float32, device="cuda") for _ in range(3)]
for _ in range(3):
    N.call("ht_terrain_fused_f32", dem.data_ptr(), n, n, ld, nodata, -32767.0, 10.0, 10.0, 1.0, 0,
           outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), ld, 0, 0, D.stream_ptr())
torch.cuda.synchronize()
print("done")
