"""Device-side tile encoder (ht_tiff_encode_tiles): kernel time and compression ratio on the
outputs of the bench raster (direction int8, accumulation fp64) and on the DEM (fp32).
    python tools/cog_bench.py [N=10000]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import torch
import hydrotools.cuda as hc
from hydrotools.cuda import synth, _native as N, _device as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dem, ld = synth.synth_dem(20261019, n, n)
fd, fa = hc.flow_direction_accumulation_arrays(dem[:, :n], csx=10.0, csy=10.0)
for name, a in (("direction int8", fd), ("accumulation fp64", fa), ("DEM fp32", dem[:, :n])):
    rows, cols, es = a.shape[0], a.shape[1], a.element_size()
    ntiles = -(-rows // 512) * -(-cols // 512)
    bound = int(N.lib.ht_tiff_encode_bound(rows, cols, es))
    out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    ws = torch.empty(int(N.lib.ht_tiff_encode_workspace_bytes(rows, cols, es)), dtype=torch.uint8, device="cuda")
    meta = torch.empty(2 * ntiles + 1, dtype=torch.int64, device="cuda")

    def run():
        N.call("ht_tiff_encode_tiles", a.data_ptr(), int(a.stride(0)) * es, rows, cols, es, 0,
               out.data_ptr(), bound, meta.data_ptr(), meta.data_ptr() + 8 * ntiles,
               meta.data_ptr() + 16 * ntiles, ws.data_ptr(), ws.numel(), D.stream_ptr())
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    total = int(meta[2 * ntiles].item())
    raw = rows * cols * es
    print(f"{name:18s} {raw / 1e6:8.1f} MB -> {total / 1e6:8.1f} MB ({raw / total:5.2f} : 1) in {ms:7.2f} ms "
          f"= {raw / ms / 1e6:6.1f} GB/s")
