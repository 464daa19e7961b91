"""Quick device-resident timings of individual kernels (development aid; the
judged numbers come from bench.py)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
from hydrotools.cuda import _native as N, _device as D, synth  # noqa: E402

PEAK = 6554.9


def time_cuda(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def terrain(n, nodata):
    dem, ld = synth.synth_dem(1, n, n, with_nulls=nodata)
    outs = [torch.empty((n, ld), dtype=torch.float32, device="cuda") for _ in range(3)]
    st = D.stream_ptr()

    def run():
        N.call("ht_terrain_fused_f32", dem.data_ptr(), n, n, ld, int(nodata), -32767.0, 10.0,
               10.0, 1.0, 0, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), ld,
               0, 0, st)
    ms = time_cuda(run)
    gbs = n * n * 16 / ms / 1e6
    print(json.dumps(dict(kernel="terrain_fused", n=n, nodata=nodata, ms=round(ms, 4),
                          gcells_s=round(n * n / ms / 1e6, 2), gbs=round(gbs, 1),
                          frac=round(gbs / PEAK, 3))))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "terrain"
    if which == "terrain":
        for n in (8192, 16384, 32768):
            terrain(n, False)
        terrain(16384, True)
