"""Per-stage device time of the weighted accumulation (mode 2) on ROWS x COLS, one GPU:
weight maximum, stage A (tile-local fixed-point sums), stage B (128-bit resolve), stage C
(walkers + fp64 output).  CUDA events on the launching stream.
    python tools/weighted_phases.py [ROWS=10000] [COLS=10000]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import torch
from hydrotools.cuda import distributed as HD, _native as N, _device as D

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
dev = D.require_cuda()
eng = HD.BandEngine(0, 1, rows, cols, rows, 10.0, 10.0)
eng.generate_synthetic(20261019)
eng.flow_direction_accumulation()
band = eng.band
w, wld = D.empty_padded(rows, cols, torch.float32, dev, 4)
N.call("ht_synth_weight_f32", w.data_ptr(), rows, cols, wld, 20261020, 0, D.stream_ptr())
band.weight, band.wld, band.wnd = w, wld, None
M = N.MODE_WEIGHTED


def timed(name, fn, iters=5):
    if os.environ.get("HT_PROF"):
        iters = 1  # under ncu: one warm and one timed launch per stage
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:28s} {ms:8.3f} ms")
    return ms


tot = 0.0
tot += timed("weight max", lambda: band.set_weight_max(band.weight_max()))
tot += timed("stage A acc_local_w", lambda: band.local(M))
tot += timed("stage B resolve<i128>", lambda: band.resolve(False))
tot += timed("stage C acc_final_w", lambda: band.final(M))
print(f"{'sum':28s} {tot:8.3f} ms = {rows * cols / tot / 1e6:.1f} Gcells/s")
for mode, name in ((N.MODE_COUNT, "count"),):
    a = timed(f"{name}: stage A", lambda: band.local(mode))
    b = timed(f"{name}: stage B", lambda: band.resolve(False))
    c = timed(f"{name}: stage C", lambda: band.final(mode))
