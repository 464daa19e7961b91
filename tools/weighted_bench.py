"""BASELINE.json configs[3] per-GPU shape on one GPU: precipitation-weighted flow
accumulation (fp64 sums of fp32 weights) over the D8 graph of a 5 000 x 40 000
band; device-resident timing and a mass-balance check."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import torch
from hydrotools.cuda import distributed as HD, _native as N, _device as D

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
eng = HD.BandEngine(0, 1, rows, cols, rows, 10.0, 10.0)
eng.generate_synthetic(20261019)
band = eng.band
w, wld = D.empty_padded(rows, cols, torch.float32, band.dev, 4)
N.call("ht_synth_weight_f32", w.data_ptr(), rows, cols, wld, 20261020, 0, D.stream_ptr())
band.weight, band.wld, band.wnd = w, wld, None
band.direction()
for _ in range(2):
    eng.accumulate(N.MODE_WEIGHTED)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.accumulate(N.MODE_WEIGHTED)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
acc = band.acc[:, :cols]
total_w = float(w[:, :cols].double().sum())
print(f"weighted accumulation {rows} x {cols}: {ms:.3f} ms, {rows * cols / ms / 1e6:.2f} Gcells/s")
# all weight ends in cells that do not pass flow on: compare with the count run's structure
eng.accumulate(N.MODE_COUNT)
cnt = band.acc[:, :cols].clone()
eng.accumulate(N.MODE_WEIGHTED)
print("max weighted / count ratio:", float((band.acc[:, :cols] / cnt).max()),
      "(weights are 1650..1700)", " min:", float((band.acc[:, :cols] / cnt).min()))
