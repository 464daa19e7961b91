"""Moving-window filter: device times on an n x n synthetic DEM for a few kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import ctypes as C
import numpy as np, torch
from hydrotools.cuda import _native as N, _device as D, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dem, ld = synth.synth_dem(1, n, n)
out = torch.empty((n, ld), dtype=torch.float32, device="cuda")
st = D.stream_ptr()


def disc(r):
    y, x = np.mgrid[-r:r + 1, -r:r + 1]
    return ((x * x + y * y) <= r * r).astype(np.uint8)


for name, k, method in [("3x3 mean", np.ones((3, 3), np.uint8), 4), ("3x3 max", np.ones((3, 3), np.uint8), 2),
                        ("7x7 std", np.ones((7, 7), np.uint8), 5), ("disc r=10 mean", disc(10), 4),
                        ("disc r=10 std", disc(10), 5), ("33x33 sum", np.ones((33, 33), np.uint8), 0)]:
    k = np.ascontiguousarray(k)

    def run():
        N.call("ht_window_filter_f32", dem.data_ptr(), n, n, ld, 0, 0.0, k.ctypes.data_as(C.c_void_p),
               k.shape[0], k.shape[1], method, out.data_ptr(), ld, 0, 0, st)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:16s} {int(k.sum()):5d} cells  {ms:8.3f} ms  {n * n / ms / 1e6:8.1f} Gcells/s  "
          f"{n * n * 8 / ms / 1e6:7.0f} GB/s (8 B/cell)  {n * n * int(k.sum()) / ms / 1e9:7.1f} Gsamples/ms")


# large discs through the public API (row prefix sums): sum / mean, any radius
import hydrotools.cuda as hc
view = dem[:, :n]
for r in (10, 50, 100):
    k = disc(r).astype(bool)
    for _ in range(2):
        hc.window_filter(view, k, "mean")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        hc.window_filter(view, k, "mean")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"disc r={r:<3d} mean (prefix sums) {int(k.sum()):6d} cells  {ms:8.3f} ms  {n * n / ms / 1e6:8.1f} Gcells/s  "
          f"{n * n * int(k.sum()) / ms / 1e9:8.1f} Gsamples/ms equivalent")


# the same through the C-ABI entry alone (no upload / allocation around it)
ws = torch.empty(int(N.lib.ht_window_filter_runs_workspace_bytes(n, n, 201)), dtype=torch.uint8, device="cuda")
for r in (5, 10, 20):
    k = np.ascontiguousarray(disc(r))

    def run():
        N.call("ht_window_filter_runs_f32", dem.data_ptr(), n, n, ld, 0, 0.0, k.ctypes.data_as(C.c_void_p),
               k.shape[0], k.shape[1], 4, out.data_ptr(), ld, 0, 0, ws.data_ptr(), ws.numel(), st)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"disc r={r:<3d} mean, ht_window_filter_runs_f32 alone {int(k.sum()):6d} cells  {ms:8.3f} ms  "
          f"{n * n / ms / 1e6:8.1f} Gcells/s  {n * n * 8 / ms / 1e6:7.0f} GB/s (8 B/cell)")
