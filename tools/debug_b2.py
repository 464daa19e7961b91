"""Development aid: flat vs two-level stage B on a banded run (run twice with HT_STAGE_B=flat / two,
then `cmp`): dumps per band the tables and the final accumulation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle
from test_watershed_gpu import _fbm
from hydrotools.cuda import distributed as HD, _native as N

if sys.argv[1] == "cmp":
    a, b = np.load("/tmp/b2_flat.npz"), np.load("/tmp/b2_two.npz")
    for k in a.files:
        x, y = a[k], b[k]
        same = np.array_equal(x, y, equal_nan=True) if x.dtype.kind == "f" else np.array_equal(x, y)
        msg = ""
        if not same:
            d = np.argwhere(~((x == y) | ((x != x) & (y != y)))) if x.dtype.kind == "f" else np.argwhere(x != y)
            msg = f" {len(d)} differ, first {d[:5].tolist()} flat {x[tuple(d[0])]} two {y[tuple(d[0])]}"
        print(k, "same" if same else "DIFFERENT" + msg)
    sys.exit(0)
raw = _fbm(7, 1100, 1400)
dem = oracle.condition_by_rank(raw, None, 10.0, 10.0)
cuts = [0, 300, 640, 1100]
P = len(cuts) - 1
bands = []
for b in range(P):
    lo, hi = cuts[b], cuts[b + 1]
    band = HD.CudaBand(hi - lo, dem.shape[1], lo, dem.shape[0], 10.0, 10.0, None, has_above=b > 0, has_below=b < P - 1)
    band.load(dem[lo:hi])
    bands.append(band)
for b in range(P):
    if b > 0:
        bands[b].set_halo_rows(top=True, rows=bands[b - 1].edge_rows(top=False, depth=2))
    if b < P - 1:
        bands[b].set_halo_rows(top=False, rows=bands[b + 1].edge_rows(top=True, depth=2))
out = {}
for b, band in enumerate(bands):
    band.direction(); band.local(0); band.mark_edges(0); band.resolve(True)
    d, l = band.band_tables(0)
    out[f"delivered{b}"] = d.cpu().numpy(); out[f"link{b}"] = l.cpu().numpy()
    tiles = ((band.rows + 63) // 64) * ((band.cols + 63) // 64)
    agg = band._ws_view(2, torch.int64)
    out[f"agg{b}"] = agg.cpu().numpy()
np.savez(f"/tmp/b2_{sys.argv[1]}.npz", **out)
print("saved", sys.argv[1])
