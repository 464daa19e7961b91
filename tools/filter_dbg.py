import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import numpy as np, torch
import hydrotools.cuda as hc
from oracle import rasterfilter as orf
rng = np.random.default_rng(1)
a = rng.normal(size=(300, 400)).astype(np.float32)
a[rng.random(a.shape) < 0.02] = np.nan
yy, xx = np.mgrid[-10:11, -10:11]
k = (xx * xx + yy * yy) <= 100
for meth in ("mean", "sum"):
    g = hc.window_filter(a, k, meth)
    e = orf.raster_filter(a, k, meth)
    m = ~np.isnan(e)
    print(meth, np.array_equal(np.isnan(g), np.isnan(e)), np.max(np.abs(g[m] - e[m]) / np.maximum(np.abs(e[m]), 1e-3)))
