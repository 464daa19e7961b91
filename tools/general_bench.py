"""General D8 path (ht_d8_general_f32) on RAW natural terrain: spectral fBm rounded to
centimetres (ties, flats after rounding, thousands of depressions).  Times the direction
kernel chain and the accumulation, prints the round / episode statistics, and compares with
the oracle when the raster is small enough for it.
    python tools/general_bench.py [N=4000] [decimals=2] [relief_m=800]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import numpy as np, torch
import hydrotools.cuda as hc
from hydrotools.cuda import watershed as W, _device as D, _native as N

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
dec = int(sys.argv[2]) if len(sys.argv) > 2 else 2
relief = float(sys.argv[3]) if len(sys.argv) > 3 else 800.0
kind = sys.argv[4] if len(sys.argv) > 4 else "fbm"
rng = np.random.default_rng(1)
dev = D.require_cuda()
if kind == "fbm":
    # isotropic spectral noise: no drainage structure at all, a third of the cells in lakes
    fy, fx = np.fft.fftfreq(n)[:, None], np.fft.fftfreq(n)[None, :]
    f = np.sqrt(fx * fx + fy * fy); f[0, 0] = 1
    z = np.fft.ifft2((rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))) / f ** 1.8).real
    dem = ((z - z.min()) / (z.max() - z.min()) * relief).round(dec).astype(np.float32)
    del z, f
else:
    # fluvial terrain (the bench's valley network) + measurement noise of `relief` metres s.d.,
    # rounded to `dec` decimals: what a raw LiDAR / SRTM tile looks like -- many small pits, ties
    from hydrotools.cuda import synth
    t, tld = synth.synth_dem(20261019, n, n)
    dem = t[:, :n].cpu().numpy()
    del t
    for r0 in range(0, n, 1000):
        dem[r0:r0 + 1000] = (dem[r0:r0 + 1000] + rng.normal(0.0, relief, dem[r0:r0 + 1000].shape)
                             ).round(dec).astype(np.float32)
buf, ld = D.upload(dem, torch.float32, 4, dev)
pits, ties = W._validate(buf, ld, n, n, None, dev)
print(f"{n} x {n} raw {kind}, {dec} decimals: {pits} pits, {ties} tie cells")
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    packed, pld, info = W.d8_general(buf, ld, n, n, None, 10.0, 10.0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    acc, ald, dout, dld = W.accumulate(packed, pld, n, n, N.MODE_COUNT, want_dir=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"direction (general) {1e3 * (t1 - t0):.1f} ms = {n * n / (t1 - t0) / 1e9:.3f} Gcells/s; "
      f"accumulation {1e3 * (t2 - t1):.2f} ms; stats {info}")
if n <= 4000:
    import oracle
    t0 = time.perf_counter()
    efd, efa = oracle.rwatershed(dem, None, 10.0, 10.0)
    t1 = time.perf_counter()
    fd = dout[:, :n].cpu().numpy(); fa = acc[:, :n].cpu().numpy()
    print(f"oracle {t1 - t0:.1f} s (1 core); direction mismatches {(fd != efd).sum()}, "
          f"accumulation mismatches {(fa != efa).sum()}")
