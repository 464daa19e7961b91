"""Multi-GPU correctness check (run under torchrun): row-band sharded D8 +
accumulation over NCCL against the oracle on the whole raster."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import numpy as np, torch, torch.distributed as dist
import oracle
from hydrotools.cuda import distributed as HD

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
rows, cols = 700, 1500
total = rows * world
for nulls in (False, True):
    eng = HD.BandEngine(rank, world, rows, cols, total, 10.0, 10.0, -32767.0 if nulls else None)
    eng.generate_synthetic(5, with_nulls=nulls)
    assert eng.validate() == (0, 0)
    eng.flow_direction_accumulation()
    fd, fa = eng.results()
    gfd = [torch.empty_like(fd.contiguous()) for _ in range(world)]
    gfa = [torch.empty_like(fa.contiguous()) for _ in range(world)]
    dist.all_gather(gfd, fd.contiguous()); dist.all_gather(gfa, fa.contiguous())
    if rank == 0:
        dem = oracle.synth_dem(5, total, cols, with_nulls=nulls)
        efd, efa = oracle.rwatershed(dem, -32767.0 if nulls else None, 10.0, 10.0)
        fd_all = torch.cat(gfd).cpu().numpy(); fa_all = torch.cat(gfa).cpu().numpy()
        print(f"world={world} nulls={nulls} dir mismatches {(fd_all != efd).sum()} acc mismatches "
              f"{(~((fa_all == efa) | (np.isnan(fa_all) & np.isnan(efa)))).sum()} max acc {np.nanmax(efa)}")
    # signed accumulation (r.watershed without -a) and weighted accumulation over bands
    eng.flow_direction_accumulation(positive_only=False)
    _, fs = eng.results()
    gfs = [torch.empty_like(fs.contiguous()) for _ in range(world)]
    dist.all_gather(gfs, fs.contiguous())
    w_all = oracle.synth_weight(5, total, cols)
    wsum = eng.flow_accumulation(torch.from_numpy(w_all[rank * rows:(rank + 1) * rows]).cuda()).clone()
    gw = [torch.empty_like(wsum) for _ in range(world)]
    dist.all_gather(gw, wsum.contiguous())
    if rank == 0:
        _, esigned = oracle.rwatershed(dem, -32767.0 if nulls else None, 10.0, 10.0, positive_only=False)
        signed = torch.cat(gfs).cpu().numpy()
        same = (signed == esigned) | (np.isnan(signed) & np.isnan(esigned))
        # weighted: r.watershed's graph (edge cells do not pass flow on), as in tests/test_bands_gpu.py
        null = np.isnan(efa)
        edge = np.zeros(dem.shape, bool); edge[[0, -1], :] = True; edge[:, [0, -1]] = True
        pad = np.pad(null, 1, constant_values=False)
        for a in (0, 1, 2):
            for b in (0, 1, 2):
                edge |= pad[a:a + dem.shape[0], b:b + dem.shape[1]]
        d = efd.copy(); d[edge & ~null] = 0
        ew = oracle.flowacc(d, w_all)
        got = torch.cat(gw).cpu().numpy()
        m = ~np.isnan(ew)
        print(f"  signed accumulation mismatches {(~same).sum()}  weighted max rel err "
              f"{np.max(np.abs(got[m] - ew[m]) / np.maximum(np.abs(ew[m]), 1e-30)):.2e}")
    # moving-window filter over bands (direct kernel and row prefix sums)
    from oracle import rasterfilter as orf
    for kern, meth in (((5, 7), "std"), ((4, 4), "max"), ("disc", "mean")):
        if kern == "disc":
            yy, xx = np.mgrid[-12:13, -12:13]
            kern = (xx * xx + yy * yy) <= 144
        fr = eng.window_filter(kern, meth)
        gfr = [torch.empty_like(fr.contiguous()) for _ in range(world)]
        dist.all_gather(gfr, fr.contiguous())
        if rank == 0:
            a = dem.astype(np.float32).copy()
            if nulls:
                a[a == -32767.0] = np.nan
            e = orf.raster_filter(a, kern, meth)
            g = torch.cat(gfr).cpu().numpy()
            m = ~np.isnan(e)
            print(f"  filter {meth}: mask equal {np.array_equal(np.isnan(g), np.isnan(e))}, max rel err "
                  f"{np.max(np.abs(g[m] - e[m]) / np.maximum(np.abs(e[m]), 1e-3)):.1e}")
    # terrain over bands
    s, a, t = eng.terrain()
    gs = [torch.empty_like(s.contiguous()) for _ in range(world)]
    dist.all_gather(gs, s.contiguous())
    if rank == 0:
        es, _, _ = oracle.gdaldem(dem, -32767.0 if nulls else None, 10.0, 10.0)
        got = torch.cat(gs).cpu().numpy()
        ok = es != -9999
        print("  terrain slope max abs err (degrees)", np.abs(got[ok] - es[ok]).max(), "mask equal", np.array_equal(got == -9999, ~ok))
dist.barrier()
dist.destroy_process_group()
