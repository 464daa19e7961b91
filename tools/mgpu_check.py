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
