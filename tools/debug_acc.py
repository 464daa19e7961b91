"""Development aid: where does the GPU accumulation differ from the oracle?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import numpy as np
import oracle
import hydrotools.cuda as hc
shape = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200, 300)
dem = oracle.synth_dem(shape[0] * 7 + shape[1], *shape)
efd, efa = oracle.rwatershed(dem, None, 10.0, 10.0)
for rep in range(2):
    fd, fa = hc.flow_direction_accumulation_arrays(dem, csx=10.0, csy=10.0)
    bad = np.argwhere(fa != efa)
    print("rep", rep, "fd mismatches", int((fd != efd).sum()), "fa mismatches", len(bad))
    for r, c in bad[:40]:
        print(f"  ({r},{c}) tile ({r//64},{c//64}) local ({r%64},{c%64}) got {fa[r,c]} exp {efa[r,c]} fd {efd[r,c]}")
    if len(bad):
        tiles = sorted({(r // 64, c // 64) for r, c in bad})
        print("  tiles:", tiles[:30])
        d = (fa - efa)[fa != efa]
        print("  diff stats: min", d.min(), "max", d.max())
