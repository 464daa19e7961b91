import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hydro-tools_b200"))
import oracle
from hydrotools.cuda import _native as N, _device as D, watershed as W
dem = oracle.synth_dem(3, 100, 150)
dev = D.require_cuda()
buf, ld = D.upload(dem, torch.float32, 4, dev)
def step(name, fn):
    try:
        r = fn(); torch.cuda.synchronize(); print("OK", name, flush=True); return r
    except Exception as e:
        print("FAIL", name, str(e).splitlines()[0], flush=True); sys.exit(1)
step("validate", lambda: W._validate(buf, ld, 100, 150, None, dev))
packed, pld = step("d8", lambda: W.d8(buf, ld, 100, 150, None, 10.0, 10.0))
ws = W._workspace(100, 150, dev)
acc, ald = D.empty_padded(100, 150, torch.float64, dev, 2)
dout, dld = D.empty_padded(100, 150, torch.int8, dev, 16)
st = D.stream_ptr()
step("local", lambda: N.call("ht_acc_local", packed.data_ptr(), pld, 100, 150, 0, 100, 0, 0, 0, None, 0, 0, 0.0, ws.data_ptr(), ws.numel(), st))
step("resolve", lambda: N.call("ht_acc_resolve", 100, 150, 0, 0, ws.data_ptr(), ws.numel(), None, st))
step("final", lambda: N.call("ht_acc_final", packed.data_ptr(), pld, 100, 150, 0, 100, 0, 0, 0, None, 0, 0, 0.0, acc.data_ptr(), ald, dout.data_ptr(), dld, ws.data_ptr(), ws.numel(), st))
efd, efa = oracle.rwatershed(dem, None, 10.0, 10.0)
fd = dout[:, :150].cpu().numpy(); fa = acc[:, :150].cpu().numpy()
print("dir mismatches", (fd != efd).sum(), "acc mismatches", (fa != efa).sum())
