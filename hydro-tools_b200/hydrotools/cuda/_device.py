"""Device-buffer plumbing (torch is used for HBM allocation, streams and
pinned host memory only; all arithmetic is in the sm_100a kernels)."""
from typing import Optional, Tuple, Union

import numpy as np
import torch

ArrayLike = Union[np.ndarray, torch.Tensor]


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "hydrotools.cuda needs an NVIDIA B200 (sm_100a); CUDA is not available "
            "and there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def empty_padded(rows: int, cols: int, dtype: torch.dtype, device, align_elems: int,
                 extra_top: int = 0, extra_bot: int = 0) -> Tuple[torch.Tensor, int]:
    """(rows+extra) x ld buffer whose pitch is a multiple of `align_elems`
    elements; returns (full buffer, ld)."""
    ld = max(round_up(cols, align_elems), align_elems)
    buf = torch.empty((rows + extra_top + extra_bot, ld), dtype=dtype, device=device)
    return buf, ld


def as_2d(a: ArrayLike) -> ArrayLike:
    """(rows, cols) view of a raster.  Lazy arrays (``dask.array`` as returned by
    ``hydrotools.raster.from_raster``, /root/reference/src/hydrotools/raster.py:773-791)
    are materialised here: the kernels want the whole band in HBM, and masked values
    are recognised by the nodata value / NaN as in the reference (raster.py:788)."""
    if not isinstance(a, (np.ndarray, torch.Tensor)) and hasattr(a, "compute"):
        a = a.compute()
    if isinstance(a, np.ma.MaskedArray):
        a = np.ma.getdata(a)
    if a.ndim == 3 and a.shape[0] == 1:
        a = a[0]
    if a.ndim != 2:
        raise ValueError(f"expected a (rows, cols) or (1, rows, cols) raster, got {tuple(a.shape)}")
    return a


def upload(a: ArrayLike, dtype: torch.dtype, align_elems: int, device,
           extra_top: int = 0, extra_bot: int = 0) -> Tuple[torch.Tensor, int]:
    """Copy a host (numpy) or device (torch) raster into a pitch-aligned HBM
    buffer.  Host arrays go through pinned memory, asynchronously."""
    a = as_2d(a)
    rows, cols = int(a.shape[0]), int(a.shape[1])
    buf, ld = empty_padded(rows, cols, dtype, device, align_elems, extra_top, extra_bot)
    view = buf[extra_top:extra_top + rows, :cols]
    if isinstance(a, torch.Tensor):
        view.copy_(a.to(dtype=dtype), non_blocking=True)
    else:
        arr = np.ascontiguousarray(np.ma.getdata(a))
        t = torch.from_numpy(arr)
        if t.dtype != dtype:
            t = t.to(dtype)
        if not t.is_pinned():
            # pageable host memory: stage through a cached pinned buffer.  The buffer is
            # reused by the next upload of the same dtype, so wait until the DMA of the
            # previous one has drained before the CPU overwrites it.
            key = ("upload", t.dtype)
            pending = _STAGE_BUSY.pop(key, None)
            if pending is not None:
                pending.synchronize()
            stage = _pinned(t.shape, t.dtype, "upload")
            stage.copy_(t)
            view.copy_(stage, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
            _STAGE_BUSY[key] = done
        else:
            view.copy_(t, non_blocking=True)
    return buf, ld


_PINNED = {}
_STAGE_BUSY = {}  # (tag, dtype) -> CUDA event recorded after the last H2D copy out of the stage


def _pinned(shape, dtype, tag: str) -> torch.Tensor:
    """Reusable pinned staging buffer (cudaHostAlloc is far slower than the copy)."""
    n = 1
    for d in shape:
        n *= int(d)
    key = (tag, dtype)
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n, 1), dtype=dtype, pin_memory=True)
        _PINNED[key] = buf
    return buf[:n].view(*shape)


def download(view: torch.Tensor, out: Optional[np.ndarray] = None) -> np.ndarray:
    """Device -> host (synchronises the current stream).  With ``out`` (a C-contiguous
    array of the right shape/dtype, ideally page-locked) the copy lands there
    directly; otherwise it goes through a cached pinned buffer into a new array."""
    if out is not None:
        t = torch.from_numpy(out)
        if tuple(t.shape) != tuple(view.shape) or t.dtype != view.dtype or not t.is_contiguous():
            raise ValueError("`out` has the wrong shape, dtype or layout")
        if t.is_pinned():
            t.copy_(view, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out
        stage = _pinned(view.shape, view.dtype, "download")
        stage.copy_(view, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        t.copy_(stage)
        return out
    stage = _pinned(view.shape, view.dtype, "download")
    stage.copy_(view, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return stage.numpy().copy()


def nodata_args(nodata: Optional[float]) -> Tuple[int, float]:
    return (0, 0.0) if nodata is None else (1, float(nodata))


def bind_to_gpu_numa_node(index: Optional[int] = None) -> dict:
    """Run this process (and place the pinned host buffers it allocates from now on) on the
    NUMA node the GPU hangs off.  With one process per GPU on a two-socket box, staging
    buffers that all sit on node 0 make half of the GPUs copy across the socket link and
    all of them share one memory controller.  Best effort: needs sysfs and permission to
    change affinity / memory policy; returns what was done."""
    import ctypes
    import os
    info = {"node": None, "cpus_bound": False, "mempolicy": False}
    try:
        idx = torch.cuda.current_device() if index is None else index
        import pynvml
        pynvml.nvmlInit()
        # NVML counts physical devices; CUDA_VISIBLE_DEVICES may renumber
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[idx]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else idx
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return info
        info["node"] = node
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        usable = cpus & os.sched_getaffinity(0)
        if usable:
            os.sched_setaffinity(0, usable)
            info["cpus_bound"] = True
        # MPOL_PREFERRED (1): allocate on `node` when possible (syscall 238 = set_mempolicy)
        mask = ctypes.c_ulong(1 << node)
        libc = ctypes.CDLL(None, use_errno=True)
        if libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask))) == 0:
            info["mempolicy"] = True
    except Exception as exc:  # no sysfs / NVML / permission: leave placement to the OS
        info["error"] = f"{type(exc).__name__}: {exc}"[:120]
    return info
