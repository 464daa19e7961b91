"""oracle.streams -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-Python restatements of the reference's numba walkers over the direction grid
(SURVEY.md 8(f) N2), each checked against tests/golden/streams.npz -- outputs of the
reference's own code, cut out with `ast` and executed by tools/make_stream_goldens.py:
  * reach_labels   `_fa_label_task`     /root/reference/src/hydrotools/streams.py:320-388
  * upstream_mean  `_upstream_compute`  /root/reference/src/hydrotools/morphology.py:50-153
and of GRASS r.water.outlet as the reference calls it
(/root/reference/src/hydrotools/watershed.py:292-318; not vendored: PARITY UNPINNED for it):
  * basin_mask     every cell whose path along the codes passes through the outlet.
Loops in Python: small rasters only.
"""
import numpy as np

DR = (0, -1, -1, -1, 0, 1, 1, 1, 0)
DC = (0, 1, 0, -1, -1, -1, 0, 1, 1)


def _recv(fd, i, j):
    d = int(fd[i, j])
    if d < 1 or d > 8:
        return None
    r, c = i + DR[d], j + DC[d]
    if r < 0 or r >= fd.shape[0] or c < 0 or c >= fd.shape[1]:
        return None
    return r, c


def reach_labels(reaches, fd, nodata=0):
    """Connected runs of one reach value along direction links (downstream and upstream),
    numbered in row-major order of discovery; out-of-raster neighbours do not exist (the
    numba code does not bounds-check; its callers keep a NULL frame)."""
    rows, cols = reaches.shape
    labels = np.zeros(reaches.shape, np.uint32)
    cur = 0
    for i in range(rows):
        for j in range(cols):
            if reaches[i, j] == nodata or labels[i, j]:
                continue
            cur += 1
            labels[i, j] = cur
            val = reaches[i, j]
            stack = [(i, j)]
            while stack:
                ci, cj = stack.pop()
                t = _recv(fd, ci, cj)
                if t is not None and labels[t] == 0 and reaches[t] == val:
                    labels[t] = cur
                    stack.append(t)
                for di in (-1, 0, 1):
                    for dj in (-1, 0, 1):
                        ni, nj = ci + di, cj + dj
                        if (di or dj) and 0 <= ni < rows and 0 <= nj < cols and \
                                reaches[ni, nj] == val and labels[ni, nj] == 0 and _recv(fd, ni, nj) == (ci, cj):
                            labels[ni, nj] = cur
                            stack.append((ni, nj))
    return labels


def upstream_mean(values, streams, fd, csx, csy, distance, fill=np.nan):
    rows, cols = values.shape
    out = np.full(values.shape, fill, np.float32)
    diag = (csx ** 2 + csy ** 2) ** 0.5
    for i in range(rows):
        for j in range(cols):
            if not streams[i, j]:
                continue
            inlet = True
            for di in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    ni, nj = i + di, j + dj
                    if (di or dj) and 0 <= ni < rows and 0 <= nj < cols and streams[ni, nj] \
                            and _recv(fd, ni, nj) == (i, j):
                        inlet = False
            if not inlet:
                continue
            vals, dists = [float(values[i, j])], [0.0]
            head, cum, total = 0, 0.0, float(values[i, j])
            ci, cj = i, j
            while True:
                t = _recv(fd, ci, cj)
                if t is None or not streams[t]:
                    break
                d = int(fd[ci, cj])
                cum += diag if (DR[d] and DC[d]) else (csx if DC[d] else csy)
                vals.append(float(values[t]))
                dists.append(cum)
                total += float(values[t])
                while cum - dists[head] > distance:
                    total -= vals[head]
                    head += 1
                out[t] = total / (len(vals) - head)
                ci, cj = t
    return out


def basin_mask(fd, row, col):
    rows, cols = fd.shape
    out = np.zeros(fd.shape, np.uint8)
    if fd[row, col] == -128:
        return out
    for i in range(rows):
        for j in range(cols):
            if fd[i, j] == -128:
                continue
            ci, cj, steps = i, j, 0
            while steps <= rows * cols:
                if (ci, cj) == (row, col):
                    out[i, j] = 1
                    break
                t = _recv(fd, ci, cj)
                if t is None or fd[t] == -128:
                    break
                ci, cj = t
                steps += 1
    return out
