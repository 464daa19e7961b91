"""CPU: the two-level boundary-graph resolve (csrc/ht_acc.cu: super_tile_kernel<1>,
the flat resolve on the OUTER nodes, super_tile_kernel<3>) restated in numpy on random
forests -- internal sums S1, promotion of large / marked nodes, level-2 graph
(base2, next2), external inflow X walked down the internal paths -- equals the plain
subtree sums of the flat resolve exactly.  The kernels themselves are compared with the
oracle in the -m gpu tests (HT_STAGE_B=two forces them on small rasters)."""
import numpy as np
import pytest

MARK = 1 << 40
BIG = 1 << 20


def flat_sums(base, nxt):
    """agg[v] = base[v] + sum of agg over the nodes whose next is v (a forest)."""
    n = len(base)
    agg = base.astype(object).copy()
    indeg = np.zeros(n, int)
    for v in range(n):
        if nxt[v] >= 0:
            indeg[nxt[v]] += 1
    stack = [v for v in range(n) if indeg[v] == 0]
    while stack:
        v = stack.pop()
        t = nxt[v]
        if t >= 0:
            agg[t] += agg[v]
            indeg[t] -= 1
            if indeg[t] == 0:
                stack.append(t)
    return agg


def two_level(base, nxt, group):
    n = len(base)
    internal = np.array([nxt[v] >= 0 and group[nxt[v]] == group[v] and group[v] >= 0 for v in range(n)])
    # K1: internal forest, promoted nodes contribute nothing to S1
    promoted = np.array([int(base[v]) >= BIG and group[v] >= 0 for v in range(n)])
    b1 = np.array([0 if promoted[v] else int(base[v]) for v in range(n)], dtype=object)
    nint = np.where(internal, nxt, -1)
    s1 = flat_sums(b1, nint)
    root1 = np.arange(n)
    for v in range(n):
        r = v
        while nint[r] >= 0:
            r = nint[r]
        root1[v] = r
    base2 = np.zeros(n, dtype=object)
    outer = np.zeros(n, bool)
    for v in range(n):
        if group[v] < 0:  # a neighbour band's cell: a level-2 node with its own value
            if int(base[v]):
                base2[v] += int(base[v]); outer[v] = True
            continue
        if promoted[v]:
            base2[v] += int(base[v]); outer[v] = True
        if root1[v] == v and nxt[v] >= 0:  # end of an internal path with a link out
            base2[nxt[v]] += s1[v]; outer[nxt[v]] = True
    assert all(int(s1[v]) < 2 ** 32 for v in range(n) if group[v] >= 0), "K1 works on 32-bit words"
    # K2: flat resolve on the outer nodes
    next2 = np.full(n, -1)
    for w in np.nonzero(outer)[0]:
        next2[w] = nxt[root1[w]] if group[w] >= 0 else -1
        assert next2[w] < 0 or outer[next2[w]]
    x = np.zeros(n, dtype=object)
    ids = np.nonzero(outer)[0]
    sub = flat_sums(base2[ids], np.array([np.searchsorted(ids, next2[w]) if next2[w] >= 0 else -1 for w in ids]))
    x[ids] = sub
    # K3: every outer node walks its X down its internal path
    agg = s1.copy()
    for w in ids:
        if group[w] < 0:
            agg[w] = x[w]
            continue
        cur = w
        while cur >= 0:
            agg[cur] += x[w]
            cur = nint[cur]
    return agg


@pytest.mark.parametrize("seed", range(6))
def test_two_level_equals_flat(seed):
    rng = np.random.default_rng(seed)
    n, groups, virtual = 600, 7, 40
    # a random forest: every node points at a later node (or ends), so there is no cycle
    nxt = np.array([rng.integers(v + 1, n) if (v + 1 < n and rng.random() < 0.9) else -1 for v in range(n)])
    group = rng.integers(0, groups, n)
    group[n - virtual:] = -1            # neighbour-band cells: no tile, no super-tile
    nxt[n - virtual:] = -1              # ... and they end every path
    base = rng.integers(0, 5000, n).astype(object)
    big = rng.random(n) < 0.05
    base[big] = [int(v) for v in rng.integers(BIG, 1 << 45, big.sum())]  # imported flows
    marked = rng.random(n) < 0.1
    for v in np.nonzero(marked)[0]:
        base[v] = int(base[v]) + MARK       # band-edge marks ride in bits 40..
    a, b = flat_sums(base, nxt), two_level(base, nxt, group)
    assert all(int(p) == int(q) for p, q in zip(a, b))
