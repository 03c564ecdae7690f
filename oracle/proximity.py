"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's proximity edge selection, row F1 of SURVEY.md section 8(f):
`FactorGraph.add_proximity_factors` (droid_slam/factor_graph.py:346-412), everything between the `video.distance` call and the final
`add_factors`.  Pinned against the unmodified reference method run on a stub graph object
(tests/golden/make_proximity_golden.py -> tests/golden/proximity.pt, tests/test_proximity_cpu.py).

The reference works on a flat distance vector d over the (i, j) grid i in [t0, t), j in [t1, t), row-major, and

  1. masks pairs with i - rad < j and distances > 100                                        (factor_graph.py:359-360)
  2. suppresses the neighbourhood of every edge already in the graph (active, bad, inactive)   (:362-373)
  3. emits the temporal-neighbour edges (i, j), (j, i) for i-rad-1 <= j < i (and (i, i) for stereo) and masks them   (:375-384)
  4. walks the remaining pairs by increasing distance, takes every pair that is still <= thresh, emits it in both directions and
     suppresses its neighbourhood (non-maximum suppression with an |di| + |dj| <= min(|i-j| - 2, nms) diamond)        (:386-409)

Quirks kept: step 3 indexes d without a bounds check, so j < t1 lands in the previous row or (negative index) wraps to the end of d
exactly like the tensor indexing it restates; `len(es) > max_factors` is tested before a pair is taken, so the list may overshoot;
NaN distances are not > thresh and are therefore taken, last (argsort puts them at the end).  `torch.argsort` is not stable: equal
finite distances may be visited in either order by the reference; this restatement (and the CUDA kernel) break ties by flat index.
"""
import numpy as np


def _suppress(d, i, j, t0, t1, t, nms):
    """factor_graph.py:365-373 / :401-409 -- the diamond around (i, j)"""
    w = max(min(abs(int(i) - int(j)) - 2, nms), 0)
    for di in range(-nms, nms + 1):
        for dj in range(-nms, nms + 1):
            if abs(di) + abs(dj) <= w:
                i1, j1 = int(i) + di, int(j) + dj
                if t0 <= i1 < t and t1 <= j1 < t:
                    d[(i1 - t0) * (t - t1) + (j1 - t1)] = np.inf


def proximity_edges(d, t0, t1, t, ii1, jj1, rad=2, nms=2, thresh=16.0, max_factors=-1, stereo=False):
    """d: float32 array-like of length (t-t0)*(t-t1) as returned by `video.distance` over the meshgrid of (:351-356); ii1/jj1: the
    edges already known to the graph.  Returns (es [n,2] int64 in emission order, d after all masking -- for inspection)."""
    d = np.array(d, dtype=np.float32).reshape(-1).copy()
    n_i, n_j = t - t0, t - t1
    assert d.shape[0] == n_i * n_j
    gi = np.repeat(np.arange(t0, t), n_j)
    gj = np.tile(np.arange(t1, t), n_i)
    d[gi - rad < gj] = np.inf
    d[d > 100] = np.inf
    for i, j in zip(np.asarray(ii1).tolist(), np.asarray(jj1).tolist()):
        _suppress(d, i, j, t0, t1, t, nms)
    es = []
    for i in range(t0, t):
        if stereo:
            es.append((i, i))
            d[(i - t0) * n_j + (i - t1)] = np.inf
        for j in range(max(i - rad - 1, 0), i):
            es.append((i, j))
            es.append((j, i))
            d[(i - t0) * n_j + (j - t1)] = np.inf          # unchecked index, like the reference (negative values wrap)
    key = np.where(np.isnan(d), np.inf, d)
    order = np.lexsort((np.arange(d.shape[0]), np.isnan(d), key))   # by distance, NaN last, ties by flat index
    for k in order.tolist():
        if d[k] > thresh:
            continue
        if max_factors > 0 and len(es) > max_factors:
            break
        i, j = int(gi[k]), int(gj[k])
        es.append((i, j))
        es.append((j, i))
        _suppress(d, i, j, t0, t1, t, nms)
    return np.asarray(es, dtype=np.int64).reshape(-1, 2), d
