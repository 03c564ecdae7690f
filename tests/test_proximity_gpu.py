"""Row F1 on the GPU: droid_backends.proximity_edges (csrc/proximity.cu, through the C ABI) against the edge lists the UNMODIFIED reference
method emitted (tests/golden/proximity.pt) and against the oracle on larger random grids -- bit-exact, order included."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle.proximity as prox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_proximity_golden as mk  # noqa: E402

pytestmark = pytest.mark.gpu
dev = "cuda"


@pytest.fixture(scope="module")
def backends():
    from droid_slam_b200 import install
    return install()


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "proximity.pt"))


def _inputs(case):
    name, kw, t, stereo, max_factors, seed, n_exist, nan = case
    d = mk.distance_matrix(kw["t0"], kw["t1"], t, seed, nan=nan)
    e = mk.existing_edges(t, n_exist, seed + 100)
    return d, torch.cat([e[0], e[2], e[4]]), torch.cat([e[1], e[3], e[5]])


@pytest.mark.parametrize("name", [c[0] for c in mk.cases()])
def test_kernel_matches_the_reference_method(backends, gold, name):
    case = [c for c in mk.cases() if c[0] == name][0]
    _, kw, t, stereo, max_factors, *_ = case
    d, ii1, jj1 = _inputs(case)
    es = backends.proximity_edges(d.to(dev), kw["t0"], kw["t1"], t, ii1.to(dev), jj1.to(dev), kw["rad"], kw["nms"], kw["thresh"], max_factors, stereo)
    assert torch.equal(es.cpu(), gold[name + "_es"])


@pytest.mark.parametrize("t,t0,t1,rad,nms,thresh,max_factors,stereo,seed", [
    (400, 0, 0, 2, 2, 22.0, 6000, False, 1),          # global BA sized grid: 160 000 pairs, the cap stops the walk
    (400, 0, 0, 2, 2, 22.0, -1, False, 2),            # no cap: thousands of accepted pairs
    (257, 200, 120, 3, 1, 16.0, -1, True, 3),         # rectangular window, stereo, ragged sizes
    (1, 0, 0, 2, 2, 16.0, -1, False, 4),              # a single frame: no pairs at all
    (2, 0, 0, 2, 2, 16.0, -1, True, 5),
])
def test_kernel_matches_oracle_on_large_grids(backends, t, t0, t1, rad, nms, thresh, max_factors, stereo, seed):
    d = mk.distance_matrix(t0, t1, t, seed, nan=3 if t > 100 else 0)
    e = mk.existing_edges(t, 500 if t > 100 else 0, seed + 100)
    ii1, jj1 = torch.cat([e[0], e[2], e[4]]), torch.cat([e[1], e[3], e[5]])
    want, _ = prox.proximity_edges(d.numpy(), t0, t1, t, ii1.numpy(), jj1.numpy(), rad=rad, nms=nms, thresh=thresh, max_factors=max_factors, stereo=stereo)
    es = backends.proximity_edges(d.to(dev), t0, t1, t, ii1.to(dev), jj1.to(dev), rad, nms, thresh, max_factors, stereo)
    assert es.shape[0] == want.shape[0]
    assert np.array_equal(es.cpu().numpy(), want)


def test_equal_distances_are_visited_in_index_order(backends):
    t = 24
    d = torch.full((t * t,), 5.0)                       # every pair at the same distance: ties everywhere
    z = torch.zeros(0, dtype=torch.long)
    want, _ = prox.proximity_edges(d.numpy(), 0, 0, t, z.numpy(), z.numpy(), rad=2, nms=2, thresh=16.0)
    es = backends.proximity_edges(d.to(dev), 0, 0, t, z.to(dev), z.to(dev), 2, 2, 16.0, -1, False)
    assert np.array_equal(es.cpu().numpy(), want)


def test_unchecked_index_of_the_reference_is_reported(backends):
    # a grid of 2 x 1 pairs whose temporal neighbours j = i - rad - 1 .. lie far below t1: the reference's unchecked index leaves the array (IndexError there)
    t, t0, t1 = 12, 10, 11
    d = mk.distance_matrix(t0, t1, t, 1)
    z = torch.zeros(0, dtype=torch.long, device=dev)
    with pytest.raises(Exception):
        prox.proximity_edges(d.numpy(), t0, t1, t, [], [], rad=2, nms=2, thresh=16.0)
    with pytest.raises(RuntimeError, match="IndexError"):
        backends.proximity_edges(d.to(dev), t0, t1, t, z, z, 2, 2, 16.0, -1, False)


def test_hook_drives_the_reference_graph_interface(backends, gold):
    """droid_slam_b200.modules.add_proximity_factors on an object with the reference FactorGraph's attributes: same add_factors call"""
    from droid_slam_b200 import modules
    for case in mk.cases():
        name, kw, t, stereo, max_factors, seed, n_exist, nan = case
        d = mk.distance_matrix(kw["t0"], kw["t1"], t, seed, nan=nan)
        edges = [x.to(dev) for x in mk.existing_edges(t, n_exist, seed + 100)]
        graph = mk._Graph(mk._Video(t, stereo, d.to(dev)), edges, max_factors)
        modules.add_proximity_factors(graph, **kw)
        (ii, jj, remove), = graph.calls
        assert torch.equal(torch.stack([ii, jj], 1).cpu(), gold[name + "_es"]) and bool(remove) == bool(gold[name + "_remove"])
