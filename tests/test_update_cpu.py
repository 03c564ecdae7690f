"""Row A6 (update operator) on CPU: the oracle against golden vectors produced by the reference module itself
(tests/golden/make_update_golden.py), and the host mirror (droid_slam_b200/update.py) against the oracle."""
import os

import pytest
import torch

import oracle
from droid_slam_b200 import synth
from droid_slam_b200.update import UpdateModule, segment_mean

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = (("a", dict(E=5, ht=6, wd=8, seed=0, n_src=3)), ("b", dict(E=7, ht=5, wd=9, seed=1, n_src=4)))
NAMES = ("net", "delta", "weight", "eta", "upmask")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "update_module.pt"))


@pytest.fixture(scope="module")
def weights():
    return synth.make_update_weights(0)


@pytest.mark.parametrize("name,kw", CASES)
def test_update_oracle_matches_reference_module(gold, weights, name, kw):
    net, inp, corr, flow, ii = synth.make_update_inputs(**kw)
    out = oracle.update_module_forward(weights, net, inp, corr, flow, ii)
    for k, t in zip(NAMES, out):
        g = gold["%s_%s" % (name, k)]
        assert t.shape == g.shape
        assert torch.allclose(t, g, rtol=1e-5, atol=1e-6), (k, float((t - g).abs().max()))      # observed: bit-identical
    out = oracle.update_module_forward(weights, net, inp, corr, None, None)
    assert len(out) == 3
    for k, t in zip(NAMES[:3], out):
        assert torch.allclose(t, gold["%s_noflow_%s" % (name, k)], rtol=1e-5, atol=1e-6)


def test_update_mirror_has_the_reference_parameter_names(gold, weights):
    mod = UpdateModule()
    assert sorted(mod.state_dict().keys()) == gold["state_dict_keys"]          # a DROID checkpoint's update.* entries load as they are
    assert mod.load_state_dict(weights, strict=True) is not None
    for k, v in mod.state_dict().items():
        assert v.shape == weights[k].shape


@pytest.mark.parametrize("name,kw", CASES)
def test_update_mirror_matches_oracle(weights, name, kw):
    mod = UpdateModule().eval()
    mod.load_state_dict(weights)
    net, inp, corr, flow, ii = synth.make_update_inputs(**kw)
    with torch.no_grad():
        got = mod(net, inp, corr, flow, ii)
        ref = oracle.update_module_forward(weights, net, inp, corr, flow, ii)
        assert len(got) == 5
        for k, a, b in zip(NAMES, got, ref):
            assert a.shape == b.shape and torch.allclose(a, b, rtol=1e-4, atol=1e-5), (k, float((a - b).abs().max()))
        got3 = mod(net, inp, corr)                                              # flow=None, ii=None like MotionFilter.track's call
        ref3 = oracle.update_module_forward(weights, net, inp, corr)
        assert len(got3) == 3
        for a, b in zip(got3, ref3):
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)


def test_segment_mean_is_scatter_mean_over_sorted_unique_sources():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 9, 3, 4, generator=g)
    ii = torch.tensor([7, 2, 7, 7, 4, 2, 9, 4, 4])
    out = segment_mean(x, ii)
    assert out.shape == (2, 4, 3, 4)
    for k, f in enumerate((2, 4, 7, 9)):
        assert torch.allclose(out[:, k], x[:, ii == f].mean(1), atol=1e-6)
    assert torch.allclose(out, oracle.scatter_mean_by_source(x, ii))
