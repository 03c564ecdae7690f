"""Golden-vector cases shared by make_golden.py (runs the REFERENCE CUDA build on a GPU) and the CPU tests that pin
the oracle against those outputs.  Inputs are regenerated from seeds; only the reference's outputs are stored."""
import torch

from droid_slam_b200 import synth

CORR_SHAPES = [(2, 6, 8, 6, 8), (2, 5, 7, 12, 16), (1, 4, 6, 24, 32)]
BA_CASES = {
    "mono_it1": dict(cfg=dict(E=24, N=8, ht=16, wd=24, stereo=False), itrs=1, motion_only=False, rgbd=False, seed=1),
    "mono_it2": dict(cfg=dict(E=24, N=8, ht=16, wd=24, stereo=False), itrs=2, motion_only=False, rgbd=False, seed=2),
    "rgbd_it2": dict(cfg=dict(E=30, N=9, ht=12, wd=16, stereo=False), itrs=2, motion_only=False, rgbd=True, seed=3),
    "stereo_it2": dict(cfg=dict(E=30, N=9, ht=12, wd=16, stereo=True), itrs=2, motion_only=False, rgbd=False, seed=4),
    "motion_only": dict(cfg=dict(E=24, N=8, ht=16, wd=24, stereo=False), itrs=2, motion_only=True, rgbd=False, seed=5),
    "backend_lm": dict(cfg=dict(E=40, N=10, ht=12, wd=16, stereo=False), itrs=3, motion_only=False, rgbd=False, seed=6, lm=1e-5, ep=1e-2),
}


def corr_case(shape, dtype, seed):
    n, h1, w1, h2, w2 = shape
    g = torch.Generator().manual_seed(1000 + seed)
    vol = torch.randn(n, h1, w1, h2, w2, generator=g).to(dtype)
    cx = torch.rand(n, 1, h1, w1, generator=g) * (w2 + 8) - 4
    cy = torch.rand(n, 1, h1, w1, generator=g) * (h2 + 8) - 4
    grad = torch.randn(n, 7, 7, h1, w1, generator=g).to(dtype)
    return vol, torch.cat([cx, cy], 1).contiguous(), grad


def altcorr_case(dtype, seed=0):
    g = torch.Generator().manual_seed(2000 + seed)
    B, N, C, H, W = 1, 4, 16, 8, 12
    fmaps = torch.randn(B, N, C, H, W, generator=g).to(dtype)
    ii = torch.tensor([0, 1, 2, 3, 0]); jj = torch.tensor([1, 2, 3, 3, 3])
    coords = torch.rand(B, 5, 2, H, W, generator=g) * torch.tensor([W + 6.0, H + 6.0]).view(1, 1, 2, 1, 1) - 3
    return fmaps, coords.contiguous(), ii, jj


def ba_scene(name):
    c = BA_CASES[name]
    cfg = dict(c["cfg"]); cfg.update(itrs=c["itrs"], lm=c.get("lm", 1e-4), ep=c.get("ep", 0.1))
    return synth.make_scene(cfg, seed=c["seed"], rgbd=c["rgbd"]), c


def geom_scene():
    return synth.make_scene(dict(E=20, N=8, ht=12, wd=16, stereo=False, itrs=1, lm=1e-4, ep=0.1), seed=7)
