"""Ours against the UNMODIFIED reference CUDA build on the same GPU, same tensors, at BASELINE's full sizes.

`oracle/_ref/droid_backends_ref*.so` = /root/reference/src/{droid.cpp,droid_kernels.cu,correlation_kernels.cu,altcorr_kernel.cu}
compiled in place for sm_100a by oracle/build_ref.sh (Eigen stand-in for the absent submodule); it travels to the GPU box with the
snapshot (git-ignored, not gpurun-ignored).  These tests call both extensions the way the reference's Python does
(depth_video.py:213-225, factor_graph.py:327-328, modules/corr.py:12,79) and compare:
  * index / lookup ops: torch.equal (bit-identical);
  * ba: poses and inverse depths after the update against BASELINE.json's 1e-4 relative bound.  Poses: every component relative to
    the pose's translation norm (unit quaternion: to 1).  Inverse depths: elementwise |a-b|/|b| with no absolute floor, evaluated at
    the 99.99th percentile, plus max|a-b| <= 1e-4 max|b| -- the worst single pixel is not a usable statistic because inverse depths
    pass through zero in these scenes (|b| as small as 1e-3) and the reference itself deviates from exact arithmetic by the same
    amount there: against the fp64 oracle the worst pixel is 6.5e-5 (ours) vs 4.1e-5 (reference) at the metric size and 9.4e-4 vs
    2.1e-4 on the stereo config, the p99.99 1e-5 for both (profiles/r2_ba_vs_reference_stats.txt).
Skipped (not failed) when the reference build is absent."""
import glob
import os
import sys

import pytest
import torch

from droid_slam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not glob.glob(os.path.join(REF_DIR, "droid_backends_ref*.so")), reason="oracle/_ref not built")]
dev = "cuda"


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF_DIR)
    import droid_backends_ref
    return droid_backends_ref


def _pose_rel(P, Pr):
    P, Pr = P.double().cpu(), Pr.double().cpu()
    et = (P[:, :3] - Pr[:, :3]).abs() / Pr[:, :3].norm(dim=1, keepdim=True).clamp(min=1e-2)
    eq = (P[:, 3:] - Pr[:, 3:]).abs()
    return float(torch.cat([et, eq], 1).max())


def _disp_rel(D, Dr, q=0.9999):
    """(q-quantile of the elementwise relative error, max abs error / max |reference|)"""
    D, Dr = D.double().cpu().flatten(), Dr.double().cpu().flatten()
    rel = torch.sort((D - Dr).abs() / Dr.abs()).values
    return float(rel[min(rel.numel() - 1, int(q * rel.numel()))]), float((D - Dr).abs().max() / Dr.abs().max())


def _ok(P, D, Pr, Dr):
    ep, (eq, ea) = _pose_rel(P, Pr), _disp_rel(D, Dr)
    return (ep < 1e-4 and eq < 1e-4 and ea < 1e-4), (ep, eq, ea)


def _ba_both(backends, ref, s, itrs, motion_only=False):
    args = [s[k].to(dev) for k in ("intrinsics", "disps_sens", "targets", "weights", "eta", "ii", "jj")]
    P, D = s["poses"].to(dev), s["disps"].to(dev)
    Pr, Dr = s["poses"].to(dev), s["disps"].to(dev)
    o = backends.ba(P, D, *args, s["t0"], s["t1"], itrs, s["lm"], s["ep"], motion_only)
    r = ref.ba(Pr, Dr, *args, s["t0"], s["t1"], itrs, s["lm"], s["ep"], motion_only)
    torch.cuda.synchronize()
    return (P, D, o), (Pr, Dr, r)


def test_corr_index_forward_bit_identical_at_metric_size(backends, ref):
    """512 edges x 48x64 f16 volumes, all four levels (CorrBlock.__call__, modules/corr.py:40-50).  The reference's 32-bit accessors
    cannot address 512 level-0 planes at once, so it is called in 128-edge chunks; ours takes the whole batch in one call."""
    s = synth.make_scene("metric")
    pyr, coords, _ = synth.make_corr_inputs(s, dtype=torch.float16, device=dev)
    for lvl, vol in enumerate(pyr):
        c = (coords / 2 ** lvl).contiguous()
        ours, = backends.corr_index_forward(vol, c, 3)
        for a in range(0, 512, 128):
            r, = ref.corr_index_forward(vol[a:a + 128], c[a:a + 128].contiguous(), 3)
            assert torch.equal(ours[a:a + 128], r), (lvl, a)
    del pyr
    torch.cuda.empty_cache()


def test_corr_index_forward_f32_and_backward_bit_identical(backends, ref):
    s = synth.make_scene("c2_frontend")                         # config 2: 128 edges, fp32 volumes
    sub = dict(s); sub["ii"] = s["ii"][:96]; sub["jj"] = s["jj"][:96]; sub["coords_gt"] = s["coords_gt"][:96]; sub["cfg"] = dict(s["cfg"], E=96)
    pyr, coords, _ = synth.make_corr_inputs(sub, dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    for lvl, vol in enumerate(pyr):
        c = (coords / 2 ** lvl).contiguous()
        assert torch.equal(backends.corr_index_forward(vol, c, 3)[0], ref.corr_index_forward(vol, c, 3)[0]), lvl
        if lvl >= 2:
            grad = torch.randn(96, 7, 7, 48, 64, device=dev, generator=g)
            assert torch.equal(backends.corr_index_backward(vol, c, grad, 3)[0], ref.corr_index_backward(vol, c, grad, 3)[0]), lvl


def test_altcorr_forward_bit_identical_at_full_resolution(backends, ref):
    """AltCorrBlock.__call__ (modules/corr.py:104-117) on 48x64 f16 feature maps, 4 levels, a chunk of 24 edges"""
    g = torch.Generator().manual_seed(5)
    N, M = 8, 24
    fmaps = torch.randn(1, N, 128, 48, 64, generator=g).half().to(dev)
    s = synth.make_scene(dict(E=M, N=N, ht=48, wd=64, stereo=False, itrs=1, lm=1e-4, ep=0.1), seed=3)
    coords = (s["coords_gt"] + 2 * torch.rand(M, 48, 64, 2, generator=g) - 1).permute(0, 3, 1, 2)[None].contiguous().to(dev)
    ii, jj = s["ii"].to(dev), s["jj"].to(dev)
    f = fmaps[0]
    for lvl in range(4):
        f2 = f[None].contiguous()
        c = (coords / 2 ** lvl).contiguous()
        a, = backends.altcorr_forward(fmaps, f2, c, ii, jj, 3)
        b, = ref.altcorr_forward(fmaps, f2, c, ii, jj, 3)
        assert a.shape == b.shape and torch.equal(a.contiguous(), b.contiguous()), lvl
        f = torch.nn.functional.avg_pool2d(f, 2, stride=2)


def test_geometry_ops_match_at_metric_size(backends, ref):
    s = synth.make_scene("metric")
    P, D, K, ii, jj = [s[k].to(dev) for k in ("poses", "disps", "intrinsics", "ii", "jj")]
    c, v = backends.projmap(P, D, K, ii, jj)
    cr, vr = ref.projmap(P, D, K, ii, jj)
    assert torch.equal(c, cr) and torch.equal(v, vr)
    assert torch.equal(backends.iproj(P, D, K), ref.iproj(P, D, K))
    ix = torch.arange(72, device=dev); th = torch.full((72,), 0.05, device=dev)
    assert torch.equal(backends.depth_filter(P, D, K, ix, th), ref.depth_filter(P, D, K, ix, th))
    # all-pairs distance like DepthVideo.distance (depth_video.py:181-211); sums of 3072 terms in a different order: 1e-5 relative
    a, b = torch.meshgrid(torch.arange(72), torch.arange(72), indexing="ij")
    a, b = a.reshape(-1).to(dev), b.reshape(-1).to(dev)
    d, dr = backends.frame_distance(P, D, K, a, b, 0.3), ref.frame_distance(P, D, K, a, b, 0.3)
    assert float(((d - dr).abs() / dr.abs().clamp(min=1e-3)).max()) < 1e-5


def test_ba_metric_size_within_1e4_relative_of_reference(backends, ref):
    s = synth.make_scene("metric")                              # 512 edges, 72 keyframes, ba(itrs=2, lm=1e-4, ep=0.1)
    (P, D, o), (Pr, Dr, r) = _ba_both(backends, ref, s, 2)
    ok, errs = _ok(P, D, Pr, Dr)
    assert ok, errs
    assert o[0].shape == r[0].shape and o[1].shape == r[1].shape


def test_ba_config4_stereo_within_1e4_relative_of_reference(backends, ref):
    s = synth.make_scene("c4_stereo")                           # 256 edges incl. one (i,i) edge per frame
    (P, D, o), (Pr, Dr, r) = _ba_both(backends, ref, s, 2)
    ok, errs = _ok(P, D, Pr, Dr)
    assert ok, errs


def test_ba_config2_rgbd_and_motion_only(backends, ref):
    s = synth.make_scene("c2_frontend", rgbd=True)
    (P, D, o), (Pr, Dr, r) = _ba_both(backends, ref, s, 2)
    ok, errs = _ok(P, D, Pr, Dr)
    assert ok, errs
    (P, D, o), (Pr, Dr, r) = _ba_both(backends, ref, s, 2, motion_only=True)
    assert _pose_rel(P, Pr) < 1e-4 and torch.equal(D, Dr)


def test_ba_config3_global_10_iterations_within_1e4_relative_of_reference(backends, ref):
    """BASELINE config 3: 2048 edges / 400 keyframes, 10 Gauss-Newton iterations, lm=1e-5, ep=1e-2 (droid_backend.py:25-42 ->
    factor_graph.py:327-328); 6P = 2394."""
    s = synth.make_scene("c3_global")
    (P, D, o), (Pr, Dr, r) = _ba_both(backends, ref, s, 10)
    ok, errs = _ok(P, D, Pr, Dr)
    assert ok, errs
