"""Hooks that put the B200-native extras under the reference's own Python classes.

The reference's `droid_slam/modules/corr.py` runs UNCHANGED on this package's `droid_backends` (its four correlation ops are the nine
drop-in callables); nothing of it is restated here.  What the reference computes in Python around those ops and this package has a
kernel for is offered as a hook:

  * `install_corr_volume_hook(corr_module)`: `CorrBlock.__init__` (modules/corr.py:24-38, 63-71: torch.matmul + 3x avg_pool2d) builds
    its four pyramid levels with the one-pass tcgen05 kernel `droid_backends.corr_volume_pyramid` instead.  Only the constructor is
    replaced; lookups, `cat` and `__getitem__` stay the reference's code.
  * `reproject(...)`: `DepthVideo.reproject` (depth_video.py:171-179 -> geom/projective_ops.py:165-198) as one kernel.
  * `add_proximity_factors(graph, ...)` / `install_proximity_hook(FactorGraph)`: the edge selection of
    `FactorGraph.add_proximity_factors` (factor_graph.py:346-412) on the device (row F1).
"""
import torch

from . import install

__all__ = ["install_corr_volume_hook", "reproject", "upsample", "add_proximity_factors", "install_proximity_hook"]


def install_corr_volume_hook(corr_module, strict=True, fused_lookup=False):
    """corr_module = the imported reference module `modules.corr`.  strict: shapes without a kernel (anything but f16, 128 channels,
    wd = 64 handled by corr_volume_pyramid) raise instead of silently taking the reference's library path.
    fused_lookup: additionally replace `CorrBlock.__call__` by the one-launch 4-level lookup `corr_lookup_pyramid`; the volumes of
    levels 0 and 1 are then kept in the tiled private layout (64-byte DRAM atoms, about a third less HBM traffic per lookup) -- same
    tensor shapes, `cat` / `__getitem__` over edges keep working, results bit-identical to the reference-layout path."""
    be = install()
    ref_init = corr_module.CorrBlock.__init__
    tiled = bool(fused_lookup)

    def __init__(self, fmap1, fmap2, num_levels=4, radius=3):
        batch, num, dim, ht, wd = fmap1.shape
        ok = fmap1.is_cuda and fmap1.dtype == torch.float16 and fmap2.dtype == torch.float16 and num_levels == 4 and be.corr_volume_supported(dim, ht, wd)
        if not ok:
            if strict:
                raise RuntimeError("corr_volume_pyramid has no kernel for fmaps %s %s, levels=%d" % (tuple(fmap1.shape), fmap1.dtype, num_levels))
            return ref_init(self, fmap1, fmap2, num_levels, radius)
        self.num_levels, self.radius = num_levels, radius
        idx = torch.arange(batch * num, device=fmap1.device)
        self.corr_pyramid = be.corr_volume_pyramid(fmap1.reshape(batch * num, dim, ht, wd).contiguous(), fmap2.reshape(batch * num, dim, ht, wd).contiguous(), idx, idx, tiled)
        self._b200_tiled = tiled

    def __call__(self, coords):
        batch, num, ht, wd, _ = coords.shape
        c = coords.permute(0, 1, 4, 2, 3).contiguous().view(batch * num, 2, ht, wd)
        return be.corr_lookup_pyramid([v.contiguous() for v in self.corr_pyramid], c, getattr(self, "_b200_tiled", False)).view(batch, num, -1, ht, wd)

    corr_module.CorrBlock.__init__ = __init__
    if fused_lookup:
        corr_module.CorrBlock.__call__ = __call__
    return corr_module


def reproject(poses, disps, intrinsics, ii, jj):
    """DepthVideo.reproject (reference droid_slam/depth_video.py:171-179) in one kernel: poses [N,7], disps [N,ht,wd],
    intrinsics [N,4], ii/jj index tensors or lists -> (coords [1,E,ht,wd,2], valid [1,E,ht,wd,1]) like the reference."""
    be = install()
    ii = torch.as_tensor(ii).to(device=poses.device, dtype=torch.long).reshape(-1)
    jj = torch.as_tensor(jj).to(device=poses.device, dtype=torch.long).reshape(-1)
    coords, valid = be.reproject(poses.contiguous(), disps.contiguous(), intrinsics.contiguous(), ii, jj)
    return coords[None], valid[None]


def upsample(disps, disps_up, ix, mask):
    """DepthVideo.upsample (reference droid_slam/depth_video.py:155-159) in one kernel: disps_up[ix] = cvx_upsample(disps[ix], mask).
    disps [N,ht,wd] f32, disps_up [N,8ht,8wd] f32 (written in place), ix index tensor, mask [1,len(ix),576,ht,wd] (the update operator's upmask)."""
    be = install()
    m = mask.reshape(-1, 576, disps.shape[1], disps.shape[2]).contiguous()
    disps_up[ix] = be.cvx_upsample(disps[ix].contiguous(), m)
    return disps_up


def add_proximity_factors(graph, t0=0, t1=0, rad=2, nms=2, beta=0.25, thresh=16.0, remove=False):
    """FactorGraph.add_proximity_factors (reference factor_graph.py:346-412) with the edge selection on the device (row F1).

    Same signature and effect as the reference method, `graph` being the reference's FactorGraph instance: the distance matrix stays on the
    GPU (`video.distance` -> droid_backends.frame_distance), masking / suppression / greedy selection run in
    `droid_backends.proximity_edges` (csrc/proximity.cu) instead of the Python triple loop over a CPU copy, and the resulting edge list
    -- identical, order included -- goes to the graph's own `add_factors`.  One host read (the number of edges) instead of the
    reference's `.cpu()` round trips."""
    be = install()
    video = graph.video
    t = video.counter.value
    dev = graph.ii.device
    ix = torch.arange(t0, t, device=dev)
    jx = torch.arange(t1, t, device=dev)
    ii, jj = torch.meshgrid(ix, jx, indexing="ij")
    d = video.distance(ii.reshape(-1), jj.reshape(-1), beta=beta).float().contiguous()
    ii1 = torch.cat([graph.ii, graph.ii_bad, graph.ii_inac], 0).to(torch.long).contiguous()
    jj1 = torch.cat([graph.jj, graph.jj_bad, graph.jj_inac], 0).to(torch.long).contiguous()
    es = be.proximity_edges(d, int(t0), int(t1), int(t), ii1, jj1, int(rad), int(nms), float(thresh), int(graph.max_factors), bool(video.stereo))
    graph.add_factors(es[:, 0].contiguous(), es[:, 1].contiguous(), remove)


def install_proximity_hook(factor_graph_class):
    """replace `add_proximity_factors` of the reference's FactorGraph class (factor_graph.py:346) by the device version"""
    factor_graph_class.add_proximity_factors = add_proximity_factors
    return factor_graph_class
