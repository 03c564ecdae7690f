"""Host-side mirror of the reference's correlation interface (droid_slam/modules/corr.py) on the B200-native kernels.

`CorrBlock(fmap1, fmap2)(coords)` and `AltCorrBlock(fmaps)(coords, ii, jj)` keep the reference's constructor arguments, call
pattern, tensor layouts and `cat` / `__getitem__` helpers, so `factor_graph.py` can use them in place of the reference
classes (`from droid_slam_b200.modules import CorrBlock, AltCorrBlock`).  The lookup goes through
`droid_backends.corr_index_forward` / `altcorr_forward` exactly like the reference; the difference is the volume build:
`CorrBlock.__init__` produces all four pyramid levels with one tcgen05/TMEM/TMA kernel (`droid_backends.corr_volume_pyramid`)
instead of `torch.matmul` + 3x `avg_pool2d` (reference modules/corr.py:24-38, 63-71).

Inference only (no autograd), like the frontend/backend use of these classes.
"""
import torch
import torch.nn.functional as F

from . import install

__all__ = ["CorrBlock", "AltCorrBlock", "reproject"]


class CorrBlock:
    """reference droid_slam/modules/corr.py:23-71"""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=3):
        self.num_levels = num_levels
        self.radius = radius
        self._be = install()
        batch, num, dim, ht, wd = fmap1.shape
        assert fmap2.shape == fmap1.shape
        f1 = fmap1.reshape(batch * num, dim, ht, wd)
        f2 = fmap2.reshape(batch * num, dim, ht, wd)
        if self.native_volume_supported(fmap1, num_levels):
            idx = torch.arange(batch * num, device=fmap1.device)
            self.corr_pyramid = self._be.corr_volume_pyramid(f1.contiguous(), f2.contiguous(), idx, idx)
        else:
            # shapes the tensor-core kernel does not cover yet (wd != 64, dim != 128, dtype != f16): the reference formula
            corr = torch.matmul((f1.reshape(batch * num, dim, ht * wd) / 4.0).transpose(1, 2), f2.reshape(batch * num, dim, ht * wd) / 4.0)
            corr = corr.reshape(batch * num * ht * wd, 1, ht, wd)
            self.corr_pyramid = []
            for i in range(num_levels):
                self.corr_pyramid.append(corr.view(batch * num, ht, wd, ht // 2 ** i, wd // 2 ** i))
                corr = F.avg_pool2d(corr, 2, stride=2)

    @staticmethod
    def native_volume_supported(fmap, num_levels=4):
        _, _, dim, ht, wd = fmap.shape
        return fmap.is_cuda and fmap.dtype == torch.float16 and dim == 128 and wd == 64 and ht % 8 == 0 and num_levels == 4

    def __call__(self, coords):
        out_pyramid = []
        batch, num, ht, wd, _ = coords.shape
        coords = coords.permute(0, 1, 4, 2, 3).contiguous().view(batch * num, 2, ht, wd)
        for i in range(self.num_levels):
            corr, = self._be.corr_index_forward(self.corr_pyramid[i], (coords / 2 ** i).contiguous(), self.radius)
            out_pyramid.append(corr.view(batch, num, -1, ht, wd))
        return torch.cat(out_pyramid, dim=2)

    def cat(self, other):
        for i in range(self.num_levels):
            self.corr_pyramid[i] = torch.cat([self.corr_pyramid[i], other.corr_pyramid[i]], 0)
        return self

    def __getitem__(self, index):
        for i in range(self.num_levels):
            self.corr_pyramid[i] = self.corr_pyramid[i][index]
        return self


class AltCorrBlock:
    """reference droid_slam/modules/corr.py:89-117"""

    def __init__(self, fmaps, num_levels=4, radius=3):
        self.num_levels = num_levels
        self.radius = radius
        self._be = install()
        B, N, C, H, W = fmaps.shape
        fmaps = fmaps.view(B * N, C, H, W)
        self.pyramid = []
        for i in range(self.num_levels):
            self.pyramid.append(fmaps.view(B, N, C, H // 2 ** i, W // 2 ** i))
            fmaps = F.avg_pool2d(fmaps, 2, stride=2)

    def __call__(self, coords, ii, jj):
        coords = coords.permute(0, 1, 4, 2, 3).contiguous()
        corr_list = []
        for i in range(self.num_levels):
            corr, = self._be.altcorr_forward(self.pyramid[0].contiguous(), self.pyramid[i].contiguous(), (coords / 2 ** i).contiguous(), ii, jj, self.radius)
            corr_list.append(corr.flatten(2, 3))
        return torch.stack(corr_list, dim=2).flatten(2, 3)


def reproject(poses, disps, intrinsics, ii, jj):
    """DepthVideo.reproject (reference droid_slam/depth_video.py:171-179) in one kernel: poses [N,7], disps [N,ht,wd],
    intrinsics [N,4], ii/jj index tensors or lists -> (coords [1,E,ht,wd,2], valid [1,E,ht,wd,1]) like the reference."""
    be = install()
    ii = torch.as_tensor(ii).to(device=poses.device, dtype=torch.long).reshape(-1)
    jj = torch.as_tensor(jj).to(device=poses.device, dtype=torch.long).reshape(-1)
    coords, valid = be.reproject(poses.contiguous(), disps.contiguous(), intrinsics.contiguous(), ii, jj)
    return coords[None], valid[None]
