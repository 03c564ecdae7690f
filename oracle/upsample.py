"""CPU restatement of the reference's convex upsampling -- TEST INFRASTRUCTURE ONLY.
cvx_upsample / upsample_disp: droid_slam/droid_net.py:21-42, used by DepthVideo.upsample (depth_video.py:155-159).
Pinned against the reference function itself (tests/golden/make_update_golden.py stores its output for seeded inputs)."""
import torch
import torch.nn.functional as F

__all__ = ["cvx_upsample", "upsample_disp"]


def cvx_upsample(data, mask):
    """droid_net.py:21-35.  data [B,ht,wd,dim], mask [B,576,ht,wd] -> [B,8ht,8wd,dim]"""
    batch, ht, wd, dim = data.shape
    data = data.permute(0, 3, 1, 2)
    mask = torch.softmax(mask.view(batch, 1, 9, 8, 8, ht, wd), dim=2)
    up = F.unfold(data, [3, 3], padding=1).view(batch, dim, 9, 1, 1, ht, wd)
    up = torch.sum(mask * up, dim=2)
    return up.permute(0, 4, 2, 5, 3, 1).reshape(batch, 8 * ht, 8 * wd, dim)


def upsample_disp(disp, mask):
    """droid_net.py:37-42.  disp [B,N,ht,wd], mask [B,N,576,ht,wd] -> [B,N,8ht,8wd]"""
    batch, num, ht, wd = disp.shape
    return cvx_upsample(disp.reshape(batch * num, ht, wd, 1), mask.reshape(batch * num, -1, ht, wd)).view(batch, num, 8 * ht, 8 * wd)
