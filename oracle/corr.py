"""Correlation lookup ops restated from src/correlation_kernels.cu and src/altcorr_kernel.cu, plus the
Python-side volume/pyramid construction of modules/corr.py.  TEST INFRASTRUCTURE ONLY."""
import torch
import torch.nn.functional as F

__all__ = ["corr_index_forward", "corr_index_backward", "altcorr_forward", "altcorr_backward",
           "corr_volume", "corr_pyramid", "fmap_pyramid", "corr_block_lookup", "altcorr_block_lookup"]


def _fma(a, b, c):
    """fma(a,b,c) in the tensor dtype.  fp32: product is exact in fp64, one extra rounding of the sum
    (double rounding is possible but ~2^-29 rare); fp64: plain a*b+c."""
    if a.dtype == torch.float32:
        return (a.double() * b.double() + c.double()).float()
    return a * b + c


def _taps(volume, coords, r):
    """gather the (2r+2)^2 taps volume[n,y,x,y1,x1], y1=floor(y0)-r+j, x1=floor(x0)-r+i; zeros outside
    (src/correlation_kernels.cu:46-55).  Returns S [N,D,D,h1,w1] indexed [i (x-offset), j (y-offset)], dx, dy."""
    N, h1, w1, h2, w2 = volume.shape
    D = 2 * r + 2
    x0 = coords[:, 0]; y0 = coords[:, 1]                     # [N,h1,w1] float32
    fx = torch.floor(x0); fy = torch.floor(y0)
    dx = x0 - fx; dy = y0 - fy
    big = 1 << 20
    fxi = torch.nan_to_num(fx, nan=0.0, posinf=big, neginf=-big).clamp(-big, big).long()
    fyi = torch.nan_to_num(fy, nan=0.0, posinf=big, neginf=-big).clamp(-big, big).long()
    off = torch.arange(D) - r
    x1 = fxi[:, None, None] + off[None, :, None, None, None]      # [N,D,1,h1,w1]
    y1 = fyi[:, None, None] + off[None, None, :, None, None]      # [N,1,D,h1,w1]
    inb = (x1 >= 0) & (x1 < w2) & (y1 >= 0) & (y1 < h2)
    lin = (y1.clamp(0, h2 - 1) * w2 + x1.clamp(0, w2 - 1))         # [N,D,D,h1,w1]
    vol = volume.reshape(N, h1, w1, h2 * w2).permute(0, 3, 1, 2)    # [N,h2*w2,h1,w1]
    S = torch.gather(vol, 1, lin.reshape(N, D * D, h1, w1)).reshape(N, D, D, h1, w1)
    S = torch.where(inb, S, torch.zeros((), dtype=volume.dtype))
    return S, dx, dy, inb


def corr_index_forward(volume, coords, radius):
    """src/correlation_kernels.cu:20-71,127-156.  volume [N,h1,w1,h2,w2] (f16/f32/f64), coords [N,2,h1,w1] f32
    -> [corr [N,2r+1,2r+1,h1,w1]] (x-offset major).  Per output the kernel accumulates, in the volume dtype and in
    this order, taps (i,j),(i,j+1),(i+1,j),(i+1,j+1) with weights (1-dx)(1-dy), (1-dx)dy, dx(1-dy), dx dy, each
    weight rounded to the volume dtype first (:56-66); fp32/fp64 `+=` of a product contracts to one FMA under
    nvcc's default -fmad=true, fp16 rounds the product and the sum separately (c10::Half operators)."""
    r = radius
    rd = 2 * r + 1
    dt = volume.dtype
    S, dx, dy, inb = _taps(volume, coords, r)
    one = torch.ones((), dtype=torch.float32)
    w00 = ((one - dx) * (one - dy)).to(dt)[:, None, None]
    w01 = ((one - dx) * dy).to(dt)[:, None, None]
    w10 = (dx * (one - dy)).to(dt)[:, None, None]
    w11 = (dx * dy).to(dt)[:, None, None]
    s00 = S[:, :rd, :rd]; s01 = S[:, :rd, 1:]; s10 = S[:, 1:, :rd]; s11 = S[:, 1:, 1:]
    i00 = inb[:, :rd, :rd]; i01 = inb[:, :rd, 1:]; i10 = inb[:, 1:, :rd]; i11 = inb[:, 1:, 1:]
    out = torch.zeros_like(s00)
    for s, w, m in ((s00, w00, i00), (s01, w01, i01), (s10, w10, i10), (s11, w11, i11)):
        if dt == torch.float16:
            new = out + s * w                     # two roundings to half
        else:
            new = _fma(s, w.expand_as(s), out)
        out = torch.where(m, new, out)            # out-of-bounds taps are skipped, not multiplied (:53)
    return [out]


def corr_index_backward(volume, coords, corr_grad, radius):
    """src/correlation_kernels.cu:74-125,158-185: volume_grad[n,y,x,y1,x1] += g, g summed in the volume dtype in
    the order (i-1,j-1),(i-1,j),(i,j-1),(i,j)."""
    r = radius
    rd = 2 * r + 1
    D = rd + 1
    dt = volume.dtype
    N, h1, w1, h2, w2 = volume.shape
    _, dx, dy, inb = _taps(volume, coords, r)
    one = torch.ones((), dtype=torch.float32)
    w11 = (dx * dy).to(dt)[:, None, None]
    w10 = (dx * (one - dy)).to(dt)[:, None, None]
    w01 = ((one - dx) * dy).to(dt)[:, None, None]
    w00 = ((one - dx) * (one - dy)).to(dt)[:, None, None]
    G = torch.zeros(N, D + 1, D + 1, h1, w1, dtype=dt)   # padded so that index -1 / rd read zero
    G[:, 1:rd + 1, 1:rd + 1] = corr_grad
    g = torch.zeros(N, D, D, h1, w1, dtype=dt)
    # tap (i,j): terms corr_grad[i-1][j-1]*dxdy, [i-1][j]*dx(1-dy), [i][j-1]*(1-dx)dy, [i][j]*(1-dx)(1-dy)
    for (oi, oj, w) in ((0, 0, w11), (0, 1, w10), (1, 0, w01), (1, 1, w00)):
        cg = G[:, oi:oi + D, oj:oj + D]
        if dt == torch.float16:
            g = g + cg * w
        elif dt == torch.float32:
            g = (cg.double() * w.double() + g.double()).float()
        else:
            g = cg * w + g
    g = torch.where(inb, g, torch.zeros((), dtype=dt))
    off = torch.arange(D) - r
    big = 1 << 20
    fxi = torch.nan_to_num(torch.floor(coords[:, 0]), nan=0.0, posinf=big, neginf=-big).clamp(-big, big).long()
    fyi = torch.nan_to_num(torch.floor(coords[:, 1]), nan=0.0, posinf=big, neginf=-big).clamp(-big, big).long()
    x1 = (fxi[:, None, None] + off[None, :, None, None, None]).clamp(0, w2 - 1)
    y1 = (fyi[:, None, None] + off[None, None, :, None, None]).clamp(0, h2 - 1)
    lin = (y1 * w2 + x1).expand(N, D, D, h1, w1).reshape(N, D * D, h1, w1)
    vg = torch.zeros(N, h2 * w2, h1, w1, dtype=dt)
    vg.scatter_add_(1, lin, g.reshape(N, D * D, h1, w1))   # taps of one pixel never collide
    return [vg.permute(0, 2, 3, 1).reshape(N, h1, w1, h2, w2).contiguous()]


def altcorr_forward(fmap1, fmap2, coords, ii, jj, radius):
    """src/altcorr_kernel.cu:24-75,132-172.  fmap1 [B,N1,C,H,W], fmap2 [B,N2,C,H2,W2], coords [B,M,2,H,W] f32,
    ii,jj [M] -> [out [B,M,2r+1,2r+1,H,W]] (x-offset major after the permute at :171)."""
    r = radius
    D = 2 * r + 2
    dt = fmap1.dtype
    B, M, _, H, W = coords.shape
    C = fmap1.shape[2]
    H2, W2 = fmap2.shape[3], fmap2.shape[4]
    f1 = (fmap1[:, ii] / 4.0)                                  # [B,M,C,H,W]   (:66)
    f2 = (fmap2[:, jj] / 4.0).reshape(B, M, C, H2 * W2)        # (:67)
    x = coords[:, :, 0]; y = coords[:, :, 1]
    big = 1 << 20
    fxi = torch.nan_to_num(torch.floor(x), nan=0.0, posinf=big, neginf=-big).clamp(-big, big).long()
    fyi = torch.nan_to_num(torch.floor(y), nan=0.0, posinf=big, neginf=-big).clamp(-big, big).long()
    corr = torch.zeros(B, M, D, D, H, W, dtype=dt)
    for a in range(D):          # a <-> ii (y offset), c <-> jj (x offset)   (:45-46,60-61)
        for c in range(D):
            i1 = fyi + (a - r); j1 = fxi + (c - r)
            inb = (i1 >= 0) & (i1 < H2) & (j1 >= 0) & (j1 < W2)
            lin = (i1.clamp(0, H2 - 1) * W2 + j1.clamp(0, W2 - 1))             # [B,M,H,W]
            g = torch.gather(f2, 3, lin.reshape(B, M, 1, H * W).expand(B, M, C, H * W)).reshape(B, M, C, H, W)
            prod = (f1 * g).float()                            # product rounded in scalar_t, then to float (:68)
            s = prod.sum(2)                                    # fp32 accumulation (:63)
            corr[:, :, a, c] = torch.where(inb, s, torch.zeros_like(s)).to(dt)
    xs = x[:, :, None, None]; ys = y[:, :, None, None]
    dx = (xs - torch.floor(xs)).to(dt)                         # (:158-161)
    dy = (ys - torch.floor(ys)).to(dt)
    out = (1 - dx) * (1 - dy) * corr[:, :, 0:D - 1, 0:D - 1]
    out = out + (dx) * (1 - dy) * corr[:, :, 0:D - 1, 1:D]
    out = out + (1 - dx) * (dy) * corr[:, :, 1:D, 0:D - 1]
    out = out + (dx) * (dy) * corr[:, :, 1:D, 1:D]
    return [out.permute(0, 1, 3, 2, 4, 5)]


def altcorr_backward(fmap1, fmap2, coords, corr_grad, ii, jj, radius):
    """src/altcorr_kernel.cu:78-129 (kernel) and :175-225 (host).  corr_grad [B,M,2r+1(x),2r+1(y),H,W] float32 is the
    gradient of the tensor altcorr_forward returned; the host un-permutes it, spreads it over the raw (2r+2)^2 window
    with the bilinear weights (g1+g2+g3+g4, fp32) and the kernel scatters g*fmap (un-scaled features, no /4;
    g and the products rounded in the feature dtype; atomics -> order not defined, compared with a tolerance)."""
    r = radius
    D = 2 * r + 2
    B, M, _, H, W = coords.shape
    C = fmap1.shape[2]
    H2, W2 = fmap2.shape[3], fmap2.shape[4]
    dt = fmap1.dtype
    grad = corr_grad.float().permute(0, 1, 3, 2, 4, 5)           # [B,M,a(y),c(x),H,W]   (:190)
    x = coords[:, :, 0, None, None]; y = coords[:, :, 1, None, None]
    dx = x - torch.floor(x); dy = y - torch.floor(y)
    g1 = torch.zeros(B, M, D, D, H, W); g2 = torch.zeros_like(g1); g3 = torch.zeros_like(g1); g4 = torch.zeros_like(g1)
    g1[:, :, 0:D - 1, 0:D - 1] = (1 - dx) * (1 - dy) * grad
    g2[:, :, 0:D - 1, 1:D] = (dx) * (1 - dy) * grad
    g3[:, :, 1:D, 0:D - 1] = (1 - dx) * (dy) * grad
    g4[:, :, 1:D, 1:D] = (dx) * (dy) * grad
    raw = g1 + g2 + g3 + g4                                        # (:203)
    g1o = torch.zeros_like(fmap1, dtype=torch.float64)
    g2o = torch.zeros_like(fmap2, dtype=torch.float64)
    big = 1 << 20
    fxi = torch.nan_to_num(torch.floor(coords[:, :, 0]), nan=0.0, posinf=big, neginf=-big).clamp(-big, big).long()
    fyi = torch.nan_to_num(torch.floor(coords[:, :, 1]), nan=0.0, posinf=big, neginf=-big).clamp(-big, big).long()
    for m in range(M):
        ix = int(ii[m]); jx = int(jj[m])
        f1 = fmap1[:, ix].reshape(B, C, H * W)
        f2 = fmap2[:, jx].reshape(B, C, H2 * W2)
        for a in range(D):
            for c in range(D):
                i1 = fyi[:, m] + (a - r); j1 = fxi[:, m] + (c - r)
                inb = ((i1 >= 0) & (i1 < H2) & (j1 >= 0) & (j1 < W2)).reshape(B, 1, H * W)
                lin = (i1.clamp(0, H2 - 1) * W2 + j1.clamp(0, W2 - 1)).reshape(B, 1, H * W)
                g = raw[:, m, a, c].to(dt).reshape(B, 1, H * W)                               # scalar_t g (:113)
                f2g = torch.gather(f2, 2, lin.expand(B, C, H * W))
                t1 = torch.where(inb, (g * f2g), torch.zeros((), dtype=dt)).double()           # product in scalar_t
                t2 = torch.where(inb, (g * f1), torch.zeros((), dtype=dt)).double()
                g1o[:, ix] += t1.reshape(B, C, H, W)
                g2o[:, jx].reshape(B, C, H2 * W2).scatter_add_(2, lin.expand(B, C, H * W), t2)
    return [g1o.to(dt), g2o.to(dt)]


# ---- Python-side construction (modules/corr.py) -----------------------------------------------

def corr_volume(fmap1, fmap2):
    """CorrBlock.corr, modules/corr.py:63-71: [B,E,C,ht,wd] x2 -> [B,E,ht,wd,ht,wd]."""
    batch, num, dim, ht, wd = fmap1.shape
    a = fmap1.reshape(batch * num, dim, ht * wd) / 4.0
    b = fmap2.reshape(batch * num, dim, ht * wd) / 4.0
    return torch.matmul(a.transpose(1, 2), b).view(batch, num, ht, wd, ht, wd)


def corr_pyramid(fmap1, fmap2, num_levels=4):
    """CorrBlock.__init__, modules/corr.py:24-38: list of [B*E,h1,w1,h2/2^l,w2/2^l]."""
    corr = corr_volume(fmap1, fmap2)
    batch, num, h1, w1, h2, w2 = corr.shape
    corr = corr.reshape(batch * num * h1 * w1, 1, h2, w2)
    pyr = []
    for i in range(num_levels):
        pyr.append(corr.view(batch * num, h1, w1, h2 // 2 ** i, w2 // 2 ** i))
        corr = F.avg_pool2d(corr, 2, stride=2)
    return pyr


def fmap_pyramid(fmaps, num_levels=4):
    """AltCorrBlock.__init__, modules/corr.py:89-101."""
    B, N, C, H, W = fmaps.shape
    f = fmaps.view(B * N, C, H, W)
    pyr = []
    for i in range(num_levels):
        pyr.append(f.view(B, N, C, H // 2 ** i, W // 2 ** i))
        f = F.avg_pool2d(f, 2, stride=2)
    return pyr


def corr_block_lookup(pyramid, coords, radius=3):
    """CorrBlock.__call__, modules/corr.py:40-50.  coords [B,E,ht,wd,2] -> [B,E,L*(2r+1)^2,ht,wd]."""
    batch, num, ht, wd, _ = coords.shape
    c = coords.permute(0, 1, 4, 2, 3).contiguous().view(batch * num, 2, ht, wd)
    outs = []
    for i, vol in enumerate(pyramid):
        corr, = corr_index_forward(vol, c / 2 ** i, radius)
        outs.append(corr.view(batch, num, -1, ht, wd))
    return torch.cat(outs, dim=2)


def altcorr_block_lookup(pyramid, coords, ii, jj, radius=3):
    """AltCorrBlock.__call__, modules/corr.py:104-117."""
    c = coords.permute(0, 1, 4, 2, 3).contiguous()
    outs = []
    for i in range(len(pyramid)):
        corr, = altcorr_forward(pyramid[0], pyramid[i], c / 2 ** i, ii, jj, radius)
        outs.append(corr.flatten(2, 3))
    return torch.stack(outs, dim=2).flatten(2, 3)
