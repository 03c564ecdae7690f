"""Test helper: the dataflow of csrc/update_op.cu restated with torch on the PACKED weights (channels-last tensors, one matrix product
per tap, the same fusions: z|r as one 256-wide product, the 7x7 flow convolution as a 196-wide K, block-diagonal heads, 3x192 upmask).
It exists to validate `pack_update_weights` and the kernel sequence on CPU against the oracle before anything runs on a GPU."""
import torch
import torch.nn.functional as F


def conv_taps(x, wpk, bias, ks):
    """x [E,H,W,K] channels-last, wpk [ks*ks, N, Kpad] -> [E,H,W,N]: sum over taps of shift(x) @ w[tap]^T (zero padding)"""
    E, H, W, K = x.shape
    kpad = wpk.shape[2]
    if kpad > K:
        x = torch.cat([x, x.new_zeros(E, H, W, kpad - K)], -1)
    pad = ks // 2
    xp = F.pad(x, (0, 0, pad, pad, pad, pad))
    out = None
    for dy in range(ks):
        for dx in range(ks):
            t = xp[:, dy:dy + H, dx:dx + W] @ wpk[dy * ks + dx].float().t()
            out = t if out is None else out + t
    return out + bias[: out.shape[-1]]


def gather9(y, no):
    """y [E,H,W,9*no] per-tap partial sums -> [E,H,W,no]: out[p] = sum_t y[p + shift_t][t*no:(t+1)*no] (zero outside)"""
    E, H, W, _ = y.shape
    yp = F.pad(y, (0, 0, 1, 1, 1, 1))
    out = 0
    for t in range(9):
        dy, dx = t // 3, t % 3
        out = out + yp[:, dy:dy + H, dx:dx + W, t * no:(t + 1) * no]
    return out


def emulate(pk, net, inp, corr, flow, seg, n_src, round16=False):
    """net/inp [E,128,H,W], corr [E,196,H,W], flow [E,4,H,W] or None, seg [E] or None -> net' [E,H,W,128], delta, weight [E,H,W,2], eta [n_src,H,W], upmask [n_src,576,H,W]"""
    r = (lambda t: t.half().float()) if round16 else (lambda t: t)
    pk = {k: v.float() for k, v in pk.items()}
    E, _, H, W = net.shape
    h = r(net.permute(0, 2, 3, 1).float())
    x_inp = r(inp.permute(0, 2, 3, 1).float())
    cc = r(corr.permute(0, 2, 3, 1).float())
    if flow is None:
        flow = torch.zeros(E, 4, H, W)
    fp = F.pad(flow.float(), (3, 3, 3, 3))
    cols = [fp[:, :, dy:dy + H, dx:dx + W].permute(0, 2, 3, 1) for dy in range(7) for dx in range(7)]
    f0 = r(torch.cat(cols, -1))                                                   # [E,H,W,196], K = (dy*7+dx)*4 + c
    c1 = r(F.relu(conv_taps(cc, pk["w_corr0"], pk["b_corr0"], 1)))
    c2 = r(F.relu(conv_taps(c1, pk["w_corr2"], pk["b_corr2"], 3)))
    f1 = r(F.relu(conv_taps(f0, pk["w_flow0"], pk["b_flow0"], 1)))
    f2 = r(F.relu(conv_taps(f1, pk["w_flow2"], pk["b_flow2"], 3)))
    x = torch.cat([x_inp, c2, f2], -1)                                            # 320 channels
    g = (torch.sigmoid(conv_taps(h, pk["w_gate"], pk["b_gate"], 1)) * h).mean((1, 2))          # [E,128]
    glo = g @ pk["w_glo"].t() + pk["b_glo"]                                        # [E,384]
    zr = torch.sigmoid(conv_taps(torch.cat([h, x], -1), pk["w_zr"], pk["b_zr"], 3) + glo[:, None, None, :256])
    z, rr = r(zr[..., :128]), zr[..., 128:]
    rh = r(rr * h)
    q = torch.tanh(conv_taps(torch.cat([rh, x], -1), pk["w_q"], pk["b_q"], 3) + glo[:, None, None, 256:])
    hn = r((1 - z) * h + z * q)
    n_stem = 384 if seg is not None else 256
    s = r(F.relu(conv_taps(hn, pk["w_stem"][:, :n_stem], pk["b_stem"][:n_stem], 3)))
    hd = gather9(conv_taps(s[..., :256], pk["w_heads"], pk["b_zero"], 1)[..., :36], 4) + pk["b_heads"]
    delta, weight = hd[..., 0:2], torch.sigmoid(hd[..., 2:4])
    if seg is None:
        return hn, delta, weight
    a1 = s[..., 256:384]
    am = torch.zeros(n_src, H, W, 128).index_add_(0, seg, a1) / torch.bincount(seg, minlength=n_src).float().view(-1, 1, 1, 1)
    am = r(am)
    b2 = r(F.relu(conv_taps(am, pk["w_agg2"], pk["b_agg2"], 3)))
    eta = 0.01 * F.softplus(gather9(conv_taps(b2, pk["w_eta"], pk["b_zero"], 1)[..., :9], 1)[..., 0] + pk["b_eta"][0])
    up = conv_taps(b2, pk["w_upmask"], pk["b_upmask"], 1).permute(0, 3, 1, 2)
    return hn, delta, weight, eta, up
