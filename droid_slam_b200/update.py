"""Host-side mirror of the reference's update operator: `UpdateModule`, `ConvGRU`, `GraphAgg` with the
reference's constructor-free interface, submodule / parameter names (a DROID checkpoint's `update.*` entries load with
`load_state_dict`) and return values (droid_slam/droid_net.py:46-143, droid_slam/modules/gru.py:5-32).

STATUS (round 1): this is the LIBRARY baseline of the row -- the convolutions go through torch (cuDNN on the GPU), not through
hand-written tensor-core kernels; what is restructured for the GPU is the launch count and the memory traffic around them:
  * `convz`/`convr` (same 448-channel input) run as ONE 256-output convolution, the three global-context 1x1 convolutions as one
    384-output matrix product on the [B,128] context vector, the two head stems (`delta.0`, `weight.0`) as one 256-output
    convolution: 19 convolutions -> 14 launches, the GRU input is concatenated once;
  * `scatter_mean` over edges with equal source frame is an index_add segment mean (no torch_scatter dependency);
  * optional channels_last + fp16 autocast like the reference's call site (factor_graph.py:214).
The fused tcgen05 implicit-GEMM GRU replaces the convolution calls behind this same interface.
The CPU oracle (oracle/update.py) is pinned bit-exactly against the reference module; tests/test_update_cpu.py holds this module to it.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["ConvGRU", "GraphAgg", "UpdateModule", "segment_mean"]


def _conv(name):
    """the convolution `name` of the update operator with the reference's shape (table in synth.UPDATE_SHAPES), 'same' padding"""
    from .synth import UPDATE_SHAPES
    co, ci, k = UPDATE_SHAPES[name]
    return nn.Conv2d(ci, co, k, padding=k // 2)


def segment_mean(x, ii):
    """mean over dim 1 of the entries with equal ii; slots ordered by ascending ii (== scatter_mean(x, unique_inverse(ii), dim=1))"""
    uniq, ix = torch.unique(ii, return_inverse=True)
    out = torch.zeros((x.shape[0], uniq.shape[0]) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
    out.index_add_(1, ix, x)
    cnt = torch.bincount(ix, minlength=uniq.shape[0]).to(x.dtype)
    return out / cnt.view(1, -1, *([1] * (x.dim() - 2)))


class ConvGRU(nn.Module):
    """modules/gru.py:5-32 (same parameter names)"""

    def __init__(self, prefix="gru."):
        super().__init__()
        for name in ("convz", "convr", "convq", "w", "convz_glo", "convr_glo", "convq_glo"):      # shapes: synth.UPDATE_SHAPES
            setattr(self, name, _conv(prefix + name))

    def forward(self, net, *inputs):
        inp = torch.cat(inputs, dim=1)
        net_inp = torch.cat([net, inp], dim=1)
        b, c, h, w = net.shape
        glo = (torch.sigmoid(self.w(net)) * net).view(b, c, h * w).mean(-1)                   # [b, c] global context (:25-26)
        # the three 1x1 convolutions of the context vector as one product
        wg = torch.cat([self.convz_glo.weight, self.convr_glo.weight, self.convq_glo.weight], 0).view(3 * c, c)
        bg = torch.cat([self.convz_glo.bias, self.convr_glo.bias, self.convq_glo.bias], 0)
        g = F.linear(glo, wg.to(glo.dtype), bg.to(glo.dtype)).view(b, 3 * c, 1, 1)
        # z and r share their input: one 2c-output convolution
        zr = F.conv2d(net_inp, torch.cat([self.convz.weight, self.convr.weight], 0), torch.cat([self.convz.bias, self.convr.bias], 0), padding=1)
        z = torch.sigmoid(zr[:, :c] + g[:, :c])                                               # :28
        r = torch.sigmoid(zr[:, c:] + g[:, c:2 * c])                                          # :29
        q = torch.tanh(self.convq(torch.cat([r * net, inp], dim=1)) + g[:, 2 * c:])          # :30
        return (1 - z) * net + z * q                                                          # :32


class GraphAgg(nn.Module):
    """droid_net.py:46-75 (same parameter names; GradientClip is the identity in the forward pass and has no parameters)"""

    def __init__(self, prefix="agg."):
        super().__init__()
        self.conv1, self.conv2 = _conv(prefix + "conv1"), _conv(prefix + "conv2")
        self.eta = nn.Sequential(_conv(prefix + "eta.0"), nn.Identity(), nn.Softplus())          # slot 1: the reference's GradientClip
        self.upmask = nn.Sequential(_conv(prefix + "upmask.0"))

    def forward(self, net, ii):
        batch, num, ch, ht, wd = net.shape
        x = F.relu(self.conv1(net.reshape(batch * num, ch, ht, wd)))
        x = segment_mean(x.view(batch, num, 128, ht, wd), ii).reshape(-1, 128, ht, wd)
        x = F.relu(self.conv2(x))
        eta = self.eta(x).view(batch, -1, ht, wd)
        upmask = self.upmask(x).view(batch, -1, 8 * 8 * 9, ht, wd)
        return .01 * eta, upmask


class UpdateModule(nn.Module):
    """droid_net.py:78-143: forward(net, inp, corr, flow=None, ii=None, jj=None) -> net, delta, weight[, eta, upmask]"""

    def __init__(self):
        super().__init__()
        relu = lambda: nn.ReLU(inplace=True)
        self.corr_encoder = nn.Sequential(_conv("corr_encoder.0"), relu(), _conv("corr_encoder.2"), relu())
        self.flow_encoder = nn.Sequential(_conv("flow_encoder.0"), relu(), _conv("flow_encoder.2"), relu())
        self.weight = nn.Sequential(_conv("weight.0"), relu(), _conv("weight.2"), nn.Identity(), nn.Sigmoid())
        self.delta = nn.Sequential(_conv("delta.0"), relu(), _conv("delta.2"), nn.Identity())
        self.gru = ConvGRU()
        self.agg = GraphAgg()

    def forward(self, net, inp, corr, flow=None, ii=None, jj=None):
        batch, num, ch, ht, wd = net.shape
        if flow is None:
            flow = torch.zeros(batch, num, 4, ht, wd, device=net.device, dtype=net.dtype)
        output_dim = (batch, num, -1, ht, wd)
        net = net.reshape(batch * num, -1, ht, wd)
        inp = inp.reshape(batch * num, -1, ht, wd)
        corr = corr.reshape(batch * num, -1, ht, wd)
        flow = flow.reshape(batch * num, -1, ht, wd)
        corr = self.corr_encoder(corr)
        flow = self.flow_encoder(flow)
        net = self.gru(net, inp, corr, flow)
        # both head stems read the same hidden state: one 256-output convolution
        stem = F.relu(F.conv2d(net, torch.cat([self.delta[0].weight, self.weight[0].weight], 0),
                               torch.cat([self.delta[0].bias, self.weight[0].bias], 0), padding=1))
        delta = self.delta[2](stem[:, :128]).view(*output_dim)
        weight = torch.sigmoid(self.weight[2](stem[:, 128:])).view(*output_dim)
        delta = delta.permute(0, 1, 3, 4, 2)[..., :2].contiguous()
        weight = weight.permute(0, 1, 3, 4, 2)[..., :2].contiguous()
        net = net.view(*output_dim)
        if ii is not None:
            eta, upmask = self.agg(net, ii.to(net.device))
            return net, delta, weight, eta, upmask
        return net, delta, weight
