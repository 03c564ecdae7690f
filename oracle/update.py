"""CPU restatement of the reference's update operator -- TEST INFRASTRUCTURE ONLY.

Follows, line by line in behaviour:
  UpdateModule.forward   droid_slam/droid_net.py:111-143   (encoders :83-93, heads :95-106)
  ConvGRU.forward        droid_slam/modules/gru.py:19-32
  GraphAgg.forward       droid_slam/droid_net.py:59-75     (torch_scatter.scatter_mean over edges with equal source frame)
  GradientClip           droid_slam/modules/clipping.py     (identity in the forward pass)
Weights are a flat dict with the reference's state_dict names (`corr_encoder.0.weight`, `gru.convz.weight`, `agg.eta.0.bias` ...).
Pinned against the reference module itself: tests/golden/make_update_golden.py imports droid_net.UpdateModule from
/root/reference (with stubs for the absent lietorch / torch_scatter packages), loads the same weights and stores its outputs in
tests/golden/update_module.pt; tests/test_update_cpu.py compares.
"""
import torch
import torch.nn.functional as F

__all__ = ["update_module_forward", "conv_gru_forward", "graph_agg_forward", "scatter_mean_by_source"]


def _conv(w, name, x, pad):
    return F.conv2d(x, w[name + ".weight"], w[name + ".bias"], padding=pad)


def conv_gru_forward(w, net, *inputs, prefix="gru."):
    """modules/gru.py:19-32.  net [B,128,h,w]; inputs concatenated along channels (inp 128, corr 128, flow 64)."""
    inp = torch.cat(inputs, dim=1)
    net_inp = torch.cat([net, inp], dim=1)
    b, c, h, wd = net.shape
    glo = torch.sigmoid(_conv(w, prefix + "w", net, 0)) * net                 # :25
    glo = glo.view(b, c, h * wd).mean(-1).view(b, c, 1, 1)                    # :26 spatial mean -> global context
    z = torch.sigmoid(_conv(w, prefix + "convz", net_inp, 1) + _conv(w, prefix + "convz_glo", glo, 0))      # :28
    r = torch.sigmoid(_conv(w, prefix + "convr", net_inp, 1) + _conv(w, prefix + "convr_glo", glo, 0))      # :29
    q = torch.tanh(_conv(w, prefix + "convq", torch.cat([r * net, inp], dim=1), 1) + _conv(w, prefix + "convq_glo", glo, 0))   # :30
    return (1 - z) * net + z * q                                               # :32


def scatter_mean_by_source(x, ii):
    """torch_scatter.scatter_mean(x, unique_inverse(ii), dim=1) (droid_net.py:64,67): x [B,E,...] -> [B,M',...], M' = #unique(ii),
    slot k = k-th smallest source frame, mean over the edges with that source."""
    uniq, ix = torch.unique(ii, return_inverse=True)
    out = torch.zeros((x.shape[0], uniq.shape[0]) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
    out.index_add_(1, ix, x)
    cnt = torch.bincount(ix, minlength=uniq.shape[0]).to(x.dtype)
    return out / cnt.view(1, -1, *([1] * (x.dim() - 2)))


def graph_agg_forward(w, net, ii, prefix="agg."):
    """droid_net.py:59-75.  net [B,E,128,h,w] -> eta [B,M',h,w] (x0.01 after softplus), upmask [B,M',576,h,w]."""
    batch, num, ch, ht, wd = net.shape
    x = F.relu(_conv(w, prefix + "conv1", net.reshape(batch * num, ch, ht, wd), 1))
    x = scatter_mean_by_source(x.view(batch, num, 128, ht, wd), ii)
    x = x.reshape(-1, 128, ht, wd)
    x = F.relu(_conv(w, prefix + "conv2", x, 1))
    eta = F.softplus(_conv(w, prefix + "eta.0", x, 1)).view(batch, -1, ht, wd)
    upmask = _conv(w, prefix + "upmask.0", x, 0).view(batch, -1, 8 * 8 * 9, ht, wd)
    return 0.01 * eta, upmask


def update_module_forward(w, net, inp, corr, flow=None, ii=None):
    """droid_net.py:111-143.  net, inp [B,E,128,h,w]; corr [B,E,196,h,w]; flow [B,E,4,h,w] (zeros if None).
    Returns net, delta [B,E,h,w,2], weight [B,E,h,w,2] (+ eta, upmask when ii is given)."""
    batch, num, ch, ht, wd = net.shape
    if flow is None:
        flow = torch.zeros(batch, num, 4, ht, wd, dtype=net.dtype, device=net.device)
    out_dim = (batch, num, -1, ht, wd)
    net = net.reshape(batch * num, -1, ht, wd)
    inp = inp.reshape(batch * num, -1, ht, wd)
    corr = corr.reshape(batch * num, -1, ht, wd)
    flow = flow.reshape(batch * num, -1, ht, wd)
    corr = F.relu(_conv(w, "corr_encoder.2", F.relu(_conv(w, "corr_encoder.0", corr, 0)), 1))      # :83-87
    flow = F.relu(_conv(w, "flow_encoder.2", F.relu(_conv(w, "flow_encoder.0", flow, 3)), 1))      # :89-93
    net = conv_gru_forward(w, net, inp, corr, flow)                                                    # :127
    delta = _conv(w, "delta.2", F.relu(_conv(w, "delta.0", net, 1)), 1).view(*out_dim)                 # :102-106,130
    weight = torch.sigmoid(_conv(w, "weight.2", F.relu(_conv(w, "weight.0", net, 1)), 1)).view(*out_dim)   # :95-100,131
    delta = delta.permute(0, 1, 3, 4, 2)[..., :2].contiguous()
    weight = weight.permute(0, 1, 3, 4, 2)[..., :2].contiguous()
    net = net.view(*out_dim)
    if ii is None:
        return net, delta, weight
    eta, upmask = graph_agg_forward(w, net, ii)
    return net, delta, weight, eta, upmask
