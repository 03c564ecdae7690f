"""First-contact diagnostics for the tensor-core update operator on a GPU box: every building-block shape separately (each in its own
try block and with a watchdog-friendly order: small first), then the full operator, printing max errors per output."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
import droid_slam_b200
from droid_slam_b200 import synth
from droid_slam_b200.update import UpdateModule, pack_update_weights, _taps
import oracle
from update_emul import emulate, conv_taps

be = droid_slam_b200.install()
DEV = "cuda:0"


def conv_case(E, ht, wd, c0, c1, ks, n, relu):
    g = torch.Generator().manual_seed(1)
    x0 = torch.randn(E, ht, wd, c0, generator=g).half()
    x1 = torch.randn(E, ht, wd, c1, generator=g).half() if c1 else None
    ct = c0 + c1
    w = (torch.randn(n, ct, ks, ks, generator=g) * (1.0 / (ct * ks * ks)) ** 0.5).half()
    b = 0.1 * torch.randn(n, generator=g)
    xin = x0 if x1 is None else torch.cat([x0, x1], -1)
    ref = F.conv2d(xin.float().permute(0, 3, 1, 2), w.float(), b, padding=ks // 2)
    ref = (F.relu(ref) if relu else ref).permute(0, 2, 3, 1)
    parts = [_taps(w[:, :c0].float(), 64 * ((c0 + 63) // 64))]
    if c1:
        parts.append(_taps(w[:, c0:].float(), 64 * ((c1 + 63) // 64)))
    wpk = torch.cat(parts, 2).half().contiguous()
    t0 = time.time()
    got = be.conv_nhwc(x0.to(DEV), x1.to(DEV) if c1 else None, wpk.to(DEV), b.to(DEV), ks, relu)
    torch.cuda.synchronize()
    d = (got.float().cpu() - ref).abs()
    print("conv E=%d %dx%d c=%d+%d k=%d N=%d: max err %.3e (ref max %.2f) nan=%d  %.2fs" % (E, ht, wd, c0, c1, ks, n, float(d.max()), float(ref.abs().max()),
          int(torch.isnan(got).sum()), time.time() - t0), flush=True)
    if float(d.max()) > 1e-2:
        bad = (d > 1e-2).nonzero()
        print("   first bad idx", bad[:5].tolist(), "n_bad", bad.shape[0], "of", d.numel(), flush=True)
        # per-position error pattern: which (y,x) and which channels
        print("   err by y:", d.amax((0, 2, 3))[:12].tolist())
        print("   err by x:", d.amax((0, 1, 3))[:12].tolist())
        print("   err by n (first 8 / last 8):", d.amax((0, 1, 2))[:8].tolist(), d.amax((0, 1, 2))[-8:].tolist())


for case in [(1, 8, 64, 64, 0, 1, 128, False), (1, 8, 64, 128, 0, 1, 128, False), (1, 8, 64, 64, 0, 3, 128, False), (2, 16, 64, 128, 0, 3, 128, True),
             (2, 16, 64, 128, 320, 3, 256, False), (2, 8, 64, 128, 0, 3, 384, True), (2, 24, 96, 128, 0, 3, 64, True), (2, 10, 40, 64, 0, 3, 32, False),
             (150, 8, 64, 64, 0, 3, 128, True)]:
    try:
        conv_case(*case)
    except Exception:
        traceback.print_exc()

for (E, ht, wd, n_src) in [(4, 16, 64, 2), (5, 24, 32, 3), (3, 48, 64, 2)]:
    try:
        w = synth.make_update_weights(0)
        net, inp, corr, flow, ii = synth.make_update_inputs(E=E, ht=ht, wd=wd, seed=E, n_src=n_src)
        mod = UpdateModule().to(DEV); mod.load_state_dict(w)
        with torch.no_grad():
            got = mod(net.half().to(DEV), inp.half().to(DEV), corr.half().to(DEV), flow.to(DEV), ii.to(DEV))
        torch.cuda.synchronize()
        ref = oracle.update_module_forward(w, net.half().float(), inp.half().float(), corr.half().float(), flow, ii)
        for k, a, b in zip(("net", "delta", "weight", "eta", "upmask"), got, ref):
            d = (a.float().cpu() - b).abs()
            print("update E=%d %dx%d %-7s max err %.3e mean %.3e (ref max %.3f) nan=%d" % (E, ht, wd, k, float(d.max()), float(d.mean()), float(b.abs().max()), int(torch.isnan(a).sum())), flush=True)
    except Exception:
        traceback.print_exc()

# timing at the metric size
try:
    E, ht, wd = int(os.environ.get("UPD_E", 512)), 48, 64
    w = synth.make_update_weights(0)
    mod = UpdateModule().to(DEV); mod.load_state_dict(w)
    g = torch.Generator(device=DEV).manual_seed(0)
    net = torch.tanh(torch.randn(1, E, 128, ht, wd, device=DEV, generator=g)).half()
    inp = torch.relu(torch.randn(1, E, 128, ht, wd, device=DEV, generator=g)).half()
    corr = torch.randn(1, E, 196, ht, wd, device=DEV, generator=g).half()
    flow = 4 * torch.randn(1, E, 4, ht, wd, device=DEV, generator=g)
    ii = (torch.arange(E, device=DEV) // 8)
    with torch.no_grad():
        for _ in range(2):
            out = mod(net, inp, corr, flow, ii)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = mod(net, inp, corr, flow, ii)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = (14.03e9 * E + 1.37e9 * int(ii.max() + 1)) * (ht * wd / 3072.0)
    print("update operator E=%d: %.3f ms/call, %.1f TFLOP/s (nan=%d)" % (E, ms, fl / ms / 1e9, int(torch.isnan(out[0]).sum())), flush=True)
except Exception:
    traceback.print_exc()
