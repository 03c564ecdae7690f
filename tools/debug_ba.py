import os, sys, ctypes, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from droid_slam_b200 import synth, c_api
from util import c_ba
L = c_api.load(); dev = "cuda"
torch.set_num_threads(16)
cfgs = {"dense16": dict(E=240, N=16, ht=16, wd=24, stereo=False, itrs=1, lm=1e-4, ep=0.1), "metric": synth.CONFIGS["metric"], "mid": dict(E=200, N=30, ht=24, wd=32, stereo=False, itrs=1, lm=1e-4, ep=0.1)}
for name in sys.argv[1:] or ["dense16", "mid", "metric"]:
    s = synth.make_scene(cfgs[name])
    N, ht, wd = s["disps"].shape; E = s["ii"].shape[0]; t0, t1 = s["t0"], s["t1"]; P = t1 - t0; n = 6 * P
    deg = torch.bincount(s["ii"], minlength=N)
    P_, D_ = s["poses"].to(dev), s["disps"].to(dev)
    dx, dz, M, st, (a, ws) = c_ba(L, P_, D_, s["intrinsics"].to(dev), s["disps_sens"].to(dev), s["targets"].to(dev), s["weights"].to(dev), s["eta"].to(dev),
                                  s["ii"].to(dev), s["jj"].to(dev), t0, t1, 0, s["lm"], s["ep"], False, s["M"])
    # iterations=0 ran prepare only; now build once and read the system
    c_api.check(L.dba_ba_build(ctypes.byref(a)), "build")
    torch.cuda.synchronize()
    off = L.dba_ba_system_offset(N, E, ht, wd, t0, t1)
    sysbuf = ws[off:off + 8 * (n * n + n)].view(torch.float64).cpu()
    H = sysbuf[:n * n].reshape(n, n); b = sysbuf[n * n:]
    p64, d64 = s["poses"].double(), s["disps"].double()
    T = oracle.ba_edge_terms(p64, d64, s["intrinsics"].double(), s["targets"].double(), s["weights"].double(), s["ii"], s["jj"])
    A, bb, aux = oracle.ba_system(T, d64, s["disps_sens"].double(), s["eta"].double(), s["ii"], s["jj"], t0, t1, False, torch.float64)
    Hl = torch.tril(H); Al = torch.tril(A)
    err = (Hl - Al).abs()
    blk = err.reshape(P, 6, P, 6).amax(dim=(1, 3))
    worst = torch.nonzero(blk > 1e-3 * Al.abs().max())
    print(name, "max deg", int(deg.max()), "M", M, "status", st, "| H err", float(err.max()), "scale", float(Al.abs().max()), "| b err", float((b - bb).abs().max()), "scale", float(bb.abs().max()))
    print("   bad blocks (pose a,b):", worst[:12].tolist(), "n bad", worst.shape[0])
    c_api.check(L.dba_ba_solve(ctypes.byref(a)), "solve"); torch.cuda.synchronize()
    x, ok = oracle.ba.__globals__["_solve"](A, bb, s["lm"], s["ep"], P)
    print("   dx err", float((dx.cpu().double() - x).abs().max()), "dx scale", float(x.abs().max()), "ok", ok)
