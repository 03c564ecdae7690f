"""scratch timing of ours vs the reference build on the metric config (not the bench)"""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import droid_slam_b200
from droid_slam_b200 import synth
ours = droid_slam_b200.install()
try:
    import droid_backends_ref as ref
except Exception as e:
    ref = None; print("no ref", e)
dev = "cuda"
cfg = sys.argv[1] if len(sys.argv) > 1 else "metric"
dt = torch.float16 if (len(sys.argv) < 3 or sys.argv[2] == "f16") else torch.float32
t = time.time(); s = synth.make_scene(cfg); print("scene", time.time() - t)
t = time.time(); pyr, coords, _ = synth.make_corr_inputs(s, dtype=dt, device=dev); torch.cuda.synchronize(); print("corr inputs", time.time() - t)

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

E = s["ii"].shape[0]; HW = s["cfg"]["ht"] * s["cfg"]["wd"]
sz = 2 if dt == torch.float16 else 4
for be, name in ((ours, "ours"), (ref, "ref")):
    if be is None: continue
    per = []
    for l, vol in enumerate(pyr):
        c = (coords / 2 ** l).contiguous()
        ms = timeit(lambda: be.corr_index_forward(vol, c, 3))
        per.append(ms)
    tot = sum(per); alg = E * HW * (452 * sz + 32)
    print(name, "corr_index levels ms", ["%.3f" % x for x in per], "total %.3f ms  alg GB/s %.1f" % (tot, alg / tot / 1e6))
args = [s[k].to(dev) for k in ("intrinsics", "disps_sens", "targets", "weights", "eta", "ii", "jj")]
res = {}
for be, name in ((ours, "ours"), (ref, "ref")):
    if be is None: continue
    def run():
        P, D = s["poses"].to(dev), s["disps"].to(dev)
        be.ba(P, D, *args, s["t0"], s["t1"], s["itrs"], s["lm"], s["ep"], False)
        return P, D
    ms = timeit(run, n=5, warm=2)
    P, D = run(); torch.cuda.synchronize(); res[name] = (P.cpu(), D.cpu())
    print(name, "ba(itrs=%d) ms %.3f (incl. 2 small H2D copies)" % (s["itrs"], ms))
if "ref" in res:
    print("ba ours vs ref: pose max abs %.3e  disp max abs %.3e" % ((res["ours"][0] - res["ref"][0]).abs().max(), (res["ours"][1] - res["ref"][1]).abs().max()))
