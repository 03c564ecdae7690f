"""experiment: corr_index_forward time per level vs cudaLimitMaxL2FetchGranularity (32/64/128)"""
import os, sys, ctypes, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import droid_slam_b200
from droid_slam_b200 import synth, c_api
be = droid_slam_b200.install(); L = c_api.load()
dev = "cuda"
dt = torch.float16 if (len(sys.argv) < 2 or sys.argv[1] == "f16") else torch.float32
E = int(sys.argv[2]) if len(sys.argv) > 2 else 512
grans = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [64, 32, 128, 32]
s = synth.make_scene(dict(E=E, N=72, ht=48, wd=64, stereo=False, itrs=2, lm=1e-4, ep=0.1))
pyr, coords, _ = synth.make_corr_inputs(s, dtype=dt, device=dev)
cl = [(coords / 2 ** l).contiguous() for l in range(4)]
def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
sz = 2 if dt == torch.float16 else 4
print("default granularity", L.dba_get_l2_fetch_granularity())
for g in grans:
    rc = L.dba_set_l2_fetch_granularity(g)
    per = [timeit(lambda l=l: be.corr_index_forward(pyr[l], cl[l], 3)) for l in range(4)]
    tot = sum(per); alg = E * 3072 * (452 * sz + 32)
    print(json.dumps({"gran_set": g, "rc": rc, "gran_now": L.dba_get_l2_fetch_granularity(), "levels_ms": per, "total_ms": tot, "alg_GBps": alg / tot / 1e6}))
