"""run ONE rank's share of the N-GPU weak-scaling BA on a single GPU (no collectives) so that it can be profiled"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from droid_slam_b200 import synth, sharded
world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = "cuda"
s = synth.make_scene(dict(E=512 * world, N=72, ht=48, wd=64, stereo=False, itrs=2, lm=1e-4, ep=0.1))
bounds = sharded.partition_frames(s["ii"], 72, world)
lo, hi = bounds[0]
idx = sharded.shard_edges(s["ii"], lo, hi)
kx = torch.unique(torch.cat([torch.arange(s["t0"], s["t1"]), s["ii"]]))
eta_f = torch.zeros(72, 48, 64); eta_f[kx] = s["eta"]
d = dict(poses=s["poses"], disps=s["disps"], K=s["intrinsics"], sens=s["disps_sens"], tg=s["targets"][idx].contiguous(), wt=s["weights"][idx].contiguous(), eta=eta_f,
         ii=s["ii"][idx].contiguous(), jj=s["jj"][idx].contiguous())
d = {k: v.to(dev) for k, v in d.items()}
print("world", world, "rank0 frames", bounds[0], "edges", idx.numel(), "max out-degree", int(torch.bincount(s["ii"], minlength=72).max()))
drv = sharded.ShardedBA(sharded.CApiEngine(dev))
for it in range(4):
    P, D = d["poses"].clone(), d["disps"].clone()
    drv.run(P, D, d["K"], d["sens"], d["tg"], d["wt"], d["eta"], d["ii"], d["jj"], s["t0"], s["t1"], 2, 1e-4, 0.1, bounds, exchange_disps=False)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(10):
    P, D = d["poses"].clone(), d["disps"].clone()
    drv.run(P, D, d["K"], d["sens"], d["tg"], d["wt"], d["eta"], d["ii"], d["jj"], s["t0"], s["t1"], 2, 1e-4, 0.1, bounds, exchange_disps=False)
e1.record(); torch.cuda.synchronize()
print("ba ms per call (no collectives)", e0.elapsed_time(e1) / 10)
