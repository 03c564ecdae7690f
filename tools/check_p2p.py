"""torchrun --nproc-per-node 2 tools/check_p2p.py : fused peer-to-peer reduction vs NCCL all-reduce vs unsharded BA"""
import os, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import droid_slam_b200
from droid_slam_b200 import synth, sharded
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
be = droid_slam_b200.install()
s = synth.make_scene(dict(E=512 * world, N=72, ht=48, wd=64, stereo=False, itrs=2, lm=1e-4, ep=0.1))
bounds = sharded.partition_frames(s["ii"], 72, world)
lo, hi = bounds[rank]; idx = sharded.shard_edges(s["ii"], lo, hi)
kx = torch.unique(torch.cat([torch.arange(s["t0"], s["t1"]), s["ii"]]))
eta_f = torch.zeros(72, 48, 64); eta_f[kx] = s["eta"]
g = {k: v.to(dev) for k, v in dict(K=s["intrinsics"], sens=s["disps_sens"], tg=s["targets"][idx].contiguous(), wt=s["weights"][idx].contiguous(), eta=eta_f,
                                   ii=s["ii"][idx].contiguous(), jj=s["jj"][idx].contiguous()).items()}
res = {}
p2p = sharded.P2PSystem(6 * (s["t1"] - s["t0"]), dev)
for mode in ("nccl", "p2p", "p2p"):
    drv = sharded.ShardedBA(sharded.CApiEngine(dev), p2p=(p2p if mode == "p2p" else None))
    def run():
        P, D = s["poses"].to(dev), s["disps"].to(dev)
        drv.run(P, D, g["K"], g["sens"], g["tg"], g["wt"], g["eta"], g["ii"], g["jj"], s["t0"], s["t1"], 2, 1e-4, 0.1, bounds, exchange_disps=True)
        return P, D
    for _ in range(3): run()
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): P, D = run()
    e1.record(); torch.cuda.synchronize()
    res[mode] = (P.cpu(), D.cpu(), e0.elapsed_time(e1) / 10)
# the same sharded step captured once into a CUDA graph (device-resident epoch) and replayed
drv = sharded.ShardedBA(sharded.CApiEngine(dev), p2p=p2p)
P0, D0 = s["poses"].to(dev), s["disps"].to(dev)
Pg, Dg = P0.clone(), D0.clone()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph, stream=side):
    Pg.copy_(P0); Dg.copy_(D0)
    drv.run(Pg, Dg, g["K"], g["sens"], g["tg"], g["wt"], g["eta"], g["ii"], g["jj"], s["t0"], s["t1"], 2, 1e-4, 0.1, bounds, exchange_disps=True)
torch.cuda.current_stream().wait_stream(side)
for _ in range(3): graph.replay()
torch.cuda.synchronize(); dist.barrier()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): graph.replay()
e1.record(); torch.cuda.synchronize()
res["p2p-graph"] = (Pg.cpu(), Dg.cpu(), e0.elapsed_time(e1) / 10)
if rank == 0:
    args = [s[k].to(dev) for k in ("intrinsics", "disps_sens", "targets", "weights", "eta", "ii", "jj")]
    P, D = s["poses"].to(dev), s["disps"].to(dev)
    be.ba(P, D, *args, s["t0"], s["t1"], 2, 1e-4, 0.1, False)
    torch.cuda.synchronize()
    for mode in ("nccl", "p2p", "p2p-graph"):
        print(mode, "ms/call %.3f" % res[mode][2], "| vs unsharded: pose %.2e disp %.2e" % (float((res[mode][0] - P.cpu()).abs().max()), float((res[mode][1] - D.cpu()).abs().max())))
    print("p2p-graph vs nccl: pose %.2e disp %.2e" % (float((res["p2p-graph"][0] - res["nccl"][0]).abs().max()), float((res["p2p-graph"][1] - res["nccl"][1]).abs().max())))
    print("p2p vs nccl: pose %.2e disp %.2e" % (float((res["p2p"][0] - res["nccl"][0]).abs().max()), float((res["p2p"][1] - res["nccl"][1]).abs().max())))
torch.cuda.synchronize(); dist.barrier()
del graph          # a captured NCCL collective must go before the communicator does
dist.destroy_process_group()
