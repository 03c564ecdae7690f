"""Timeline of one CTA of ba_schur_tc_kernel (debug build tools/bin/libdroid_b200_timing.so, -DDBA_TC_TIMING): per-chunk phases
of producer warp 0 and of the MMA-issuing thread, in microseconds.  Run on the GPU box."""
import os, sys, ctypes, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from droid_slam_b200 import c_api, synth
c_api.lib_path = lambda: os.path.join(ROOT, "tools", "bin", "libdroid_b200_timing.so")
L = c_api.load()
import util
s = synth.make_scene(synth.CONFIGS["metric"])
dev = "cuda"
g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in s.items()}
for it in range(3):
    P, D = g["poses"].clone(), g["disps"].clone()
    util.c_ba(L, P, D, g["intrinsics"], g["disps_sens"], g["targets"], g["weights"], g["eta"], g["ii"], g["jj"], s["t0"], s["t1"], 1, s["lm"], s["ep"], False, s["M"])
    torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 8192)()
L.dba_debug_tc_timing(buf, 8192)
t = list(buf)
nch, R6 = int(t[1]), int(t[2])
t0 = t[7]
us = lambda x: (x - t0) / 1e3
print("nchunks %d R6 %d | row list done 0.0, setup done %.1f, loop start %.1f, drains done %.1f, epilogue done %.1f, exit barrier %.1f" % (nch, R6, us(t[0]), us(t[3]), us(t[4]), us(t[5]), us(t[6])))
print("producer warp 0: chunk: start | copy-wait  empty-wait  transform  fence+arrive  issue  drain   || MMA thread: full-wait  acc-wait  issue+commit (at)")
for c in range(nch):
    b = 16 + 8 * c; mm = 4096 + 4 * c
    p = [t[b + k] for k in range(7)]
    q = [t[mm + k] for k in range(4)]
    print("%3d: %7.1f | %5.2f %5.2f %5.2f %5.2f %5.2f %5.2f || %5.2f %5.2f %5.2f (%.1f)" % (c, us(p[0]), (p[1]-p[0])/1e3, (p[2]-p[1])/1e3, (p[3]-p[2])/1e3, (p[4]-p[3])/1e3, (p[5]-p[4])/1e3, (p[6]-p[5])/1e3,
          (q[1]-q[0])/1e3, (q[2]-q[1])/1e3, (q[3]-q[2])/1e3, us(q[3])))
