"""world_size-2 `gloo` test (CPU) of the edge-sharded BA host logic: partitioning by source frame, one all-reduce of
the reduced pose system per Gauss-Newton iteration, replicated solve, owner-local depth update, final depth exchange.
The per-rank numerical engine here is the CPU oracle (test infrastructure); on GPUs it is the C ABI (CApiEngine)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
import importlib
oba = importlib.import_module("oracle.ba")   # the module (the package attribute `oracle.ba` is the function)
from droid_slam_b200 import sharded, synth  # noqa: E402


class OracleEngine:
    """oracle-backed stand-in for CApiEngine with the same three phases"""

    def setup(self, poses, disps, intrinsics, disps_sens, targets, weights, eta_by_frame, ii, jj, t0, t1, lm, ep, own):
        self.s = dict(poses=poses, disps=disps, K=intrinsics, sens=disps_sens, tg=targets, wt=weights, eta=eta_by_frame, ii=ii, jj=jj,
                      t0=t0, t1=t1, lm=lm, ep=ep, own=own)

    def build(self):
        s = self.s
        p, d = s["poses"].double(), s["disps"].double()
        self.terms = oracle.ba_edge_terms(p, d, s["K"].double(), s["tg"].double(), s["wt"].double(), s["ii"], s["jj"])
        _, _, _, kx, _ = oracle.ba_graph(s["ii"], s["jj"], s["t0"], s["t1"])
        A, b, self.aux = oracle.ba_system(self.terms, d, s["sens"].double(), s["eta"][kx].double(), s["ii"], s["jj"], s["t0"], s["t1"], False,
                                          torch.float64)
        self.system = torch.cat([A.reshape(-1), b])
        return self.system

    def solve(self):
        s = self.s
        P = s["t1"] - s["t0"]
        n = 6 * P
        A = self.system[:n * n].reshape(n, n); b = self.system[n * n:]
        x, ok = oba._solve(A, b, s["lm"], s["ep"], P)
        aux = self.aux
        kx, Q, w, Erows, pose = aux["kx"], aux["Q"], aux["w"], aux["Erows"], aux["pose"]
        valid = (pose > 0) & (pose < P)
        dw = torch.zeros(Erows.shape[0], Q.shape[1], dtype=torch.float64)
        dw[valid] = torch.einsum("nap,na->np", Erows[valid], x[pose[valid]])
        dz = Q * (w - oba._segsum(dw, aux["ii_exp"], kx))
        lo, hi = s["own"]
        owned = (kx >= lo) & (kx < hi)
        ht, wd = s["disps"].shape[1:]
        s["disps"][kx[owned]] += dz[owned].reshape(-1, ht, wd).to(s["disps"].dtype)
        t_new, q_new = oracle.retr_se3(x, s["poses"][s["t0"]:s["t1"], :3].double(), s["poses"][s["t0"]:s["t1"], 3:].double())
        s["poses"][s["t0"]:s["t1"]] = torch.cat([t_new, q_new], -1).to(s["poses"].dtype)


def _scene():
    return synth.make_scene(dict(E=36, N=9, ht=8, wd=12, stereo=False, itrs=2, lm=1e-4, ep=0.1), seed=5, rgbd=True)


def _eta_by_frame(s):
    kx = torch.unique(torch.cat([torch.arange(s["t0"], s["t1"]), s["ii"]]))
    eta = torch.zeros(s["cfg"]["N"], s["cfg"]["ht"], s["cfg"]["wd"])
    eta[kx] = s["eta"]
    return eta


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = _scene()
    bounds = sharded.partition_frames(s["ii"], s["cfg"]["N"], world)
    lo, hi = bounds[rank]
    idx = sharded.shard_edges(s["ii"], lo, hi)
    poses, disps = s["poses"].double(), s["disps"].double()
    drv = sharded.ShardedBA(OracleEngine())
    drv.run(poses, disps, s["intrinsics"], s["disps_sens"], s["targets"][idx], s["weights"][idx], _eta_by_frame(s), s["ii"][idx], s["jj"][idx],
            s["t0"], s["t1"], 2, s["lm"], s["ep"], bounds)
    out[rank] = (poses, disps, drv.allreduce_bytes, bounds, int(idx.numel()))
    dist.destroy_process_group()


def test_partition_is_contiguous_balanced_and_complete():
    s = _scene()
    for world in (1, 2, 3, 4):
        b = sharded.partition_frames(s["ii"], 9, world)
        assert b[0][0] == 0 and b[-1][1] == 9 and all(b[r][1] == b[r + 1][0] for r in range(world - 1))
        counts = [int(sharded.shard_edges(s["ii"], lo, hi).numel()) for lo, hi in b]
        assert sum(counts) == 36
        assert max(counts) - min(counts) <= 8          # out-degree <= 8 here: balanced to within one frame
    ii = torch.tensor([0, 0, 0, 5])
    assert sharded.partition_frames(ii, 6, 4)[-1][1] == 6


def test_sharded_ba_world2_matches_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    s = _scene()
    P, D = s["poses"].double(), s["disps"].double()
    oracle.ba(P, D, s["intrinsics"], s["disps_sens"], s["targets"], s["weights"], s["eta"], s["ii"], s["jj"], s["t0"], s["t1"], 2, s["lm"],
              s["ep"], False, dtype=torch.float64)
    n = 6 * (s["t1"] - s["t0"])
    for r in range(world):
        poses, disps, nbytes, bounds, ne = out[r]
        assert nbytes == 2 * 8 * (n * n + n)                       # one all-reduce per GN iteration, 8*(36P^2+6P) bytes each
        assert (poses - P).abs().max() < 1e-9                      # replicated solve -> identical poses everywhere
        assert (disps - D).abs().max() < 1e-9                      # owners' depths exchanged after the last iteration
        assert 0 < ne < 36
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])


# ---- property tests of the partitioning host logic (hypothesis) --------------------------------------------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 40), st.integers(1, 8), st.lists(st.integers(0, 39), min_size=0, max_size=200))
def test_partition_properties(n_frames, world, src):
    ii = torch.tensor([s % n_frames for s in src], dtype=torch.long)
    b = sharded.partition_frames(ii, n_frames, world)
    assert len(b) == world and b[0][0] == 0 and b[-1][1] == n_frames
    assert all(lo <= hi for lo, hi in b) and all(b[r][1] == b[r + 1][0] for r in range(world - 1))       # contiguous, ordered, complete
    assert b == sharded.partition_frames(ii.clone(), n_frames, world)                                      # deterministic (every rank computes it)
    shards = [sharded.shard_edges(ii, lo, hi) for lo, hi in b]
    allidx = torch.cat(shards) if shards else torch.zeros(0, dtype=torch.long)
    assert sorted(allidx.tolist()) == list(range(ii.numel()))                                              # every edge owned exactly once
    for (lo, hi), idx in zip(b, shards):
        assert torch.equal(idx, torch.sort(idx).values)                                                    # original edge order kept
        assert bool(((ii[idx] >= lo) & (ii[idx] < hi)).all())
    if ii.numel() and world > 1:
        deg = torch.bincount(ii, minlength=n_frames)
        counts = [int(i.numel()) for i in shards]
        # greedy prefix cuts: no shard exceeds its fair share by more than one frame's out-degree (+1 for the cut rule)
        assert max(counts) <= ii.numel() / world + int(deg.max()) + 1
