import os, sys, ctypes, torch
os.environ["DBA_CHOL_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from droid_slam_b200 import c_api
L = c_api.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 426
g = torch.Generator().manual_seed(0)
A = torch.randn(n, n + 8, generator=g, dtype=torch.float64); H = (A @ A.t() + 1e-3 * torch.eye(n, dtype=torch.float64)).cuda(); b = torch.randn(n, generator=g, dtype=torch.float64).cuda()
ws = torch.empty(L.dba_solve_workspace_bytes(n), dtype=torch.uint8, device="cuda"); x = torch.zeros(n, device="cuda"); fail = torch.zeros(1, dtype=torch.int32, device="cuda")
for it in range(3):
    L.dba_solve_spd(ctypes.c_void_p(H.data_ptr()), ctypes.c_void_p(b.data_ptr()), n, 1e-4, 0.1, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(fail.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws.numel(), None)
    torch.cuda.synchronize()
