"""Debug: does a solve at pcg_exit_tol t from lambda0 = 0 depend on what ran before it?  (bench.py tolerance_sweep)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mpcgpu_b200 import _capi
from bench import load_ring
L = _capi.lib()
n, N = 14, 128
host, data = load_ring(N, 16)
dS, dP, dg = (torch.from_numpy(host[k]).cuda() for k in ("S", "Pinv", "gamma"))
lam = torch.zeros(16, n * N, device="cuda")
it = torch.zeros(16, dtype=torch.int32, device="cuda")
fl = torch.zeros(16, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
print("data", data, "stream", stream)
for tol in (1e-5, 5e-5, 1e-4, 5e-4, 1e-3, 1e-4):
    lam.zero_()
    for i in range(16):
        rc = L.gbd_pcg_solve_f32(n, N, dS[i].data_ptr(), dP[i].data_ptr(), dg[i].data_ptr(), lam[i].data_ptr(), 0, 0, 0, 0,
                                 it[i:].data_ptr(), fl[i:].data_ptr(), 167, tol, stream)
        assert rc == 0
    torch.cuda.synchronize()
    print("tol", tol, "iters", it.cpu().numpy().tolist(), "capped", fl.cpu().numpy().tolist(), "|lam|max", float(lam.abs().max()))
