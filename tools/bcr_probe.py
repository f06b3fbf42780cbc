import sys, glob, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import mpcgpu_b200 as mp
from mpcgpu_b200 import synth
from oracle import pcg as op
def solve(n,N,S,g):
    dS, dg = torch.from_numpy(np.ascontiguousarray(S)).cuda(), torch.from_numpy(np.ascontiguousarray(g)).cuda()
    lam = torch.zeros(n*N, device="cuda"); mp.solve_direct(n,N,dS,dg,lam); torch.cuda.synchronize(); return lam.cpu().numpy()
for N in (8,16,32,64,128,256,512):
    d = synth.make_systems(14, N, batch=2, seed=40+N, nan_pads=True)
    for i in range(2):
        lam = solve(14,N,d["S"][i],d["gamma"][i]); tr = op.solve_f64(d["S"][i], d["gamma"][i], 14, N)
        A = op.band_to_dense(np.nan_to_num(d["S"][i]),14,N) if N<=128 else None
        print(N, i, "rel", np.abs(lam-tr).max()/np.abs(tr).max(), "res", op.rel_residual(d["S"][i], d["gamma"][i], lam, 14, N), "cond", np.linalg.cond(A) if A is not None else None)
for path in sorted(glob.glob("/root/repo/tests/golden/iiwa_*.npz")):
    z=np.load(path); n,N=int(z["n"]),int(z["N"])
    lam=solve(n,N,z["S"],z["gamma"]); S0=np.nan_to_num(z["S"])
    tr=op.solve_f64(S0,z["gamma"],n,N)
    print(os.path.basename(path), "direct res", op.rel_residual(S0,z["gamma"],lam,n,N), "rel", np.abs(lam-tr).max()/np.abs(tr).max(), "ref pcg res", op.rel_residual(S0,z["gamma"],z["run0_lam"],n,N), "pcg rel", np.abs(z["run0_lam"]-tr).max()/np.abs(tr).max(), "cond", np.linalg.cond(op.band_to_dense(S0,n,N)) if N<=128 else None)
