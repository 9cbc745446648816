"""Runs one operator configuration a few times (for ncu captures).
usage: python tools/prof_one.py --kind snh --dtype f32 --ops 11 --scatter 0 --ld 3 --n 58 --reps 3"""
import argparse, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from bench import build_mesh, cuda_potential

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="snh"); ap.add_argument("--dtype", default="f32")
ap.add_argument("--ops", type=int, default=11); ap.add_argument("--scatter", type=int, default=0)
ap.add_argument("--ld", type=int, default=3); ap.add_argument("--n", type=int, default=58)
ap.add_argument("--reps", type=int, default=3); ap.add_argument("--setup", default="host", choices=["host", "device"])
a = ap.parse_args()
dt = torch.float32 if a.dtype == "f32" else torch.float64
from apple_b200 import _lib, config
import time
t0 = time.perf_counter()
if a.setup == "device":
    # mesh, materials and fields generated on the GPU, potential built by apl_fem_create_from_mesh
    from bench import device_potentials, device_workload
    wl = device_workload(a.n, 1, 0, torch.device("cuda", 0), dt)
    pot = list(device_potentials(wl, ["snh", "arap"] if a.kind == "fused" else [a.kind], dt, fuse=True).values())[0]
    V, u, p = wl.mesh.n_points, wl.u.cpu().numpy(), wl.p.cpu().numpy()
    n_cells = wl.mesh.n_cells
else:
    mesh, u, p = build_mesh(a.n)
    V, n_cells = mesh.n_points, mesh.n_cells
    if a.kind == "fused":
        from apple_b200.warp.fem import fuse_potentials
        pot = list(fuse_potentials({k: cuda_potential(k, mesh, dt, name=k) for k in ("snh", "arap")}).values())[0]
    else:
        pot = cuda_potential(a.kind, mesh, dt)
torch.cuda.synchronize()
print(f"setup ({a.setup}): {time.perf_counter() - t0:.2f} s for {n_cells} tets")
ud = torch.zeros((V, a.ld), dtype=dt, device="cuda"); ud[:, :3] = torch.as_tensor(u, dtype=dt)
pd = torch.zeros((V, a.ld), dtype=dt, device="cuda"); pd[:, :3] = torch.as_tensor(p, dtype=dt)
outs = {k: torch.zeros((V, a.ld), dtype=dt, device="cuda") for k in ("grad", "diag", "prod")}
fun = torch.zeros(1, dtype=dt, device="cuda"); quad = torch.zeros(1, dtype=dt, device="cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
ts = []
for i in range(a.reps):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pot.eval(a.ops, ud, pd, fun=fun, quad=quad, grad=outs["grad"], diag=outs["diag"], prod=outs["prod"], scatter=a.scatter)
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(a, "ms:", ts, "Gtets/s:", n_cells / min(ts) / 1e6)
