"""Runs a few eager (plain-launch) fused PNCG iterations on the config-2 style model (for ncu captures / launch lists).
usage: python tools/prof_pncg.py --n 58 --iters 3 [--graph 0|1|2]"""
import argparse, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from bench import build_mesh, cuda_potential

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=58); ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--graph", type=int, default=0); ap.add_argument("--potentials", default="snh,arap")
ap.add_argument("--dtype", default="f32")
a = ap.parse_args()
dtype = torch.float32 if a.dtype == "f32" else torch.float64
dev = torch.device("cuda", 0)
from apple_b200.common import FIXED_MASK, FIXED_VALUE
from apple_b200.forward import Forward, ModelBuilder
from apple_b200.optim import PNCG
from apple_b200.optim.pncg import ConvergenceCriteria
from apple_b200.warp.fem import fuse_potentials
t0 = time.perf_counter()
mesh, u, p = build_mesh(a.n)
V = mesh.n_points
pots = fuse_potentials({k: cuda_potential(k, mesh, dtype, name=k) for k in a.potentials.split(",")})
print(f"setup: {time.perf_counter() - t0:.2f} s for {mesh.n_cells} tets")
builder = ModelBuilder(dtype=dtype, device=dev)
builder.add_vertices(mesh)
fixed = np.zeros((V, 3), dtype=bool); fixed[mesh.points[:, 2] == 0.0] = True
mesh.point_data[FIXED_MASK.vtk] = fixed
mesh.point_data[FIXED_VALUE.vtk] = np.zeros((V, 3))
builder.add_fixed(mesh)
for pot in pots.values():
    builder.add_potential(pot)
model = builder.finalize()
h = 1.0 / a.n
u0 = np.ascontiguousarray(0.05 * h * np.sin(7.0 * mesh.points[:, [1, 2, 0]] + 0.3)); u0[fixed] = 0.0
crit = ConvergenceCriteria(max_steps=a.iters + 20, target_relative_gradient_norm=0.0)
fwd = Forward(model, optimizer=PNCG(criteria=crit, check_every=a.iters, use_graph=a.graph))
fwd.state.u = torch.as_tensor(u0, dtype=dtype, device=dev)
opt_state = fwd.optimizer.init(fwd.problem, fwd.state, fwd.free)
state = opt_state.step(fwd.problem, fwd.state, a.iters)
torch.cuda.synchronize()
print("pncg iterations done:", a.iters, "accepted", opt_state.n_accepted)
