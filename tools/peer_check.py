"""Peer-memory exchange with MORE RANKS THAN GPUs: world ranks (gloo for the host plumbing) share the visible GPUs, so a
4- or 8-rank slab partition -- middle ranks with two neighbours, end ranks with one -- can be validated on a 1-GPU box
(CUDA IPC works between processes on one device).  Sharded fused evaluation and sharded PNCG on device-generated slabs
against the whole cube on one GPU.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 tools/peer_check.py"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch, torch.distributed as dist
from bench import device_potentials, device_workload
from apple_b200 import _lib
from apple_b200.dist import ShardedOperators, ShardedPNCG
from apple_b200.optim.pncg import ConvergenceCriteria
from apple_b200.warp.model import WarpModel

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", rank % torch.cuda.device_count())
torch.cuda.set_device(dev)
dist.init_process_group("gloo")
n = int(os.environ.get("PEER_CHECK_N", "24"))
OPS = _lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_PROD
ok = True
for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 1e-5)):
    wl = device_workload(n, world, rank, dev, dtype)
    pots = device_potentials(wl, ["snh", "arap"], dtype)
    ops = ShardedOperators(WarpModel(pots), wl.shard, dev, dtype, transport="peer", peer_overlap=dtype == torch.float64)
    whole = device_workload(n, 1, 0, dev, dtype)
    wp = device_potentials(whole, ["snh", "arap"], dtype)
    V = whole.mesh.n_points
    fun = torch.zeros(1, dtype=dtype, device=dev)
    g, h = (torch.zeros((V, 3), dtype=dtype, device=dev) for _ in range(2))
    WarpModel(wp).eval(OPS, whole.u, whole.p, fun=fun, grad=g, prod=h)
    inv = torch.empty(V, dtype=torch.int64, device=dev)
    inv[whole.mesh.vertex_gid] = torch.arange(V, device=dev)
    gid = inv[wl.mesh.vertex_gid]             # this rank's vertices in the whole cube's (Morton) numbering
    for rep in range(3):                      # alternating receive buffers
        res = ops.eval(OPS, wl.u, wl.p)
        e = (abs(float(res["fun"]) - float(fun)) / abs(float(fun)), float((res["grad"] - g[gid]).abs().max() / g.abs().max()),
             float((res["prod"] - h[gid]).abs().max() / h.abs().max()))
        ok = ok and max(e) < 5 * tol
    print(f"rank {rank}/{world} {dtype} sharded eval vs whole cube: {e[0]:.2e} {e[1]:.2e} {e[2]:.2e}", flush=True)
    # sharded PNCG (device-side exchanges, WHILE-node graph) vs single GPU
    X = wl.mesh.points
    free = torch.ones((wl.mesh.n_points, 3), dtype=torch.bool, device=dev); free[X[:, 2] == 0.0] = False
    u0 = wl.u.clone(); u0[~free] = 0.0
    crit = ConvergenceCriteria(max_steps=10 ** 6, target_relative_gradient_norm=0.0)
    sp = ShardedPNCG(list(pots.values()), [], wl.shard, free, u0, criteria=crit, use_graph=2)
    sp.iterate(20)
    Xw = whole.mesh.points
    freew = torch.ones((V, 3), dtype=torch.bool, device=dev); freew[Xw[:, 2] == 0.0] = False
    u0w = whole.u.clone(); u0w[~freew] = 0.0
    s1 = ShardedPNCG(list(wp.values()), [], whole.shard, freew, u0w, criteria=crit, use_graph=2)
    s1.iterate(20)
    err = float((sp.x[:, :3] - s1.x[:, :3][gid]).abs().max() / s1.x[:, :3].abs().max())
    fa, fb = float(sp._read()[_lib.S_F]), float(s1._read()[_lib.S_F])
    good = err < (1e-8 if dtype == torch.float64 else 2e-4) and abs(fa - fb) <= 1e-4 * abs(fb)
    ok = ok and good
    print(f"rank {rank}/{world} {dtype} sharded PNCG 20 its: |u-u1|/|u1| = {err:.2e}, f = {fa:.8e} vs {fb:.8e} {'OK' if good else 'FAIL'}", flush=True)
    # block Jacobi: three fields (g', diag', block off-diagonals) travel in one exchange
    sb = ShardedPNCG(list(pots.values()), [], wl.shard, free, u0, criteria=crit, use_graph=2, preconditioner="block")
    sb.iterate(10)
    s1b = ShardedPNCG(list(wp.values()), [], whole.shard, freew, u0w, criteria=crit, use_graph=2, preconditioner="block")
    s1b.iterate(10)
    errb = float((sb.x[:, :3] - s1b.x[:, :3][gid]).abs().max() / s1b.x[:, :3].abs().max())
    goodb = errb < (1e-8 if dtype == torch.float64 else 2e-4)
    ok = ok and goodb
    print(f"rank {rank}/{world} {dtype} sharded block-Jacobi PNCG 10 its: |u-u1|/|u1| = {errb:.2e} {'OK' if goodb else 'FAIL'}", flush=True)
    del sp, s1, sb, s1b, ops
flag = torch.tensor([1.0 if ok else 0.0])
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PEER_CHECK PASS" if flag.item() == 1.0 else "PEER_CHECK FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
