"""Multi-GPU parity check (run under torchrun, one rank per GPU):
   sharded fused evaluation and sharded PNCG vs the single-GPU path on rank 0.
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch, torch.distributed as dist
from helpers import make_case, cuda_potential, rel_err
from apple_b200 import _lib
from apple_b200.dist import ShardedOperators, ShardedPNCG, partition_mesh
from apple_b200.warp.model import WarpModel
from apple_b200.optim.pncg import ConvergenceCriteria

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ok = True
for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 1e-5)):
    mesh, u, p = make_case(n=10, seed=5, grading=1.0)
    mesh.cell_data.pop("Fraction")
    shard = partition_mesh(mesh, world, rank)
    pots = {k: cuda_potential(k, shard.mesh, dtype, name=k) for k in ("snh", "arap")}
    ops = ShardedOperators(WarpModel(pots), shard, dev, dtype, transport="nccl")
    peer = ShardedOperators(WarpModel(pots), shard, dev, dtype)           # default on CUDA: peer memory (apl_xchg_*)
    assert peer.transport == "peer" and not peer.overlap
    ul = torch.as_tensor(u[shard.l2g], dtype=dtype, device=dev).contiguous()
    pl = torch.as_tensor(p[shard.l2g], dtype=dtype, device=dev).contiguous()
    f, g, h = ops.fun_grad_hess_prod(ul, pl)
    # reference: the whole mesh on this rank's GPU
    full = {k: cuda_potential(k, mesh, dtype, name=k) for k in ("snh", "arap")}
    from apple_b200.warp.model import WarpModelAdapter
    ad = WarpModelAdapter(WarpModel(full), mesh.n_points)
    fr, gr, hr = ad.fun_grad_hess_prod(torch.as_tensor(u, dtype=dtype, device=dev), torch.as_tensor(p, dtype=dtype, device=dev))
    idx = torch.as_tensor(shard.l2g, device=dev)
    errs = (rel_err(f.cpu(), fr.cpu()), float((g - gr[idx]).abs().max() / gr.abs().max()), float((h - hr[idx]).abs().max() / hr.abs().max()))
    good = max(errs) < 5 * tol
    # split evaluation (boundary tiles, exchange overlapped with interior tiles) vs the plain sequence
    assert ops.overlap and ops.n_boundary_tiles > 0
    ops.overlap = False
    f2, g2, h2 = ops.fun_grad_hess_prod(ul, pl)
    ops.overlap = True
    e2 = (rel_err(f.cpu(), f2.cpu()), float((g - g2).abs().max() / g2.abs().max()), float((h - h2).abs().max() / h2.abs().max()))
    good = good and max(e2) < 5 * tol
    print(f"rank {rank} {dtype} split vs plain evaluation: {e2[0]:.2e} {e2[1]:.2e} {e2[2]:.2e}, {ops.n_boundary_tiles} boundary tiles", flush=True)
    # peer-memory exchange: same plan and summation order as pack / all-to-all / unpack -> the halo-summed rows agree
    # to rounding of the element pass (fp atomics order), shared rows are bit-identical across the two sharers, and
    # repeated evaluations (alternating buffer parity) stay consistent
    for rep in range(3):
        f3, g3, h3 = peer.fun_grad_hess_prod(ul, pl)
    e3 = (rel_err(f3.cpu(), fr.cpu()), float((g3 - gr[idx]).abs().max() / gr.abs().max()), float((h3 - hr[idx]).abs().max() / hr.abs().max()))
    good = good and max(e3) < 5 * tol
    other = 1 - rank if world == 2 else None
    if other is not None and other in shard.neighbors:
        mine = g3[torch.as_tensor(shard.neighbors[other], device=dev)].contiguous()
        theirs = torch.empty_like(mine)
        reqs = [dist.isend(mine, other), dist.irecv(theirs, other)] if rank == 0 else [dist.irecv(theirs, other), dist.isend(mine, other)]
        for r in reqs:
            r.wait()
        same = bool(torch.equal(mine, theirs))
        good = good and same
        print(f"rank {rank} {dtype} peer exchange: shared rows bit-identical on both sharers: {same}", flush=True)
    print(f"rank {rank} {dtype} peer-memory exchange vs 1-GPU: {e3[0]:.2e} {e3[1]:.2e} {e3[2]:.2e}", flush=True)
    ok &= good
    print(f"rank {rank} {dtype} fused eval vs 1-GPU: energy/grad/hvp rel err {errs[0]:.2e} {errs[1]:.2e} {errs[2]:.2e} {'OK' if good else 'FAIL'}", flush=True)

    # slabs generated and set up on the device (bench.py's workload) vs the whole cube on this GPU
    from bench import device_potentials, device_workload
    nn = 12
    wl = device_workload(nn, world, rank, dev, dtype)
    sops = ShardedOperators(WarpModel(device_potentials(wl, ["snh", "arap"], dtype)), wl.shard, dev, dtype)
    fs, gs, hs = sops.fun_grad_hess_prod(wl.u, wl.p)
    w1 = device_workload(nn, 1, 0, dev, dtype)
    ad1 = WarpModelAdapter(WarpModel(device_potentials(w1, ["snh", "arap"], dtype)), w1.mesh.n_points)
    f1, g1, h1 = ad1.fun_grad_hess_prod(w1.u, w1.p)
    o1 = torch.argsort(w1.mesh.vertex_gid)
    rows = o1[torch.searchsorted(w1.mesh.vertex_gid[o1], wl.mesh.vertex_gid)]
    es = (rel_err(fs.cpu(), f1.cpu()), float((gs - g1[rows]).abs().max() / g1.abs().max()), float((hs - h1[rows]).abs().max() / h1.abs().max()))
    good = max(es) < 5 * tol
    ok &= good
    print(f"rank {rank} {dtype} device-generated slabs vs whole cube: {es[0]:.2e} {es[1]:.2e} {es[2]:.2e} {'OK' if good else 'FAIL'}", flush=True)

    # PNCG: sharded vs single-GPU fused, fixed iteration count
    from apple_b200.common import FIXED_MASK, FIXED_VALUE
    from apple_b200.forward import Forward, ModelBuilder
    from apple_b200.optim import PNCG
    V = mesh.n_points
    fixed = np.zeros((V, 3), bool); fixed[mesh.points[:, 2] == 0.0] = True
    iters = 30
    crit = ConvergenceCriteria(max_steps=iters, target_relative_gradient_norm=0.0)
    b = ModelBuilder(dtype=dtype, device=dev); b.add_vertices(mesh)
    mesh.point_data[FIXED_MASK.vtk] = fixed; mesh.point_data[FIXED_VALUE.vtk] = np.zeros((V, 3))
    b.add_fixed(mesh)
    for pot in full.values(): b.add_potential(pot)
    fwd = Forward(b.finalize(), optimizer=PNCG(criteria=crit, check_every=iters))
    X = mesh.points
    u0 = np.ascontiguousarray(0.02 * np.sin(5.0 * X[:, [1, 2, 0]])); u0[fixed] = 0.0
    fwd.state.u = torch.as_tensor(u0, dtype=dtype, device=dev)
    sol = fwd.step()
    free_local = torch.as_tensor(~fixed[shard.l2g], device=dev)
    ptol = 1e-8 if dtype == torch.float64 else 2e-4
    # device-side sharded iteration (peer-memory exchanges inside apl_pncg_iterate; plain launches, static graph, WHILE
    # graph) and the host-driven NCCL path
    for transport, graph in (("peer", 0), ("peer", 1), ("peer", 2), ("nccl", 0)):
        sp = ShardedPNCG(list(pots.values()), [], shard, free_local, torch.as_tensor(u0[shard.l2g], dtype=dtype, device=dev),
                         criteria=crit, transport=transport, use_graph=graph)
        res = sp.solve(check_every=iters)
        eu = float((sp.u_local - fwd.state.u[idx]).abs().max() / fwd.state.u.abs().max())
        good = eu < ptol and res["n_steps"] == sol.stats["n_steps"] and abs(res["fun"] - sol.stats["fun"]) <= 10 * ptol * abs(sol.stats["fun"])
        ok &= good
        print(f"rank {rank} {dtype} PNCG[{transport}, graph {graph}] {res['n_steps']} its: |u-u1|/|u1| = {eu:.2e}, f = {res['fun']:.10e} vs {sol.stats['fun']:.10e}, accepted {res['n_accepted']} vs {sol.stats['n_accepted']} {'OK' if good else 'FAIL'}", flush=True)
        del sp
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
if rank == 0:
    print("DIST_CHECK " + ("PASS" if flag.item() == 1.0 else "FAIL"), flush=True)
sys.exit(0 if flag.item() == 1.0 else 1)
