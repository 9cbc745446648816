#!/usr/bin/env python
"""Recorded runs of BASELINE.json configs[2] and configs[3] (the default `bench.py` line is configs[4] with configs[1]
as a secondary key).  Prints ONE JSON line; the product path only (no oracle import).

  --config 3   117^3 x 5 = 8,008,065 tets, rectilinear coordinates graded by 1.02 per layer, Stable Neo-Hookean,
               PNCG with the reference-style point Jacobi AND the opt-in 3x3 block Jacobi, sharded across the ranks of
               the torchrun launch (1 / 2 / 4 / 8 GPUs): iterations/s, energy and gradient norm after the same number
               of iterations.
  --config 4   74^3 x 5 = 2,026,120 tets, three potentials on one mesh in the pattern of
               exp/2026/05/06/toy/src/21-smas-prestrain-stable-neo-hookean-muscle.py:228-240 (fat SNH with
               Fraction = 1 - s, muscle SNH x 1e3 with Fraction = s and an activation field with a prestrained slab),
               base fixed: forward PNCG solve, then the adjoint solve H p = -dL/du by the fused Jacobi-PCG on hess_prod
               to a relative residual of 1e-5 (exp/2025/09/24/inverse-grin/src/35-inverse-small-reg.py:223-260) and
               the mixed derivative products d(grad E . p)/d(activation, mu, lambda).  One GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402


def measured_peak():
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def config3(args):
    import torch.distributed as dist

    from apple_b200 import _lib
    from apple_b200.common import lame_converter
    from apple_b200.dist import ShardedPNCG, slab_shard_device
    from apple_b200.mesh import hash_uniform_device
    from apple_b200.optim.pncg import ConvergenceCriteria
    from apple_b200.warp.fem import StableNeoHookean

    world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    w = 4 if dtype == torch.float32 else 8
    n = args.n or 117
    t0 = time.perf_counter()
    shard = slab_shard_device(n, world, rank, dev, grading=1.02)
    dm = shard.mesh
    E = 10.0 ** (4.0 + hash_uniform_device(dm.cell_gid, 1))
    nu = 0.3 + 0.15 * hash_uniform_device(dm.cell_gid, 2)
    la, mu = lame_converter(E, nu)
    pot = StableNeoHookean.from_device_mesh(dm.cells, dm.points, mu=mu.to(dtype), lambda_=la.to(dtype), dtype=dtype, name="snh")
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    X = dm.points
    free = torch.ones((dm.n_points, 3), dtype=torch.bool, device=dev)
    free[X[:, 2] == 0.0] = False
    h = 1.0 / n
    u0 = (0.05 * h * torch.sin(7.0 * X[:, [1, 2, 0]] + 0.3)).to(dtype)
    u0[~free] = 0.0

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    T_total, V_total = 5 * n ** 3, (n + 1) ** 3
    n_free = 3 * (V_total - (n + 1) ** 2)
    iters = args.iters
    runs = {}
    for name, kw in (("point_jacobi", {}), ("block_jacobi", {"preconditioner": "block"}),
                     ("block_jacobi_psd", {"preconditioner": "block", "psd": True})):
        crit = ConvergenceCriteria(max_steps=10 ** 6, target_relative_gradient_norm=0.0)
        sp = ShardedPNCG([pot], [], shard, free, u0, criteria=crit, use_graph=2, **kw)
        sp.iterate(5)                                        # warm-up: graph capture
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sp.iterate(iters)
        e1.record()
        torch.cuda.synchronize()
        dt = allmax(e0.elapsed_time(e1) * 1e-3)
        s = sp._read()
        # algorithmic bytes of one iteration: two element passes + the vector kernels
        b_t = 16 + 9 * w + w + 2 * w
        vec = 15 * V_total * w + 8 * n_free * w + (3 * V_total * w if "block" in name else 0)
        bytes_iter = 2 * T_total * b_t + vec
        runs[name] = {
            "iters": iters, "seconds": dt, "iters_per_s": iters / dt, "accepted": int(s[_lib.S_N_ACCEPTED]) - 5,
            "energy": float(s[_lib.S_F]), "rel_grad_norm": float(np.sqrt(s[_lib.S_GNORM2] / s[_lib.S_GNORM2_FIRST])),
            "algorithmic_gbs_per_gpu": bytes_iter * iters / dt / 1e9 / world, "transport": sp.transport}
        del sp
        torch.cuda.empty_cache()
    peak, src = measured_peak()
    for r in runs.values():
        r["roofline_frac"] = r["algorithmic_gbs_per_gpu"] / peak
    if rank == 0:
        print(json.dumps({
            "config": "BASELINE.json configs[2]: 8M-tet graded mesh, Stable Neo-Hookean, PNCG point- and block-Jacobi, sharded",
            "workload": f"cube {n}^3x5 = {T_total} tets / {V_total} verts, grading 1.02 per layer, SNH, base z = 0 fixed, "
                        f"u0 = 0.05 h sin perturbation, {iters} PNCG iterations after 5 warm-up iterations",
            "n_gpus": world, "dtype": args.dtype, "setup_s": setup_s, "pncg": runs, "hbm_peak_gbs": peak, "peak_source": src,
            "note": "CUDA events, max over ranks; WHILE-node CUDA graph per iteration; exchanges are device-side peer-memory "
                    "kernels; algorithmic bytes = 2 T B_t + 15 V w + 8 n w (+ 3 V w for the block off-diagonals)"}))
    if world > 1:
        dist.destroy_process_group()


def config4(args):
    from apple_b200 import _lib
    from apple_b200.common import FIXED_MASK, FIXED_VALUE
    from apple_b200.forward import Forward, ModelBuilder
    from apple_b200.mesh import cube_tet_mesh
    from apple_b200.optim import PNCG, adjoint_solve
    from apple_b200.optim.pncg import ConvergenceCriteria
    from apple_b200.warp.fem import StableNeoHookean, StableNeoHookeanMuscle

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    w = 4 if dtype == torch.float32 else 8
    n = args.n or 74
    t0 = time.perf_counter()
    mesh = cube_tet_mesh(n, morton=True)
    T, V = mesh.n_cells, mesh.n_points
    rng = np.random.default_rng(1)
    cx = mesh.points[mesh.cells].mean(axis=1)
    s = np.clip(0.5 + 0.5 * np.sin(6.0 * cx[:, 0]) * np.cos(5.0 * cx[:, 1]), 0.05, 0.95)
    mu = 10.0 ** rng.uniform(3.0, 4.0, T)
    la = 10.0 ** rng.uniform(3.5, 4.5, T)
    act = 0.05 * rng.standard_normal((T, 6))
    slab = np.abs(cx[:, 2] - 0.5) < 0.1
    act[slab, :3] = np.array([1.2, 1.3 ** -2, 1.2]) - 1.0
    act[slab, 3:] = 0.0
    fixed = np.zeros((V, 3), dtype=bool); fixed[mesh.points[:, 2] == 0.0] = True
    mesh.point_data[FIXED_MASK.vtk] = fixed
    mesh.point_data[FIXED_VALUE.vtk] = np.zeros((V, 3))
    fat_m, mus_m = mesh.copy(), mesh.copy()
    fat_m.cell_data.update({"mu": mu, "lambda": la, "Fraction": 1.0 - s})
    mus_m.cell_data.update({"mu": 1e3 * mu, "lambda": 1e3 * la, "Fraction": s, "activation": act})
    b = ModelBuilder()
    b.add_vertices(mesh)
    b.add_fixed(mesh)
    b.add_potential(StableNeoHookean.from_pyvista(fat_m, dtype=dtype, name="fat"))
    b.add_potential(StableNeoHookeanMuscle.from_pyvista(mus_m, dtype=dtype, name="muscle"))
    model = b.finalize()
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    # forward solve: the prestrained slab contracts the cube (no external load needed)
    iters = args.iters
    crit = ConvergenceCriteria(max_steps=10 ** 7, target_relative_gradient_norm=args.forward_rtol)
    fwd = Forward(model, optimizer=PNCG(criteria=crit, use_graph=2, check_every=iters))
    opt = fwd.optimizer.init(fwd.problem, fwd.state, fwd.free)
    state = opt.step(fwd.problem, fwd.state, 5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    state = opt.step(fwd.problem, state, iters - 5)
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
    n_free = int(model.n_free)
    b_t = (16 + 9 * w + w + 2 * w) + (16 + 9 * w + w + 8 * w)          # fat + muscle records (two passes over the cells)
    bytes_iter = 2 * T * b_t + 2 * 15 * V * w // 2 + 8 * n_free * w
    forward = {"iters": iters - 5, "seconds": dt, "iters_per_s": (iters - 5) / dt, "accepted": opt.n_accepted,
               "energy": opt.line_search_state.f_alpha, "rel_grad_norm": opt.relative_grad_norm,
               "algorithmic_gbs": bytes_iter * (iters - 5) / dt / 1e9}

    # continue (untimed) towards the equilibrium: the adjoint system is defined there
    t1 = time.perf_counter()
    extra = 0
    while opt.done_code == 0.0 and extra < args.max_forward_iters:
        state = opt.step(fwd.problem, state, 500)       # no-ops once the device-side criterion is met
        extra += 500
    torch.cuda.synchronize()
    forward["to_equilibrium"] = {"total_iters": opt.n_steps, "done_code": opt.done_code, "seconds": time.perf_counter() - t1, "rel_grad_norm": opt.relative_grad_norm,
                                 "energy": opt.line_search_state.f_alpha}

    # adjoint: dL/du of L = 1/2 |u - u_target|^2 with u_target = 0.9 u  ->  H p = -(u - u_target) on the free DOFs
    u = state.u
    rhs = -(0.1 * model.dof_map.to_free(u)).contiguous()
    wm = model.warp_model.__wrapped__

    def true_residual(p, psd):
        pf = model.dof_map.to_full_grad(p)
        out = torch.zeros_like(pf)
        (wm.hess_prod_psd if psd else wm.hess_prod)(u, pf, out)
        res = model.dof_map.to_free_grad(out) - rhs
        return float(torch.linalg.vector_norm(res) / torch.linalg.vector_norm(rhs))

    out = {}
    p_adj = None
    for name, kw in (("fused_graph", dict(fused=True, use_graph=True)), ("fused_psd", dict(fused=True, psd=True)),
                     ("host_driven_torch", dict(fused=False))):
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        p, info = adjoint_solve(fwd.problem, state, rhs, tol=1e-5, maxiter=min(max(n_free // 10, 10), args.max_cg_iters),
                                check_every=32, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t1
        psd = bool(kw.get("psd"))
        bytes_it = T * ((16 + 9 * w + w + 2 * w) + (16 + 9 * w + w + 8 * w)) + 2 * 9 * V * w // 2 + 13 * n_free * w
        out[name] = {"iters": info.n_iters, "converged": bool(info.converged), "seconds": dt,
                     "iters_per_s": info.n_iters / dt if dt > 0 else None,
                     "true_relative_residual": true_residual(p, psd), "operator": "PSD-projected H" if psd else "H",
                     "algorithmic_gbs": bytes_it * info.n_iters / dt / 1e9}
        if info.converged and p_adj is None:
            p_adj = p
    if p_adj is None:
        p_adj = p
    # mixed derivative products with the adjoint (what the inverse problems add to dL/dq)
    t2 = time.perf_counter()
    md = model.warp_model.__wrapped__.mixed_derivative_prod(u, model.dof_map.to_full_grad(p_adj))
    torch.cuda.synchronize()
    md_s = time.perf_counter() - t2
    mixed = {pot: {k: float(torch.linalg.vector_norm(v)) for k, v in d.items()} for pot, d in md.items()}
    peak, src = measured_peak()
    forward["roofline_frac"] = forward["algorithmic_gbs"] / peak
    for r in out.values():
        r["roofline_frac"] = r["algorithmic_gbs"] / peak
    print(json.dumps({
        "config": "BASELINE.json configs[3]: ~2M-tet muscle-driven mesh, heterogeneous materials + activation, forward solve + adjoint HVP gradient",
        "workload": f"cube {n}^3x5 = {T} tets / {V} verts; fat SNH (Fraction 1-s) + muscle SNH x1e3 (Fraction s, activation with a "
                    f"prestrained slab); base fixed; {iters} PNCG iterations, then Jacobi-PCG on hess_prod to 1e-5",
        "n_gpus": 1, "dtype": args.dtype, "setup_s": setup_s, "n_free": n_free, "forward_pncg": forward, "adjoint_pcg": out,
        "mixed_derivative_prod": {"seconds": md_s, "norms": mixed}, "hbm_peak_gbs": peak, "peak_source": src,
        "note": "forward: CUDA events around the fused WHILE-node graph iterations; adjoint: wall clock around adjoint_solve "
                "(setup of the work vectors, hess_diag and the host checks every 32 iterations included); "
                "true_relative_residual is re-evaluated after the solve with the product's own hess_prod (PSD-projected for the psd run); a "
                "plain CG that meets a direction of non-positive curvature stops with converged = false (the reference then falls back "
                "to NormalCG, 35-inverse-small-reg.py:262-283)"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 4])
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--forward-rtol", type=float, default=1e-4, help="config 4: continue the forward solve to this relative gradient norm")
    ap.add_argument("--max-forward-iters", type=int, default=20000)
    ap.add_argument("--max-cg-iters", type=int, default=20000)
    args = ap.parse_args()
    (config3 if args.config == 3 else config4)(args)


if __name__ == "__main__":
    main()
