#!/usr/bin/env python
"""Benchmark of the FEM-elasticity hot path (see BASELINE.json / DESIGN.md section "Measurement").

Workload (N = 1): BASELINE.json configs[1] -- synthetic 58^3 x 5 = 975,560-tet cube, Stable
Neo-Hookean + ARAP on the same mesh, fp32.  One *step* = one fused energy + gradient +
Hessian-vector-product evaluation of the whole model (every potential, one pass each).

  value      tets/s, inputs resident in HBM, L2 flushed between timed steps
  e2e        same metric through WarpModelAdapter.fun_grad_hess_prod_host with HOST (pinned) u, p:
             H2D copies, kernels, D2H of energy + gradient + HVP inside the timed region
  roofline   dominant kernel: algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json
  pncg       200 fused PNCG iterations on the same model (iterations/s)
  cpu_baseline  the oracle (numpy restatement of the reference) on a bounded sample, host cores

`--impl reference` times the CPU restatement of the reference (the reference itself needs Warp/JAX,
which are not installable here) on the same workload definition.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "tets_per_s_fused_energy_grad_hvp"
UNIT = "tets/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=0, help="hexes per cube edge (5 tets per hex); 0 = 58 * gpus^(1/3)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1 with --n 0: weak = ~975k tets per GPU (mesh grows with N), strong = the 58^3 mesh")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--scatter", default="tile", choices=["tile", "atomic", "tile_simple"])
    ap.add_argument("--layout", default="tet", choices=["tet", "pair"],
                    help="pair = EXPERIMENTAL: one consumer thread per pair of face-adjacent tets (tile assembly only)")
    ap.add_argument("--potentials", default="snh,arap")
    ap.add_argument("--pncg-iters", type=int, default=200)
    ap.add_argument("--no-fuse", action="store_true", help="one pass per potential (the reference's structure)")
    ap.add_argument("--graph", action="store_true", help="N>1, experimental: replay each step as one CUDA graph")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N>1: do not overlap the halo exchange with the interior tiles (one launch, then exchange)")
    ap.add_argument("--no-probe", action="store_true",
                    help="N=1: do not run the experimental pair layout in a time-limited subprocess after the measurement")
    ap.add_argument("--probe-parity", action="store_true",
                    help="also compare energy / gradient / HVP of the model with the C oracle (used by the pair-layout probe)")
    ap.add_argument("--slab", action="store_true",
                    help="config-5 style strong scaling: every rank GENERATES only its slab of the --n cube (no global "
                         "mesh on any rank; fields are functions of the global ids); operators only, also at 1 GPU")
    ap.add_argument("--no-pncg", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="also time every operator / variant (stderr table)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.n == 0:
        args.n = 58 if (world == 1 or args.scaling == "strong") else int(round(58 * world ** (1.0 / 3.0)))
    else:
        args.scaling = "strong" if world > 1 else args.scaling
    return args


def build_mesh(n, seed=0):
    from apple_b200.common import lame_converter
    from apple_b200.mesh import cube_tet_mesh

    mesh = cube_tet_mesh(n, morton=True)
    rng = np.random.default_rng(seed)
    T = mesh.n_cells
    E = 10.0 ** rng.uniform(4.0, 5.0, T)
    nu = rng.uniform(0.3, 0.45, T)
    la, mu = lame_converter(E, nu)
    mesh.cell_data["mu"] = mu
    mesh.cell_data["lambda"] = la
    h = 1.0 / n
    X = mesh.points
    u = 0.05 * h * np.sin(7.0 * X[:, [1, 2, 0]] + 0.3) + 0.02 * h * rng.uniform(-1, 1, X.shape)
    p = rng.uniform(-1, 1, X.shape)
    return mesh, np.ascontiguousarray(u), np.ascontiguousarray(p)


def cuda_potential(kind, mesh, dtype, **kw):
    """A potential of the product (`apple_b200.warp.fem`) on `mesh`; nothing of oracle/ is involved."""
    from apple_b200.warp.fem import Arap, StableNeoHookean, StableNeoHookeanMuscle

    cls = {"snh": StableNeoHookean, "arap": Arap, "muscle": StableNeoHookeanMuscle}[kind]
    return cls.from_pyvista(mesh, dtype=dtype, **kw)


def build_slab(n, world, rank):
    """This rank's slab of the n^3 x 5 cube (apple_b200.dist.slab_shard) with the fields of `build_mesh` redefined as
    functions of the GLOBAL vertex / cell ids, so that every partition of the same cube evaluates the same model."""
    from apple_b200.common import lame_converter
    from apple_b200.dist import slab_shard
    from apple_b200.mesh import hash_uniform

    shard = slab_shard(n, world, rank)
    mesh = shard.mesh
    cg, vg = mesh.cell_data["gid"], mesh.point_data["gid"]
    E = 10.0 ** (4.0 + hash_uniform(cg, 1))
    nu = 0.3 + 0.15 * hash_uniform(cg, 2)
    la, mu = lame_converter(E, nu)
    mesh.cell_data["mu"] = mu
    mesh.cell_data["lambda"] = la
    h = 1.0 / n
    X = mesh.points
    noise = np.stack([hash_uniform(3 * vg + k, 3) for k in range(3)], axis=1)
    u = 0.05 * h * np.sin(7.0 * X[:, [1, 2, 0]] + 0.3) + 0.02 * h * (2.0 * noise - 1.0)
    p = 2.0 * np.stack([hash_uniform(3 * vg + k, 4) for k in range(3)], axis=1) - 1.0
    return shard, mesh, np.ascontiguousarray(u), np.ascontiguousarray(p)


def workload_name(args, mesh, kinds):
    """config.workload, shared by both arms."""
    T, V = (5 * args.n ** 3, (args.n + 1) ** 3) if getattr(args, "slab", False) else (mesh.n_cells, mesh.n_points)
    return (f"cube {args.n}^3x5 = {T} tets / {V} verts, {'+'.join(kinds)}, fused energy+grad+HVP"
            + (" (per-rank slab generation)" if getattr(args, "slab", False) else ""))


def algorithmic_bytes_per_tet(kind, w, v_over_t, per_vertex_words):
    m = {"snh": 2, "arap": 1, "muscle": 8}[kind]
    return 16 + 9 * w + w + m * w + v_over_t * per_vertex_words * w


class ClockSampler:
    """Samples SM clock and clock-event (throttle) reasons while the benchmark runs: NVML every 2 ms when
    `pynvml` works, else the profiling recipe's `nvidia-smi --query-gpu` line back to back.  `summary`
    reports the median over the samples that fall inside the timed windows."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index=0):
        self.samples, self._stop, self.index = [], threading.Event(), index  # (t, sm_mhz, max_mhz, [reasons])
        self.windows = []
        self.backend = "nvidia-smi"
        self._nvml = None
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
            int(get_reasons(h))
            self._nvml = (pynvml, h, mx, get_reasons)
            self.backend = "nvml"
        except Exception:
            self._nvml = None
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            t = time.perf_counter()
            try:
                if self._nvml is not None:
                    pynvml, h, mx, get_reasons = self._nvml
                    sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    mask = int(get_reasons(h))
                    self.samples.append((t, sm, mx, [n for n, bit in self.REASONS if mask & bit]))
                    self._stop.wait(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 9:
                    names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
                    self.samples.append((t, float(f[1]), float(f[2]),
                                         [n for n, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
            except Exception:
                self._stop.wait(0.05)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def summary(self):
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in self.windows)]
        use = inside or self.samples
        reasons = sorted({r for s in use for r in s[3]})
        return {"sm_mhz": float(np.median([s[1] for s in use])) if use else None,
                "sm_max_mhz": max((s[2] for s in use), default=None), "reasons": reasons,
                "samples": len(use), "samples_in_timed_region": len(inside), "source": self.backend}


def measured_peak_gbs():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------ reference arm


def oracle_model(mesh, kinds, n_tets=None):
    """The reference's operators restated in C (oracle/c, all host threads), one pass per operator per
    potential exactly like WarpModel (warp/model/_model.py:13-36)."""
    from helpers import oracle_potential
    from apple_b200.mesh import TetMesh
    from oracle import cbind

    if n_tets is not None and n_tets < mesh.n_cells:
        sub = TetMesh(mesh.points, mesh.cells[:n_tets], cell_data={k: v[:n_tets] for k, v in mesh.cell_data.items()})
    else:
        sub = mesh
    pots = []
    for k in kinds:
        ref = oracle_potential(k, sub)  # numpy oracle: only used here to assemble dhdX / dV / materials
        pots.append(cbind.CPotential(k, ref.cells, ref.dhdX, ref.dV, ref.materials["mu"], ref.materials.get("lambda_"),
                                     ref.materials.get("activation")))
    return pots, sub.n_cells, cbind.num_threads()


def time_oracle(mesh, u, p, kinds, sample_tets, steps, warmup):
    pots, T, threads = oracle_model(mesh, kinds, sample_tets)
    V = mesh.n_points
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        e, g, h = np.zeros(1), np.zeros((V, 3)), np.zeros((V, 3))
        for pot in pots:  # fun, grad, hess_prod: three passes per potential, as the reference launches them
            pot.fun(u, e)
        for pot in pots:
            pot.grad(u, g)
        for pot in pots:
            pot.hess_prod(u, p, h)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return T * len(times) / sum(times), T, float(np.mean(times)), threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kinds = args.potentials.split(",")
    mesh, u, p = build_mesh(args.n)
    sample = min(mesh.n_cells, 1_000_000)
    steps, warmup = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    value, T, dt, threads = time_oracle(mesh, u, p, kinds, sample, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": args.scaling if int(os.environ.get("WORLD_SIZE", "1")) > 1 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, mesh, kinds),
                   "reference_arm": "energy + gradient + HVP as three operator passes per potential, exactly as the "
                                    "reference launches them (warp/model/_model.py:13-36); CPU restatement of the "
                                    "reference's Warp kernels in C (oracle/c, pthreads, fp64) -- the reference itself needs "
                                    "warp-lang / JAX, which cannot be installed here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"first {T} tets of the Morton-ordered mesh, {steps} evaluations"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist

    from apple_b200 import _lib, config
    from apple_b200.mesh import TetMesh
    from apple_b200.warp.fem import fuse_potentials
    from apple_b200.warp.model import WarpModel, WarpModelAdapter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        try:   # NCCL kernels on a high-priority stream: the halo exchange overlaps the interior element pass
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group("nccl", device_id=dev, pg_options=opts)
        except Exception:
            if dist.is_initialized():
                raise
            dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    w = 4 if args.dtype == "f32" else 8
    config.scatter = {"tile": _lib.SCATTER_TILE, "atomic": _lib.SCATTER_ATOMIC,
                      "tile_simple": _lib.SCATTER_TILE_SIMPLE}[args.scatter]
    config.layout = {"tet": _lib.LAYOUT_TET, "pair": _lib.LAYOUT_PAIR}[args.layout]
    kinds = args.potentials.split(",")

    shard = None
    if args.slab:
        shard, mesh, u, p = build_slab(args.n, world, rank)     # `mesh`, `u`, `p` are this rank's slab only
        T_total, V = shard.n_global_cells, shard.n_global_points
        args.no_pncg = args.no_cpu_baseline = args.no_probe = True
        args.scaling = "strong"
    else:
        mesh, u, p = build_mesh(args.n)
        T_total, V = mesh.n_cells, mesh.n_points
    OPS = _lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_PROD
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def flush():
        if not args.no_flush:
            flush_buf.fill_(1)

    if world == 1 and not args.slab:
        lo, hi = 0, T_total
        pots = {k: cuda_potential(k, mesh, dtype, name=k) for k in kinds}
        if not args.no_fuse:
            pots = fuse_potentials(pots)
        model = WarpModel(pots)
        adapter = WarpModelAdapter(model, n_points=V)
        ud = torch.as_tensor(u, dtype=dtype, device=dev).contiguous()
        pd = torch.as_tensor(p, dtype=dtype, device=dev).contiguous()
        from apple_b200.warp.model._adapter import zeros_block

        out_block, (grad, prod), (fun,) = zeros_block(V, 2, 1, dtype, dev)   # all outputs of a step in one buffer

        def step():
            out_block.zero_()      # the operators accumulate (warp/model/_model.py:13-36 zeroes first): one memset
            model.eval(OPS, ud, pd, fun=fun, grad=grad, prod=prod, zero=False)
    else:
        # strong scaling: rank r owns a contiguous chunk of the Morton-ordered tets and a local copy of
        # the vertices they touch; one halo sum (all-to-all of the shared rows) + one scalar all-reduce
        from apple_b200.dist import ShardedOperators, partition_mesh

        if shard is None:
            shard = partition_mesh(mesh, world, rank)
            u, p = u[shard.l2g], p[shard.l2g]
        lo, hi = shard.cell_range
        pots = {k: cuda_potential(k, shard.mesh, dtype, name=k) for k in kinds}
        if not args.no_fuse:
            pots = fuse_potentials(pots)
        sharded = ShardedOperators(WarpModel(pots), shard, dev, dtype, overlap=not args.no_overlap)
        ud = torch.as_tensor(u, dtype=dtype, device=dev).contiguous()
        pd = torch.as_tensor(p, dtype=dtype, device=dev).contiguous()
        fun = torch.zeros(1, dtype=dtype, device=dev)
        grad = torch.zeros((shard.n_local, 3), dtype=dtype, device=dev)
        prod = torch.zeros((shard.n_local, 3), dtype=dtype, device=dev)

        def eager_step():
            return sharded.eval(OPS, ud, pd)

        # The split evaluation (boundary tiles -> halo exchange overlapped with the interior tiles) must give
        # what the plain sequence gives; if it does not on this box, the plain sequence is benchmarked.
        overlap_note = None
        if sharded.overlap:
            try:
                r_split = eager_step()
            except Exception as exc:  # pragma: no cover - a host-side error is the same on every rank
                r_split, overlap_note = None, f"split evaluation failed ({type(exc).__name__}: {exc}); disabled"
            sharded.overlap = False
            r_plain = eager_step()
            sharded.overlap = r_split is not None
            torch.cuda.synchronize()
            err = float("inf") if r_split is None else max(
                float((r_split[k] - r_plain[k]).abs().max() / r_plain[k].abs().max().clamp_min(1e-30))
                for k in ("fun", "grad", "prod"))
            bad = torch.tensor([1.0 if not (err < 1e-4) else 0.0], device=dev)
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
            if float(bad.item()) > 0:
                sharded.overlap = False
                overlap_note = overlap_note or (f"split evaluation disagreed with the plain one on some rank (rel. err "
                                                f"{err:.2e} on rank 0); disabled")

        # EXPERIMENTAL, opt-in (--graph): replay the step (kernels + the two NCCL collectives) as one CUDA
        # graph.  Not validated: the one attempt on 8 GPUs hung during capture, so the default is plain
        # launches.
        graph = None
        if args.graph:
            for _ in range(3):
                eager_step()
            torch.cuda.synchronize()
            dist.barrier()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                graph_out = eager_step()

        def step():
            if graph is not None:
                graph.replay()
            else:
                eager_step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from before the warm-up until after the last timed GPU phase
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    for _ in range(args.warmup):
        flush(); step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_w0 = time.perf_counter()
    for a, b in ev:
        flush()
        a.record(); step(); b.record()
    barrier()
    clocks.window(t_w0, time.perf_counter())
    t_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([t_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_ms = float(t.item())
    ms_per_step = t_ms / args.steps
    value = T_total * args.steps / (t_ms * 1e-3)

    # ---- per-kernel roofline (each potential's fused kernel timed alone, L2 flushed) ----
    peak, peak_src = measured_peak_gbs()
    v_over_t = (shard.n_local if shard is not None else V) / max(hi - lo, 1)
    kern = {}
    for k, pot in pots.items():
        ts = []
        for _ in range(max(5, min(args.steps, 20))):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); pot.eval(OPS, ud, pd, fun=fun, grad=grad, prod=prod); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        dt = float(np.mean(ts)) * 1e-3
        bpt = sum(algorithmic_bytes_per_tet(part, w, 0.0, 0) - 16 - 9 * w for part in k.split("+")) + 16 + 9 * w \
            + v_over_t * 12 * w  # connectivity, Dm^-1 and the nodal fields are touched once per pass
        kern[k] = {"ms": dt * 1e3, "bytes_per_tet": bpt, "gbs": bpt * (hi - lo) / dt / 1e9,
                   "gtets_per_s": (hi - lo) / dt / 1e9}
    dom = max(kern, key=lambda k: kern[k]["ms"])
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists() and world == 1:
        try:
            traffic = json.loads(tf.read_text()).get(f"{args.dtype}:{dom}:n{args.n}")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["gbs"] / peak, "traffic": traffic, "kernel": f"fem_pipe_kernel<{args.dtype},{dom},fun|grad|hess_prod>",
                "peak_source": peak_src, "frac_of_nominal_8TBs": kern[dom]["gbs"] / 8000.0, "per_kernel": kern}

    parity = None
    if world == 1 and args.probe_parity and not args.slab:
        # the whole model once against the C restatement of the reference (fp64) on the full mesh
        pots_o, _, _ = oracle_model(mesh, kinds)
        e_o, g_o, h_o = np.zeros(1), np.zeros((V, 3)), np.zeros((V, 3))
        for po in pots_o:
            po.fun(u, e_o); po.grad(u, g_o); po.hess_prod(u, p, h_o)
        step(); torch.cuda.synchronize()
        rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())  # noqa: E731
        parity = {"energy": rel(fun.cpu().numpy(), e_o), "grad": rel(grad.cpu().numpy(), g_o),
                  "hess_prod": rel(prod.cpu().numpy(), h_o), "against": "C oracle (fp64) on the full mesh"}

    # ---- e2e: host buffers through the public adapter API ----
    e2e = None
    if world == 1 and not args.slab:
        uh = torch.as_tensor(u, dtype=dtype).pin_memory()
        ph = torch.as_tensor(p, dtype=dtype).pin_memory()
        gh = torch.empty((V, 3), dtype=dtype).pin_memory()
        hh = torch.empty((V, 3), dtype=dtype).pin_memory()
        fh = torch.empty(1, dtype=dtype).pin_memory()

        def step_streamed():
            # the host-buffer entry point of the plugin: H2D of u, p and D2H of energy, gradient, HVP are
            # inside, overlapped with two element passes on side streams (see its docstring)
            adapter.fun_grad_hess_prod_host(uh, ph, out=(fh, gh, hh))

        def step_serial():
            # the same call sequence a caller with host state would write by hand: copy in, one fused
            # pass, copy out, all on the current stream
            u_d = uh.to(dev, non_blocking=True); p_d = ph.to(dev, non_blocking=True)
            f, g, h = adapter.fun_grad_hess_prod(u_d, p_d)
            fh.copy_(f.reshape(1), non_blocking=True); gh.copy_(g, non_blocking=True); hh.copy_(h, non_blocking=True)

        def time_e2e(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            tt = 0.0
            for _ in range(args.steps):
                flush()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                torch.cuda.synchronize()
                tt += a.elapsed_time(b)
            return tt

        try:
            # the serial form is timed first and kept as the check of the streamed one (same energy / gradient /
            # HVP in the host buffers) and as the fallback should the streamed entry point fail on this box
            tt_serial = time_e2e(step_serial)
            ref = (fh.clone(), gh.clone(), hh.clone())
            api, note, launches = "serial", None, len(pots)
            tt = tt_serial
            try:
                fh.zero_(); gh.zero_(); hh.zero_()
                tt_streamed = time_e2e(step_streamed)
                err = max(float((x - y).abs().max() / y.abs().max()) for x, y in zip((fh, gh, hh), ref))
                if err < 1e-4 and tt_streamed <= tt_serial:
                    api, tt, launches = "streamed", tt_streamed, 2 * len(pots)
                elif err < 1e-4:
                    note = f"streamed entry point verified but slower here ({tt_streamed / args.steps:.4f} ms per step); serial number reported"
                else:
                    note = f"streamed entry point disagreed with the serial one (rel. err {err:.2e}); serial number reported"
            except Exception as exc:  # pragma: no cover - robustness of the benchmark line
                note = f"streamed entry point failed ({type(exc).__name__}: {exc}); serial number reported"
            e2e = {"value": T_total * args.steps / (tt * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": int(2 * V * 3 * w), "d2h_bytes_per_step": int(2 * V * 3 * w + w),
                   "ms_per_step": tt / args.steps,
                   "api": {"streamed": "WarpModelAdapter.fun_grad_hess_prod_host(u_host, p_host, out=host tensors): copies "
                                       "overlapped with a fun+grad pass and a hess_prod pass on side streams",
                           "serial": "u.to(device); p.to(device); WarpModelAdapter.fun_grad_hess_prod; copy_ to host"}[api],
                   "gpu_launches_per_step": launches, "serial_ms_per_step": tt_serial / args.steps, "note": note}
        except Exception as exc:  # pragma: no cover - keep the headline line
            e2e = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- PNCG iterations/s on the same model (config 1/2 style solve: fixed base, fused path) ----
    pncg = None
    if world == 1 and not args.no_pncg:
        try:
            pncg = bench_pncg(args, mesh, pots, dtype, dev, w)
        except Exception as exc:  # pragma: no cover - the headline line must survive a failure of a secondary section
            pncg = {"error": f"{type(exc).__name__}: {exc}"}

    clocks.__exit__(None, None, None)

    # ---- CPU baseline: the oracle on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            val, Ts, dt, threads = time_oracle(mesh, u, p, kinds, min(T_total, 1_000_000), 5, 1)
            cpu = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"first {Ts} tets of the same mesh, 5 evaluations (fun, grad, hess_prod passes per "
                             f"potential), C restatement of the reference kernels (oracle/c), fp64"}
        except Exception as exc:  # pragma: no cover
            cpu = {"error": f"{type(exc).__name__}: {exc}"}

    if args.sweep and rank == 0 and world == 1:
        sweep(args, mesh, u, p, dtype, dev, flush)

    # ---- experimental pair layout, in a time-limited subprocess (a kernel that has never run must not be able
    #      to stall or crash the benchmark of the product layout) ----
    probe = None
    if rank == 0 and world == 1 and args.layout == "tet" and not args.no_probe:
        cmd = [sys.executable, str(ROOT / "bench.py"), "--layout", "pair", "--steps", "10", "--warmup", "3", "--no-pncg",
               "--no-cpu-baseline", "--no-probe", "--probe-parity", "--n", str(args.n), "--dtype", args.dtype,
               "--potentials", args.potentials]
        try:
            pr = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
            last = [ln for ln in pr.stdout.strip().splitlines() if ln.startswith("{")]
            if pr.returncode == 0 and last:
                d = json.loads(last[-1])
                probe = {"status": "ran", "value": d["value"], "ms_per_step": d["ms_per_step"],
                         "roofline_frac": d["roofline"]["frac"], "parity_vs_oracle": d.get("parity"),
                         "e2e": (d.get("e2e") or {}).get("value"), "speedup_vs_tet_layout": d["value"] / value}
            else:
                probe = {"status": f"failed (rc {pr.returncode})", "stderr_tail": pr.stderr[-400:]}
        except subprocess.TimeoutExpired:
            probe = {"status": "timeout (240 s): killed"}
        except Exception as exc:  # pragma: no cover
            probe = {"status": f"not run ({type(exc).__name__}: {exc})"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling if (world > 1 or args.slab) else "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_name(args, mesh, kinds), "assembly": args.scatter, "layout": args.layout,
                       "l2": "256 MiB flush write between timed steps" if not args.no_flush else "no flush",
                       "parallelism": (f"{world} ranks x " + ("slab of hex layers, generated per rank" if args.slab else
                                                               "contiguous Morton chunk of tets") +
                                       f" (~{T_total // world} tets per "
                                       f"GPU, {args.scaling} scaling); halo sum of grad+HVP (NCCL all-to-all of "
                                       f"shared rows) + scalar all-reduce per step, "
                                       f"{'replayed as one CUDA graph' if args.graph else 'plain launches'}; "
                                       + (f"boundary tiles ({sharded.n_boundary_tiles} on rank 0) first, exchange "
                                          f"overlapped with the interior tiles" if sharded.overlap else
                                          "one element launch, then the exchange") +
                                       (f" [{overlap_note}]" if overlap_note else ""))
                       if world > 1 else "1 GPU"},
            "clocks": clocks.summary(), "e2e": e2e,
            "gpu_launches": args.steps * len(pots) * (2 if (world > 1 and sharded.overlap) else 1),
            "roofline": roofline, "cpu_baseline": cpu, "pncg": pncg,
        }
        if parity is not None:
            line["parity"] = parity
        if probe is not None:
            line["experimental_pair_layout"] = probe
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_pncg(args, mesh, pots, dtype, dev, w):
    import torch

    from apple_b200.common import FIXED_MASK, FIXED_VALUE
    from apple_b200.forward import Forward, ModelBuilder
    from apple_b200.optim import PNCG
    from apple_b200.optim.pncg import ConvergenceCriteria

    V, T = mesh.n_points, mesh.n_cells
    builder = ModelBuilder(dtype=dtype, device=dev)
    builder.add_vertices(mesh)
    fixed = np.zeros((V, 3), dtype=bool); fixed[mesh.points[:, 2] == 0.0] = True
    mesh.point_data[FIXED_MASK.vtk] = fixed
    mesh.point_data[FIXED_VALUE.vtk] = np.zeros((V, 3))
    builder.add_fixed(mesh)
    for pot in pots.values():
        builder.add_potential(pot)
    model = builder.finalize()
    h = 1.0 / args.n
    X = mesh.points
    u0 = np.ascontiguousarray(0.05 * h * np.sin(7.0 * X[:, [1, 2, 0]] + 0.3))
    u0[fixed] = 0.0
    out = {}
    for graph in (2, 1, 0):
        iters = args.pncg_iters
        crit = ConvergenceCriteria(max_steps=iters + 20, target_relative_gradient_norm=0.0)
        fwd = Forward(model, optimizer=PNCG(criteria=crit, check_every=iters, use_graph=graph))
        fwd.state.u = torch.as_tensor(u0, dtype=dtype, device=dev)
        problem, state = fwd.problem, fwd.state
        opt_state = fwd.optimizer.init(problem, state, fwd.free)      # workspace + initial pass (untimed)
        state = opt_state.step(problem, state, 20)                    # warm-up: graph capture, clocks
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        state = opt_state.step(problem, state, iters)                 # `iters` iterations, one host sync at the end
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dt = e0.elapsed_time(e1) * 1e-3
        n_free = model.n_free
        n_pots = len(pots)
        per_tet = sum(16 + 9 * w + sum(w + {"snh": 2, "arap": 1, "muscle": 8}[part] * w for part in k.split("+"))
                      for k in pots)
        alg = 2 * T * per_tet + n_pots * 15 * V * w + 8 * n_free * w
        out[{2: "graph_while", 1: "graph", 0: "eager"}[graph]] = {
            "iters": iters, "accepted_total": opt_state.n_accepted, "device_seconds": dt, "wall_seconds": wall,
            "iters_per_s": iters / max(dt, wall), "energy": opt_state.line_search_state.f_alpha,
            "rel_grad_norm": opt_state.relative_grad_norm, "algorithmic_gbs": alg * iters / dt / 1e9,
        }
    out["iters_per_s"] = out["graph_while"]["iters_per_s"]
    out["note"] = ("graph_while = one CUDA graph per iteration with a conditional WHILE node for the line search; "
                   "graph = all trials as flag-guarded launches; eager = plain launches.  200 iterations after 20 "
                   "warm-up iterations, CUDA events + wall clock around the enqueue-and-sync; no L2 flush "
                   "(iterations run back to back)")
    return out


def sweep(args, mesh, u, p, dtype, dev, flush):
    """Per-operator / per-variant timings (stderr): evidence for the assembly-strategy choice."""
    import torch

    from apple_b200 import _lib

    V, T = mesh.n_points, mesh.n_cells
    rows = []
    for dt_name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        for kind in ("snh", "arap"):
            pot = cuda_potential(kind, mesh, dt)
            for ld in (3, 4):
                ud = torch.zeros((V, ld), dtype=dt, device=dev); ud[:, :3] = torch.as_tensor(u, dtype=dt)
                pd = torch.zeros((V, ld), dtype=dt, device=dev); pd[:, :3] = torch.as_tensor(p, dtype=dt)
                outs = {k: torch.zeros((V, ld), dtype=dt, device=dev) for k in ("grad", "diag", "prod")}
                fun = torch.zeros(1, dtype=dt, device=dev); quad = torch.zeros(1, dtype=dt, device=dev)
                for ops in (1, 2, 4, 8, 16, 7, 11, 15):
                    for scatter in ((0,) if args.layout == "pair" else (0, 2, 1)):
                        ts = []
                        for i in range(8):
                            flush()
                            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            a.record()
                            pot.eval(ops, ud, pd, fun=fun, quad=quad, grad=outs["grad"], diag=outs["diag"],
                                     prod=outs["prod"], scatter=scatter)
                            b.record(); torch.cuda.synchronize()
                            if i >= 2:
                                ts.append(a.elapsed_time(b))
                        ms = float(np.mean(ts))
                        rows.append((dt_name, kind, ld, ops, {0: "tile", 1: "atomic", 2: "simple"}[scatter], ms, T / ms / 1e6))
    print("dtype kind ld ops scatter ms Gtets/s", file=sys.stderr)
    for r in rows:
        print(f"{r[0]} {r[1]} {r[2]} {r[3]:2d} {r[4]:6s} {r[5]:8.4f} {r[6]:8.2f}", file=sys.stderr)


if __name__ == "__main__":
    main()
