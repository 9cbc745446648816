#!/usr/bin/env python
"""Benchmark of the FEM-elasticity hot path (BASELINE.json / DESIGN.md section "Measurement").

Workload, at every N: BASELINE.json configs[4] -- the synthetic 234^3 x 5 = 64,064,520-tet cube (12,977,875 vertices),
Stable Neo-Hookean + ARAP on the same cells (fused into one pass), fp32, STRONG scaling: the mesh is fixed and rank r
of N generates, orders, tiles and packs its slab of hex layers entirely on its GPU (nothing but the connectivity of
the slab visits the host).  One *step* = one fused energy + gradient + Hessian-vector-product evaluation of the whole
model, halo sums and the energy reduction included (peer memory over NVLink, `apl_xchg_*`).

  value       tets/s, inputs resident in HBM, L2 flushed between timed steps (the working set is >> L2 anyway)
  e2e         the same metric from HOST (pinned) u, p to HOST energy / gradient / HVP: H2D, kernels, exchange, D2H inside
  roofline    dominant kernel: algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json
  parity      rank 0's gradient / HVP rows on three grid planes around its first slab interface (the shared plane
              included) and the energy of a 4-layer sample against the C restatement of the reference (oracle/c, fp64)
  hvp         Hessian-vector product alone, fp32 and fp64 (configs[4]: "fp32 and fp64 HVP throughput vs HBM roofline")
  config2     N = 1 only: BASELINE.json configs[1] (58^3 x 5 = 975,560 tets): operators, roofline, 200 PNCG iterations
  cpu_baseline  the C restatement of the reference on a bounded sample (4 hex layers ~ 1.1 M tets), all host threads

`--impl reference` times that C restatement (the reference itself needs Warp / JAX, not installable here) on the
same workload definition, in the same dtype, honouring --steps / --warmup; rank 0 only.
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
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "tets_per_s_fused_energy_grad_hvp"
UNIT = "tets/s"
N_HEADLINE = 234     # BASELINE.json configs[4]
N_CONFIG2 = 58       # BASELINE.json configs[1]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_HEADLINE, help="hexes per cube edge (5 tets per hex); fixed for every N")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--scatter", default="tile", choices=["tile", "atomic", "tile_simple"])
    ap.add_argument("--potentials", default="snh,arap")
    ap.add_argument("--pncg-iters", type=int, default=200)
    ap.add_argument("--no-fuse", action="store_true", help="one pass per potential (the reference's structure)")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="N>1: halo sums through peer memory (apl_xchg_*) or pack / NCCL all-to-all / unpack / all-reduce")
    ap.add_argument("--graph", action="store_true", help="N>1: replay each step as one CUDA graph (peer transport only)")
    ap.add_argument("--peer-overlap", action="store_true",
                    help="N>1, peer transport: boundary tiles, push, interior tiles, pull (default: one pass, then the exchange)")
    ap.add_argument("--no-pncg", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-hvp", action="store_true")
    ap.add_argument("--no-config2", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-probe", action="store_true", help=argparse.SUPPRESS)   # accepted for old command lines
    ap.add_argument("--sweep", action="store_true", help="N=1: also time every operator / variant at config 2 (stderr table)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ workload
# Materials and fields are functions of the GLOBAL cell / vertex ids (splitmix64 hashes), so that every partition of the
# cube -- and the host sample handed to the oracle -- evaluates the same model.

def _fields(xp, hashf, X, vg, cg, n):
    """xp = numpy or torch; X (V, 3) rest positions, vg / cg global vertex / cell ids."""
    from apple_b200.common import lame_converter

    E = 10.0 ** (4.0 + hashf(cg, 1))
    nu = 0.3 + 0.15 * hashf(cg, 2)
    la, mu = lame_converter(E, nu)
    h = 1.0 / n
    noise = xp.stack([hashf(3 * vg + k, 3) for k in range(3)], 1)
    u = 0.05 * h * xp.sin(7.0 * X[:, [1, 2, 0]] + 0.3) + 0.02 * h * (2.0 * noise - 1.0)
    p = 2.0 * xp.stack([hashf(3 * vg + k, 4) for k in range(3)], 1) - 1.0
    return mu, la, u, p


def device_workload(n, world, rank, device, dtype):
    """This rank's slab of the n^3 x 5 cube, generated on the device, with materials and fields."""
    import torch

    from apple_b200.dist import slab_shard_device
    from apple_b200.mesh import hash_uniform_device

    shard = slab_shard_device(n, world, rank, device)
    dm = shard.mesh
    mu, la, u, p = _fields(torch, hash_uniform_device, dm.points, dm.vertex_gid, dm.cell_gid, n)
    return SimpleNamespace(shard=shard, mesh=dm, mu=mu.to(dtype), la=la.to(dtype), u=u.to(dtype).contiguous(),
                           p=p.to(dtype).contiguous())


def device_potentials(w, kinds, dtype, fuse=True):
    """The model's potentials built by apl_fem_create_from_mesh (setup on the GPU)."""
    from apple_b200.warp.fem import Arap, FusedSnhArap, StableNeoHookean

    dm = w.mesh
    if fuse and sorted(kinds) == ["arap", "snh"]:
        return {"snh+arap": FusedSnhArap.from_device_mesh(dm.cells, dm.points, mu=w.mu, lambda_=w.la, mu_arap=w.mu,
                                                          dtype=dtype, name="snh+arap")}
    pots = {}
    for k in kinds:
        if k == "snh":
            pots[k] = StableNeoHookean.from_device_mesh(dm.cells, dm.points, mu=w.mu, lambda_=w.la, dtype=dtype, name=k)
        elif k == "arap":
            pots[k] = Arap.from_device_mesh(dm.cells, dm.points, mu=w.mu, dtype=dtype, name=k)
        else:
            raise KeyError(f"bench.py builds snh / arap potentials, not {k!r}")
    return pots


def host_sample(n, i0, i1):
    """Hex layers [i0, i1) of the same cube as numpy arrays (for the oracle): TetMesh + fields, local vertex numbering,
    `layer` = grid index i of every local vertex."""
    from apple_b200.mesh import TetMesh, cube_tet_slab, hash_uniform

    pts, cells, vg, cg = cube_tet_slab(n, i0, i1)
    mu, la, u, p = _fields(np, hash_uniform, pts, vg, cg, n)
    mesh = TetMesh(pts, cells, point_data={"gid": vg}, cell_data={"gid": cg, "mu": mu, "lambda": la})
    return SimpleNamespace(mesh=mesh, u=np.ascontiguousarray(u), p=np.ascontiguousarray(p), vgid=vg,
                           layer=vg // ((n + 1) * (n + 1)))


def sample_layers(n, world):
    """The 4 hex layers around rank 0's upper slab interface (N > 1; its shared plane is the third of the five grid
    planes) or the first 4 layers (N = 1)."""
    k = min(4, n)
    if world == 1:
        return 0, k
    i1 = n // world                       # rank 0 owns hex layers [0, i1): plane i1 is shared with rank 1
    a = max(0, min(i1 - 2, n - k))
    return a, a + k


def build_mesh(n, seed=0):
    """Config-2 style HOST mesh (numpy; Morton-ordered) for the PNCG section and the tools."""
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
    """A potential of the product (`apple_b200.warp.fem`) on a host `mesh`; nothing of oracle/ is involved."""
    from apple_b200.warp.fem import Arap, StableNeoHookean, StableNeoHookeanMuscle

    cls = {"snh": StableNeoHookean, "arap": Arap, "muscle": StableNeoHookeanMuscle}[kind]
    return cls.from_pyvista(mesh, dtype=dtype, **kw)


def workload_name(n, kinds):
    return (f"cube {n}^3x5 = {5 * n ** 3} tets / {(n + 1) ** 3} verts, {'+'.join(kinds)}, fused energy+grad+HVP"
            + (" (BASELINE.json configs[4])" if n == N_HEADLINE else ""))


def bytes_per_tet(kinds, w, v_over_t, vertex_words):
    """Algorithmic bytes per tet of one fused pass over potentials that share their cells: connectivity
    16, Dm^-1 9w, then per potential vol + materials, plus the nodal fields touched once per vertex."""
    m = {"snh": 2, "arap": 1, "muscle": 8}
    return 16 + 9 * w + sum(w + m[k] * w for k in kinds) + v_over_t * vertex_words * w


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
# (oracle/ is imported below this line only: CPU baseline, reference arm and the parity check of the benchmark)


def oracle_model(sample, kinds, np_dtype=np.float64):
    """The reference's operators restated in C (oracle/c, all host threads), one pass per operator per potential
    exactly like WarpModel (warp/model/_model.py:13-36), on a host sample of the workload."""
    from oracle import cbind
    from oracle import region as oregion

    mesh = sample.mesh
    dhdX, dV = oregion.compute_grad(mesh.points, mesh.cells, None, dtype=np.float64)
    pots = []
    for k in kinds:
        pots.append(cbind.CPotential(k, mesh.cells, dhdX, dV, mesh.cell_data["mu"],
                                     mesh.cell_data["lambda"] if k != "arap" else None, None, dtype=np_dtype))
    return pots, cbind.num_threads()


def oracle_eval(pots, sample, np_dtype=np.float64):
    """energy, gradient, HVP of the sample: three operator passes per potential, as the reference launches them."""
    V = sample.mesh.n_points
    e, g, h = np.zeros(1), np.zeros((V, 3), np_dtype), np.zeros((V, 3), np_dtype)
    for pot in pots:
        pot.fun(sample.u, e)
    for pot in pots:
        pot.grad(sample.u, g)
    for pot in pots:
        pot.hess_prod(sample.u, sample.p, h)
    return e, g, h


def time_oracle(sample, kinds, steps, warmup, np_dtype):
    pots, threads = oracle_model(sample, kinds, np_dtype)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        oracle_eval(pots, sample, np_dtype)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    T = sample.mesh.n_cells
    return T * len(times) / sum(times), T, float(np.mean(times)), threads


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    kinds = args.potentials.split(",")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    a, b = sample_layers(args.n, 1)
    sample = host_sample(args.n, a, b)
    np_dtype = np.float32 if args.dtype == "f32" else np.float64
    value, T, dt, threads = time_oracle(sample, kinds, args.steps, args.warmup, np_dtype)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload_name(args.n, kinds),
                   "reference_arm": "energy + gradient + HVP as three operator passes per potential, exactly as the "
                                    "reference launches them (warp/model/_model.py:13-36); CPU restatement of the "
                                    f"reference's Warp kernels in C (oracle/c, pthreads, {args.dtype}) on a bounded sample "
                                    "-- the reference itself needs warp-lang / JAX, which cannot be installed here",
                   "launched_ranks": world},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"hex layers [{a}, {b}) of the cube = {T} tets per step, {args.steps} steps"},
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
    from apple_b200.warp.model import WarpModel, WarpModelAdapter
    from apple_b200.warp.model._adapter import zeros_block

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    w = 4 if args.dtype == "f32" else 8
    config.scatter = {"tile": _lib.SCATTER_TILE, "atomic": _lib.SCATTER_ATOMIC,
                      "tile_simple": _lib.SCATTER_TILE_SIMPLE}[args.scatter]
    kinds = args.potentials.split(",")
    n = args.n
    T_total, V_total = 5 * n ** 3, (n + 1) ** 3
    OPS = _lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_PROD
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def flush():
        if not args.no_flush:
            flush_buf.fill_(1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- setup: mesh, fields, potentials -- all on the device ----
    barrier()
    t_setup = time.perf_counter()
    wl = device_workload(n, world, rank, dev, dtype)
    shard = wl.shard
    T_local, V_local = wl.mesh.n_cells, wl.mesh.n_points
    pots = device_potentials(wl, kinds, dtype, fuse=not args.no_fuse)
    model = WarpModel(pots)
    ud, pd = wl.u, wl.p
    sharded = None
    if world > 1:
        from apple_b200.dist import ShardedOperators

        sharded = ShardedOperators(model, shard, dev, dtype, transport=args.transport, overlap=args.transport == "nccl",
                                   peer_overlap=args.peer_overlap)
    torch.cuda.synchronize()
    setup_s = max_over_ranks(time.perf_counter() - t_setup)
    sharded_peer_overlap = bool(sharded is not None and sharded.peer_overlap)

    if world == 1:
        adapter = WarpModelAdapter(model, n_points=V_local)
        out_block, (grad, prod), (fun,) = zeros_block(V_local, 2, 1, dtype, dev)   # all outputs of a step in one buffer

        def step():
            out_block.zero_()      # the operators accumulate (warp/model/_model.py:13-36 zeroes first): one memset
            model.eval(OPS, ud, pd, fun=fun, grad=grad, prod=prod, zero=False)
            return {"fun": fun, "grad": grad, "prod": prod}
    else:
        def eager_step():
            return sharded.eval(OPS, ud, pd)

        graph = None
        if args.graph and args.transport == "peer":
            for _ in range(3):
                eager_step()
            barrier()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                graph_out = eager_step()

        def step():
            if graph is not None:
                graph.replay()
                return graph_out
            return eager_step()

    if world == 1:
        launches_per_step = len(pots)
    elif args.transport == "peer":    # element pass(es) + push + pull (+ the scalar push + pull of the overlapped form)
        launches_per_step = len(pots) * (2 if sharded.peer_overlap else 1) + (4 if sharded.peer_overlap else 2)
    else:
        launches_per_step = len(pots) + 2 + len(pots) * int(sharded.overlap)

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
        a.record(); res = step(); b.record()
    barrier()
    clocks.window(t_w0, time.perf_counter())
    t_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    ms_per_step = t_ms / args.steps
    value = T_total * args.steps / (t_ms * 1e-3)

    # ---- per-kernel roofline (each potential's fused kernel timed alone on this rank's slab, L2 flushed) ----
    peak, peak_src = measured_peak_gbs()
    v_over_t = V_local / max(T_local, 1)
    fun1 = torch.zeros(1, dtype=dtype, device=dev)
    g1, h1 = (torch.zeros((V_local, 3), dtype=dtype, device=dev) for _ in range(2))
    kern = {}
    for k, pot in pots.items():
        ts = []
        for _ in range(max(5, min(args.steps, 20))):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); pot.eval(OPS, ud, pd, fun=fun1, grad=g1, prod=h1); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        dt = max_over_ranks(float(np.mean(ts))) * 1e-3      # the slowest rank's kernel
        bpt = bytes_per_tet(k.split("+"), w, v_over_t, 12)
        kern[k] = {"ms": dt * 1e3, "bytes_per_tet": bpt, "gbs": bpt * T_local / dt / 1e9, "gtets_per_s": T_local / dt / 1e9,
                   "tets": T_local}
    dom = max(kern, key=lambda k: kern[k]["ms"])
    # the Stable Neo-Hookean kernel ALONE on the same slab (the bandwidth-bound member of the fused pair), reported beside
    # the headline kernel; never the headline itself
    if "snh" in kinds and "snh" not in pots and not args.no_hvp:
        try:
            from apple_b200.warp.fem import StableNeoHookean

            alone = StableNeoHookean.from_device_mesh(wl.mesh.cells, wl.mesh.points, mu=wl.mu, lambda_=wl.la, dtype=dtype, name="snh")
            ts = []
            for _ in range(max(5, min(args.steps, 20))):
                flush()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); alone.eval(OPS, ud, pd, fun=fun1, grad=g1, prod=h1); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            dt = max_over_ranks(float(np.mean(ts[1:]))) * 1e-3
            bpt = bytes_per_tet(["snh"], w, v_over_t, 12)
            kern["snh (alone, same mesh)"] = {"ms": dt * 1e3, "bytes_per_tet": bpt, "gbs": bpt * T_local / dt / 1e9,
                                              "gtets_per_s": T_local / dt / 1e9, "tets": T_local, "frac": bpt * T_local / dt / 1e9 / peak}
            del alone
            torch.cuda.empty_cache()
        except Exception as exc:  # pragma: no cover - secondary record
            kern["snh (alone, same mesh)"] = {"error": f"{type(exc).__name__}: {exc}"}
    del g1, h1
    # measured DRAM traffic of the dominant kernel (ncu --set full capture, profiles/traffic.json): per launch when the
    # capture was taken at this size on one GPU, else the capture's bytes per tet x this rank's tets
    traffic, traffic_note = None, None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            tj = json.loads(tf.read_text())
            if world == 1 and f"{args.dtype}:{dom}:n{n}" in tj:
                traffic = tj[f"{args.dtype}:{dom}:n{n}"]
                traffic_note = "dram__bytes_read.sum + dram__bytes_write.sum of one launch at this size (ncu --set full)"
            else:
                per_tet = tj.get("per_tet", {}).get(f"{args.dtype}:{dom}")
                if per_tet is not None:
                    traffic = per_tet * T_local
                    traffic_note = (f"{per_tet:.1f} B/tet measured by ncu at the size named in profiles/traffic.json, times this "
                                    f"rank's {T_local} tets")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["gbs"] / peak, "traffic": traffic, "traffic_note": traffic_note,
                "kernel": f"fem_pipe_kernel<{args.dtype},{dom},fun|grad|hess_prod>", "peak_source": peak_src,
                "frac_of_nominal_8TBs": kern[dom]["gbs"] / 8000.0,
                "note": "per GPU: this rank's slab, slowest rank" if world > 1 else None, "per_kernel": kern}

    # ---- parity at size: rank 0 against the C restatement of the reference on 4 hex layers ----
    parity = None
    if not args.no_parity:
        try:
            parity = parity_check(args, wl, res, kinds, dtype, dev, world, rank)
        except Exception as exc:  # pragma: no cover - keep the headline line
            parity = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- e2e: host buffers through the public API ----
    e2e = None
    if not args.no_e2e:
        try:
            e2e = bench_e2e(args, wl, model, sharded, adapter if world == 1 else None, dtype, dev, w, T_total, flush,
                            max_over_ranks, barrier, len(pots))
        except Exception as exc:  # pragma: no cover - keep the headline line
            e2e = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- HVP alone, fp32 and fp64 (configs[4]) ----
    hvp = None
    if not args.no_hvp:
        try:
            hvp = bench_hvp(args, wl, kinds, dev, world, flush, max_over_ranks, T_total, peak)
        except Exception as exc:  # pragma: no cover
            hvp = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- PNCG iterations/s on the same (sharded) mesh ----
    pncg_big = None
    if not args.no_pncg:
        try:
            pncg_big = bench_pncg_sharded(args, wl, pots, dtype, dev, w, world, max_over_ranks, barrier)
        except Exception as exc:  # pragma: no cover
            pncg_big = {"error": f"{type(exc).__name__}: {exc}"}

    clocks.__exit__(None, None, None)
    # release the headline model before the secondary sections
    del model, pots, sharded, res
    if world == 1:
        del adapter, out_block, grad, prod, fun
    torch.cuda.empty_cache()

    # ---- config 2 (N = 1): 975,560 tets, operators + roofline + 200 PNCG iterations ----
    config2 = None
    if world == 1 and not args.no_config2:
        try:
            config2 = bench_config2(args, dtype, dev, w, flush, peak)
        except Exception as exc:  # pragma: no cover
            config2 = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- CPU baseline: the C restatement of the reference on a bounded sample, host cores of this box ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            a, b = sample_layers(n, 1)
            np_dtype = np.float32 if args.dtype == "f32" else np.float64
            val, Ts, dt, threads = time_oracle(host_sample(n, a, b), kinds, 5, 1, np_dtype)
            cpu = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"hex layers [{a}, {b}) of the same cube = {Ts} tets, 5 evaluations (fun, grad, hess_prod passes "
                             f"per potential), C restatement of the reference kernels (oracle/c), {args.dtype}"}
        except Exception as exc:  # pragma: no cover
            cpu = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_name(n, kinds), "assembly": args.scatter,
                       "l2": "256 MiB flush write between timed steps (working set per GPU >> 126 MB L2)" if not args.no_flush else "no flush",
                       "setup": f"mesh, Morton order, tiling and static planes built on the GPU: {setup_s:.2f} s (slowest rank)",
                       "parallelism": ("1 GPU" if world == 1 else
                                       f"{world} ranks x slab of {n // world}+ hex layers (~{T_total // world} tets per GPU, STRONG "
                                       f"scaling, fixed mesh); per step one element pass, then the halo sum of grad+HVP and the "
                                       f"energy reduction "
                                       + (("through peer memory over NVLink (apl_xchg_*): boundary tiles, push, interior tiles, pull, "
                                           "then one push + pull of the partial scalars" if sharded_peer_overlap else
                                           "through peer memory over NVLink: one push + one pull kernel (apl_xchg_*)")
                                          if args.transport == "peer" else
                                          "as pack / NCCL all-to-all / unpack / all-reduce")
                                       + ("; step replayed as one CUDA graph" if (world > 1 and graph is not None) else ""))},
            "setup_s": setup_s,
            "clocks": clocks.summary(), "e2e": e2e,
            "gpu_launches": args.steps * launches_per_step,
            "roofline": roofline, "parity": parity, "hvp": hvp, "cpu_baseline": cpu, "config2": config2,
            "pncg": {"headline_mesh": pncg_big, "config2": (config2 or {}).get("pncg") if isinstance(config2, dict) else None},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def parity_check(args, wl, res, kinds, dtype, dev, world, rank):
    """Rank 0: gradient / HVP of the timed step on the grid planes whose sums are complete inside a 4-layer host sample
    (for N > 1 the sample straddles rank 0's slab interface: its shared plane is halo-summed), and the energy of the
    sample through a second GPU handle, against the C restatement of the reference (fp64)."""
    import torch

    from apple_b200 import _lib
    from apple_b200.warp.model import WarpModel

    if rank != 0:
        return None
    n = args.n
    a, b = sample_layers(n, world)
    sample = host_sample(n, a, b)
    pots_o, _ = oracle_model(sample, kinds)
    e_o, g_o, h_o = oracle_eval(pots_o, sample)
    # planes a+1 .. b-1 have every incident tet inside the sample; rank 0 holds planes <= its upper interface
    top = n if world == 1 else n // world
    keep = (sample.layer > a) & (sample.layer < b) & (sample.layer <= top)
    if a == 0:
        keep |= sample.layer == 0
    gid = sample.vgid[keep]
    # rows of rank 0's result with those global ids
    vg = wl.mesh.vertex_gid
    order = torch.argsort(vg)
    pos = torch.searchsorted(vg[order], torch.as_tensor(gid, device=dev))
    rows = order[pos]
    assert bool((vg[rows] == torch.as_tensor(gid, device=dev)).all())
    rel = lambda x, y: float(np.abs(np.asarray(x, np.float64) - y).max() / np.abs(y).max())  # noqa: E731
    out = {"grad": rel(res["grad"][rows].cpu().numpy(), g_o[keep]),
           "hess_prod": rel(res["prod"][rows].cpu().numpy(), h_o[keep]),
           "rows": int(keep.sum()), "planes": [int(x) for x in np.unique(sample.layer[keep])],
           "halo_plane_included": bool(world > 1 and top in set(np.unique(sample.layer[keep]).tolist()))}
    # energy: the sample as its own small model on this GPU
    sw = SimpleNamespace(mesh=SimpleNamespace(cells=torch.as_tensor(sample.mesh.cells, device=dev),
                                              points=torch.as_tensor(sample.mesh.points, device=dev)),
                         mu=torch.as_tensor(sample.mesh.cell_data["mu"], dtype=dtype, device=dev),
                         la=torch.as_tensor(sample.mesh.cell_data["lambda"], dtype=dtype, device=dev))
    spots = device_potentials(sw, kinds, dtype, fuse=not args.no_fuse)
    f = torch.zeros(1, dtype=dtype, device=dev)
    WarpModel(spots).eval(_lib.OP_FUN, torch.as_tensor(sample.u, dtype=dtype, device=dev), None, fun=f)
    out["energy"] = rel(f.cpu().numpy(), e_o)
    out["against"] = (f"C restatement of the reference (oracle/c, fp64) on hex layers [{a}, {b}) = {sample.mesh.n_cells} tets; "
                      f"tolerance of the north star: 1e-5 relative in fp32, 1e-10 in fp64")
    out["within_tolerance"] = bool(max(out["grad"], out["hess_prod"], out["energy"]) < (1e-5 if dtype == torch.float32 else 1e-10))
    return out


def bench_e2e(args, wl, model, sharded, adapter, dtype, dev, w, T_total, flush, max_over_ranks, barrier, n_pots):
    """Host (pinned) u, p in -> host energy, gradient, HVP out, copies inside the timed region; per rank its slab."""
    import torch

    from apple_b200 import _lib

    V = wl.mesh.n_points
    OPS = _lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_PROD
    uh, ph = wl.u.cpu().pin_memory(), wl.p.cpu().pin_memory()
    gh, hh = torch.empty((V, 3), dtype=dtype).pin_memory(), torch.empty((V, 3), dtype=dtype).pin_memory()
    fh = torch.empty(1, dtype=dtype).pin_memory()

    if sharded is None:
        def step_streamed():
            adapter.fun_grad_hess_prod_host(uh, ph, out=(fh, gh, hh))

        def step_serial():
            u_d = uh.to(dev, non_blocking=True); p_d = ph.to(dev, non_blocking=True)
            f, g, h = adapter.fun_grad_hess_prod(u_d, p_d)
            fh.copy_(f.reshape(1), non_blocking=True); gh.copy_(g, non_blocking=True); hh.copy_(h, non_blocking=True)
    else:
        def step_serial():
            u_d = uh.to(dev, non_blocking=True); p_d = ph.to(dev, non_blocking=True)
            r = sharded.eval(OPS, u_d, p_d)
            fh.copy_(r["fun"].reshape(1), non_blocking=True); gh.copy_(r["grad"], non_blocking=True)
            hh.copy_(r["prod"], non_blocking=True)
        step_streamed = None

    steps = max(3, min(args.steps, 10))

    def time_it(fn):
        for _ in range(2):
            fn()
        barrier()
        tt = 0.0
        for _ in range(steps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            tt += a.elapsed_time(b)
        return max_over_ranks(tt)

    tt_serial = time_it(step_serial)
    api, note, tt, launches = "serial", None, tt_serial, n_pots + (2 if sharded is not None else 0)
    if step_streamed is not None:
        ref = (fh.clone(), gh.clone(), hh.clone())
        try:
            fh.zero_(); gh.zero_(); hh.zero_()
            tt_streamed = time_it(step_streamed)
            err = max(float((x - y).abs().max() / y.abs().max()) for x, y in zip((fh, gh, hh), ref))
            if err < 1e-4 and tt_streamed <= tt_serial:
                api, tt, launches = "streamed", tt_streamed, 2 * n_pots
            elif err < 1e-4:
                note = f"streamed entry point verified but slower here ({tt_streamed / steps:.4f} ms per step); serial number reported"
            else:
                note = f"streamed entry point disagreed with the serial one (rel. err {err:.2e}); serial number reported"
        except Exception as exc:  # pragma: no cover
            note = f"streamed entry point failed ({type(exc).__name__}: {exc}); serial number reported"
    return {"value": T_total * steps / (tt * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": int(2 * V * 3 * w), "d2h_bytes_per_step": int(2 * V * 3 * w + w),
            "bytes_are": "per rank (its slab)" if sharded is not None else "whole mesh",
            "ms_per_step": tt / steps, "steps": steps,
            "api": {"streamed": "WarpModelAdapter.fun_grad_hess_prod_host(u_host, p_host, out=host tensors): copies "
                                "overlapped with a fun+grad pass and a hess_prod pass on side streams",
                    "serial": ("u.to(device); p.to(device); " + ("ShardedOperators.eval" if sharded is not None else
                               "WarpModelAdapter.fun_grad_hess_prod") + "; copy_ to host")}[api],
            "gpu_launches_per_step": launches, "serial_ms_per_step": tt_serial / steps, "note": note}


def bench_hvp(args, wl, kinds, dev, world, flush, max_over_ranks, T_total, peak):
    """hess_prod alone (the kernel an adjoint solve repeats), fp32 and fp64, same mesh and partition."""
    import torch

    from apple_b200 import _lib
    from apple_b200.warp.model import WarpModel

    out = {}
    V, T_local = wl.mesh.n_points, wl.mesh.n_cells
    for name, dt, w in (("f32", torch.float32, 4), ("f64", torch.float64, 8)):
        w2 = SimpleNamespace(shard=wl.shard, mesh=wl.mesh, mu=wl.mu.to(dt), la=wl.la.to(dt))
        pots = device_potentials(w2, kinds, dt, fuse=not args.no_fuse)
        model = WarpModel(pots)
        u, p = wl.u.to(dt).contiguous(), wl.p.to(dt).contiguous()
        if world > 1:
            from apple_b200.dist import ShardedOperators

            sh = ShardedOperators(model, wl.shard, dev, dt, transport=args.transport, overlap=False)
            fn = lambda: sh.eval(_lib.OP_HESS_PROD, u, p)  # noqa: E731
        else:
            prod = torch.zeros((V, 3), dtype=dt, device=dev)

            def fn():
                prod.zero_()
                model.eval(_lib.OP_HESS_PROD, u, p, prod=prod, zero=False)
        for _ in range(3):
            flush(); fn()
        torch.cuda.synchronize()
        reps, tt = max(5, min(args.steps, 10)), 0.0
        for _ in range(reps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            tt += a.elapsed_time(b)
        dtm = max_over_ranks(tt / reps) * 1e-3
        bpt = sum(bytes_per_tet(k.split("+"), w, V / max(T_local, 1), 9) for k in pots)   # read u, p; write Hp
        out[name] = {"value": T_total / dtm, "unit": "tets/s", "ms_per_step": dtm * 1e3, "bytes_per_tet": bpt,
                     "gbs_per_gpu": bpt * T_local / dtm / 1e9, "roofline_frac": bpt * T_local / dtm / 1e9 / peak}
        del pots, model
        torch.cuda.empty_cache()
    out["note"] = ("one hess_prod evaluation of the whole model per step (halo sum included for N > 1), CUDA events, L2 "
                   "flushed; algorithmic bytes = 16 + 9w + materials + (V/T) 9w per tet")
    return out


def bench_pncg_sharded(args, wl, pots, dtype, dev, w, world, max_over_ranks, barrier):
    """PNCG iterations/s on the benchmark mesh itself (all ranks): fixed base z = 0, device-side iteration with the
    peer-memory exchanges inside the native workspace, WHILE-node CUDA graph per iteration."""
    import torch

    from apple_b200.dist import ShardedPNCG
    from apple_b200.optim.pncg import ConvergenceCriteria

    n = args.n
    X = wl.mesh.points
    free = torch.ones((wl.mesh.n_points, 3), dtype=torch.bool, device=dev)
    free[X[:, 2] == 0.0] = False
    h = 1.0 / n
    u0 = (0.05 * h * torch.sin(7.0 * X[:, [1, 2, 0]] + 0.3)).to(dtype)
    u0[~free] = 0.0
    iters = max(10, min(args.pncg_iters, 50))
    crit = ConvergenceCriteria(max_steps=iters + 30, target_relative_gradient_norm=0.0)
    out = {}
    for graph in (2, 0):
        sp = ShardedPNCG(list(pots.values()), [], wl.shard, free, u0, criteria=crit, transport=args.transport if world > 1 else None,
                         use_graph=graph)
        sp.iterate(10)                                  # warm-up: graph capture
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        sp.iterate(iters)                               # one host read of the scalars at the end
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dt = max_over_ranks(max(e0.elapsed_time(e1) * 1e-3, wall))
        s = sp._read()
        from apple_b200 import _lib

        out[{2: "graph_while", 0: "eager"}[graph]] = {
            "iters": iters, "seconds": dt, "iters_per_s": iters / dt, "accepted_total": int(s[_lib.S_N_ACCEPTED]),
            "energy": float(s[_lib.S_F]), "transport": sp.transport}
        del sp
        torch.cuda.empty_cache()
    out["iters_per_s"] = max(v["iters_per_s"] for v in out.values())
    T_total = 5 * n ** 3
    out["tets_per_s_equivalent"] = out["iters_per_s"] * 2 * T_total     # two element passes per iteration (A and B)
    out["note"] = (f"{iters} iterations after 10 warm-up iterations on the {T_total}-tet mesh; max over ranks of CUDA-event / "
                   "wall time around enqueue-and-sync; every exchange of the iteration is a device-side peer-memory kernel pair")
    return out


def bench_config2(args, dtype, dev, w, flush, peak):
    """BASELINE.json configs[1] on one GPU: 58^3 x 5 tets, SNH + ARAP fused, operators + 200 PNCG iterations."""
    import torch

    from apple_b200 import _lib
    from apple_b200.warp.fem import fuse_potentials
    from apple_b200.warp.model import WarpModel
    from apple_b200.warp.model._adapter import zeros_block

    kinds = args.potentials.split(",")
    mesh, u, p = build_mesh(N_CONFIG2)
    T, V = mesh.n_cells, mesh.n_points
    pots = {k: cuda_potential(k, mesh, dtype, name=k) for k in kinds}
    if not args.no_fuse:
        pots = fuse_potentials(pots)
    model = WarpModel(pots)
    ud = torch.as_tensor(u, dtype=dtype, device=dev).contiguous()
    pd = torch.as_tensor(p, dtype=dtype, device=dev).contiguous()
    OPS = _lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_PROD
    out_block, (grad, prod), (fun,) = zeros_block(V, 2, 1, dtype, dev)

    def step():
        out_block.zero_()
        model.eval(OPS, ud, pd, fun=fun, grad=grad, prod=prod, zero=False)

    for _ in range(5):
        flush(); step()
    torch.cuda.synchronize()
    tt = 0.0
    for _ in range(args.steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record()
        torch.cuda.synchronize()
        tt += a.elapsed_time(b)
    ms = tt / args.steps
    bpt = sum(bytes_per_tet(k.split("+"), w, V / T, 12) for k in pots)
    out = {"workload": f"cube {N_CONFIG2}^3x5 = {T} tets / {V} verts, {'+'.join(kinds)}, fused energy+grad+HVP (BASELINE.json configs[1])",
           "value": T / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "bytes_per_tet": bpt,
           "roofline_frac": bpt * T / (ms * 1e-3) / 1e9 / peak}
    if not args.no_pncg:
        try:
            out["pncg"] = bench_pncg(args, mesh, pots, dtype, dev, w)
        except Exception as exc:  # pragma: no cover
            out["pncg"] = {"error": f"{type(exc).__name__}: {exc}"}
    if args.sweep:
        sweep(args, mesh, u, p, dtype, dev, flush)
    return out


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
    h = 1.0 / N_CONFIG2
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
                    for scatter in (0, 2, 1):
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
