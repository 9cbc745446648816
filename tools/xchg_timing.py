"""Where the sharded step's time goes: CUDA events around the output memset, the element pass, the push and the pull of
one fused evaluation on the benchmark mesh (torchrun, one rank per GPU).  Prints the median over the steps of every phase
on every rank.   torchrun --nproc-per-node 2 tools/xchg_timing.py [n]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch, torch.distributed as dist
from bench import device_potentials, device_workload
from apple_b200 import _lib
from apple_b200.dist import ShardedOperators
from apple_b200.warp.model import WarpModel
from apple_b200.warp.model._adapter import zeros_block

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 234
dtype = torch.float32
wl = device_workload(n, world, rank, dev, dtype)
model = WarpModel(device_potentials(wl, ["snh", "arap"], dtype))
ops = ShardedOperators(model, wl.shard, dev, dtype, transport="peer")
OPS = _lib.OP_FUN | _lib.OP_GRAD | _lib.OP_HESS_PROD
V = wl.mesh.n_points
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
rows = []
for step in range(25):
    flush.fill_(1)
    dist.barrier(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    buf, (grad, prod), (fun,) = zeros_block(V, 2, 1, dtype, dev)
    ev[1].record()
    model.eval(OPS, wl.u, wl.p, fun=fun, grad=grad, prod=prod, zero=False)
    ev[2].record()
    scal = buf[buf.numel() - 1:]
    ops.halo.push((grad, prod), scal)
    ev[3].record()
    ops.halo.pull((grad, prod), scal)
    ev[4].record()
    torch.cuda.synchronize()
    if step >= 5:
        rows.append([ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(4)] + [ev[0].elapsed_time(ev[4]) * 1e3])
med = np.median(np.array(rows), axis=0)
out = [None] * world
dist.all_gather_object(out, [float(x) for x in med])
if rank == 0:
    print(f"n = {n}, {world} ranks; median microseconds per phase (alloc+memset, element pass, push, pull, total)")
    for r, m in enumerate(out):
        print(f"  rank {r}: " + "  ".join(f"{x:9.1f}" for x in m))
dist.destroy_process_group()
