"""Dry run of bench.py's N>1 flow: 2 processes over gloo, torch.cuda mocked, oracle-backed fake potentials that
implement mark_boundary / part by cell subsets."""
import sys, os, contextlib, json, io, time, copy, numpy as np, torch, torch.distributed as dist
ROOT=str(__import__('pathlib').Path(__file__).resolve().parents[2]); sys.path.insert(0,ROOT); sys.path.insert(0,ROOT+'/tests')
real_device=torch.device
class Ev:
    def __init__(self,enable_timing=False): self.t=None
    def record(self): self.t=time.perf_counter()
    def elapsed_time(self,o): return max((o.t-self.t)*1e3,1e-3)
class FakeStream:
    def __init__(self,*a,**k): pass
    def wait_stream(self,s): pass
    def wait_event(self,e): pass
    def record_event(self): return Ev()
    def synchronize(self): pass
    cuda_stream=0
for name,val in dict(set_device=lambda *a,**k: None, synchronize=lambda *a,**k: None, Event=Ev, Stream=FakeStream,
                     current_stream=lambda *a,**k: FakeStream(), stream=lambda s: contextlib.nullcontext(),
                     device=lambda d: contextlib.nullcontext(), is_available=lambda: True, current_device=lambda: 0).items():
    setattr(torch.cuda,name,val)
torch.Tensor.pin_memory=lambda self: self
torch.device=lambda *a,**k: real_device("cpu")
_init=dist.init_process_group
dist.init_process_group=lambda backend,**kw: _init("gloo")
import helpers
from apple_b200 import _lib
class FakePot:
    def __init__(self,kind,mesh,dtype,name=None,**kw):
        self.o={0:helpers.oracle_potential(kind,mesh)}; self.device=real_device("cpu"); self.dtype=dtype; self.name=name or kind; self.n_points=mesh.n_points
    def mark_boundary(self,flags):
        flags=np.asarray(flags).astype(bool); pot=self.o[0]; touch=flags[pot.cells].any(1)
        for part,sel in ((1,touch),(2,~touch)):
            sub=copy.copy(pot); sub.cells,sub.dhdX,sub.dV=pot.cells[sel],pot.dhdX[sel],pot.dV[sel]
            sub.materials={k:np.asarray(v)[sel] for k,v in pot.materials.items()}; self.o[part]=sub if sel.any() else None
        return int(touch.sum())
    def eval(self,ops,u,p=None,*,fun=None,quad=None,grad=None,diag=None,prod=None,scatter=None,part=0):
        o=self.o[part]
        if o is None: return
        un=u.numpy().astype(np.float64); pn=None if p is None else p.numpy().astype(np.float64); V=self.n_points
        if ops&_lib.OP_FUN: e=np.zeros(1); o.fun(un,e); fun+=float(e[0])
        if ops&_lib.OP_GRAD: g=np.zeros((V,3)); o.grad(un,g); grad+=torch.from_numpy(g).to(grad.dtype)
        if ops&_lib.OP_HESS_PROD: h=np.zeros((V,3)); o.hess_prod(un,pn,h); prod+=torch.from_numpy(h).to(prod.dtype)
_fake=lambda kind,mesh,dtype,**kw: FakePot(kind,mesh,dtype,**kw)
import apple_b200.warp.fem as wf
wf.fuse_potentials=lambda pots: pots
import bench
bench.cuda_potential=_fake
mode=sys.argv[1]
sys.argv=["bench.py","--gpus","2","--n","5","--steps","2","--warmup","1","--no-flush"]+(["--slab"] if mode=="slab" else [])
buf=io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
if int(os.environ["RANK"])==0:
    line=json.loads(buf.getvalue().strip().splitlines()[-1])
    print(mode, "value",line["value"],"scaling",line["scaling"],"launches",line["gpu_launches"]); print(line["config"]["parallelism"]); print(line["config"]["workload"])
