"""Dry run of bench.py's N=1 flow on the CPU: torch.cuda mocked, potentials replaced by oracle-backed fakes.
Catches Python-level mistakes in the flow and in the JSON assembly (not performance, not kernels)."""
import sys, contextlib, json, io, time, numpy as np, torch
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
torch.cuda.set_device=lambda *a,**k: None
torch.cuda.synchronize=lambda *a,**k: None
torch.cuda.Event=Ev
torch.cuda.Stream=FakeStream
torch.cuda.current_stream=lambda *a,**k: FakeStream()
torch.cuda.stream=lambda s: contextlib.nullcontext()
torch.cuda.device=lambda d: contextlib.nullcontext()
torch.cuda.is_available=lambda: True
torch.cuda.current_device=lambda: 0
class Props: uuid="00000000-0000-0000-0000-000000000000"
torch.cuda.get_device_properties=lambda i: Props()
torch.Tensor.pin_memory=lambda self: self
class DevProxy:
    def __call__(self,*a,**k): return real_device("cpu")
import builtins
torch.device=lambda *a,**k: real_device("cpu")
import helpers
from oracle import fem as ofem
from apple_b200 import _lib
class FakePot:
    def __init__(self,kind,mesh,dtype,name=None,**kw):
        self.o=helpers.oracle_potential(kind,mesh); self.device=real_device("cpu"); self.dtype=dtype; self.name=name or kind; self.n_points=mesh.n_points
    def eval(self,ops,u,p=None,*,fun=None,quad=None,grad=None,diag=None,prod=None,scatter=None,part=0):
        un=u.numpy().astype(np.float64); pn=None if p is None else p.numpy().astype(np.float64)
        V=self.n_points
        if ops&_lib.OP_FUN: e=np.zeros(1); self.o.fun(un,e); fun+=float(e[0])
        if ops&_lib.OP_GRAD: g=np.zeros((V,3)); self.o.grad(un,g); grad+=torch.from_numpy(g).to(grad.dtype)
        if ops&_lib.OP_HESS_PROD: h=np.zeros((V,3)); self.o.hess_prod(un,pn,h); prod+=torch.from_numpy(h).to(prod.dtype)
_fake=lambda kind,mesh,dtype,**kw: FakePot(kind,mesh,dtype,**kw)
import apple_b200.warp.fem as wf
wf.fuse_potentials=lambda pots: pots
import bench
bench.cuda_potential=_fake
sys.argv=["bench.py","--n","5","--steps","2","--warmup","1","--no-pncg","--no-flush"]
buf=io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
line=json.loads(buf.getvalue().strip().splitlines()[-1])
print(json.dumps({k:(v if not isinstance(v,dict) else {kk:(vv if not isinstance(vv,(dict,list)) else '...') for kk,vv in v.items()}) for k,v in line.items()},indent=1)[:3500])
