"""Dry run of WarpModelAdapter.fun_grad_hess_prod_host on the CPU with torch.cuda mocked: catches Python-level
mistakes (names, shapes, kwargs) in the stream choreography that otherwise only a GPU run would reveal."""
import sys, types, contextlib, numpy as np, torch
ROOT=str(__import__('pathlib').Path(__file__).resolve().parents[2]); sys.path.insert(0,ROOT); sys.path.insert(0,ROOT+'/tests')
import apple_b200.warp.model._adapter as A
from apple_b200 import _lib
from helpers import make_case, oracle_potential
from oracle import fem as ofem

class Ev: pass
class FakeStream:
    def __init__(self,*a,**k): self.log=[]
    def wait_stream(self,s): pass
    def wait_event(self,e): pass
    def record_event(self): return Ev()
class FakeCuda:
    Stream=FakeStream
    @staticmethod
    def current_stream(dev=None): return FakeStream()
    @staticmethod
    def stream(s): return contextlib.nullcontext()
    @staticmethod
    def device(d): return contextlib.nullcontext()
class FakeDev:
    type="cuda"; index=0
class TorchProxy:
    def __getattr__(self,k): return getattr(torch,k)
    cuda=FakeCuda
    @staticmethod
    def device(x): return FakeDev() if not isinstance(x,FakeDev) else x
    @staticmethod
    def empty(*a,device=None,**k): return torch.empty(*a,**k)
    @staticmethod
    def zeros(*a,device=None,**k): return torch.zeros(*a,**k)
A.torch=TorchProxy()
torch.Tensor.pin_memory=lambda self: self

mesh,u,p=make_case(n=3,seed=1); V=mesh.n_points
om=ofem.Model([oracle_potential("snh",mesh),oracle_potential("arap",mesh)],V)
class FakePot:
    device="cuda"
class FakeModel:
    potentials={"a":FakePot()}
    def eval(self,ops,u,p=None,*,fun=None,quad=None,grad=None,diag=None,prod=None,scatter=None,part=0,zero=True):
        if zero:
            for o in (fun,quad,grad,diag,prod):
                if o is not None: o.zero_()
        un=u.numpy().astype(np.float64); pn=None if p is None else p.numpy().astype(np.float64)
        if ops&_lib.OP_FUN: fun+=float(om.fun(un))
        if ops&_lib.OP_GRAD: grad+=torch.from_numpy(om.grad(un))
        if ops&_lib.OP_HESS_PROD: prod+=torch.from_numpy(om.hess_prod(un,pn))
ad=A.WarpModelAdapter(FakeModel(),V)
uh=torch.from_numpy(u); ph=torch.from_numpy(p)
for rep in range(2):
    f,g,h=ad.fun_grad_hess_prod_host(uh,ph)
    assert abs(float(f)-om.fun(u))<1e-12*abs(om.fun(u)) and np.allclose(g.numpy(),om.grad(u)) and np.allclose(h.numpy(),om.hess_prod(u,p)), rep
f2,g2,h2=ad.fun_grad_hess_prod(uh,ph)
assert np.allclose(g2.numpy(),om.grad(u)) and abs(float(f2)-om.fun(u))<1e-12*abs(om.fun(u))
f3,g3,d3=ad.fun_grad_hess_diag(uh) if False else (None,None,None)
print("dry run ok: host entry point and fused forms produce the oracle's results twice in a row")
