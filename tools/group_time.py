import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biem_helmholtz_sphere_b200 import _ops
from biem_helmholtz_sphere_b200.geometry import grid_centers
S=4; dev=torch.device("cuda",0); n_end=16
cen=torch.as_tensor(grid_centers(2,3),device=dev); B=cen.shape[0]; rad=torch.ones(B,dtype=torch.float64,device=dev); N=B*n_end*n_end
ks=torch.linspace(0.5,8.0,256,dtype=torch.float64,device=dev)[:S].contiguous(); eta=torch.ones(S,dtype=torch.float64,device=dev)
dirv=torch.tensor([1.0,0.0,0.0],dtype=torch.float64,device=dev)
A=torch.empty((S,N,N),dtype=torch.complex128,device=dev); bufs=_ops.SolveBuffers(N,1,S)
work=_ops._work(_ops.load().bhs_assemble_workspace(_ops.get_plan(3,n_end).handle,B,S))
f=_ops.rhs_expand(3,n_end,centers=cen,radii=rad,k_in=ks,direction=dirv)
def one():
    _ops.assemble(3,n_end,cen,rad,ks,eta,out=A,work=work)
    r=f.reshape(S,N).clone()
    _ops.launch_count(reset=True)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); _ops.zgesv_batched_(A,r,bufs); e1.record(); torch.cuda.synchronize()
    return _ops.launch_count(reset=True), e0.elapsed_time(e1), A
one()
n,ms,A=one()
v=A.view(torch.float64)
print(os.environ.get("BHS_LU_SKIP"), os.environ.get("BHS_LU_GEMM_ONLY"), "launches", n, f"{ms:.2f} ms per group; nan frac {torch.isnan(v).double().mean().item():.3f} inf frac {torch.isinf(v).double().mean().item():.3f} absmax finite {v[torch.isfinite(v)].abs().max().item():.3e}")
