import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, torch, sys
from mdcraft_b200 import _capi, synthetic
N=100000; F=int(sys.argv[1]) if len(sys.argv)>1 else 2
Lg=float(synthetic.cubic_box_length(N,2.5)); Ls=float(synthetic.cubic_box_length(N,0.8))
gr=np.stack([synthetic.uniform_frame(N,Lg,2000+i) for i in range(F)]); sq=np.stack([synthetic.lj_like_frame(N,Ls,3000+i) for i in range(F)])
gd=torch.from_numpy(gr).cuda(); sd=torch.from_numpy(sq).cuda(); box=np.array([Lg,Lg,Lg,90,90,90],np.float32)
A=_capi.Context(0); B=_capi.Context(0, high_priority=True)
for name,(ca,cb) in (("same context",(A,A)),("two contexts",(A,B))):
    rdf=_capi.RdfPlan(ca,200,(0.0,Lg/2),N,None,(1,1))
    sqp=_capi.SqPlan(cb,[N],[(None,None)],n_points=161,L0=[Ls]*3,q_max=20.0)
    for order in ("rdf first","sq first"):
        def step():
            if order=="rdf first":
                rdf.push(gd,None,box,n_frames=F); sqp.push(sd,n_frames=F)
            else:
                sqp.push(sd,n_frames=F); rdf.push(gd,None,box,n_frames=F)
        for _ in range(3): step()
        ca.synchronize(); cb.synchronize()
        t0=time.perf_counter()
        for _ in range(10): step()
        ca.synchronize(); cb.synchronize()
        dt=(time.perf_counter()-t0)/10
        t0=time.perf_counter()
        for _ in range(10):
            step(); ca.synchronize(); cb.synchronize()
        dt2=(time.perf_counter()-t0)/10
        print(f"{name:14s} {order:10s}: {dt*1e3:.2f} ms/step pipelined = {F/dt:.1f} frames/s; {dt2*1e3:.2f} ms/step with a barrier per step = {F/dt2:.1f} frames/s")
    rdf.close(); sqp.close()
