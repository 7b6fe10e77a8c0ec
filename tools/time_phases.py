import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import pertrenderer_b200 as pb
from pertrenderer_b200 import ops, _cabi
from pertrenderer_b200.dist import CudaStages
dev="cuda:0"
N,HW,K,S=1,128,50,4096
fr,col=pb.synthetic_fragments(N,HW,HW,K,kind="realistic",sigma=1e-3,seed=0,device=dev)
G=torch.randn((N,HW,HW,4),device=dev)
pr=ops.ShadeProblem(pix_to_face=fr.pix_to_face,zbuf=fr.zbuf,dists=fr.dists,colors=col,znear=1.0,zfar=100.0,background=(1.,1.,1.),sigma=1e-3,gamma=1e-2,alpha=1.0,eps=1e-10,S_rast=S,S_agg=S,seed_rast=1,seed_agg=2,s_rast=(0,S//2),s_agg=(0,S//2))
def ev(): return torch.cuda.Event(enable_timing=True)
def run(timed=False):
    st=CudaStages(pr); marks=[]
    def mark(name):
        if timed: e=ev(); e.record(); marks.append((name,e))
    mark("start")
    counts,rsum=st.rast(); mark("rast")
    packed=torch.stack((counts,rsum)); mark("stack")
    hist=st.agg(packed[0],packed[1]); mark("agg")
    img=st.blend(hist); mark("blend")
    p=st.bwd_sample(G); mark("bwd_sample")
    out=st.bwd_finish(G,p); mark("bwd_finish")
    return marks
for _ in range(5): run()
torch.cuda.synchronize()
t0=time.perf_counter(); 
for _ in range(20): run()
torch.cuda.synchronize(); print("wall per step ms", (time.perf_counter()-t0)/20*1e3)
m=run(True); torch.cuda.synchronize()
for (n0,e0),(n1,e1) in zip(m,m[1:]): print(f"{n1:12s} {e0.elapsed_time(e1):.3f} ms")
