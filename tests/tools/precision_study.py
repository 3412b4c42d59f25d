"""CPU study: how far a float32 / TF32-operand restatement of the reference drifts from the float64
reference over the 100-step BASELINE config-1 run (tests/golden/traj_c1.npz).  Pure numpy; test
infrastructure (imports oracle/).  Results are quoted in DESIGN.md and tests/test_gpu_parity.py."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pycnn_oracle as O
from tests.helpers import load_golden, synth_images, synth_labels
g = load_golden("traj_c1.npz")
S,C,bs,N,steps = 32,10,32,2048,100
layers=[256,128,64]
feats = O.extract_batch(synth_images(N,S,0))
y = O.one_hot(synth_labels(N,C,1),C)
def tf(x, mode):
    if mode is None: return x
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    if mode == 'trunc': u = u & np.uint32(0xFFFFE000)
    else: u = (u + np.uint32(0x0FFF) + ((u >> 13) & 1)) & np.uint32(0xFFFFE000)
    return u.view(np.float32)
def mm(a, b, mode):
    return (tf(a,mode).astype(np.float32) @ tf(b,mode).astype(np.float32)) if mode else a @ b
def run(dtype, opt, lr, mode=None):
    w,b = O.init_layers(feats.shape[1], layers, C, seed=42)
    w={k:v.astype(dtype) for k,v in w.items()}; b={k:v.astype(dtype) for k,v in b.items()}
    m={k:np.zeros_like(v) for k,v in {**w,**b}.items()}; v={k:np.zeros_like(x) for k,x in {**w,**b}.items()}
    X=feats.astype(dtype); losses=[]
    step=0
    while step<steps:
        perm=np.random.permutation(N)
        for s in range(0,N,bs):
            idx=perm[s:s+bs]; x=X[idx]
            acts=[x]; cur=x
            for i in range(3):
                z=mm(cur,w[f"w{i+1}"].T,mode)+b[f"b{i+1}"]; z[z<0]=0; cur=z; acts.append(cur)
            lg=mm(cur,w["wo"].T,mode)+b["bo"]
            e=np.exp(lg-lg.max(1,keepdims=True)); p=e/e.sum(1,keepdims=True)
            losses.append(float(-(y[idx]*np.log(p.astype(np.float64)+1e-9)).sum()))
            dz=(p-y[idx]).astype(dtype); gw={};gb={}
            gw["wo"]=mm(dz.T,acts[-1],mode); gb["bo"]=dz.sum(0); da=mm(dz,w["wo"],mode)
            for i in range(2,-1,-1):
                dz=da.copy(); dz[acts[i+1]<=0]=0
                gw[f"w{i+1}"]=mm(dz.T,acts[i],mode); gb[f"b{i+1}"]=dz.sum(0)
                if i>0: da=mm(dz,w[f"w{i+1}"],mode)
            for P,G in ((w,gw),(b,gb)):
                for k in G:
                    gg=G[k]; n=np.linalg.norm(gg.astype(np.float64))
                    if n>5: gg=gg*dtype(5/n)
                    if opt=="sgd": P[k]=P[k]-dtype(lr)*gg
                    else:
                        m[k]=m[k]*dtype(0.9)+dtype(0.1)*gg; v[k]=v[k]*dtype(0.999)+dtype(0.001)*gg*gg
                        mh=m[k]/dtype(0.1); vh=np.maximum(v[k]/dtype(0.001), dtype(1e-16))
                        P[k]=P[k]-dtype(lr)*mh/(np.sqrt(vh)+dtype(1e-8))
            step+=1
            if step>=steps: break
    return np.array(losses)

for opt,lr in (("sgd",1e-3),("adam",1e-4)):
    for dt in (np.float64, np.float32):
        l=run(dt,opt,lr); want=g[f"{opt}_step_losses"]
        err=np.abs(l-want)/want
        print(opt, dt.__name__, "max rel dev", err.max(), "at", err.argmax(), "dev@20", err[:20].max(), "dev@50", err[:50].max())
    for mode in ("trunc","rne"):
        l=run(np.float32,opt,lr,mode); want=g[f"{opt}_step_losses"]
        err=np.abs(l-want)/want
        print(opt, "tf32-"+mode, "max rel dev", err.max(), "at", err.argmax(), "dev@20", err[:20].max(), "dev@50", err[:50].max(), "mean", err.mean())
