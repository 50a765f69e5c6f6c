import sys; sys.path.insert(0,'.')
import numpy as np, yalla_b200 as yb
from yalla_b200 import workloads
new, ref = yb.product(), yb.reference()
n=30000
rng=np.random.default_rng(n)
X=np.zeros((n,5),dtype=np.float32); X[:,:5]=workloads.polarized_ball(n,0.8,rng,lattice=True)
gs=workloads.grid_size_for(n,0.8)
sims=[lib.sim("epithelium",n,gs,1.0) for lib in (new,ref)]
for s in sims: s.set_state(X)
for step in range(1,13):
    outs=[]
    for s in sims:
        s.step(0.05,1); outs.append(s.get_state())
    d=np.abs(outs[0].astype(np.float64)-outs[1]); 
    i=np.unravel_index(np.argmax(d), d.shape)
    print(step, "max err", d.max(), "at", i, "lane max", d.max(axis=0), "scale", np.abs(outs[1]).max())
