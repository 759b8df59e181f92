import sys,time; sys.path.insert(0,'/root/repo')
import numpy as np, gmmloc_b200 as g
from gmmloc_b200 import synth
ctx=g.Context(0)
for K,p in ((10000,0.5),(50000,0.95),(200000,0.9)):
    mean,cov=synth.make_map(K,p,seed=1)
    t0=time.perf_counter(); m=g.GMM(mean,cov); t1=time.perf_counter()-t0
    print(K, 'create %.3f s'%t1, flush=True)
z=np.load('/root/repo/tests/golden/v1_gmm.npz'); t0=time.perf_counter(); m=g.GMM(z['mean'],z['cov']); print('v1', time.perf_counter()-t0)
