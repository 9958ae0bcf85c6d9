import sys; sys.path.insert(0,'/root/repo')
from same_b200 import _lib as L
import numpy as np
from same_b200.device import Section
a=np.random.rand(200000,2); p=np.random.rand(200000,3)
with Section(a,a,p,p) as s, s.batch() as b:
    b.candidates(0.01,8,False,1.0); b.sync()
print("after work", L.mempool_stats(0))
L.mempool_reserve(0, 512<<20)
print("after reserve 512M", L.mempool_stats(0))
with Section(a,a,p,p) as s, s.batch() as b:
    b.candidates(0.01,8,False,1.0); b.sync()
print("after more work", L.mempool_stats(0))
