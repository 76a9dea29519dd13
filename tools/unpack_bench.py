import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np
from recastnavigation_b200 import capi, geom
n=8214
fields=np.zeros(n, geom.FIELD_DTYPE); fields["width"]=76; fields["height"]=76
C=n*5776
rng=np.random.default_rng(0)
cnt=capi.pinned_empty(C,np.uint8); cnt[:]=rng.integers(0,3,C).astype(np.uint8)
mask=capi.pinned_empty(C//32+2,np.uint32); mask[:]=0xffffffff
out=capi.pinned_empty(C,np.uint32)
out2=np.zeros(C,np.uint32)
for thr in (1,2,4,8,16):
    for o,name in ((out,"pinned"),(out2,"pageable")):
        capi.unpack_cells(cnt,mask,fields,o,thr)
        t0=time.perf_counter()
        for _ in range(3): capi.unpack_cells(cnt,mask,fields,o,thr)
        dt=(time.perf_counter()-t0)/3
        print("threads",thr,name,"%.1f ms  %.1f GB/s out"%(dt*1e3, C*4/dt/1e9))
print("cpus",os.cpu_count())
