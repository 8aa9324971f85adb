import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import ubu_b200 as U
from ubu_b200 import workloads as WL, sharding as SH
n=1_000_000; chunk=250_000
w=WL.uniform("cfg2", n, 76, 500, seed=1002, paired=True)
al=U.SwAligner(device=0, slots=3)
def pin(a):
    p=U.pinned_empty(a.shape,a.dtype); p[...]=a; return p
def pin_set(sq): return U.SeqSet(pin(sq.packed), pin(sq.off), pin(sq.len), None, None)
bounds=[(lo,min(lo+chunk,n)) for lo in range(0,n,chunk)]
chunks=[]
for lo,hi in bounds:
    wsl,ix=SH.window_slice(w.windows,w.pair_ref_idx,lo,hi)
    chunks.append((pin_set(w.reads.slice(lo,hi)),pin_set(wsl),pin(ix)))
res=U.pinned_empty((n,),U.RESULT_DTYPE); pools=[U.pinned_empty((6*chunk+4096,),np.uint32) for _ in bounds]
for rep in range(6):
    ts=[]; tw=[]; t0=time.perf_counter(); infl=[]
    for ci,(lo,hi) in enumerate(bounds):
        if len(infl)==al.slots:
            a=time.perf_counter(); al.wait(infl.pop(0)); tw.append(time.perf_counter()-a)
        a=time.perf_counter(); infl.append(al.submit(chunks[ci][0],chunks[ci][1],chunks[ci][2],res[lo:hi],pools[ci])); ts.append(time.perf_counter()-a)
    for t in infl:
        a=time.perf_counter(); al.wait(t); tw.append(time.perf_counter()-a)
    tot=time.perf_counter()-t0
    print(f"total {tot*1e3:.2f} ms submit {[round(x*1e3,2) for x in ts]} wait {[round(x*1e3,2) for x in tw]}")
