import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import parity
from mpr_b200 import capi
for model,dim,size in [("hello_world",2,1024),("hello_world",2,2048),("hello_world",2,4096),("prospero",2,1024),("prospero",2,256),("hello_world",3,512)]:
    ctx=capi.Context(size,num_subtapes=6400000); tape=capi.Tape(parity.load_tape(model))
    f=(lambda: ctx.render2D(tape)) if dim==2 else (lambda: ctx.render3D(tape))
    for _ in range(5): f()
    ws=[];gs=[]
    for _ in range(30):
        t0=time.perf_counter(); f(); ws.append((time.perf_counter()-t0)*1e3); gs.append(ctx.stats().gpu_ms)
    print(model,dim,size,"wall mean %.3f min %.3f max %.3f | gpu mean %.3f min %.3f max %.3f"%(np.mean(ws),np.min(ws),np.max(ws),np.mean(gs),np.min(gs),np.max(gs)), flush=True)
    ctx.close()
