import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from optimal_conv_b200 import hec, params as PR, synth
Q,P=PR.Q_SET6[:2],PR.P_PACK; N=1<<16; B=16
w=synth.conv_workload(Q,P,16,B,1)
ctx=hec.Context(16,Q,P)
mono=np.zeros((16,N),dtype=np.uint64)
for i in range(16):
    m=np.zeros(N,dtype=np.uint64); m[1<<i]=1; mono[i]=ctx.ntt(m,0)
ker=[ctx.upload_pt(w["pt_ker"][i],PR.SCALE) for i in range(B)]
idx=[ctx.upload_pt(mono[i:i+1],1.0) for i in range(16)]
bias=ctx.upload_pt(w["bias"][None,:],PR.SCALE)
for j,k in w["keys"].items(): ctx.upload_swk((1<<(j+1))+1,k,0)
ct=ctx.upload_ct(*w["ct"][0],PR.SCALE)
plan=ctx.plan(ker,1,PR.SCALE,PR.SCALE,idx,bias,1)
for _ in range(3): plan.run([ct])
ctx.sync()
import statistics
ts=[]
for _ in range(20):
    ctx.timer_start(); plan.run([ct]); ts.append(ctx.timer_stop_ms())
print("graph latency ms", statistics.median(ts))
pr=plan.profile([ct]); pr=plan.profile([ct])
print([ (n, round(ms*1000,1)) for n,ms in pr]); print("sum us", sum(ms for _,ms in pr)*1000)
