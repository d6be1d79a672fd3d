import ctypes as ct, numpy as np, torch, sys
sys.path.insert(0,'.')
import bench
from infercnv_b200.device import Engine
from infercnv_b200 import _lib, dist as shard
eng=Engine(0)
G,C=10000,10000
cs,cl=bench.chr_layout(G)
refs=bench.ref_groups_global(C)
plan=shard.plan_shards(C,refs,1)[0]
X=eng.synth(G,cs,cl,plan.local_cells,C,bench.SEED)
lib=_lib.load()
out=(ct.c_ulonglong*16)()
lib.icnv_debug_stats(out,1)
Y,f=eng.smooth_block(X,cs,cl,plan.local_ref_groups(),plan.ref_sizes,plan.max_chunks)
torch.cuda.synchronize()
lib.icnv_debug_stats(out,1)
tot=sum(out[4:10]) or 1
print("CTA0 cycles by stage [wait, A, B scans, B outputs, C median, D]:", [round(100*out[4+i]/tot,1) for i in range(6)], "total cycles", tot)
mt=sum(out[10:15]) or 1
print("median sub-stages [stats, pivots, count, reduce, gather+rank] %:", [round(100*out[10+i]/mt,1) for i in range(5)], "cycles", mt)
print("histogram medians", out[10]); print("median rounds",out[0],"medians",out[1],"rounds/median",out[0]/max(1,out[1]),"split exits",out[2],"gather exits",out[3])
xs=X[5000].cpu().numpy(); print("cell 5000 raw stats", xs.mean(), (xs==0).mean())
