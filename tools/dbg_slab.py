import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from animal2vec_b200 import gemm
def rel(a,b): return ((a.float()-b.float()).norm()/(b.float().norm()+1e-20)).item()
for (bsz,t,ng,groups,taps,pad) in [(2,300,64,4,19,9),(1,2000,64,16,19,9),(24,2000,64,16,19,9),(2,257,48,3,7,3)]:
    g=torch.Generator(device="cuda").manual_seed(1)
    x=(torch.randn(bsz,t,groups*64,device="cuda",generator=g)).bfloat16()
    wt=(torch.randn(groups*ng,64,taps,device="cuda",generator=g)*0.05).bfloat16()
    ref=F.conv1d(x.float().transpose(1,2),wt.float(),None,padding=pad,groups=groups).transpose(1,2)
    w=wt.permute(0,2,1).reshape(groups*ng,taps*64).contiguous()
    for dt in (torch.float32, torch.bfloat16, torch.bfloat16, torch.float32):
        o=gemm.conv_slab(x,w,taps=taps,pad=pad,groups=groups,out_dtype=dt)
        torch.cuda.synchronize()
        e=rel(o,ref)
        print((bsz,t,ng,groups,taps), dt, "rel", e)
        if e>1e-2:
            err=(o.float()-ref).abs()
            # per (b, m-tile of 128, group)
            for b in range(min(bsz,2)):
                print(" b",b,[[round(err[b,m*128:(m+1)*128,gg*ng:(gg+1)*ng].max().item(),2) for gg in range(groups)] for m in range((t+127)//128)])
