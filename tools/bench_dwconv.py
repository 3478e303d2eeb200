import os, sys, torch
sys.path.insert(0, os.getcwd())
from cmgan_b200.ops import call
B,T,F=4,321,101
M=B*T*F
g=torch.randn(M,256,device='cuda'); dz=torch.randn(M,128,device='cuda'); w=torch.randn(128,31,device='cuda')
dg=torch.empty(M,256,device='cuda'); dw=torch.zeros(128,31,device='cuda'); db=torch.zeros(128,device='cuda')
out=torch.empty(M,128,device='cuda'); bias=torch.zeros(128,device='cuda'); sums=torch.zeros(128,2,dtype=torch.float64,device='cuda')
for axis in (0,1):
    for _ in range(3): call("cmgan_glu_dwconv_fwd", g, w, bias, B,T,F,axis, out, sums)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): call("cmgan_glu_dwconv_fwd", g, w, bias, B,T,F,axis, out, sums)
    e1.record(); torch.cuda.synchronize()
    print("fwd axis",axis, e0.elapsed_time(e1)/10*1e3,"us")
    for _ in range(3): call("cmgan_glu_dwconv_bwd", g, dz, w, B,T,F,axis, dg, dw, db)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): call("cmgan_glu_dwconv_bwd", g, dz, w, B,T,F,axis, dg, dw, db)
    e1.record(); torch.cuda.synchronize()
    print("bwd axis",axis, e0.elapsed_time(e1)/10*1e3,"us")
