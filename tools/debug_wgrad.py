import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from cmgan_b200.ops import gemm
torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
DEV = "cuda"
M, K, N = 64, 64, 64
for m0 in (0, 1, 9):
    A = torch.zeros(M, K, device=DEV); D = torch.zeros(M, N, device=DEV)
    A[m0] = torch.arange(1, K + 1, device=DEV).float()
    D[m0] = torch.arange(1, N + 1, device=DEV).float() * 100
    for prec in (0, 1):
        dw = torch.zeros(N, K, device=DEV)
        gemm(wgrad=True, W=None, C=dw, ldc=0, A=A, lda=K, Cin=K, D=D, ldd=N, N=N, sb_k=1, sb_n=K, M=M, precision=prec)
        torch.cuda.synchronize()
        print("m0", m0, "prec", prec, "nonzero", int((dw != 0).sum()), "sum", dw.sum().item())
        print(dw[:4, :8])
        print(dw[30:34, 28:36])
