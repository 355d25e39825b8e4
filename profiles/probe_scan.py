import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from pesto_b200 import _lib
from test_gpu_umma import probe
torch.manual_seed(0)
for K,N in [(32,128),(64,64),(32,16),(128,128)]:
    A=(torch.randn(128,K)*3).cuda(); B=torch.randn(N,K).cuda()
    ref1=(A.bfloat16().double()@B.bfloat16().double().T); ref=(A.double()@B.double().T)
    for name,(lbo,sbo) in {'lbo=N*16,sbo=128':(N*16,128),'lbo=128,sbo=N*16':(128,N*16)}.items():
        try:
            D1=probe(A,B,0,lbo,sbo); D3=probe(A,B,1,lbo,sbo)
            print(K,N,name,'x1 err',(D1.double()-ref1).abs().max().item(),'x3 err',(D3.double()-ref).abs().max().item(),'scale',ref.abs().max().item(), flush=True)
        except Exception as e:
            print(K,N,name,'EXC',e, flush=True)
