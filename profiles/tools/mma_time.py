"""Which shared-memory operand layouts the tensor core reads at full speed: cycles per tcgen05.mma (M = 128, K = 16, fp16) in a
chain of 64, for A in TMEM / K-major / MN-major with and without swizzle (pesto_debug_mma_time).  GPU box only."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pesto_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
NM = 64
def run(label, N, a, b, a_tmem=0):
    # a / b = (major_mn, layout, lbo, sbo, kstep)
    for _ in range(2):
        _lib.check(lib.pesto_debug_mma_time(NM, N, *a, *b, a_tmem, out.data_ptr(), None), label)
        torch.cuda.synchronize()
    o = out.cpu().tolist()
    print(f"{label:78s} N={N:3d}: issue {o[0] / NM:6.1f}  total {o[1] / NM:6.1f} cycles per MMA")
KB = lambda N: (0, 0, N * 16, 128, 2 * N * 16)                       # B K-major, no swizzle: [K/8][N][8] (the weight images)
for N in (16, 32, 64, 128):
    run("A TMEM, B K-major none (the weight GEMMs)", N, (0, 0, 0, 0, 0), KB(N), 1)
    run("A K-major none [K/8][128][8], B K-major none", N, (0, 0, 2048, 128, 4096), KB(N))
for N in (16, 32):
    BMN = (1, 0, 128, 2048, 256)                                     # B MN-major none: chunk (n/8) stride 2048, k-group stride 128
    run("A MN none, chunk stride 2048 (first RMMA kernel: V16), B MN none", N, (1, 0, 128, 2048, 256), BMN)
    run("A MN none, chunk stride 512 (first RMMA kernel: ring), B MN none", N, (1, 0, 128, 512, 256), BMN)
    run("A MN none, chunk stride 2048+128 (padded), B MN none", N, (1, 0, 128, 2176, 256), BMN)
    run("A MN SW128 (64 ch x 8 k atoms; lbo = next 64 ch, sbo = 1024), B MN none", N, (1, 2, 16384, 1024, 2048), BMN)
    run("A MN SW64 (32 ch x 8 k atoms; lbo = 8192, sbo = 512) [probe], B MN none", N, (1, 4, 8192, 512, 1024), BMN)
    run("A MN SW128, B K-major none", N, (1, 2, 16384, 1024, 2048), KB(N))
    run("A K-major none, B MN none", N, (0, 0, 2048, 128, 4096), BMN)
    run("A TMEM, B MN none", N, (0, 0, 0, 0, 0), BMN, 1)
    run("A TMEM, B MN SW32 (16 kinds x 8 k atoms, sbo = 256)", 16, (0, 0, 0, 0, 0), (1, 6, 4096, 256, 512), 1)
