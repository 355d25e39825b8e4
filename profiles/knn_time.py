import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from pesto_b200.data_encoding import batch_topology
from pesto_b200.synth import synth_structure
g = dict(np.load("tests/golden/pdbs_test_53.npz"))
X = torch.from_numpy(g["X"]).cuda(); sizes = [int(s) for s in g["sizes"]]
def timed(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("pdbs_test 53 structures:", round(timed(lambda: batch_topology(X, sizes, 64)), 3), "ms")
Xs = synth_structure(32768, 7)[0].cuda()
print("synthetic 32768 chain:", round(timed(lambda: batch_topology(Xs, [32768], 64)), 3), "ms")
Xb = torch.cat([synth_structure(8192, 100 + s)[0] for s in range(32)]).cuda()
print("32 x 8192:", round(timed(lambda: batch_topology(Xb, [8192] * 32, 64)), 3), "ms")
