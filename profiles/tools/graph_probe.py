"""How much of a single small structure's forward is launch overhead: the same forward replayed from a CUDA graph
(torch.cuda.CUDAGraph around Model.forward) against eager launches.  GPU box only."""
import json, os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from pesto_b200.model import Model
from pesto_b200.data_encoding import extract_topology
from pesto_b200.synth import synth_structure, one_hot_features
g = os.path.join(REPO, "tests", "golden")
model = Model(json.load(open(os.path.join(g, "config_i_v4_1.json"))), mode="f16x3")
model.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(g, "weights_i_v4_1.npz")).items()})
model = model.eval().cuda()
for n in (600, 2386, 8192, 32768):
    X, el, rid = synth_structure(n, 7)
    Xd = X.cuda(); ids1 = extract_topology(Xd, 64)[0] + 1; q0 = one_hot_features(el).cuda(); ridd = rid.int().cuda(); nr = int(rid.max()) + 1
    def ev(fn, reps):
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
    with torch.no_grad():
        for _ in range(3): z = model(Xd, ids1, q0, ridd, n_res=nr)
        eager = ev(lambda: model(Xd, ids1, q0, ridd, n_res=nr), 30)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(2): model(Xd, ids1, q0, ridd, n_res=nr)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            zg = model(Xd, ids1, q0, ridd, n_res=nr)
        gr.replay(); torch.cuda.synchronize()
        graph = ev(gr.replay, 30)
        print(f"N={n}: eager {eager:.3f} ms ({n / eager / 1e3:.2f} M atoms/s), graph {graph:.3f} ms ({n / graph / 1e3:.2f} M atoms/s), equal {bool(torch.equal(z, zg))}", flush=True)
