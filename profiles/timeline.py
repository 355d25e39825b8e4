#!/usr/bin/env python
"""Phase timeline of the tensor-core edge kernel (debug build aid): runs one forward on a synthetic structure with
pesto_debug_edge_timeline switched on and prints, for CTA 0, the mean clock cycles each phase of a tile takes
(last layer of the model, i.e. nn = 64 for i_v4_1)."""
import argparse, json, os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pesto_b200 import _lib                                          # noqa: E402
from pesto_b200.model import Model                                   # noqa: E402
from pesto_b200.data_encoding import extract_topology                # noqa: E402
from pesto_b200.synth import synth_structure, one_hot_features       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--atoms", type=int, default=32768)
ap.add_argument("--mode", default="f16x3")
ap.add_argument("--tiles", type=int, default=48)
a = ap.parse_args()
g = os.path.join(REPO, "tests", "golden")
model = Model(json.load(open(os.path.join(g, "config_i_v4_1.json"))), mode=a.mode)
model.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(g, "weights_i_v4_1.npz")).items()})
model = model.eval().cuda()
X, el, rid = synth_structure(a.atoms, 20230419)
Xd = X.cuda()
ids1 = extract_topology(Xd, 64)[0] + 1
q0, ridd = one_hot_features(el).cuda(), rid.int().cuda()
lib = _lib.load()
NS = 19
buf = torch.zeros((a.tiles, 2, 2, NS), dtype=torch.int64, device="cuda")
z = model(Xd, ids1, q0, ridd, n_res=int(rid.max()) + 1)              # warm-up
lib.pesto_debug_edge_timeline(buf.data_ptr(), a.tiles)
z = model(Xd, ids1, q0, ridd, n_res=int(rid.max()) + 1)
torch.cuda.synchronize()
lib.pesto_debug_edge_timeline(None, 0)
t = buf.cpu().numpy().astype(np.float64)
names = ["S0 compute", "barrier A", "(T_j: cp.async, no wait)", "wait M1", "E1 compute", "barrier B", "wait M2", "E2 compute", "barrier C",
         "wait M3", "E3 compute(+pj issue)", "barrier D", "R loop", "barrier E + P", "combine", "barrier G + T issue"]
valid = t[..., 0] > 0
n_ok = int(valid.any(axis=(1, 2)).sum())
t = t[2:n_ok]
t[t == 0] = np.nan                              # a half that ran fewer tiles (or was switched off in an experiment)
extra = t[..., 17:]
t = t[..., :17]
d = np.diff(t, axis=-1)                      # [tiles, half, grp, 16]
print(f"tiles used: {t.shape[0]}; cycles per tile (mean over tiles), CTA 0")
print(f"{'phase':28s} " + " ".join(f"H{h}g{gg:1d}".rjust(8) for h in range(2) for gg in range(2)))
for k, nm in enumerate(names):
    print(f"{nm:28s} " + " ".join(f"{np.nanmean(d[:, h, gg, k]):8.0f}" for h in range(2) for gg in range(2)))
print(f"{'  E3: p_j prefetch issue':28s} " + " ".join(f"{np.nanmean((extra[:, h, gg, 0] - t[:, h, gg, 10])):8.0f}" for h in range(2) for gg in range(2)))
print(f"{'  E3: arithmetic':28s} " + " ".join(f"{np.nanmean((extra[:, h, gg, 1] - extra[:, h, gg, 0])):8.0f}" for h in range(2) for gg in range(2)))
tot = t[:, :, :, -1] - t[:, :, :, 0]
print(f"{'tile total':28s} " + " ".join(f"{np.nanmean(tot[:, h, gg]):8.0f}" for h in range(2) for gg in range(2)))
per = np.diff(t[:, :, :, 0], axis=0)
print(f"{'tile period':28s} " + " ".join(f"{np.nanmean(per[:, h, gg]):8.0f}" for h in range(2) for gg in range(2)))
