#!/usr/bin/env python
"""Phase timeline of the tensor-core edge kernel, generic version (works for both kernel variants): runs forwards on a
synthetic structure with pesto_debug_edge_timeline on and prints, for CTA 0 and the model's last layer (nn = 64), every
clock stamp of a tile relative to the tile's start, in time order, plus the aux warp's state stamps (RMMA kernel)."""
import argparse, json, os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from pesto_b200 import _lib                                          # noqa: E402
from pesto_b200.model import Model                                   # noqa: E402
from pesto_b200.data_encoding import extract_topology                # noqa: E402
from pesto_b200.synth import synth_structure, one_hot_features       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--atoms", type=int, default=32768)
ap.add_argument("--tiles", type=int, default=48)
ap.add_argument("--layers", type=int, default=32, help="use the first LAYERS layers of i_v4_1: the timeline is the LAST launch's "
                "(8 -> nn = 8, 16 -> nn = 16, 24 -> nn = 32; needs a -DPESTO_PROF_ALL_NN build, profiles/variants.py)")
a = ap.parse_args()
g = os.path.join(REPO, "tests", "golden")
cfg = json.load(open(os.path.join(g, "config_i_v4_1.json")))
cfg["sum"] = cfg["sum"][:a.layers]
model = Model(cfg, mode="f16x3")
model.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(g, "weights_i_v4_1.npz")).items()
                       if not k.startswith("sum.") or int(k.split(".")[1]) < a.layers})
print(f"last layer: nn = {cfg['sum'][-1]['nn']}")
model = model.eval().cuda()
X, el, rid = synth_structure(a.atoms, 20230419)
Xd = X.cuda()
ids1 = extract_topology(Xd, 64)[0] + 1
q0, ridd = one_hot_features(el).cuda(), rid.int().cuda()
lib = _lib.load()
NS = 19
buf = torch.zeros(a.tiles * 4 * NS + a.tiles * 2 * 10, dtype=torch.int64, device="cuda")
z = model(Xd, ids1, q0, ridd, n_res=int(rid.max()) + 1)              # warm-up
lib.pesto_debug_edge_timeline(buf.data_ptr(), a.tiles)
z = model(Xd, ids1, q0, ridd, n_res=int(rid.max()) + 1)
torch.cuda.synchronize()
lib.pesto_debug_edge_timeline(None, 0)
raw = buf.cpu().numpy().astype(np.float64)
t = raw[:a.tiles * 4 * NS].reshape(a.tiles, 2, 2, NS)
aux = raw[a.tiles * 4 * NS:].reshape(a.tiles, 2, 10)
labels = {0: "tile start", 4: "T_j loads issued", 5: "E1 done", 6: "barrier B passed", 7: "M2 issued", 8: "E2 done", 9: "barrier C passed",
          10: "M3 results (first chunk) there", 18: "E3 done", 11: "S0 stores of next tile done", 12: "barrier D passed",
          13: "E1 chunks done (before EP)", 14: "EP of previous tile done", 15: "M1 of next tile issued, T prefetched (tile end)",
          16: "tile end (CUDA-core reduction kernel)"}
n_ok = int((t[..., 0] > 0).any(axis=(1, 2)).sum())
t = t[2:n_ok - 1]
aux = aux[2:n_ok - 1]
t[t == 0] = np.nan
aux[aux == 0] = np.nan
print(f"tiles used {t.shape[0]}; mean cycles since the tile's start, CTA 0")
print(f"{'stamp':52s} " + " ".join(f"H{h}g{gg}".rjust(8) for h in range(2) for gg in range(2)))
rel = t - t[..., :1]
order = np.argsort(np.nan_to_num(np.nanmean(rel[:, 0, 0, :], axis=0), nan=1e18))
for k in order:
    if k not in labels or np.isnan(rel[..., k]).all():
        continue
    print(f"{k:2d} {labels[k]:49s} " + " ".join(f"{np.nanmean(rel[:, h, gg, k]):8.0f}" for h in range(2) for gg in range(2)))
per = np.diff(t[:, :, :, 0], axis=0)
print(f"{'tile period':52s} " + " ".join(f"{np.nanmean(per[:, h, gg]):8.0f}" for h in range(2) for gg in range(2)))
if not np.isnan(aux).all():
    names = ["ids loaded", "gather 0 issued", "gather 1 issued", "Ra Rb Rc0 issued", "Rc1 issued", "gather 2 issued", "gather 3 issued",
             "Rc2 issued", "Rc3 issued (dr_full commit)"]
    print("aux warp, cycles relative to the E3-done stamp (18) of the SAME tile's group 0 (negative = before):")
    for k, nm in enumerate(names):
        print(f"   {nm:32s} " + " ".join(f"{np.nanmean(aux[:, h, k] - t[:, h, 0, 18]):9.0f}" for h in range(2)))
    print("   EP of that tile done (group 1)   " + " ".join(f"{np.nanmean(t[1:, h, 1, 14] - t[:-1, h, 0, 18]):9.0f}" for h in range(2)))
