#!/usr/bin/env python
"""Small profiling target: kNN + N forwards of i_v4_1 on one synthetic structure (default 8192 atoms, the
BASELINE north_star size).  Used under ncu; never a bench number."""
import argparse
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pesto_b200.model import Model                                   # noqa: E402
from pesto_b200.data_encoding import extract_topology                # noqa: E402
from pesto_b200.synth import synth_structure, one_hot_features       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--atoms", type=int, default=8192)
ap.add_argument("--forwards", type=int, default=1)
ap.add_argument("--mode", default="fp32")
ap.add_argument("--tag", default="i_v4_1")
a = ap.parse_args()
g = os.path.join(REPO, "tests", "golden")
model = Model(json.load(open(os.path.join(g, f"config_{a.tag}.json"))), mode=a.mode)
model.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(g, f"weights_{a.tag}.npz")).items()})
model = model.eval().cuda()
X, el, rid = synth_structure(a.atoms, 20230419)
Xd = X.cuda()
ids1 = extract_topology(Xd, 64)[0] + 1
q0, ridd = one_hot_features(el).cuda(), rid.int().cuda()
for _ in range(a.forwards):
    z = model(Xd, ids1, q0, ridd, n_res=int(rid.max()) + 1)
torch.cuda.synchronize()
print("ok", tuple(z.shape), float(z.abs().max()))
