#!/usr/bin/env python
"""Throughput of the BASELINE.json configs that are not the bench line (they are parity-test cases; these numbers are
context for profiles/README.md).  config 3: i_v4_0 over 32 synthetic structures of 8192 atoms in one batch;
config 4: i_v4_1 over one synthetic chain of 32 768 atoms.  CUDA events, 3 warm-up + 10 timed forwards, inputs
resident, topology timed separately."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pesto_b200.model import Model                                   # noqa: E402
from pesto_b200.data_encoding import batch_topology                  # noqa: E402
from pesto_b200.synth import synth_structure, one_hot_features, BASE_SEED       # noqa: E402

G = os.path.join(REPO, "tests", "golden")


def load(tag, mode):
    m = Model(json.load(open(os.path.join(G, f"config_{tag}.json"))), mode=mode)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, f"weights_{tag}.npz")).items()})
    return m.eval().cuda()


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run(name, tag, n, n_struct, mode):
    parts = [synth_structure(n, BASE_SEED + s) for s in range(n_struct)]
    X = torch.cat([p[0] for p in parts]).cuda()
    q0 = one_hot_features(torch.cat([p[1] for p in parts])).cuda()
    rid = torch.cat([p[2] + s * ((n + 7) // 8) for s, p in enumerate(parts)]).int().cuda()
    n_res = n_struct * ((n + 7) // 8)
    model = load(tag, mode)
    ms_knn = timed(lambda: batch_topology(X, [n] * n_struct, 64))
    ids1 = batch_topology(X, [n] * n_struct, 64)
    with torch.no_grad():
        ms = timed(lambda: model(X, ids1, q0, rid, n_res=n_res))
    atoms = n * n_struct
    print(json.dumps({"config": name, "model": tag, "mode": mode, "atoms": atoms, "structures": n_struct, "forward_ms": ms,
                      "atoms_per_s": atoms / ms * 1e3, "topology_ms": ms_knn}), flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
    run("configs[2]: i_v4_0, 32 x 8192 synthetic atoms, one batch", "i_v4_0", 8192, 32, mode)
    run("configs[3]: i_v4_1, one synthetic chain of 32768 atoms", "i_v4_1", 32768, 1, mode)
    run("configs[0]-sized: i_v4_1, one structure of 2386 atoms (launch-bound regime)", "i_v4_1", 2386, 1, mode)
