#!/usr/bin/env python
"""Edge-kernel time per nn on the bench workload (bench.edge_kernel_ms: back-to-back launches between CUDA events), for the
library PESTO_B200_LIB points at; status words are not checked, so timing-only variants with wrong results can be measured."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench                                                           # noqa: E402
from pesto_b200 import _lib                                            # noqa: E402
from pesto_b200.data_encoding import batch_topology                    # noqa: E402
from pesto_b200.model import Model                                     # noqa: E402
from pesto_b200.synth import one_hot_features                          # noqa: E402

wl = bench.load_workload()
dev = torch.device("cuda", 0)
with open(os.path.join(bench.GOLDEN, "config_i_v4_1.json")) as fh:
    model = Model(json.load(fh), mode="f16x3")
model.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights().items()})
model = model.eval().to(dev)
Xd = wl["X"].to(dev)
ids1 = batch_topology(Xd, [int(n) for n in wl["sizes"]], 64)
q0d = one_hot_features(wl["el"]).to(dev)
lib = _lib.load()
print(os.environ.get("PESTO_B200_LIB", "default"),
      " ".join(f"{nn}:{bench.edge_kernel_ms(lib, model, dev, Xd, ids1, q0d, _lib.MODES['f16x3'], nn=nn):.4f}" for nn in (8, 16, 32, 64)))
