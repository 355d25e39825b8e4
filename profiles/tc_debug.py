#!/usr/bin/env python
"""Debug aid: per-layer max error of the tensor-core edge kernel against the FFMA kernel on the same input state."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from conftest import load_case, load_config, load_weights
from pesto_b200 import _lib
from pesto_b200.model import Model
from test_gpu_parity import _staged, run_case

m = Model(load_config("i_v4_1")); m.load_state_dict({k: torch.from_numpy(v) for k, v in load_weights("i_v4_1").items()}); m = m.eval().cuda()
c = load_case(sys.argv[1] if len(sys.argv) > 1 else "2CUA_A")
lib, h, n, st0, ids32, geom, node = _staged(m, c)
for mode in (1, 2):
    cur = st0
    for layer in range(lib.pesto_model_num_layers(h)):
        ref = torch.empty_like(cur); out = torch.full_like(cur, float("nan"))
        _lib.check(lib.pesto_state_update(h, layer, n, ids32.data_ptr(), geom.data_ptr(), cur.data_ptr(), ref.data_ptr(), node.data_ptr(), 0, None), "fp32")
        _lib.check(lib.pesto_state_update(h, layer, n, ids32.data_ptr(), geom.data_ptr(), cur.data_ptr(), out.data_ptr(), node.data_ptr(), mode, None), "tc")
        torch.cuda.synchronize()
        d = (out - ref).abs()
        print(f"mode {mode} layer {layer:2d} nn {lib.pesto_model_layer_nn(h, layer):2d} |ref|max {ref.abs().max().item():8.3f} "
              f"err q {d[:, :32].max().item():.3e} p {d[:, 32:].max().item():.3e} nan {int(torch.isnan(out).sum())}", flush=True)
        cur = ref
    z = run_case(m, c, mode={1: "bf16x3", 2: "bf16"}[mode]).cpu()
    print("mode", mode, "logit err vs reference", (z - torch.from_numpy(c["z_i_v4_1"])).abs().max().item(), flush=True)
